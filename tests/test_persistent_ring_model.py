"""CPU model of the TMA ring of reverse_fused_persistent_kernel (adept-2_b200/csrc/fused_persistent_kernels.cuh).

A CTA of the resident kernel pulls the records of its items through a ring of kStages shared-memory slots, one
mbarrier per slot, and fetches the first piece of its NEXT item while it works on the current one.  Items are of three
kinds: a chunk that is one piece; a long chunk (several pieces, streamed once per set of column blocks) owned by the
first of its `split` parts; and the other parts of a long chunk, which take no piece at all.  The model replays the
kernel's issue/wait sequence for random item sequences and checks what the hardware would not forgive: every wait has
exactly one matching copy in flight in its slot, with the right data, in the phase its parity says, and no slot is
given a second transaction before the first one was consumed (the bug this model found: the prologue fetched a piece
for a first item that takes none, and the look-ahead then armed the same mbarrier again).
"""
import random

K = 4   # kStages


def run(seed):
    rng = random.Random(seed)
    items = []
    for _ in range(rng.randint(1, 60)):
        kind = rng.choice(["single", "long", "pass", "long", "pass"])
        if kind == "single":
            items.append(("single", 1, 1))
        elif kind == "long":
            items.append(("long", rng.randint(2, 6), rng.randint(1, 3)))     # pieces, sets of column blocks
        else:
            items.append(("pass", rng.randint(2, 6), 1))
    phase = [0] * K          # completed phases of every slot's mbarrier
    pending = [None] * K     # (sequence number, what was copied) in flight in the slot

    def issue(seq, tag):
        st = seq % K
        assert pending[st] is None, f"slot {st} armed twice: {pending[st]} and {seq}"
        assert seq // K == phase[st], f"piece {seq} is use {seq // K} of slot {st}, whose barrier completed {phase[st]} phases"
        pending[st] = (seq, tag)

    def wait(seq, tag):
        st = seq % K
        assert pending[st] is not None and pending[st][0] == seq, f"wait for piece {seq}: nothing in flight"
        assert pending[st][1] == tag, "the slot holds another item's records"
        pending[st] = None
        phase[st] += 1

    pc, s_pre = 0, -1
    if items[0][0] != "pass":                       # prologue
        issue(0, (0, 0))
        s_pre = 0
    for k, (kind, n_pieces, n_sets) in enumerate(items):
        total = 1 if kind == "single" else (n_pieces * n_sets if kind == "long" else 0)
        npc = 1 if kind == "single" else n_pieces
        for it in range(min(total, K)):             # my own pieces, less the one already in flight
            if not (it == 0 and s_pre == pc):
                issue(pc + it, (k, it % npc))

        def look_ahead():
            nonlocal s_pre
            if k + 1 < len(items) and total < K and items[k + 1][0] != "pass":
                issue(pc + total, (k + 1, 0))
                s_pre = pc + total

        if kind == "single":
            wait(pc, (k, 0))
            look_ahead()                            # after the activity bytes are in
        elif kind == "pass":
            look_ahead()
        else:
            look_ahead()
            for it in range(total):
                wait(pc + it, (k, it % npc))
                if it + K < total:
                    issue(pc + it + K, (k, (it + K) % npc))
        pc += total
    assert all(p is None for p in pending), "a copy outlives the CTA"


def test_ring_protocol_random_item_sequences():
    for seed in range(4000):
        run(seed)
