"""CPU model of the item enumeration and the wait/arrive protocol of reverse_fused_persistent_kernel
(adept-2_b200/csrc/fused_persistent_kernels.cuh: PCursor, pcursor_advance, the PItem transitions).

The resident kernel has no grid-wide barrier: every CTA enumerates its own (chunk, column group) items of the whole
sweep, counts itself in at cnt[l] after its last item of step l and, before its first item of a step, waits for the
participants of the non-empty step above.  This model runs G such CTAs under random interleavings and checks that
 * every item of every step is processed exactly once,
 * no item of step l starts before every item of every step above l is finished (the dependency the level schedule
   needs: steps run downwards in a reverse sweep),
 * nobody waits forever (the CTA working on the highest unfinished step can always proceed).
"""
import random

import pytest


class Cursor:
    """Mirror of PCursor + pcursor_advance for CTA b of G."""

    def __init__(self, b, G, fchunk, gy, n_levels):
        self.b, self.G, self.fchunk, self.gy = b, G, fchunk, gy
        self.c0 = self.n_items = self.it = 0
        self.l = n_levels + 1
        self.l_prev = 0
        self.parts_prev = 0
        self.rot = 0
        self.advance()

    def advance(self):
        G = self.G
        self.it += G
        if self.it < self.n_items:
            return
        while True:
            if self.n_items > 0:
                self.l_prev = self.l
                self.parts_prev = min(self.n_items, G)
                if self.n_items >= G:
                    self.rot = (self.rot + self.n_items) % G
            l = self.l - 1
            c0 = c1 = 0
            while l >= 1:
                c0, c1 = self.fchunk[l], self.fchunk[l + 1]
                if c1 > c0:
                    break
                l -= 1
            if l < 1:
                self.l, self.n_items = 0, 0
                return
            self.l, self.c0 = l, c0
            self.n_items = (c1 - c0) * self.gy
            self.it = (self.b + G - self.rot) % G
            if self.it < self.n_items:
                return


def cta_program(b, G, fchunk, gy, n_levels):
    """The CTA's sequence of actions, as the kernel derives them: ('wait', l, n), ('item', l, c, cby), ('arrive', l)."""
    cu = Cursor(b, G, fchunk, gy, n_levels)
    prog = []
    if cu.l and cu.parts_prev:
        prog.append(("wait", cu.l_prev, cu.parts_prev))
    while cu.l:
        l = cu.l
        prog.append(("item", l, cu.c0 + cu.it // gy, cu.it % gy))
        cu.advance()
        if cu.l != l:
            prog.append(("arrive", l))
            alone = cu.l_prev == l and cu.parts_prev == 1
            if cu.l != 0 and cu.parts_prev != 0 and not alone:
                prog.append(("wait", cu.l_prev, cu.parts_prev))
    return prog


def run(G, widths, gy, seed):
    """widths[l-1] = chunks of step l (l = 1 .. n_levels); steps are swept from n_levels down to 1."""
    n_levels = len(widths)
    fchunk = [0] * (n_levels + 2)
    for l in range(1, n_levels + 1):
        fchunk[l + 1] = fchunk[l] + widths[l - 1]
    progs = [cta_program(b, G, fchunk, gy, n_levels) for b in range(G)]
    pos = [0] * G
    cnt = [0] * (n_levels + 2)
    done_items = {}          # (l, c, cby) -> CTA
    finished_in_step = [0] * (n_levels + 2)
    rng = random.Random(seed)
    live = [b for b in range(G) if progs[b]]
    while live:
        movable = []
        for b in live:
            op = progs[b][pos[b]]
            if op[0] != "wait" or cnt[op[1]] >= op[2]:
                movable.append(b)
        assert movable, "deadlock: every unfinished CTA waits"
        b = rng.choice(movable)
        op = progs[b][pos[b]]
        if op[0] == "item":
            _, l, c, cby = op
            assert (l, c, cby) not in done_items, "item processed twice"
            assert fchunk[l] <= c < fchunk[l + 1] and 0 <= cby < gy
            for up in range(l + 1, n_levels + 1):      # everything above is complete
                assert finished_in_step[up] == widths[up - 1] * gy, f"step {l} started before step {up} was finished"
            done_items[(l, c, cby)] = b
            finished_in_step[l] += 1
        elif op[0] == "arrive":
            cnt[op[1]] += 1
        pos[b] += 1
        if pos[b] == len(progs[b]):
            live.remove(b)
    assert len(done_items) == sum(widths) * gy
    for l in range(1, n_levels + 1):                   # the counters end at parts(l) = min(G, items)
        assert cnt[l] == min(G, widths[l - 1] * gy)


@pytest.mark.parametrize("G", [1, 2, 7, 16, 61])
@pytest.mark.parametrize("gy", [1, 2, 5])
def test_random_step_widths(G, gy):
    rng = random.Random(G * 100 + gy)
    for trial in range(6):
        n_levels = rng.choice([1, 2, 9, 40])
        widths = [rng.choice([0, 0, 1, 1, 2, 3, 8, 30, 100]) for _ in range(n_levels)]
        run(G, widths, gy, seed=trial)


def test_scalar_chains_between_wide_steps_need_no_global_wait():
    # steps of ONE item between wide steps (Toon's boundary statements): the CTA that runs a chain of them was the only
    # participant of the step above and must not wait for anybody
    G, gy = 8, 1
    widths = [20, 1, 1, 1, 20, 1, 20]
    n_levels = len(widths)
    fchunk = [0] * (n_levels + 2)
    for l in range(1, n_levels + 1):
        fchunk[l + 1] = fchunk[l] + widths[l - 1]
    waits_on_single = 0
    for b in range(G):
        for op in cta_program(b, G, fchunk, gy, n_levels):
            if op[0] == "wait" and op[2] == 1:
                waits_on_single += 1
    # steps 4 -> 3 -> 2 and 6 are run by the SAME CTA, which never waits on itself; only the other CTAs of the wide
    # steps 5 and 1 wait for the single participant of the step above them
    assert waits_on_single == 2 * (G - 1)
    run(G, widths, gy, seed=3)


def test_toon_like_shape():
    # config 2's shape: thousands of steps of ~225 chunks, two column groups, 592 resident CTAs
    rng = random.Random(5)
    widths = [rng.choice([225, 226, 224, 2, 1]) for _ in range(60)]
    run(592, widths, 2, seed=1)
