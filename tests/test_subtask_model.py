"""CPU model of how the fused reverse kernels cut a staged piece into sub-tasks and share it between CTAs
(adept-2_b200/csrc/replay_kernels.cuh: fused_first_head, fused_subtask_range, fused_single_piece).

A piece is a sequence of target groups (a head record followed by its contributions).  Sub-task s is
[first head >= s*sub, first head >= (s+1)*sub); CTA `part` of `nparts` owns the sub-tasks [part*nsub/nparts,
(part+1)*nsub/nparts).  Whatever the group lengths, sub and nparts: the sub-tasks must partition the piece at group
boundaries (a group's running sum lives in one warp's registers), and the CTAs' record ranges -- for which each CTA
fetches the activity bytes and clears the ZMARK bytes -- must be exactly the union of their sub-tasks.
"""
import random


def head_mask_words(heads, n):
    words = [0] * ((n + 31) // 32)
    for i in range(n):
        if heads[i]:
            words[i >> 5] |= 1 << (i & 31)
    return words


def first_head(headm, pos, n):                     # fused_first_head
    blk = pos >> 5
    while (blk << 5) < n:
        w = headm[blk]
        if blk == (pos >> 5):
            w &= (0xFFFFFFFF << (pos & 31)) & 0xFFFFFFFF
        if w:
            i = (blk << 5) + ((w & -w).bit_length() - 1)
            return min(i, n)
        blk += 1
    return n


def subtask_range(headm, n, sub, s):               # fused_subtask_range
    i0 = 0 if s == 0 else first_head(headm, s * sub, n)
    i1 = n if (s + 1) * sub >= n else first_head(headm, (s + 1) * sub, n)
    return i0, i1


def cta_range(headm, n, sub, part, nparts):        # fused_single_piece, nparts > 1
    nsub = (n + sub - 1) // sub
    s_lo, s_hi = part * nsub // nparts, (part + 1) * nsub // nparts
    i0 = 0 if s_lo == 0 else first_head(headm, s_lo * sub, n)
    i1 = n if s_hi * sub >= n else first_head(headm, s_hi * sub, n)
    if s_hi <= s_lo:
        i1 = i0
    return s_lo, s_hi, i0, i1


def test_subtasks_partition_a_piece_at_group_heads():
    rng = random.Random(11)
    for trial in range(3000):
        n = rng.randint(1, 256)
        heads = [False] * n
        i = 0
        while i < n:                               # the piece begins with a head; groups of 1 .. 40 records
            heads[i] = True
            i += rng.choice([1, 1, 2, 3, 5, 6, 9, 40])
        headm = head_mask_words(heads, n)
        sub = rng.choice([4, 5, 8, 16, 32, 64, 256])
        nsub = (n + sub - 1) // sub
        covered = 0
        for s in range(nsub):
            i0, i1 = subtask_range(headm, n, sub, s)
            assert i0 == covered or i0 >= i1, "sub-tasks leave a gap or overlap"
            if i0 < i1:
                assert heads[i0] and (i1 == n or heads[i1]), "a sub-task does not begin and end at a head"
                covered = i1
        assert covered == n
        nparts = rng.choice([1, 2, 3, 4, 8, 16])
        covered = 0
        for part in range(nparts):
            s_lo, s_hi, i0, i1 = cta_range(headm, n, sub, part, nparts)
            lo = None
            for s in range(s_lo, s_hi):            # the CTA's record range is the union of its sub-tasks
                j0, j1 = subtask_range(headm, n, sub, s)
                if j0 < j1:
                    assert i0 <= j0 and j1 <= i1
                    lo = j0 if lo is None else lo
            assert i0 == covered or i0 >= i1
            if i0 < i1:
                covered = i1
        assert covered == n
