"""CPU model of the round-2 window walk (adept-2_b200/csrc/tape_kernels.cuh, window_levels_kernel): the warp probes FOUR
consecutive statements per round against the per-slot (writer step, reader step) table as it was when the round began,
and commits the longest prefix in which no statement shares a slot with an earlier statement of the round; the rest is
probed again.  The claim is that this gives exactly the steps of the one-at-a-time rule (oracle_levels mode 2):
    step(s) = max( writer(x)+1 and readers(x) over operands x, max(writer, readers)(lhs)+1, 1 )
Here the rounds are replayed sequentially in numpy/Python and compared with the oracle on tapes with random slots (no
conflicts to speak of), stencils (every neighbour conflicts), self references, repeated operands and long statements."""
import numpy as np
import pytest


def walk_in_rounds(stmt, idx, group=4, max_ops=7):
    stmt = np.asarray(stmt, np.int64)
    S = stmt.shape[0] - 1
    wl, rl = {}, {}                      # slot -> step of the last writer / max step of the readers of the live version
    level = np.zeros(S + 1, np.int64)
    rounds = 0

    def one(s, W, R):
        ops = idx[stmt[s - 1, 1]:stmt[s, 1]]
        lhs = int(stmt[s, 0])
        v = 1
        for x in ops:
            v = max(v, W.get(int(x), -1) + 1, R.get(int(x), -1))
        v = max(v, max(W.get(lhs, -1), R.get(lhs, -1)) + 1)
        return v, [int(x) for x in ops], lhs

    s = 1
    while s <= S:
        rounds += 1
        if stmt[s, 1] - stmt[s - 1, 1] > max_ops:      # a long statement takes the whole warp, alone
            v, ops, lhs = one(s, wl, rl)
            for x in ops:
                rl[x] = v
            wl[lhs], rl[lhs] = v, 0
            level[s] = v
            s += 1
            continue
        batch = []
        for g in range(group):
            if s + g > S or stmt[s + g, 1] - stmt[s + g - 1, 1] > max_ops:
                break
            batch.append(s + g)
        probed = [one(t, wl, rl) for t in batch]            # every probe sees the table as the round found it
        seen, ncommit = set(), 0
        for (v, ops, lhs) in probed:
            mine = set(ops) | {lhs}
            if ncommit > 0 and (mine & seen):
                break
            seen |= mine
            ncommit += 1
        for t, (v, ops, lhs) in zip(batch[:ncommit], probed[:ncommit]):
            for x in ops:
                rl[x] = v
            wl[lhs], rl[lhs] = v, 0
            level[t] = v
        s += ncommit
    return level, rounds


def _tapes(pkg):
    rng = np.random.default_rng(3)
    yield "synthetic", pkg.tapegen.synthetic_tape(4000, 12, 32, seed=4)
    yield "toon", pkg.tapegen.advection_tape("toon", 40, 12)
    yield "lw", pkg.tapegen.advection_tape("lw", 33, 9)
    yield "rosenbrock", pkg.tapegen.rosenbrock_tape(300)
    for k in range(4):
        yield f"random{k}", pkg.tapegen.random_small_tape(rng, 1500, 25 + 40 * k, max_ops=9, n_indep=4, n_dep=3, lhs_any_prob=0.2)


def test_rounds_of_four_give_the_sequential_steps(po, pkg):
    for name, t in _tapes(pkg):
        S = t["stmt"].shape[0] - 1
        _, want = po.levels(t["stmt"], t["idx"], t["max_gradient"], 2, window=max(S + 1, 64), huge=1 << 30, pack=0)
        got, rounds = walk_in_rounds(t["stmt"], t["idx"])
        assert np.array_equal(got, want), name
        assert rounds <= S
        if name == "synthetic":
            assert rounds < 0.35 * S        # random slots: nearly four statements per round
