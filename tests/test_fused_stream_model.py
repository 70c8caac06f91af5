"""CPU model of the FUSED reverse stream ("versioned rows") that adept-2_b200/csrc builds on the device
(tape_kernels.cuh: fused_* kernels; replay_kernels.cuh: reverse_fused_kernel), checked against the oracle.

The level-scheduled reverse sweep used to need two kernels per step: one parks a_s = g[lhs_s] and zeroes
g[lhs_s] for every statement of the step, one gathers the contributions per target slot.  The fused form
needs ONE: every slot owns two physical rows, row(X, p) = X + p*G, and the row that holds the live value
flips each time the sweep passes a statement whose LHS is X.  A step then only READS live rows and WRITES
either the target's own live row in place (X is not an LHS of the step, so nobody else reads it) or the
other row of an LHS slot (so the a_s every other target of the step still needs stays intact):

    ZMARK(row(lhs_s, p^1))  OP(m, row(lhs_s, p))*      target X = lhs_s: starts from +0.0 (self references follow)
    MARK (row(X, p))        OP(m, row(lhs_s', p'))*    any other target: starts from its live value

with p = number of statements later in the tape that write the slot, mod 2.  Everything here is the
sequential statement of what the device does with two radix sorts and three scans; the arithmetic order
per slot is the reference's (adept/jacobian.cpp:410-428), so the result is bit-identical.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, same_bits

OP, MARK, ZMARK = 0, 1, 2


def fused_stream(stmt, mult, idx, G, level):
    """Returns (kind, row, m, step_first) of the fused stream and fpar (final row parity per slot)."""
    stmt = np.asarray(stmt, np.int64)
    idx = np.asarray(idx, np.int64)
    S, O = stmt.shape[0] - 1, idx.size
    R = S + O
    end = stmt[:, 1]
    o = np.arange(O)
    s_of_op = np.searchsorted(end, o, side="right")                 # smallest s with end[s] > o
    s_ids = np.arange(1, S + 1)
    # position in the reference's reverse sweep: HEAD(lhs_s) then the operands of s ascending, s descending
    pos_op = (S - s_of_op) + (O - end[s_of_op]) + 1 + (o - end[s_of_op - 1])
    pos_l = (S - s_ids) + (O - end[s_ids])
    key = np.empty(R, np.int64)
    val = np.empty(R, np.int64)                                       # element id: op o -> o, LHS of s -> O + s - 1
    key[pos_op], val[pos_op] = idx, o
    key[pos_l], val[pos_l] = stmt[1:, 0], O + s_ids - 1
    # sort 1: by slot (stable) -> per slot, references in sweep order; count the writers already passed
    order = np.argsort(key, kind="stable")
    key, val = key[order], val[order]
    isl = (val >= O).astype(np.int64)
    E = np.cumsum(isl) - isl
    first = np.r_[True, key[1:] != key[:-1]]
    last = np.r_[key[1:] != key[:-1], True]
    e0 = np.zeros(G, np.int64)
    e0[key[first]] = E[first]
    cnt = E - e0[key]
    par = np.zeros(R, np.int64)
    par[val] = cnt & 1
    fpar = np.zeros(G, np.int64)
    fpar[key[last]] = (cnt[last] + isl[last]) & 1
    # sort 2: by step (stable) -> (step, slot, sweep order)
    s_of_val = np.where(val < O, s_of_op[np.minimum(val, max(O - 1, 0))] if O else 0, val - O + 1)
    step = np.asarray(level, np.int64)[s_of_val]
    order = np.argsort(step, kind="stable")
    step, val, slot, s_of_val = step[order], val[order], key[order], s_of_val[order]
    flag = np.r_[True, (step[1:] != step[:-1]) | (slot[1:] != slot[:-1])]
    is_op = val < O
    assert np.all(flag[~is_op]), "an LHS reference must open its (step, slot) group"
    hflag = flag & is_op
    pos = np.arange(R) + np.cumsum(hflag)
    n_rec = R + int(hflag.sum())
    kind = np.empty(n_rec, np.int64)
    row = np.empty(n_rec, np.int64)
    m = np.zeros(n_rec)
    rstep = np.empty(n_rec, np.int64)
    lhs = stmt[s_of_val, 0]
    lhs_par = par[O + s_of_val - 1]
    # operations: read a_s from the live row of their statement's LHS
    kind[pos[is_op]] = OP
    row[pos[is_op]] = lhs[is_op] + lhs_par[is_op] * G
    m[pos[is_op]] = np.asarray(mult)[val[is_op]]
    # inserted heads: the target's live row
    kind[pos[hflag] - 1] = MARK
    row[pos[hflag] - 1] = slot[hflag] + par[val[hflag]] * G
    rstep[pos[hflag] - 1] = step[hflag]
    # LHS references: the other row, starting from zero
    kind[pos[~is_op]] = ZMARK
    row[pos[~is_op]] = lhs[~is_op] + (lhs_par[~is_op] ^ 1) * G
    rstep[pos] = step
    return kind, row, m, rstep, fpar


def replay_fused(kind, row, m, rstep, fpar, G, g0):
    """One adjoint sweep of gradient vector(s) g0[G, ...] through the fused stream, with the device's
    activity-flag semantics (replay_kernels.cuh, reverse_fused_kernel): act[row] == False means "the row is
    all +0.0 whatever memory holds".  A ZMARK only clears the flag of the row it opens; a group is stored
    (and its flag set) only when an ACTIVE contribution reached it; every read -- the sweep's and the final
    extraction's -- goes through the flag.  Memory rows that are never stored keep stale data on purpose."""
    g = np.zeros((2 * G,) + g0.shape[1:])
    g[:G] = g0
    act = np.zeros(2 * G, bool)
    act[:G] = np.any(g0.reshape(G, -1) != 0.0, axis=1)      # the seeds
    stale = np.full_like(g[0], 777.0)                        # what an unguarded read of a dead row would see
    for st in np.unique(rstep)[::-1]:
        sel = np.nonzero(rstep == st)[0]
        old, old_act = g.copy(), act.copy()   # every read of a step sees rows and flags as they were when it began
        cur = acc = None
        dirty = False

        def retire():
            if cur is not None and dirty:
                g[cur] = acc
                act[cur] = True

        for i in sel:
            if kind[i] != OP:
                retire()
                cur = row[i]
                dirty = False
                if kind[i] == ZMARK:
                    act[cur] = False          # new version of an LHS slot: +0.0 until a contribution lands
                    acc = np.zeros_like(old[cur])
                else:
                    acc = old[cur].copy() if old_act[cur] else np.zeros_like(old[cur])
            elif old_act[row[i]]:
                acc = acc + m[i] * old[row[i]]   # multiply, round, add, round (numpy does not contract)
                dirty = True
        retire()
    rows = np.arange(G) + fpar * G
    out = g[rows]
    out[~act[rows]] = 0.0                     # extract_kernel consults the flag
    g[~act] = stale                           # (the stale memory must not be what the caller sees)
    return out


def reference_adjoint(po, t, g0):
    return po.adjoint(t["stmt"], t["mult"], t["idx"], g0)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("window", [0, 64])
def test_fused_stream_reproduces_the_reference_adjoint(po, name, window):
    g = load_golden(name)
    if g["stmt"].shape[0] > 20000:
        pytest.skip("python model: small tapes only")
    G = g["max_gradient"]
    _, level = po.levels(g["stmt"], g["idx"], G, 2, window=window or 65536, huge=65536, pack=1)
    ks = fused_stream(g["stmt"], g["mult"], g["idx"], G, level)
    got = replay_fused(*ks, G, np.asarray(g["g_ad_in"], np.float64))
    assert same_bits(got, g["g_ad_out"])


@pytest.mark.parametrize("seed", range(4))
def test_fused_stream_random_tapes(po, pkg, seed):
    rng = np.random.default_rng(500 + seed)
    t = pkg.tapegen.random_small_tape(rng, int(rng.integers(50, 1500)), int(rng.integers(8, 120)),
                                      n_indep=4, n_dep=3)
    G = t["max_gradient"]
    _, level = po.levels(t["stmt"], t["idx"], G, 2, window=97, huge=65536, pack=1)
    ks = fused_stream(t["stmt"], t["mult"], t["idx"], G, level)
    g0 = rng.uniform(-1, 1, (G, 3))          # three directions at once
    got = replay_fused(*ks, G, g0)
    for d in range(3):
        assert same_bits(got[:, d], reference_adjoint(po, t, np.ascontiguousarray(g0[:, d])))


def overwritten_independent_tape():
    """X is an independent AND an LHS, overwritten three times with a break in the self-reference
    chain:  s1 X = 2*X;  s2 X = const;  s3 X = 3*X;  s4 Y = X.  dY/dX(initial) = 0, but a fused sweep that
    extracts row memory instead of (flag ? memory : 0) returns the stale 3 of two versions earlier."""
    X, Y = 0, 1
    stmt = np.array([[-1, 0], [X, 1], [X, 1], [X, 2], [Y, 3]], np.int32)
    return dict(stmt=stmt, mult=np.array([2.0, 3.0, 1.0]), idx=np.array([X, X, X], np.int32), max_gradient=3,
                indep=np.array([X], np.int32), dep=np.array([Y], np.int32))


def test_fused_stream_keeps_dead_rows_dead(po):
    t = overwritten_independent_tape()
    G = t["max_gradient"]
    _, level = po.levels(t["stmt"], t["idx"], G, 2, window=65536, huge=65536, pack=1)
    ks = fused_stream(t["stmt"], t["mult"], t["idx"], G, level)
    g0 = np.zeros(G)
    g0[1] = 1.0
    got = replay_fused(*ks, G, g0)
    want = reference_adjoint(po, t, g0)
    assert want[0] == 0.0 and same_bits(got, want)


@pytest.mark.parametrize("seed", range(6))
def test_fused_stream_random_tapes_with_overwritten_independents(po, pkg, seed):
    """Unit seeds (as a Jacobian sweep has them) on tapes whose LHS may be ANY slot, independents included."""
    rng = np.random.default_rng(900 + seed)
    t = pkg.tapegen.random_small_tape(rng, int(rng.integers(30, 600)), int(rng.integers(8, 40)), n_indep=4, n_dep=3,
                                      zero_op_prob=0.3, lhs_any_prob=0.5)
    G = t["max_gradient"]
    _, level = po.levels(t["stmt"], t["idx"], G, 2, window=53, huge=65536, pack=1)
    ks = fused_stream(t["stmt"], t["mult"], t["idx"], G, level)
    for d in t["dep"]:
        g0 = np.zeros(G)
        g0[d] = 1.0
        assert same_bits(replay_fused(*ks, G, g0), reference_adjoint(po, t, g0))
