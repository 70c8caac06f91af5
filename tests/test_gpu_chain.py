"""GPU (-m gpu): checkpoint-chained replay (reference test/test_checkpoint.cpp:166-225).

The reference differentiates a long simulation block by block, last block first: re-record the block, seed its
outputs with the gradients the previous block produced at its inputs (adept::set_gradients), Stack::reverse(),
read the gradients at the block's inputs (adept::get_gradients).  adept_cuda_chain_* keeps that hand-off vector on
the device and pipelines the blocks.  The result must equal, bit for bit, the same loop done with the oracle's
compute_adjoint and the hand-off in host memory.
"""
import numpy as np
import pytest

from conftest import same_bits

pytestmark = pytest.mark.gpu


def make_blocks(pkg, nx, nt, nblocks):
    """Forward pass in blocks (test_checkpoint.cpp:160-164), then one tape per block (the re-recordings)."""
    q = pkg.tapegen.advection_q0(nx)
    blocks = []
    for _ in range(nblocks):
        t = pkg.tapegen.advection_tape("toon", nx, nt, q0=q)
        blocks.append(t)
        q = t["q_final"]
    return blocks, q


def oracle_chain(po, blocks, h):
    for t in reversed(blocks):
        g = np.zeros(t["max_gradient"])
        g[t["dep"]] = h
        g = po.adjoint(t["stmt"], t["mult"], t["idx"], g)
        h = g[t["indep"]].copy()
    return h


@pytest.mark.parametrize("nx,nt,nblocks,schedule", [(100, 10, 6, 0), (100, 10, 6, 2), (128, 25, 5, 0), (384, 12, 9, 0)])
def test_chain_matches_block_by_block_oracle(engine, po, pkg, nx, nt, nblocks, schedule):
    blocks, q_final = make_blocks(pkg, nx, nt, nblocks)
    h0 = 2.0 * (q_final - pkg.tapegen.advection_q0(nx))         # dJ/dq_final of J = sum (q - q_init)^2
    want = oracle_chain(po, blocks, h0)
    engine.set_option("schedule", schedule)
    try:
        engine.chain_begin(nx, h0)
        for t in reversed(blocks):
            # the arrays are handed over and then scribbled on, as a re-recording Stack would do
            stmt, mult, idx = t["stmt"].copy(), t["mult"].copy(), t["idx"].copy()
            engine.chain_push(stmt, mult, idx, t["max_gradient"], t["dep"], t["indep"])
            stmt[:] = -7; mult[:] = np.nan; idx[:] = -7
        got = engine.chain_end(nx)
        assert engine.stat("chain_blocks") == nblocks
        assert same_bits(got, want)
    finally:
        engine.set_option("schedule", 0)
    # the context is usable afterwards
    t = blocks[0]
    engine.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
    g0 = np.zeros(t["max_gradient"]); g0[t["dep"]] = 1.0
    assert same_bits(engine.adjoint(g0), po.adjoint(t["stmt"], t["mult"], t["idx"], g0))


def test_chain_first_block_with_explicit_seeds_and_duplicate_slots(engine, po, pkg):
    """test_checkpoint.cpp:184-199: the first block is seeded with J.set_gradient(1.0), not with a hand-off vector;
    and set_gradients semantics: a slot named twice keeps the LAST value."""
    rng = np.random.default_rng(5)
    t = pkg.tapegen.random_small_tape(rng, 3000, 80, n_indep=6, n_dep=4, lhs_any_prob=0.2)
    seeds = np.array([t["dep"][0], t["dep"][1], t["dep"][0]], np.int32)          # dep[0] named twice
    vals = np.array([1.0, -2.5, 7.0])
    g = np.zeros(t["max_gradient"]); g[seeds[1]] = -2.5; g[seeds[0]] = 7.0
    want1 = po.adjoint(t["stmt"], t["mult"], t["idx"], g)[t["indep"]]
    t2 = pkg.tapegen.random_small_tape(rng, 500, 60, n_indep=6, n_dep=6)
    g = np.zeros(t2["max_gradient"]); g[t2["dep"]] = want1
    want2 = po.adjoint(t2["stmt"], t2["mult"], t2["idx"], g)[t2["indep"]]
    for schedule in (2, 1):
        engine.set_option("schedule", schedule)
        engine.chain_begin(6)
        engine.chain_push(t["stmt"], t["mult"], t["idx"], t["max_gradient"], seeds, t["indep"], seed_values=vals)
        engine.chain_push(t2["stmt"], t2["mult"], t2["idx"], t2["max_gradient"], t2["dep"], t2["indep"])
        assert same_bits(engine.chain_end(6), want2)
    engine.set_option("schedule", 0)


def test_chain_errors(pkg):
    eng = pkg.Engine([0])
    t = pkg.tapegen.radiance_tape()
    n = t["indep"].size
    with pytest.raises(pkg.cuda_replay_error):                       # push without begin
        eng.chain_push(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["dep"], t["indep"])
    eng.chain_begin(n)
    with pytest.raises(pkg.cuda_replay_error):                       # a second begin
        eng.chain_begin(n)
    with pytest.raises(pkg.cuda_replay_error):                       # hand-off length mismatch
        eng.chain_push(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["dep"], t["indep"][:3], seed_values=np.ones(2))
    with pytest.raises(pkg.gradient_out_of_range):
        eng.chain_push(t["stmt"], t["mult"], t["idx"], t["max_gradient"], [t["max_gradient"] + 3, 0], t["indep"], seed_values=np.ones(2))
    with pytest.raises(pkg.cuda_replay_error):                       # other calls are refused while the chain is open
        eng.jacobian(1, t["indep"], t["dep"])
    bad = t["idx"].copy(); bad[3] = t["max_gradient"] + 100            # a malformed block: reported at the end (or next push)
    eng.chain_push(t["stmt"], t["mult"], bad, t["max_gradient"], t["dep"], t["indep"], seed_values=np.ones(2))
    with pytest.raises(pkg.gradient_out_of_range):
        eng.chain_end(n)
    eng.chain_begin(n)                                               # and the context recovers
    eng.chain_push(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["dep"], t["indep"], seed_values=np.ones(2))
    assert np.isfinite(eng.chain_end(n)).all()
    eng.close()
