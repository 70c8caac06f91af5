"""GPU (-m gpu): the fused reverse sweep as ONE resident kernel (option "fused_persistent",
adept-2_b200/csrc/fused_persistent_kernels.cuh) against the reference's bits.

Same record stream, same device functions per item as the one-launch-per-step sweep; what is new is the item
enumeration and the per-step participant counters (CPU model: tests/test_persistent_cursor_model.py).  The cases
below vary what that protocol sees: grids smaller and larger than a step, one and several column groups, several
passes (the counters are reused), steps of a single item, chunks longer than one staged piece.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, same_bits
from test_gpu_parity import LEVELS, broadcast_tape, check_jacobians, upload

pytestmark = pytest.mark.gpu


@pytest.fixture()
def resident(engine):
    engine.set_option("fused_persistent", 1)
    engine.set_option("smallg", 0)
    yield engine
    engine.set_option("fused_persistent", 0)
    engine.set_option("cbs", 0)
    engine.set_option("pass_dirs", 0)
    engine.set_option("chunk_recs", 0)
    engine.set_option("split", 0)
    engine.set_option("sub", 0)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("cbs", [0, 1, 3])
def test_golden_reverse_jacobians(resident, name, cbs):
    g = load_golden(name)
    for window in (0, 50):
        upload(resident, g, LEVELS, window)
        resident.set_option("cbs", cbs)
        for pass_dirs in (0, 64):
            resident.set_option("pass_dirs", pass_dirs)
            assert same_bits(resident.jacobian(1, g["indep"], g["dep"]), g["jac_rev"]), (window, pass_dirs)
            if g["idx"].size:
                assert resident.stat("used_fused_persistent") == 1
    assert same_bits(resident.jacobian(0, g["indep"], g["dep"]), g["jac_fwd"])    # forward is unaffected


@pytest.mark.parametrize("cbs", [0, 32, 3])
def test_target_groups_longer_than_a_staged_piece(resident, po, cbs):
    t = broadcast_tape()
    upload(resident, t, LEVELS)
    resident.set_option("cbs", cbs)
    check_jacobians(resident, po, t, f"broadcast, resident kernel, cbs {cbs}")
    assert resident.stat("used_fused_persistent") == 1


@pytest.mark.parametrize("chunk_recs", [16, 64, 256])
def test_stencil_wider_than_the_grid(resident, po, pkg, chunk_recs):
    """A Toon tape whose steps hold more items than CTAs fit the device (16-record chunks, one column block per item):
    every CTA takes several items per step and the start rotates from step to step."""
    t = pkg.tapegen.advection_tape("toon", 1500, 6)
    resident.set_option("chunk_recs", chunk_recs)
    upload(resident, t, LEVELS)
    sub = dict(t)
    sub["dep"] = t["dep"][::7].copy()                        # 215 reverse directions = 4 column blocks
    for cbs in (1, 0):
        resident.set_option("cbs", cbs)
        got = resident.jacobian(1, sub["indep"], sub["dep"])
        rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], sub["indep"], sub["dep"], 1, n_threads=0)
        assert rc == 0 and same_bits(got, want), (chunk_recs, cbs)
        assert resident.stat("used_fused_persistent") == 1
    resident.set_option("fused_persistent", 0)               # and the per-step sweep gives the same bits
    assert same_bits(resident.jacobian(1, sub["indep"], sub["dep"]), want)
    assert resident.stat("used_fused_persistent") == 0


@pytest.mark.parametrize("seed", range(6))
def test_random_tapes_whose_independents_are_overwritten(resident, po, pkg, seed):
    rng = np.random.default_rng(900 + seed)
    t = pkg.tapegen.random_small_tape(rng, int(rng.integers(30, 3000)), int(rng.integers(12, 200)),
                                      n_indep=int(rng.integers(1, 12)), n_dep=int(rng.integers(1, 7)),
                                      zero_op_prob=0.3, lhs_any_prob=0.5)
    for window in (0, 53):
        upload(resident, t, LEVELS, window)
        check_jacobians(resident, po, t, f"seed {seed} window {window}")
        resident.set_option("pass_dirs", 32)
        check_jacobians(resident, po, t, f"seed {seed} window {window} passes")
        resident.set_option("pass_dirs", 0)


def test_non_finite_multipliers_follow_the_reference_block_rule(resident, po, pkg):
    rng = np.random.default_rng(4)
    t = pkg.tapegen.random_small_tape(rng, 1500, 90, n_indep=9, n_dep=70)
    t["mult"][::41] = np.inf
    t["mult"][5::97] = np.nan
    for P in (1, 2, 4, 8):
        upload(resident, t, LEVELS)
        resident.set_option("ref_block", P)
        got = resident.jacobian(1, t["indep"], t["dep"])
        rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], 1, P=P)
        assert rc == 0 and same_bits(got, want), P
    resident.set_option("ref_block", 2)


# ---- work distribution (options "split", "sub") of BOTH fused reverse kernels ---------------------------------
@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("persistent", [0, 1])
@pytest.mark.parametrize("split,sub", [(2, 32), (8, 8), (3, 4), (16, 5), (1, 8)])
def test_chunks_shared_by_several_ctas(engine, name, persistent, split, sub):
    """A chunk's sub-tasks (cut at target-group heads) are owned by `split` CTAs, the warps of a CTA take (sub-task,
    active column block) pairs: any split, any sub-task size, one launch per step or the resident kernel -- same bits."""
    g = load_golden(name)
    try:
        engine.set_option("fused_persistent", persistent)
        engine.set_option("smallg", 0)
        engine.set_option("split", split)
        engine.set_option("sub", sub)
        for window, cbs in ((0, 0), (50, 3)):
            upload(engine, g, LEVELS, window)
            engine.set_option("cbs", cbs)
            for pass_dirs in (0, 64):
                engine.set_option("pass_dirs", pass_dirs)
                assert same_bits(engine.jacobian(1, g["indep"], g["dep"]), g["jac_rev"]), (window, cbs, pass_dirs)
            if g["idx"].size:
                assert engine.stat("fused_split") == split and engine.stat("fused_sub") == sub
                assert engine.stat("used_fused_persistent") == persistent
    finally:
        for k in ("fused_persistent", "split", "sub", "cbs", "pass_dirs"):
            engine.set_option(k, 0)


@pytest.mark.parametrize("persistent", [0, 1])
def test_long_chunks_are_not_split(engine, po, persistent):
    t = broadcast_tape()          # target groups of 700 records: chunks of several staged pieces
    try:
        engine.set_option("fused_persistent", persistent)
        engine.set_option("split", 4)
        engine.set_option("sub", 8)
        upload(engine, t, LEVELS)
        check_jacobians(engine, po, t, f"broadcast, split 4, persistent {persistent}")
    finally:
        for k in ("fused_persistent", "split", "sub"):
            engine.set_option(k, 0)


def test_automatic_split_with_few_directions(engine, po, pkg):
    """Option auto_split: few directions on the device (what one rank of eight holds) and narrow steps -> two CTAs per
    chunk, chosen by the engine; not when a chunk is longer than a staged piece."""
    t = pkg.tapegen.advection_tape("toon", 1200, 8)
    sub = dict(t)
    sub["dep"] = t["dep"][100:228].copy()                   # 128 directions = 2 column blocks
    try:
        engine.set_option("auto_split", 1)
        for persistent in (0, 1):
            engine.set_option("fused_persistent", persistent)
            engine.set_option("smallg", 0)
            upload(engine, t, LEVELS)
            got = engine.jacobian(1, sub["indep"], sub["dep"])
            rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], sub["indep"], sub["dep"], 1, n_threads=0)
            assert rc == 0 and same_bits(got, want)
            assert engine.stat("fused_split") == 2 and engine.stat("fused_sub") == 32
        engine.set_option("fused_persistent", 0)
        b = broadcast_tape()
        upload(engine, b, LEVELS)
        check_jacobians(engine, po, b, "broadcast, auto_split")
        assert engine.stat("max_fused_chunk_recs") > 256 and engine.stat("fused_split") == 1
    finally:
        engine.set_option("fused_persistent", 0)
        engine.set_option("auto_split", 0)


@pytest.mark.parametrize("persistent", [0, 1])
def test_every_chunk_longer_than_a_piece_and_split(engine, po, pkg, persistent):
    """chunk_recs = 256: nearly every chunk spills a few records into a second piece, so with split > 1 most items of
    the resident kernel (and most CTAs of the per-step kernel) pass their chunk by -- also the FIRST item of a CTA (the
    resident kernel once fetched a piece for it that nobody waited for; CPU model tests/test_persistent_ring_model.py)."""
    t = pkg.tapegen.advection_tape("toon", 700, 5)
    sub = dict(t)
    sub["dep"] = t["dep"][::5].copy()
    try:
        engine.set_option("fused_persistent", persistent)
        engine.set_option("smallg", 0)
        engine.set_option("chunk_recs", 256)
        upload(engine, t, LEVELS)
        for split, cbs in ((8, 1), (3, 0), (1, 0)):
            engine.set_option("split", split)
            engine.set_option("cbs", cbs)
            got = engine.jacobian(1, sub["indep"], sub["dep"])
            rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], sub["indep"], sub["dep"], 1, n_threads=0)
            assert rc == 0 and same_bits(got, want), (split, cbs)
        assert engine.stat("max_fused_chunk_recs") > 256
    finally:
        for k in ("fused_persistent", "split", "cbs", "chunk_recs"):
            engine.set_option(k, 0)
