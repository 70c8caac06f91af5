"""GPU (-m gpu): parity at the sizes and on the code paths bench.py measures.

* BASELINE config 2 at FULL size (Toon nx=4096, nt=1000: 8 195 096 statements): the whole 4096x4096 reverse
  Jacobian is computed by the fused level sweep and 256 full rows (every 16th dependent) are compared, bit for
  bit, with the oracle -- plus 64 full columns of the forward Jacobian.
* BASELINE config 4's shape on the TWO-PASS fused reverse path the full size takes: a 2^17-slot synthetic tape
  with the workspace capped so that two rows per slot do not fit one pass; full-matrix parity, both directions.
* the multi-pass forward path and the two-kernel reverse form on the same tape.
The oracle is the pinned restatement (oracle/replay_oracle.c) run on all host cores.
"""
import numpy as np
import pytest

from conftest import same_bits

pytestmark = pytest.mark.gpu


def _opts(engine, **kw):
    base = dict(schedule=0, smallg=1, fused=1, window=0, pass_dirs=0, workspace_bytes=0, ref_block=2, streams=4, tail=1)
    base.update(kw)
    for k, v in base.items():
        engine.set_option(k, v)


def test_toon4096_full_size_rows_against_the_oracle(engine, po, pkg):
    t = pkg.tapegen.advection_tape("toon", 4096, 1000)
    assert (t["stmt"].shape[0] - 1, t["idx"].size, t["max_gradient"]) == (8195096, 36858096, 12287)
    _opts(engine)
    engine.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
    try:
        jr = engine.jacobian(1, t["indep"], t["dep"]).reshape(4096, 4096)          # [indep, dep]
        assert engine.stat("used_schedule") == 2 and engine.stat("used_fused") == 1
        rows = np.arange(5, 4096, 16).astype(np.int32)                               # 256 dependents
        rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"][rows], 1,
                               n_threads=0)
        assert rc == 0 and same_bits(np.ascontiguousarray(jr[:, rows]), want)
        jf = engine.jacobian(0, t["indep"], t["dep"]).reshape(4096, 4096)
        cols = np.arange(11, 4096, 64).astype(np.int32)                              # 64 independents
        rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"][cols], t["dep"], 0,
                               n_threads=0)
        assert rc == 0 and same_bits(np.ascontiguousarray(jf[cols, :]), want)
        # (no forward-vs-reverse RMSD check here: over 1000 steps the Toon recurrence amplifies the rounding difference
        # between the two summation orders far beyond the 1e-5 that test/test_thread_safe.cpp:86-113 uses at nt=200;
        # both sweeps equal the reference's own bits, which is the bar)
    finally:
        _opts(engine)


@pytest.mark.parametrize("fused", [1, 0])
def test_two_pass_reverse_sweep_on_a_tape_too_big_for_l2(engine, po, pkg, fused):
    """2e6 statements / 2^17 slots / 512x512: 131 073 rows x 512 directions x 8 B = 537 MB per row set (>> L2); the
    fused sweep needs two row sets, and with the workspace capped at 0.6 GB it must take two passes of 256."""
    t = pkg.tapegen.synthetic_tape(2_000_000, 17, 512, seed=7)
    _opts(engine, fused=fused, workspace_bytes=600_000_000 if fused else 300_000_000)
    try:
        engine.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
        got = engine.jacobian(1, t["indep"], t["dep"])
        assert engine.stat("n_passes") >= 2 and engine.stat("used_schedule") == 2
        assert engine.stat("used_fused") == fused
        rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], 1, n_threads=0)
        assert rc == 0 and same_bits(got, want)
        _opts(engine, fused=fused, workspace_bytes=300_000_000)
        got = engine.jacobian(0, t["indep"], t["dep"])
        assert engine.stat("n_passes") >= 2
        rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], 0, n_threads=0)
        assert rc == 0 and same_bits(got, want)
    finally:
        _opts(engine)


def test_config4_shape_direction_subsets_match_the_full_jacobian(engine, po, pkg):
    """What one rank of an 8-GPU run computes: a 1/8 block of directions must equal the same block of the
    full Jacobian, and 16 of its directions the oracle's."""
    t = pkg.tapegen.synthetic_tape(1_000_000, 16, 1024, seed=42)
    _opts(engine)
    engine.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
    for mode in (0, 1):
        full = engine.jacobian(mode, t["indep"], t["dep"]).reshape(1024, 1024)
        blk = slice(384, 512)
        if mode == 0:
            part = engine.jacobian(0, t["indep"][blk], t["dep"]).reshape(128, 1024)
            assert same_bits(part, full[blk, :])
            sub = np.arange(384, 512, 8).astype(np.int32)
            rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"][sub], t["dep"], 0, n_threads=0)
            assert rc == 0 and same_bits(np.ascontiguousarray(full[sub, :]), want)
        else:
            part = engine.jacobian(1, t["indep"], t["dep"][blk]).reshape(1024, 128)
            assert same_bits(part, np.ascontiguousarray(full[:, blk]))
            sub = np.arange(384, 512, 8).astype(np.int32)
            rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"][sub], 1, n_threads=0)
            assert rc == 0 and same_bits(np.ascontiguousarray(full[:, sub]), want)
