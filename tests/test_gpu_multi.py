"""GPU (-m gpu), >= 2 devices: direction-partitioned Jacobians.  One process, several
devices: the tape is ncclBroadcast from devices[0], every device sweeps its contiguous block of
directions, blocks are gathered on devices[0].  Results must equal the single-device ones bit for bit."""
import numpy as np
import pytest

from conftest import load_golden, same_bits

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_single_process_multi_device(pkg, po, ndev):
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    eng = pkg.Engine(list(range(ndev)))
    try:
        for name in ("toon_nx128_nt20", "adversarial_2", "synthetic_s3000"):
            g = load_golden(name)
            for schedule in (1, 2):
                eng.set_option("schedule", schedule)
                eng.upload_tape(g["stmt"], g["mult"], g["idx"], g["max_gradient"])
                assert same_bits(eng.jacobian(0, g["indep"], g["dep"]), g["jac_fwd"])
                assert same_bits(eng.jacobian(1, g["indep"], g["dep"]), g["jac_rev"])
        # strided layouts go through the gather + re-layout path
        g = load_golden("adversarial_2")
        n, m = g["indep"].size, g["dep"].size
        eng.upload_tape(g["stmt"], g["mult"], g["idx"], g["max_gradient"])
        for mode in (0, 1):
            for do, io in ((n, 1), (1, m + 2), (2, 2 * m + 1)):
                size = 1 + (m - 1) * do + (n - 1) * io
                got = eng.jacobian(mode, g["indep"], g["dep"], out=np.full(size, np.nan), dep_offset=do, indep_offset=io)
                rc, want = po.jacobian(g["stmt"], g["mult"], g["idx"], g["max_gradient"], g["indep"], g["dep"], mode,
                                       dep_offset=do, indep_offset=io, out=np.full(size, np.nan))
                assert rc == 0 and same_bits(got, want), (mode, do, io)
        # more directions than 32 * devices so that every device gets work
        t = pkg.tapegen.advection_tape("toon", 256, 60)
        eng.set_option("schedule", 0)
        eng.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
        for mode in (0, 1):
            rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], mode, n_threads=0)
            assert same_bits(eng.jacobian(mode, t["indep"], t["dep"]), want)
    finally:
        eng.close()


@pytest.mark.parametrize("ndev", [2, 8])
def test_ensemble_members_sharded_over_devices(pkg, po, ndev):
    """BASELINE config 3's partition: ensemble members are block-sharded over the devices, no
    collective at all; every member's Jacobian must equal the oracle's for that member's tape."""
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    eng = pkg.Engine(list(range(ndev)))
    try:
        t = pkg.tapegen.radiance_tape()
        for B in (ndev - 1, 1003):        # fewer members than devices (idle shards), and a ragged split
            rng = np.random.default_rng(B)
            mm = t["mult"][:, None] * (1.0 + 0.05 * rng.standard_normal((t["mult"].size, B)))
            for mode in (1, 0):
                got = eng.jacobian_batch(mode, t["stmt"], t["idx"], t["max_gradient"], mm, t["indep"], t["dep"])
                assert got.shape == (B, t["indep"].size, t["dep"].size) and not np.isnan(got).any()
                for b in sorted(set([0, B - 1, B // 2] + list(rng.integers(0, B, 16)))):
                    rc, want = po.jacobian(t["stmt"], np.ascontiguousarray(mm[:, b]), t["idx"], t["max_gradient"],
                                           t["indep"], t["dep"], mode)
                    assert rc == 0 and same_bits(got[b], want), (B, mode, b)
    finally:
        eng.close()


def test_direct_d2h_layouts(pkg, po):
    """One process, several devices, host output: in the two compact layouts every device copies its block of
    directions straight into the caller's buffer (no gather through device 0); any other stride goes through the
    NCCL gather + re-layout.  All must give the reference's bits."""
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    eng = pkg.Engine([0, 1])
    try:
        t = pkg.tapegen.advection_tape("toon", 192, 30)
        n = m = 192
        eng.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
        for mode in (0, 1):
            for do, io in ((1, 0), (0, 1), (n, 1), (1, m), (1, m + 3)):
                size = 1 + (m - 1) * (do if do > 0 else n) + (n - 1) * (io if io > 0 else m)
                got = eng.jacobian(mode, t["indep"], t["dep"], out=np.full(size, np.nan), dep_offset=do, indep_offset=io)
                rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], mode,
                                       dep_offset=do, indep_offset=io, out=np.full(size, np.nan), n_threads=0)
                assert rc == 0 and same_bits(got, want), (mode, do, io)
    finally:
        eng.close()
