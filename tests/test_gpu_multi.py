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
