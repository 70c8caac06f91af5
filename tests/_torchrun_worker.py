"""Worker of tests/test_gpu_torchrun.py: one process per GPU (torchrun), the deployment bench.py uses at N > 1.

    adept_cuda_create([local_rank]) -> adept_cuda_unique_id / adept_cuda_join -> rank 0: adept_cuda_upload_tape ->
    all: adept_cuda_share_tape -> all: adept_cuda_jacobian[_device] (collective; rank 0 receives the matrix)

Checks (rank 0 compares, every rank must survive): golden bits for both sweep directions and both schedules,
strided output layouts through the gather + re-layout path, a tape with more directions than 32 x ranks, and the
ERROR paths: a root without a tape, an out-of-range slot on one rank only, a malformed tape -- every rank must
return an error, nobody may be left waiting in a collective.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_golden, same_bits
    import pyoracle as po

    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    pkg = importlib.import_module("adept-2_b200")
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    eng = pkg.Engine([local])
    ids = [eng.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.join(ids[0], rank, world)
    checked = 0

    def expect_error(fn, what):
        try:
            fn()
        except pkg.AdeptException as e:
            return str(e)
        raise AssertionError(f"rank {rank}: {what}: no error raised")

    # ---- a root without a tape: everybody gets NO_TAPE from the header, nobody waits for arrays ----
    msg = expect_error(lambda: eng.share_tape(0), "share_tape without a tape")
    assert "no tape" in msg, msg

    def share(t):
        if rank == 0:
            eng.upload_tape(t["stmt"], t["mult"], t["idx"], t["max_gradient"])
        eng.share_tape(0)

    def collective_jacobian(t, mode, do=1, io=0):
        n, m = t["indep"].size, t["dep"].size
        size = 1 + (m - 1) * (do if do > 0 else n) + (n - 1) * (io if io > 0 else m)
        if rank == 0:
            return eng.jacobian(mode, t["indep"], t["dep"], out=np.full(size, np.nan), dep_offset=do, indep_offset=io)
        eng.jacobian_device(mode, t["indep"], t["dep"], 0, dep_offset=do, indep_offset=io)
        return None

    # ---- golden vectors, both schedules, host and device output ----
    for name in ("toon_nx128_nt20", "adversarial_2", "synthetic_s3000"):
        g = load_golden(name)
        for schedule in (1, 2):
            eng.set_option("schedule", schedule)
            share(g)
            for mode, key in ((0, "jac_fwd"), (1, "jac_rev")):
                got = collective_jacobian(g, mode)
                if rank == 0:
                    assert same_bits(got, g[key]), (name, schedule, mode)
                    checked += 1
                out = torch.full((g["indep"].size * g["dep"].size,), float("nan"), dtype=torch.float64, device="cuda") if rank == 0 else None
                eng.jacobian_device(mode, g["indep"], g["dep"], out.data_ptr() if rank == 0 else 0)
                if rank == 0:
                    assert same_bits(out.cpu().numpy(), g[key]), (name, schedule, mode, "device output")
                    checked += 1
    # ---- strided layouts: gather + re-layout on rank 0 ----
    g = load_golden("adversarial_2")
    n, m = g["indep"].size, g["dep"].size
    eng.set_option("schedule", 0)
    share(g)
    for mode in (0, 1):
        for do, io in ((n, 1), (1, m + 2), (2, 2 * m + 1)):
            got = collective_jacobian(g, mode, do, io)
            if rank == 0:
                size = 1 + (m - 1) * do + (n - 1) * io
                rc, want = po.jacobian(g["stmt"], g["mult"], g["idx"], g["max_gradient"], g["indep"], g["dep"], mode,
                                       dep_offset=do, indep_offset=io, out=np.full(size, np.nan))
                assert rc == 0 and same_bits(got, want), (mode, do, io)
                checked += 1
    # ---- more directions than 32 x ranks, level schedule, fused reverse; sizes agree on every rank ----
    t = pkg.tapegen.advection_tape("toon", 512, 60)
    share(t)
    assert eng.stat("n_levels") > 0 and eng.stat("analysis_ms") is not None     # every rank analysed its own copy
    for mode in (0, 1):
        got = collective_jacobian(t, mode)
        if rank == 0:
            rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], mode, n_threads=0)
            assert rc == 0 and same_bits(got, want), ("toon512", mode)
            checked += 1
    # ---- error on ONE rank only: an out-of-range independent on rank 0; every rank must fail, none may hang ----
    bad = t["indep"].copy()
    if rank == 0:
        bad[3] = t["max_gradient"] + 7
    expect_error(lambda: eng.jacobian_device(0, bad, t["dep"], 0), "out-of-range slot on rank 0 only")
    # ... and the communicator still works afterwards
    got = collective_jacobian(t, 1)
    if rank == 0:
        rc, want = po.jacobian(t["stmt"], t["mult"], t["idx"], t["max_gradient"], t["indep"], t["dep"], 1, n_threads=0)
        assert rc == 0 and same_bits(got, want)
        checked += 1
    # ---- a malformed tape is refused by every rank alike ----
    r = pkg.tapegen.radiance_tape()
    badidx = r["idx"].copy()
    badidx[3] = r["max_gradient"] + 100
    if rank == 0:
        eng.upload_tape(r["stmt"], r["mult"], badidx, r["max_gradient"])     # root: H2D only, analysis is collective
    expect_error(lambda: eng.share_tape(0), "malformed tape")
    share(r)
    got = collective_jacobian(r, 1)
    if rank == 0:
        rc, want = po.jacobian(r["stmt"], r["mult"], r["idx"], r["max_gradient"], r["indep"], r["dep"], 1)
        assert rc == 0 and same_bits(got, want)
        checked += 1
    dist.barrier()
    eng.close()
    if rank == 0:
        print(f"torchrun worker: world {world}: {checked} comparisons bit-identical, error paths clean")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
