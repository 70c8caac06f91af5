"""CPU, world_size 2 over gloo: the host-side contract of the multi-rank Jacobian -- contiguous blocks
of directions in multiples of 32 per rank, direction-major blocks gathered on rank 0 -- exercised with
the oracle standing in for the per-rank sweep (no GPU here).  The CUDA engine implements the same
partition in do_jacobian (adept-2_b200/csrc/engine.cu)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def partition(D, world):
    """Rank r sweeps directions [lo(r), lo(r+1)); block size is ceil(D/world) rounded up to 32."""
    blk = -(-D // world)
    blk = -(-blk // 32) * 32
    return [min(D, r * blk) for r in range(world + 1)]


def _worker(rank, world, port, name, mode, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pyoracle as po
    import torch
    from conftest import load_golden
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden(name)
    D = g["indep"].size if mode == 0 else g["dep"].size
    n_out = g["dep"].size if mode == 0 else g["indep"].size
    bounds = partition(D, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    ind = g["indep"][lo:hi] if mode == 0 else g["indep"]
    dep = g["dep"] if mode == 0 else g["dep"][lo:hi]
    if hi > lo:   # direction-major block: forward = default layout; reverse = dependents as the slow index
        do, io = (1, n_out) if mode == 0 else (n_out, 1)
        rc, blk = po.jacobian(g["stmt"], g["mult"], g["idx"], g["max_gradient"], ind, dep, mode, dep_offset=do, indep_offset=io)
        assert rc == 0
    else:
        blk = np.zeros(0)
    blocks = [None] * world
    dist.gather_object(blk, blocks if rank == 0 else None, dst=0)
    if rank == 0:
        dm = np.concatenate(blocks).reshape(D, n_out)          # [direction][out index]
        q.put(dm)
    dist.destroy_process_group()


@pytest.mark.parametrize("name,mode", [("toon_nx128_nt20", 0), ("toon_nx128_nt20", 1), ("synthetic_s3000", 1)])
def test_two_ranks_reassemble_the_jacobian(name, mode):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    dm = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = load_golden(name)
    n, m = g["indep"].size, g["dep"].size
    ref = (g["jac_fwd"] if mode == 0 else g["jac_rev"]).reshape(n, m)   # [indep][dep]
    want = ref if mode == 0 else ref.T
    assert np.array_equal(dm.view(np.uint64), np.ascontiguousarray(want).view(np.uint64))


def test_partition_properties():
    for D in (1, 31, 32, 33, 100, 4096, 8192, 8191):
        for world in (1, 2, 4, 8):
            b = partition(D, world)
            assert b[0] == 0 and b[-1] == D and all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % 32 == 0 or x == D for x in b)     # the reference's P-lane blocks stay whole
