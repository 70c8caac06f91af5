"""The host-side ensemble recording helper (adept-2_b200/dropin/adept_cuda_ensemble.h): members
recorded with the plain Adept-2 API on all host cores, one adept::Stack per thread, structures checked, multipliers
packed [operation][member] for adept_cuda_jacobian_batch.

CPU part (no GPU needed): the program adept-2_b200/dropin/_build/ensemble_check records every member twice -- through
the helper and serially on one Stack with no helper code -- and the two must agree bit for bit; the tapes must
reproduce the model's analytic Jacobian through the oracle; a model with a data-dependent branch must be refused.
GPU part (-m gpu): the same program calls adept_cuda_jacobian_batch with the packed ensemble; every member's
Jacobian must equal the oracle's for that member's tape, bit for bit, in both sweep directions.
The binary is built where /root/reference exists (dropin/Makefile) and travels; nothing here reads the reference tree."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, same_bits

B_DIR = os.path.join(ROOT, "adept-2_b200", "dropin", "_build")
EXE = os.path.join(B_DIR, "ensemble_check")
needs_build = pytest.mark.skipif(not os.path.exists(EXE), reason="ensemble_check not built (needs /root/reference at build time)")
NX, NY = 8, 3


def run_check(tmp_path, n_members, mode, threads=0):
    out = tmp_path / f"ens_{mode}_{threads}.bin"
    r = subprocess.run([EXE, str(n_members), str(out), mode, str(threads)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    hdr = np.frombuffer(raw, np.int64, 12)
    B, S, O, G, ni, nd, refused, serial_same, G0, njac, gpu_rc, S0 = (int(x) for x in hdr)
    off = 96

    def take(dt, n):
        nonlocal off
        a = np.frombuffer(raw, dt, n, off)
        off += a.nbytes
        return a
    d = dict(B=B, S=S, O=O, G=G, refused=refused, serial_same=serial_same, G0=G0, gpu_rc=gpu_rc)
    d["stmt"] = take(np.int32, 2 * S).reshape(-1, 2)
    d["idx"] = take(np.int32, O)
    d["indep"], d["dep"] = take(np.int32, ni), take(np.int32, nd)
    d["mult"] = take(np.float64, O * B).reshape(O, B)                 # [operation][member]
    d["stmt0"] = take(np.int32, 2 * S0).reshape(-1, 2)
    d["idx0"] = take(np.int32, O)
    d["serial_mult"] = take(np.float64, O * B).reshape(B, O)          # [member][operation]
    d["analytic"] = take(np.float64, B * NX * NY).reshape(B, NX, NY)  # per member column-major NY x NX: [j][k]
    d["jac_rev"] = take(np.float64, njac).reshape(-1, NX, NY) if njac else None
    d["jac_fwd"] = take(np.float64, njac).reshape(-1, NX, NY) if njac else None
    assert off == len(raw)
    return d


@needs_build
@pytest.mark.parametrize("n_members,threads", [(1, 0), (37, 0), (1000, 0), (200, 1), (200, 3)])
def test_helper_packs_what_a_serial_recording_gives(po, tmp_path, n_members, threads):
    d = run_check(tmp_path, n_members, "cpu", threads)
    assert d["B"] == n_members and d["refused"] == 1 and d["serial_same"] == 1
    assert (len(d["indep"]), len(d["dep"])) == (NX, NY)
    assert np.array_equal(d["stmt"], d["stmt0"]) and np.array_equal(d["idx"], d["idx0"]) and d["G"] == d["G0"]
    assert d["stmt"][0].tolist() == [-1, 0] and d["stmt"][-1][1] == d["O"]
    assert same_bits(d["mult"].T, d["serial_mult"])                    # the packing, bit for bit
    if n_members > 1:
        assert not np.array_equal(d["mult"][:, 0], d["mult"][:, -1])   # the model is nonlinear: members differ
    # the recorded tapes mean what the model says: oracle Jacobians against the analytic ones
    for b in sorted({0, n_members // 2, n_members - 1}):
        m = np.ascontiguousarray(d["mult"][:, b])
        for mode in (0, 1):
            rc, jac = po.jacobian(d["stmt"], m, d["idx"], d["G"], d["indep"], d["dep"], mode)
            assert rc == 0
            np.testing.assert_allclose(jac.reshape(NX, NY), d["analytic"][b], rtol=1e-12, atol=1e-14)


@needs_build
def test_thread_count_does_not_change_the_result(tmp_path):
    a = run_check(tmp_path, 333, "cpu", 1)
    b = run_check(tmp_path, 333, "cpu", 0)
    assert same_bits(a["mult"], b["mult"]) and np.array_equal(a["stmt"], b["stmt"]) and np.array_equal(a["idx"], b["idx"])


@needs_build
@pytest.mark.gpu
@pytest.mark.parametrize("n_members", [1, 257])
def test_zz_helper_feeds_the_batch_entry_point(po, tmp_path, n_members):
    d = run_check(tmp_path, n_members, "gpu")
    assert d["gpu_rc"] == 0 and d["jac_rev"] is not None
    for b in range(n_members):
        m = np.ascontiguousarray(d["mult"][:, b])
        for mode, got in ((1, d["jac_rev"]), (0, d["jac_fwd"])):
            rc, want = po.jacobian(d["stmt"], m, d["idx"], d["G"], d["indep"], d["dep"], mode)
            assert rc == 0 and same_bits(got[b], want), (mode, b)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(B_DIR, "ensemble_bench")), reason="ensemble_bench not built")
@pytest.mark.parametrize("n_members,chunks", [(300, 4), (5003, 1), (5003, 4)])
def test_zz_radiance_ensemble_end_to_end_every_member(po, n_members, chunks):
    """BASELINE config 3 end to end (bench.py --workload ensemble1e5 at a smaller size): the reference's radiance model
    recorded for every member with real Adept templates on all host cores, packed into page-locked memory, replayed
    by the pipelined batch entry point on every visible GPU -- EVERY member's 2x26 Jacobian must equal, bit for bit,
    what the reference's own simulate_radiances() (record + Stack::jacobian) returns for that state vector."""
    import sys
    import torch
    sys.path.insert(0, ROOT)
    import bench
    if not (po.ref_available("") and hasattr(po._ref(""), "ref_ensemble_radiances")):
        pytest.skip("oracle/_ref without ref_ensemble_radiances")
    state = bench.ensemble_state(n_members, seed=n_members)
    want, _ = po.ref_ensemble_radiances(state, 0, "")
    for n_dev in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        jac, info = bench.run_ensemble_bench(state, n_dev, reps=1, chunks=chunks)
        assert jac is not None, info
        assert jac.shape == (n_members, 26, 2) and info["devices"] == n_dev
        assert same_bits(jac, want), f"{n_dev} device(s)"
