"""GPU (-m gpu): the C++ drop-in.  Programs written against the plain Adept-2 API -- including the
reference's OWN test programs, compiled unmodified -- run against adept-2_b200/dropin/_build/
libadept_cuda_dropin.so, i.e. the reference library with Stack::jacobian*/compute_* replaced by the
CUDA engine.  The binaries are built where /root/reference exists (adept-2_b200/dropin/Makefile) and
travel to the GPU box; this test never reads the reference tree."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, same_bits

pytestmark = pytest.mark.gpu
B = os.path.join(ROOT, "adept-2_b200", "dropin", "_build")
needs_build = pytest.mark.skipif(not os.path.exists(os.path.join(B, "dropin_check")),
                                 reason="drop-in binaries not built (needs /root/reference at build time)")


def _read_dump(path, nx):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, np.int64, 5)
    ns, no, g, ni, nd = (int(x) for x in hdr)
    off = 40
    def take(dt, n):
        nonlocal off
        a = np.frombuffer(raw, dt, n, off)
        off += a.nbytes
        return a
    d = dict(stmt=take(np.int32, 2 * ns).reshape(-1, 2), mult=take(np.float64, no), idx=take(np.int32, no),
             indep=take(np.int32, ni), dep=take(np.int32, nd), max_gradient=g)
    d["jf"], d["jr"], d["jauto"] = take(np.float64, nx * nx), take(np.float64, nx * nx), take(np.float64, nx * nx)
    d["gad"], d["gtl"] = take(np.float64, nx), take(np.float64, nx)
    d["errors_ok"], d["rmsd_ok"] = (int(x) for x in take(np.int64, 2))
    return d


@needs_build
@pytest.mark.parametrize("nx,nt", [(128, 200), (40, 7)])
def test_user_program_through_the_dropin(po, tmp_path, nx, nt):
    out = tmp_path / "dump.bin"
    r = subprocess.run([os.path.join(B, "dropin_check"), str(nx), str(nt), str(out)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    d = _read_dump(out, nx)
    assert d["errors_ok"] == 7 and d["rmsd_ok"] == 1          # the reference's exceptions, its fwd==rev criterion
    args = (d["stmt"], d["mult"], d["idx"], d["max_gradient"], d["indep"], d["dep"])
    assert same_bits(d["jf"], po.jacobian(*args, 0)[1])
    assert same_bits(d["jr"], po.jacobian(*args, 1)[1])
    assert same_bits(d["jauto"], d["jf"])                      # square: Stack::jacobian() picks forward
    g = np.zeros(d["max_gradient"]); g[d["dep"]] = 1.0 + 0.001 * np.arange(nx)
    assert same_bits(d["gad"], po.adjoint(d["stmt"], d["mult"], d["idx"], g)[d["indep"]])
    g = np.zeros(d["max_gradient"]); g[d["indep"]] = np.cos(0.1 * np.arange(nx))
    assert same_bits(d["gtl"], po.tangent_linear(d["stmt"], d["mult"], d["idx"], g)[d["dep"]])


@needs_build
def test_reference_test_programs_run_unmodified():
    # test/test_thread_safe.cpp: exits 1 if forward and reverse Jacobians disagree (RMSD > 1e-5)
    for mode in ("-serial", "-parallel"):
        r = subprocess.run([os.path.join(B, "test_thread_safe"), mode], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="4"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "CORRECT BEHAVIOUR" in r.stdout and "ERROR" not in r.stdout
    # test/test_adept.cpp: prints dy/dx from reverse(), forward() and jacobian(); expected 61.3944, 552.55 ... for x = {2, 3}
    r = subprocess.run([os.path.join(B, "test_adept")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    vals = {}
    for line in r.stdout.splitlines():
        if line.strip().startswith("dy_dx") and "=" in line:
            k, v = line.split("=")
            vals[k.strip()] = float(v)
    for x in ("dy_dx0", "dy_dx1"):
        assert vals[f"{x}[adjoint]"] == vals[f"{x}[jacobian]"]                      # reverse() vs jacobian()
        assert abs(vals[f"{x}[adjoint]"] - vals[f"{x}[numerical]"]) <= 1e-3 * abs(vals[f"{x}[numerical]"])


def _epoch_dump(exe, tmp_path, n):
    out = tmp_path / (os.path.basename(exe) + ".bin")
    r = subprocess.run([exe, str(n), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    uploads = [int(x) for x in np.frombuffer(raw, np.uint64, 4)]
    vals = np.frombuffer(raw, np.float64, 8 * n, 32).reshape(8, n)
    return uploads, vals[:4], vals[4:]


@pytest.mark.skipif(not os.path.exists(os.path.join(B, "epoch_check")), reason="epoch_check not built")
def test_exact_tape_cache_with_the_recording_epoch_patch(tmp_path):
    """SURVEY 8f-2 / VERDICT r1 #7: optimiser-style loops that sweep ONE tape many times
    (test/test_derivatives.cpp:25-29).  With adept_recording_epoch.patch applied to the reference's Stack.h the
    drop-in's tape cache is exact: 2n sweeps of one recording cost ONE upload + schedule build; a re-recording of
    identical size is detected (new values used); a recording that grew after a sweep is re-uploaded once.  Built
    against the stock headers the same program re-uploads on every sweep.  Results are the analytic derivatives."""
    n = 24
    uploads, got, want = _epoch_dump(os.path.join(B, "epoch_check"), tmp_path, n)
    assert uploads == [1, 1, 1, 1], uploads
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)
    assert not np.allclose(got[0], got[2])                         # the re-recording really has other multipliers
    uploads_s, got_s, want_s = _epoch_dump(os.path.join(B, "epoch_check_stock"), tmp_path, n)
    assert uploads_s == [2 * n, n, n, 0], uploads_s
    assert same_bits(got_s, got)                                   # cached or not: the same bits
