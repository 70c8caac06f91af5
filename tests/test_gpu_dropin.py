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
