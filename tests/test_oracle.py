"""CPU: pin the oracle.  The C restatement (oracle/replay_oracle.c) must reproduce, bit for bit,
(1) the committed golden vectors, which the REAL reference produced (oracle/make_golden.py), and
(2) the real reference itself (oracle/_ref, built from /root/reference) wherever that library is
present, on the reference's own test workloads."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, same_bits


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("P", [1, 2, 4, 8, 32])
def test_oracle_reproduces_golden_jacobians(po, name, P):
    g = load_golden(name)
    for mode, key in ((0, "jac_fwd"), (1, "jac_rev")):
        rc, out = po.jacobian(g["stmt"], g["mult"], g["idx"], g["max_gradient"], g["indep"], g["dep"], mode, P=P)
        assert rc == 0
        assert same_bits(out, g[key]), f"{name} mode {mode} P {P}"


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden_sweeps(po, name):
    g = load_golden(name)
    assert same_bits(po.tangent_linear(g["stmt"], g["mult"], g["idx"], g["g_tl_in"]), g["g_tl_out"])
    assert same_bits(po.adjoint(g["stmt"], g["mult"], g["idx"], g["g_ad_in"]), g["g_ad_out"])


def test_golden_forward_and_reverse_agree_like_the_reference_test():
    # test/test_thread_safe.cpp:86-113: RMSD(forward, reverse) <= 1e-5
    g = load_golden("toon_nx128_nt20")
    rmsd = np.sqrt(np.mean((g["jac_fwd"] - g["jac_rev"]) ** 2))
    assert rmsd <= 1e-5


def test_oracle_openmp_threads_do_not_change_bits(po):
    g = load_golden("toon_nx128_nt20")
    for mode in (0, 1):
        a = po.jacobian(g["stmt"], g["mult"], g["idx"], g["max_gradient"], g["indep"], g["dep"], mode, n_threads=1)[1]
        b = po.jacobian(g["stmt"], g["mult"], g["idx"], g["max_gradient"], g["indep"], g["dep"], mode, n_threads=4)[1]
        assert same_bits(a, b)


def test_oracle_offsets_and_errors(po):
    g = load_golden("adversarial_1")
    n, m = g["indep"].size, g["dep"].size
    args = (g["stmt"], g["mult"], g["idx"], g["max_gradient"], g["indep"], g["dep"])
    for mode in (0, 1):
        col = po.jacobian(*args, mode)[1].reshape(n, m)                    # default: dependents contiguous
        row = po.jacobian(*args, mode, dep_offset=n, indep_offset=1)[1].reshape(m, n)
        assert same_bits(col.T, row)
        padded = np.full((n, m + 3), np.nan)
        po.jacobian(*args, mode, dep_offset=1, indep_offset=m + 3, out=padded.reshape(-1))
        assert same_bits(padded[:, :m], col) and np.all(np.isnan(padded[:, m:]))
        # empty lists -> dependents_or_independents_not_identified (adept/jacobian.cpp:208-210)
        assert po.jacobian(g["stmt"], g["mult"], g["idx"], g["max_gradient"], [], g["dep"], mode)[0] == 10
        assert po.jacobian(g["stmt"], g["mult"], g["idx"], g["max_gradient"], g["indep"], [], mode)[0] == 10


def test_level_rules(po, pkg):
    """The dependency rule the device analysis implements (mode 2) is never looser than RAW/WAR/WAW
    (mode 0) and never tighter than full per-slot serialisation (mode 1); windows only add steps."""
    t = pkg.tapegen.advection_tape("toon", 64, 30)
    d0, l0 = po.levels(t["stmt"], t["idx"], t["max_gradient"], 0)
    d2, l2 = po.levels(t["stmt"], t["idx"], t["max_gradient"], 2)
    d1, l1 = po.levels(t["stmt"], t["idx"], t["max_gradient"], 1)
    assert d0 == 3 * 30 + 1 and d2 == d0          # Toon: 3 phases per step + the copy-in
    assert np.all(l0 <= l2) and np.all(l2 <= l1)
    dw, lw = po.levels(t["stmt"], t["idx"], t["max_gradient"], 2, window=500)
    assert dw >= d2
    dp, lp = po.levels(t["stmt"], t["idx"], t["max_gradient"], 2, window=500, pack=1)
    assert d2 <= dp <= dw                       # packing windows never beats the global rule, never loses to stacking
    r = pkg.tapegen.rosenbrock_tape(5000)       # independent windows collapse onto a few steps
    assert po.levels(r["stmt"], r["idx"], r["max_gradient"], 2, window=512, pack=1)[0] <= 8
    assert po.levels(r["stmt"], r["idx"], r["max_gradient"], 2, window=512)[0] > 30


# ---- against the real reference (oracle/_ref) -------------------------------------------------
needs_ref = pytest.mark.skipif(not __import__("pyoracle").ref_available(""),
                               reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("variant", ["", "_avx2", "_avx512"])
def test_reference_builds_agree_with_oracle(po, variant):
    import pyoracle
    if not pyoracle.ref_available(variant):
        pytest.skip("variant not built")
    flags = open("/proc/cpuinfo").read()
    if (variant == "_avx512" and " avx512f" not in flags) or (variant == "_avx2" and " avx2" not in flags):
        pytest.skip("host CPU cannot run this build")
    t = pyoracle.RefTape(variant)
    t.record_advection("toon", 128, 200)            # the test_thread_safe workload: 51128 / 228528 / 383
    assert t.sizes()[:3] == (51129, 228528, 383)
    a = t.arrays()
    for mode in (0, 1):
        ref1 = t.jacobian(mode, n_threads=1)
        refN = t.jacobian(mode, n_threads=0)        # OpenMP path (adept/jacobian.cpp:127-194, :331-452)
        rc, orc = po.jacobian(a["stmt"], a["mult"], a["idx"], a["max_gradient"], a["indep"], a["dep"], mode,
                              P=t.packet_size, n_threads=2)
        assert rc == 0 and same_bits(ref1, refN) and same_bits(ref1, orc)


@needs_ref
def test_reference_small_programs(po):
    import pyoracle
    t = pyoracle.RefTape()
    y = t.record_algorithm(1.5, 0.7)                # test/algorithm.cpp via test/test_adept.cpp
    assert np.isfinite(y) and t.sizes() == (5, 6, 4, 2, 1)
    a = t.arrays()
    assert same_bits(t.jacobian(2), po.jacobian(a["stmt"], a["mult"], a["idx"], a["max_gradient"], a["indep"],
                                                a["dep"], 1)[1])   # 2 independents > 1 dependent -> reverse
    g = np.zeros(a["max_gradient"]); g[a["dep"][0]] = 1.0
    assert same_bits(t.sweep(True, g), po.adjoint(a["stmt"], a["mult"], a["idx"], g))
    g = np.zeros(a["max_gradient"]); g[a["indep"]] = [1.0, -0.5]
    assert same_bits(t.sweep(False, g), po.tangent_linear(a["stmt"], a["mult"], a["idx"], g))
    # gradients_not_initialized (adept/Stack.cpp:161-163, :187-189)
    assert t.sweep_uninitialised(True) == 11 and t.sweep_uninitialised(False) == 11


@needs_ref
def test_reference_injected_tape_offsets(po, pkg):
    import pyoracle
    rng = np.random.default_rng(5)
    r = pkg.tapegen.random_small_tape(rng, 400, 40, n_indep=7, n_dep=5)
    t = pyoracle.RefTape().inject(r["stmt"], r["mult"], r["idx"], r["max_gradient"] - 1, r["indep"], r["dep"])
    n, m = 7, 5
    for mode in (0, 1):
        for do, io in ((1, 0), (0, 1), (n, 1), (1, m + 2), (0, 0)):
            size = 1 + (m - 1) * (do if do > 0 else n) + (n - 1) * (io if io > 0 else m)
            ref = t.jacobian(mode, dep_offset=do, indep_offset=io, out=np.full(size, np.nan))
            rc, orc = po.jacobian(r["stmt"], r["mult"], r["idx"], r["max_gradient"], r["indep"], r["dep"], mode,
                                  dep_offset=do, indep_offset=io, out=np.full(size, np.nan))
            assert rc == 0 and same_bits(ref, orc), (mode, do, io)
