"""CPU: the tape writers used by bench.py / the GPU tests produce the tapes Adept itself records:
identical structure, multipliers to rounding (checked against golden tapes recorded by the real
reference and, where oracle/_ref is present, against fresh recordings)."""
import numpy as np
import pytest

from conftest import load_golden

PROFILE = [290.0, 285.0, 279.0, 273.0, 267.0, 261.0, 255.0, 248.0, 242.0, 235.0, 229.0, 222.0, 216.0,
           216.0, 216.0, 216.0, 216.0, 216.0, 217.0, 218.0, 219.0, 220.0, 222.0, 223.0, 224.0]


def _same_structure(a, b):
    assert np.array_equal(a["stmt"], b["stmt"])
    assert np.array_equal(a["idx"], b["idx"])
    assert int(a["max_gradient"]) == int(b["max_gradient"])
    assert np.array_equal(a["indep"], b["indep"]) and np.array_equal(a["dep"], b["dep"])


def test_advection_tapes_match_golden(pkg):
    for name, scheme, nx, nt, tol in (("toon_nx128_nt20", "toon", 128, 20, 1e-4), ("lw_nx100_nt30", "lw", 100, 30, 0.0)):
        # Toon is ill-conditioned where q[i] ~ q[i+1] (benchmark/advection_schemes.h:41-44 says as much):
        # rounding differences in the recorded values grow step by step; step 0 agrees to 1e-12.
        g = load_golden(name)
        t = pkg.tapegen.advection_tape(scheme, nx, nt)
        _same_structure(g, t)
        rel = np.max(np.abs(g["mult"] - t["mult"]) / np.abs(g["mult"]))
        assert rel <= tol, (name, rel)


def test_advection_sizes_follow_the_survey_formulas(pkg):
    for scheme, per in (("lw", 7), ("toon", 9)):
        nx, nt = 50, 13
        t = pkg.tapegen.advection_tape(scheme, nx, nt)
        assert t["stmt"].shape[0] - 1 == nt * (2 * nx - 1) + nx
        assert t["idx"].size == nt * (per * nx - (per + 1)) + nx
        assert t["max_gradient"] == 3 * nx - 1


def test_rosenbrock_tape_matches_golden(pkg):
    g = load_golden("rosenbrock_n400")
    t = pkg.tapegen.rosenbrock_tape(400)
    _same_structure(g, t)
    assert np.array_equal(g["mult"], t["mult"])
    n = 400
    assert t["stmt"].shape[0] - 1 == 4 * (n - 1) + 2 and t["idx"].size == 10 * (n - 1) + 1


def test_radiance_tape_matches_golden(pkg):
    g = load_golden("radiances")
    t = pkg.tapegen.radiance_tape(294.0, PROFILE)
    _same_structure(g, t)
    assert np.max(np.abs(g["mult"] - t["mult"]) / np.abs(g["mult"])) < 1e-14


def test_synthetic_tape_contract(pkg):
    t = pkg.tapegen.synthetic_tape(20000, 12, 64, seed=3)
    S = t["stmt"].shape[0] - 1
    G = 1 << 12
    nops = np.diff(t["stmt"][:, 1])
    assert S == 20000 and nops.min() >= 1 and nops.max() <= 7 and abs(nops.mean() - 4.0) < 0.1
    lhs = t["stmt"][1:, 0]
    assert lhs.min() >= 64                                   # independents are never written
    assert np.array_equal(lhs[-64:], G - 64 + np.arange(64))  # last statements write the dependents
    assert lhs[:-64].max() < G - 64
    assert t["idx"].min() >= 0 and t["idx"].max() < G - 64
    assert np.abs(t["mult"]).max() <= 0.866 and t["max_gradient"] == G + 1
    t2 = pkg.tapegen.synthetic_tape(20000, 12, 64, seed=3)
    assert np.array_equal(t["mult"], t2["mult"]) and np.array_equal(t["idx"], t2["idx"])   # frozen seed


@pytest.mark.skipif(not __import__("pyoracle").ref_available(""), reason="oracle/_ref not built")
def test_advection_tape_matches_a_fresh_reference_recording(pkg, po):
    t = po.RefTape().record_advection("toon", 100, 25)
    a = t.arrays()
    b = pkg.tapegen.advection_tape("toon", 100, 25)
    _same_structure(a, b)
    assert np.max(np.abs(a["mult"] - b["mult"]) / np.abs(a["mult"])) < 1e-4
    n0 = 100 + 6 * 99                      # first time step: before any amplification
    assert np.max(np.abs(a["mult"][:n0] - b["mult"][:n0]) / np.abs(a["mult"][:n0])) < 1e-11
