"""oracle/make_golden.py -- generates tests/golden/*.npz from the REAL reference.

Run in the build container (needs oracle/_ref/libadept_ref.so, i.e. /root/reference):
    python oracle/make_golden.py
Each fixture holds a tape (stmt, mult, idx, max_gradient, indep, dep) exactly as adept::Stack
held it, and what the unmodified reference computed from it with its default build flags
(-O2, no FMA contraction, P = 2): jac_fwd / jac_rev (Stack::jacobian_forward/_reverse, flat,
default offsets), and one tangent-linear and one adjoint sweep from seeded gradient vectors.
The fixtures are small on purpose; full-size cases run against oracle/_ref at test time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import pyoracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

PROFILE = [294.0, 290.0, 285.0, 279.0, 273.0, 267.0, 261.0, 255.0, 248.0, 242.0, 235.0, 229.0, 222.0,
           216.0, 216.0, 216.0, 216.0, 216.0, 216.0, 217.0, 218.0, 219.0, 220.0, 222.0, 223.0, 224.0]


def finish(name, t: po.RefTape, seed):
    a = t.arrays()
    rng = np.random.default_rng(seed)
    G = a["max_gradient"]
    jf = t.jacobian(0, n_threads=1)
    jr = t.jacobian(1, n_threads=1)
    g_tl = np.zeros(G)
    g_tl[a["indep"]] = rng.uniform(-1, 1, a["indep"].size)
    g_ad = np.zeros(G)
    g_ad[a["dep"]] = rng.uniform(-1, 1, a["dep"].size)
    tl = t.sweep(False, g_tl)
    ad = t.sweep(True, g_ad)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, stmt=a["stmt"], mult=a["mult"], idx=a["idx"], max_gradient=np.int64(G),
                        indep=a["indep"], dep=a["dep"], jac_fwd=jf, jac_rev=jr, g_tl_in=g_tl, g_tl_out=tl,
                        g_ad_in=g_ad, g_ad_out=ad, packet_size=np.int64(t.packet_size))
    print(f"{name}: stmts {a['stmt'].shape[0]} ops {a['idx'].size} slots {G} "
          f"jac {a['dep'].size}x{a['indep'].size} -> {os.path.getsize(path)} bytes")


def main():
    os.makedirs(OUT, exist_ok=True)
    importlib = __import__("importlib")
    tg = importlib.import_module("adept-2_b200").tapegen

    t = po.RefTape(); t.record_algorithm(1.5, 0.7); finish("algorithm", t, 1)           # test/algorithm.cpp
    t = po.RefTape(); t.record_radiances(PROFILE[0], np.array(PROFILE[1:])); finish("radiances", t, 2)
    t = po.RefTape(); t.record_advection("toon", 128, 20); finish("toon_nx128_nt20", t, 3)   # test_thread_safe shape
    t = po.RefTape(); t.record_advection("lw", 100, 30); finish("lw_nx100_nt30", t, 4)       # autodiff_benchmark shape
    t = po.RefTape(); t.record_rosenbrock(400, -3.0); finish("rosenbrock_n400", t, 5)        # one 1596-operand statement
    s = tg.synthetic_tape(3000, 9, 40, seed=7)                                            # config-4 shape, tiny
    t = po.RefTape(); t.inject(s["stmt"], s["mult"], s["idx"], s["max_gradient"] - 1, s["indep"], s["dep"])
    finish("synthetic_s3000", t, 6)
    rng = np.random.default_rng(11)
    for k in range(3):                                                                    # adversarial tapes
        r = tg.random_small_tape(rng, 300 + 150 * k, 24 + 20 * k, n_indep=5 + k, n_dep=4 + 2 * k)
        t = po.RefTape(); t.inject(r["stmt"], r["mult"], r["idx"], r["max_gradient"] - 1, r["indep"], r["dep"])
        finish(f"adversarial_{k}", t, 20 + k)


if __name__ == "__main__":
    main()
