"""CPU: the C-ABI shared library loads and exports every symbol include/adept_cuda.h declares
(no compute calls: there is no GPU here), and fails loudly -- not silently -- without a device."""
import ctypes as C
import os
import subprocess

import pytest


def test_library_is_built(pkg):
    assert os.path.exists(pkg._abi.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound(pkg):
    declared = pkg._abi.declared_symbols()
    assert len(declared) >= 15
    L = C.CDLL(pkg._abi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/adept_cuda.h but not exported"
    assert sorted(pkg._abi.SIGNATURES) == declared, "ctypes binding and header disagree"


def test_library_is_sm_100a_only_and_uses_tma(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg._abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.run(["cuobjdump", "-sass", pkg._abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "replay kernels must stage the tape with TMA bulk copies"
    assert "DFMA" not in sass, "fp64 multiply-add must not be contracted (bit parity with the reference)"


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.cuda_replay_error):
        pkg.Engine([0])
    s = pkg.Stack()
    s.new_recording()
    s.push_rhs(2.0, 0); s.push_lhs(1)
    s.independent(0); s.dependent(1)
    with pytest.raises(pkg.cuda_replay_error):      # replay needs the CUDA engine: no silent host path
        s.jacobian()


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for dirpath, _, files in os.walk(os.path.join(root, "adept-2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if "pyoracle" in text or "liboracle" in text or "libadept_ref" in text:
                    bad.append(f)
    assert not bad, f"product files reference the oracle: {bad}"


def test_every_documented_option_is_implemented_and_vice_versa():
    """include/adept_cuda.h documents the option keys of adept_cuda_set_option; the engine must accept exactly
    those (an option that exists in only one of the two places is a documentation or a dead-code bug)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "adept_cuda.h")).read()
    block = header[header.index("/* Options (key, value):"):header.index("int  adept_cuda_set_option")]
    documented = set(re.findall(r'^\s{5}"([a-z_]+)"', block, flags=re.M))
    engine = open(os.path.join(root, "adept-2_b200", "csrc", "engine.cu")).read()
    body = engine[engine.index("int adept_cuda_set_option"):engine.index("int adept_cuda_get_stat")]
    implemented = set(re.findall(r'k == "([a-z_]+)"', body))
    assert documented == implemented, (sorted(documented - implemented), sorted(implemented - documented))


def test_fused_kernel_is_built_in_all_three_skip_rule_variants(pkg):
    """reverse_fused_kernel<V=2, LP>: LP = 0 (P = 1), 1 (P = 2), 2 (P >= 4) -- the reference's block widths
    (adept/jacobian.cpp:392-407) -- plus the two-kernel form it can fall back to."""
    out = subprocess.run(["cuobjdump", "-sass", pkg._abi.LIB_PATH], capture_output=True, text=True).stdout
    for lp in (0, 1, 2):
        assert f"reverse_fused_kernelILi2ELi{lp}EE" in out
    assert "reverse_target_kernelILi2EE" in out and "reverse_snapshot_kernel" in out


def test_traffic_file_names_the_kernels_the_bench_reports():
    """profiles/traffic.json feeds roofline.traffic; bench.py only copies an entry whose kernel name matches
    the kernel it reports for that workload."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tr = json.load(open(os.path.join(root, "profiles", "traffic.json")))
    bench = open(os.path.join(root, "bench.py")).read()
    for wl in ("toon4096", "synthetic1e7_rev"):
        assert tr[wl]["kernel"] in bench and tr[wl]["dram_bytes_per_launch"] > 0
