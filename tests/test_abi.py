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
