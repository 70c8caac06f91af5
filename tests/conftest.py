import glob
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("adept-2_b200")


@pytest.fixture(scope="session")
def po():
    import pyoracle
    pyoracle._lib()  # builds liboracle.so if needed (gcc only)
    return pyoracle


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["max_gradient"] = int(d["max_gradient"])
    return d


@pytest.fixture(scope="session")
def engine(pkg):
    """One engine on cuda:0 for the GPU tests.  Fails loudly when the CUDA path is unavailable."""
    eng = pkg.Engine([0])
    yield eng
    eng.close()


def same_bits(a, b):
    """Bit-for-bit equality with all NaNs treated as one class (NaN payloads differ x86 vs NVIDIA)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64))
