"""adept-2_b200 -- B200-native replay of Adept-2 derivative tapes.

The product is ``libadept_cuda.so`` (C-ABI in include/adept_cuda.h, CUDA sources in csrc/):
hand-written sm_100a kernels that replace the replay half of ``adept::Stack`` -- Jacobians,
adjoint and tangent-linear sweeps (reference adept/Stack.cpp:139-190, adept/jacobian.cpp).
This package is the Python face of that library: ``Engine`` (one context) and ``Stack`` (a mirror
of the reference's Stack API for this path).  The directory name carries a hyphen, so import it
with ``importlib.import_module("adept-2_b200")`` (or the ``adept2_b200`` alias at the repo root).

There is no CPU fallback: importing works anywhere, but creating an Engine requires the built
library and a CUDA device.
"""
from . import _abi, tapegen
from ._abi import FORWARD, REVERSE, SWEEP_ORDERED
from .stack import (AdeptException, Engine, Stack, cuda_replay_error,
                    dependents_or_independents_not_identified, gradient_out_of_range,
                    gradients_not_initialized, size_mismatch)

__all__ = [
    "Engine", "Stack", "tapegen", "FORWARD", "REVERSE", "SWEEP_ORDERED", "AdeptException",
    "cuda_replay_error", "dependents_or_independents_not_identified", "gradient_out_of_range",
    "gradients_not_initialized", "size_mismatch",
]
__version__ = "0.1"
