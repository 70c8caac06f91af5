"""ctypes binding of libadept_cuda.so (include/adept_cuda.h).

The library is the product; this module only declares its signatures.  Loading fails
loudly when the shared object is missing -- there is no Python or CPU fallback for replay.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# $ADEPT_CUDA_LIB: another build of the same library, e.g. libadept_cuda_check.so (`make -C csrc CHECK=1`: device-side
# bounds checks compiled in) to run the GPU test suite against
LIB_PATH = os.environ.get("ADEPT_CUDA_LIB") or os.path.join(HERE, "libadept_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "adept_cuda.h")

OK = 0
ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_NO_TAPE, ERR_RANGE, ERR_LIMIT = 1, 2, 3, 4, 5, 6
ERR_NOT_IDENTIFIED, ERR_GRAD_NOT_INIT = 10, 11
FORWARD, REVERSE = 0, 1
SWEEP_ORDERED = 1

_vp = C.c_void_p
_i64 = C.c_int64

SIGNATURES = {
    "adept_cuda_create": (C.c_int, [C.POINTER(_vp), C.POINTER(C.c_int), C.c_int]),
    "adept_cuda_destroy": (None, [_vp]),
    "adept_cuda_unique_id": (C.c_int, [_vp]),
    "adept_cuda_join": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "adept_cuda_share_tape": (C.c_int, [_vp, C.c_int]),
    "adept_cuda_upload_tape": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, C.c_uint64]),
    "adept_cuda_tape_epoch": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(_i64), C.POINTER(_i64)]),
    "adept_cuda_jacobian": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _i64, _vp, _i64, _i64]),
    "adept_cuda_jacobian_device": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _i64, _vp, _i64, _i64]),
    "adept_cuda_tangent_linear": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "adept_cuda_adjoint": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "adept_cuda_tangent_linear_device": (C.c_int, [_vp, _vp, C.c_int]),
    "adept_cuda_adjoint_device": (C.c_int, [_vp, _vp, C.c_int]),
    "adept_cuda_jacobian_batch": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _i64, _i64, _vp, _i64,
                                            _vp, _i64, _vp, _i64, _vp]),
    "adept_cuda_host_alloc": (_vp, [C.c_size_t]),
    "adept_cuda_host_free": (None, [_vp]),
    "adept_cuda_chain_begin": (C.c_int, [_vp, _i64, _vp]),
    "adept_cuda_chain_push": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _i64]),
    "adept_cuda_chain_end": (C.c_int, [_vp, _vp]),
    "adept_cuda_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "adept_cuda_get_stat": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_double)]),
    "adept_cuda_get_levels": (C.c_int, [_vp, _vp, _i64]),
    "adept_cuda_last_error": (C.c_char_p, [_vp]),
    "adept_cuda_version": (C.c_char_p, []),
}


def declared_symbols(header: str = HEADER_PATH):
    """Every function name include/adept_cuda.h declares (used by the ABI test)."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adept_cuda_[a-z_0-9]+)\s*\(", text)))


_lib = None


def lib():
    """Load libadept_cuda.so once.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C adept-2_b200/csrc`.  There is no CPU fallback for tape replay.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
