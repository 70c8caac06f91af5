"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-ends for the two CPU checkers built by oracle/Makefile:

* ``liboracle.so``  -- the plain-C restatement of the reference's replay
  algorithms (oracle/replay_oracle.c; cites adept/Stack.cpp:139-190 and
  adept/jacobian.cpp:39-647).
* ``_ref/libadept_ref*.so`` -- the UNMODIFIED reference (Adept 2.1.3) compiled
  from /root/reference plus the harness oracle/ref_harness.cpp.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` leg may import this module.  The product package
(adept-2_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Run oracle/Makefile (idempotent).  ``ref`` also builds _ref/ when /root/reference exists."""
    target = ["all"] if ref else ["liboracle.so"]
    subprocess.run(["make", "-C", HERE, "-j4"] + target, check=True,
                   stdout=subprocess.DEVNULL)


# ----------------------------------------------------------------------------
# C restatement
# ----------------------------------------------------------------------------
_oracle = None


def _lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.oracle_tangent_linear.argtypes = [_i32p, C.c_int64, _f64p, _i32p, _f64p]
        L.oracle_tangent_linear.restype = None
        L.oracle_adjoint.argtypes = [_i32p, C.c_int64, _f64p, _i32p, _f64p]
        L.oracle_adjoint.restype = None
        jac = [_i32p, C.c_int64, _f64p, _i32p, C.c_int64, _i32p, C.c_int64, _i32p, C.c_int64,
               _f64p, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.oracle_jacobian_forward.argtypes = jac
        L.oracle_jacobian_forward.restype = C.c_int
        L.oracle_jacobian_reverse.argtypes = jac
        L.oracle_jacobian_reverse.restype = C.c_int
        L.oracle_levels.argtypes = [_i32p, C.c_int64, _i32p, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_int, _i32p]
        L.oracle_levels.restype = C.c_int64
        _oracle = L
    return _oracle


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def tangent_linear(stmt, mult, idx, gradient):
    """In-place forward sweep (adept/Stack.cpp:170-190). Returns the gradient vector."""
    stmt = _c(stmt, np.int32).reshape(-1)
    g = np.array(gradient, dtype=np.float64, copy=True)
    _lib().oracle_tangent_linear(stmt, stmt.size // 2, _c(mult, np.float64), _c(idx, np.int32), g)
    return g


def adjoint(stmt, mult, idx, gradient):
    """In-place reverse sweep (adept/Stack.cpp:139-164). Returns the gradient vector."""
    stmt = _c(stmt, np.int32).reshape(-1)
    g = np.array(gradient, dtype=np.float64, copy=True)
    _lib().oracle_adjoint(stmt, stmt.size // 2, _c(mult, np.float64), _c(idx, np.int32), g)
    return g


def jacobian(stmt, mult, idx, max_gradient, indep, dep, mode, dep_offset=1, indep_offset=0,
             P=2, n_threads=1, out=None):
    """mode 0 = forward, 1 = reverse (adept/jacobian.cpp).  Returns (rc, flat out buffer).

    ``out`` defaults to a dense n_dep*n_indep buffer pre-filled with NaN so that untouched
    elements are visible."""
    stmt = _c(stmt, np.int32).reshape(-1)
    indep = _c(indep, np.int32)
    dep = _c(dep, np.int32)
    if out is None:
        out = np.full(max(1, indep.size * dep.size), np.nan)
    fn = _lib().oracle_jacobian_forward if mode == 0 else _lib().oracle_jacobian_reverse
    rc = fn(stmt, stmt.size // 2, _c(mult, np.float64), _c(idx, np.int32), int(max_gradient),
            indep, indep.size, dep, dep.size, out, int(dep_offset), int(indep_offset), int(P),
            int(n_threads))
    return rc, out


def levels(stmt, idx, max_gradient, mode, window=0, huge=0, pack=0):
    """Sequential reference of the device-side dependency analysis. Returns (n_steps, global step of each statement)."""
    stmt = _c(stmt, np.int32).reshape(-1)
    lv = np.zeros(stmt.size // 2, dtype=np.int32)
    steps = _lib().oracle_levels(stmt, stmt.size // 2, _c(idx, np.int32), int(max_gradient), int(mode),
                                 int(window), int(huge), int(pack), lv)
    return int(steps), lv


# ----------------------------------------------------------------------------
# the real reference
# ----------------------------------------------------------------------------
_ref_libs = {}


def ref_available(variant: str = "") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libadept_ref{variant}.so"))


def best_ref_variant() -> str:
    """Widest SIMD build the host CPU can run ('' = -O2 SSE2 P=2, the parity build of record)."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return ""
    if " avx512f" in flags and ref_available("_avx512"):
        return "_avx512"
    if " avx2" in flags and ref_available("_avx2"):
        return "_avx2"
    return ""


def _ref(variant: str = ""):
    if variant not in _ref_libs:
        path = os.path.join(HERE, "_ref", f"libadept_ref{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle` where /root/reference exists)")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.ref_packet_size.restype = C.c_int
        L.ref_max_threads.restype = C.c_int
        L.ref_version.restype = C.c_char_p
        L.ref_tape_new.restype = vp
        L.ref_tape_free.argtypes = [vp]
        L.ref_tape_error.argtypes = [vp]
        L.ref_tape_error.restype = C.c_char_p
        L.ref_tape_inject.argtypes = [vp, _i32p, C.c_int64, _f64p, _i32p, C.c_int64, C.c_int64,
                                      _i32p, C.c_int64, _i32p, C.c_int64]
        L.ref_tape_set_io.argtypes = [vp, _i32p, C.c_int64, _i32p, C.c_int64]
        L.ref_tape_record_advection.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.ref_tape_record_algorithm.argtypes = [vp, C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.ref_tape_record_radiances.argtypes = [vp, C.c_int, C.c_double, _f64p, _f64p]
        L.ref_tape_record_rosenbrock.argtypes = [vp, C.c_int, C.c_double, C.POINTER(C.c_double)]
        L.ref_tape_sizes.argtypes = [vp] + [C.POINTER(C.c_int64)] * 5
        L.ref_tape_copy.argtypes = [vp, _i32p, _f64p, _i32p, _i32p, _i32p]
        L.ref_tape_jacobian.argtypes = [vp, C.c_int, _f64p, C.c_int64, C.c_int64, C.c_int,
                                        C.POINTER(C.c_double)]
        L.ref_tape_sweep.argtypes = [vp, C.c_int, _f64p, C.c_int, C.POINTER(C.c_double)]
        L.ref_tape_sweep_uninitialised.argtypes = [vp, C.c_int]
        if hasattr(L, "ref_ensemble_radiances"):
            L.ref_ensemble_radiances.argtypes = [C.c_int64, C.c_int, _f64p, _f64p, C.c_int, C.POINTER(C.c_double)]
            L.ref_ensemble_radiances.restype = C.c_int
        _ref_libs[variant] = L
    return _ref_libs[variant]


def ref_ensemble_radiances(state, n_threads=0, variant=""):
    """BASELINE config 3, CPU arm: the reference's own simulate_radiances() (record + Stack::jacobian per member) under
    OpenMP, one Stack per call.  state: [members, n+1].  Returns (jacobians [members, n+1, 2], seconds)."""
    L = _ref(variant)
    state = _c(state, np.float64)
    B, n1 = state.shape
    out = np.full((B, n1, 2), np.nan)
    sec = C.c_double()
    rc = L.ref_ensemble_radiances(B, n1 - 1, state.reshape(-1), out.reshape(-1), int(n_threads), C.byref(sec))
    if rc:
        raise RuntimeError("ref_ensemble_radiances failed")
    return out, sec.value


class RefError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"reference returned {rc}: {msg}")
        self.rc = rc


class RefTape:
    """One adept::Stack of the real reference, recorded or injected, replayed by the reference."""

    def __init__(self, variant: str = ""):
        self.L = _ref(variant)
        self.h = self.L.ref_tape_new()
        if not self.h:
            raise MemoryError("ref_tape_new")
        self.last_seconds = 0.0

    def close(self):
        if self.h:
            self.L.ref_tape_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise RefError(rc, self.L.ref_tape_error(self.h).decode())

    @property
    def packet_size(self):
        return self.L.ref_packet_size()

    # -- construction ---------------------------------------------------
    def inject(self, stmt, mult, idx, n_slots, indep, dep):
        """n_slots gradient slots are registered; max_gradient becomes n_slots+1."""
        stmt = _c(stmt, np.int32).reshape(-1)
        indep = _c(indep, np.int32)
        dep = _c(dep, np.int32)
        mult = _c(mult, np.float64)
        idx = _c(idx, np.int32)
        self._chk(self.L.ref_tape_inject(self.h, stmt, stmt.size // 2, mult, idx, idx.size,
                                         int(n_slots), indep, indep.size, dep, dep.size))
        return self

    def set_io(self, indep, dep):
        indep = _c(indep, np.int32)
        dep = _c(dep, np.int32)
        self._chk(self.L.ref_tape_set_io(self.h, indep, indep.size, dep, dep.size))
        return self

    def record_advection(self, scheme, nx, nt, c=0.125, q0=None):
        q = None if q0 is None else _c(q0, np.float64)
        ptr = None if q is None else q.ctypes.data_as(C.c_void_p)
        self._chk(self.L.ref_tape_record_advection(self.h, {"lw": 0, "toon": 1}[scheme], nx, nt, c, ptr))
        return self

    def record_algorithm(self, x0, x1):
        y = C.c_double()
        self._chk(self.L.ref_tape_record_algorithm(self.h, x0, x1, C.byref(y)))
        return y.value

    def record_radiances(self, surface_temperature, temperature):
        t = _c(temperature, np.float64)
        r = np.zeros(2)
        self._chk(self.L.ref_tape_record_radiances(self.h, t.size, surface_temperature, t, r))
        return r

    def record_rosenbrock(self, n, x0=-3.0):
        c = C.c_double()
        self._chk(self.L.ref_tape_record_rosenbrock(self.h, n, x0, C.byref(c)))
        return c.value

    # -- read-out ---------------------------------------------------------
    def sizes(self):
        v = [C.c_int64() for _ in range(5)]
        self.L.ref_tape_sizes(self.h, *[C.byref(x) for x in v])
        return tuple(int(x.value) for x in v)   # n_stmt, n_ops, max_gradient, n_indep, n_dep

    def arrays(self):
        ns, no, g, ni, nd = self.sizes()
        stmt = np.zeros(2 * ns, np.int32)
        mult = np.zeros(max(no, 1), np.float64)
        idx = np.zeros(max(no, 1), np.int32)
        indep = np.zeros(max(ni, 1), np.int32)
        dep = np.zeros(max(nd, 1), np.int32)
        self.L.ref_tape_copy(self.h, stmt, mult, idx, indep, dep)
        return dict(stmt=stmt.reshape(-1, 2), mult=mult[:no], idx=idx[:no], max_gradient=g,
                    indep=indep[:ni], dep=dep[:nd])

    # -- replay ---------------------------------------------------------
    def jacobian(self, mode, dep_offset=1, indep_offset=0, n_threads=1, out=None):
        ns, no, g, ni, nd = self.sizes()
        if out is None:
            out = np.full(max(1, ni * nd), np.nan)
        sec = C.c_double()
        self._chk(self.L.ref_tape_jacobian(self.h, mode, out, dep_offset, indep_offset, n_threads,
                                           C.byref(sec)))
        self.last_seconds = sec.value
        return out

    def sweep(self, adjoint, gradient, clear=True):
        g = np.array(gradient, dtype=np.float64, copy=True)
        sec = C.c_double()
        self._chk(self.L.ref_tape_sweep(self.h, int(bool(adjoint)), g, int(bool(clear)), C.byref(sec)))
        self.last_seconds = sec.value
        return g

    def sweep_uninitialised(self, adjoint):
        return self.L.ref_tape_sweep_uninitialised(self.h, int(bool(adjoint)))
