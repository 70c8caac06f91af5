"""Host-side mirror of the reference's ``adept::Stack`` replay interface, over the C-ABI.

``Engine`` is a thin handle on one ``adept_cuda_ctx``.  ``Stack`` mirrors the part of
``adept::Stack`` (reference include/adept/Stack.h) that sits on the tape-replay path -- same
member names, argument meaning and error behaviour -- so that parity tests read like the
reference's own tests (test/test_thread_safe.cpp:36-113, test/test_adept.cpp:147-195):

    s = Stack(); s.new_recording(); s.push_rhs(m, i) ...; s.push_lhs(j)
    s.independent(x); s.dependent(y); J = s.jacobian()
    s.set_gradients(i0, i1, v); s.compute_adjoint(); s.get_gradients(i0, i1)

Recording proper (expression templates) is C++ and stays in the reference; here the tape is
either pushed statement by statement (small tests) or loaded as arrays (``load_tape``).
Every replay call goes to the CUDA engine; there is no host fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


# -- exceptions named after the reference's (include/adept/exception.h:30-100) ------------
class AdeptException(RuntimeError):
    pass


class gradients_not_initialized(AdeptException):
    pass


class gradient_out_of_range(AdeptException):
    pass


class dependents_or_independents_not_identified(AdeptException):
    pass


class size_mismatch(AdeptException):
    pass


class cuda_replay_error(AdeptException):
    """Anything the reference has no name for: CUDA/NCCL failures, engine limits."""


_ERR = {
    _abi.ERR_NOT_IDENTIFIED: dependents_or_independents_not_identified,
    _abi.ERR_GRAD_NOT_INIT: gradients_not_initialized,
    _abi.ERR_RANGE: gradient_out_of_range,
}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One adept_cuda_ctx.  ``devices``: list of CUDA ordinals of this process (default [0])."""

    def __init__(self, devices=None):
        self.L = _abi.lib()
        devs = [0] if devices is None else list(devices)
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        rc = self.L.adept_cuda_create(C.byref(h), arr, len(devs))
        if rc:
            raise cuda_replay_error(f"adept_cuda_create failed with code {rc} (no usable CUDA device?)")
        self.h = h
        self.devices = devs
        self.tape_shape = None

    def close(self):
        if getattr(self, "h", None):
            self.L.adept_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            msg = self.L.adept_cuda_last_error(self.h).decode()
            raise _ERR.get(rc, cuda_replay_error)(f"[{rc}] {msg}")

    # -- multi-process (one rank per GPU) ------------------------------------------------
    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._chk(self.L.adept_cuda_unique_id(buf))
        return buf.raw

    def join(self, uid: bytes, rank: int, world: int):
        buf = C.create_string_buffer(uid, 128)
        self._chk(self.L.adept_cuda_join(self.h, buf, rank, world))

    def share_tape(self, root: int = 0):
        self._chk(self.L.adept_cuda_share_tape(self.h, root))

    # -- tape -----------------------------------------------------------------------------
    def upload_tape(self, stmt, mult, idx, max_gradient, epoch=0):
        """stmt: int32 [n_stmt,2] incl. the {-1,0} sentinel; mult float64 [n_ops]; idx int32 [n_ops].
        The arrays may be numpy arrays or pinned torch tensors' numpy views."""
        stmt = np.ascontiguousarray(stmt, dtype=np.int32).reshape(-1, 2)
        mult = np.ascontiguousarray(mult, dtype=np.float64)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        if mult.shape[0] != idx.shape[0]:
            raise ValueError("multiplier and index arrays differ in length")
        self._chk(self.L.adept_cuda_upload_tape(self.h, _ptr(stmt), stmt.shape[0], _ptr(mult), _ptr(idx),
                                                idx.shape[0], int(max_gradient), int(epoch)))
        self.tape_shape = (stmt.shape[0], idx.shape[0], int(max_gradient))

    def tape_tag(self):
        """The caller tag (``epoch``) of the tape this context holds, or None when it holds none."""
        e = C.c_uint64()
        rc = self.L.adept_cuda_tape_epoch(self.h, C.byref(e), None, None)
        return None if rc else int(e.value)

    def set_option(self, key: str, value: int):
        self._chk(self.L.adept_cuda_set_option(self.h, key.encode(), int(value)))

    def stat(self, key: str, default=None):
        v = C.c_double()
        rc = self.L.adept_cuda_get_stat(self.h, key.encode(), C.byref(v))
        if rc:
            return default
        return v.value

    def levels(self):
        n_stmt = self.tape_shape[0]
        out = np.zeros(n_stmt, np.int32)
        self._chk(self.L.adept_cuda_get_levels(self.h, _ptr(out), n_stmt))
        return out

    # -- replay ---------------------------------------------------------------------------
    def jacobian(self, mode, indep, dep, out=None, dep_offset=1, indep_offset=0):
        """Host-output Jacobian.  ``out`` is a flat float64 buffer (default: dense n_dep*n_indep,
        NaN-prefilled so that untouched elements show)."""
        indep = np.ascontiguousarray(indep, dtype=np.int32)
        dep = np.ascontiguousarray(dep, dtype=np.int32)
        if out is None:
            out = np.full(max(1, indep.size * dep.size), np.nan)
        self._chk(self.L.adept_cuda_jacobian(self.h, int(mode), _ptr(indep), indep.size, _ptr(dep), dep.size,
                                             _ptr(out), int(dep_offset), int(indep_offset)))
        return out

    def jacobian_device(self, mode, indep, dep, out_ptr: int, dep_offset=1, indep_offset=0):
        """Device-output Jacobian: ``out_ptr`` is a raw device pointer on devices[0]."""
        indep = np.ascontiguousarray(indep, dtype=np.int32)
        dep = np.ascontiguousarray(dep, dtype=np.int32)
        self._chk(self.L.adept_cuda_jacobian_device(self.h, int(mode), _ptr(indep), indep.size, _ptr(dep),
                                                    dep.size, C.c_void_p(out_ptr), int(dep_offset),
                                                    int(indep_offset)))

    def tangent_linear(self, gradient, initialized=True, flags=0):
        g = np.array(gradient, dtype=np.float64, copy=True)
        self._chk(self.L.adept_cuda_tangent_linear(self.h, _ptr(g), int(bool(initialized)), int(flags)))
        return g

    def adjoint(self, gradient, initialized=True, flags=0):
        g = np.array(gradient, dtype=np.float64, copy=True)
        self._chk(self.L.adept_cuda_adjoint(self.h, _ptr(g), int(bool(initialized)), int(flags)))
        return g

    def adjoint_device(self, gradient_ptr: int, flags=0):
        """In-place reverse sweep of a device-resident gradient vector (raw device pointer on devices[0])."""
        self._chk(self.L.adept_cuda_adjoint_device(self.h, C.c_void_p(gradient_ptr), int(flags)))

    def tangent_linear_device(self, gradient_ptr: int, flags=0):
        """In-place forward sweep of a device-resident gradient vector."""
        self._chk(self.L.adept_cuda_tangent_linear_device(self.h, C.c_void_p(gradient_ptr), int(flags)))

    # -- checkpoint-chained replay (reference test/test_checkpoint.cpp:166-225) -------------------
    def chain_begin(self, n_handoff: int, handoff=None):
        h = None if handoff is None else np.ascontiguousarray(handoff, dtype=np.float64)
        self._chk(self.L.adept_cuda_chain_begin(self.h, int(n_handoff), _ptr(h)))

    def chain_push(self, stmt, mult, idx, max_gradient, seed_slots, out_slots, seed_values=None):
        """One block: g = 0; g[seed_slots] = seed_values (None: the device-resident hand-off vector); reverse sweep;
        hand-off = g[out_slots].  Returns once the block is staged; the arrays may be reused at once."""
        stmt = np.ascontiguousarray(stmt, dtype=np.int32).reshape(-1, 2)
        mult = np.ascontiguousarray(mult, dtype=np.float64)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        ss = np.ascontiguousarray(seed_slots, dtype=np.int32)
        os_ = np.ascontiguousarray(out_slots, dtype=np.int32)
        sv = None if seed_values is None else np.ascontiguousarray(seed_values, dtype=np.float64)
        self._chk(self.L.adept_cuda_chain_push(self.h, _ptr(stmt), stmt.shape[0], _ptr(mult), _ptr(idx), idx.shape[0],
                                               int(max_gradient), _ptr(ss), ss.size, _ptr(sv), _ptr(os_), os_.size))

    def chain_end(self, n_handoff: int):
        out = np.full(int(n_handoff), np.nan)
        self._chk(self.L.adept_cuda_chain_end(self.h, _ptr(out)))
        return out

    def jacobian_batch(self, mode, stmt, idx, max_gradient, multipliers, indep, dep):
        """multipliers: float64 [n_ops, n_members].  Returns [n_members, n_indep, n_dep] (each member
        column-major n_dep x n_indep, the reference's default layout)."""
        stmt = np.ascontiguousarray(stmt, dtype=np.int32).reshape(-1, 2)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        mm = np.ascontiguousarray(multipliers, dtype=np.float64)
        indep = np.ascontiguousarray(indep, dtype=np.int32)
        dep = np.ascontiguousarray(dep, dtype=np.int32)
        B = mm.shape[1] if mm.ndim == 2 else 0
        out = np.full((B, indep.size, dep.size), np.nan)
        self._chk(self.L.adept_cuda_jacobian_batch(self.h, int(mode), _ptr(stmt), stmt.shape[0], _ptr(idx),
                                                   idx.size, int(max_gradient), _ptr(mm), B, _ptr(indep),
                                                   indep.size, _ptr(dep), dep.size, _ptr(out)))
        return out


_next_tag = 0   # tags of uploaded tapes, unique in this process (0 is never used)


class Stack:
    """Mirror of adept::Stack for the replay path (reference include/adept/Stack.h:97-808)."""

    def __init__(self, engine: Engine | None = None, devices=None):
        # the engine (and with it the CUDA context) is created on the first replay call, so that
        # recording-side bookkeeping and the reference's pre-checks work -- and are testable --
        # before any device is touched
        self._engine_obj = engine
        self._devices = [0] if devices is None else list(devices)
        self._stmt = np.zeros((16, 2), np.int32)
        self._mult = np.zeros(16, np.float64)
        self._idx = np.zeros(16, np.int32)
        self._n_stmt = 0
        self._n_ops = 0
        self._i_gradient = 0
        self._max_gradient = 0
        self._gradient = None
        self._gradients_initialized = False
        self._independent = []
        self._dependent = []
        self._epoch = 0
        self._uploaded = None
        self._tag = None
        self.new_recording()

    @property
    def _engine(self) -> Engine:
        if self._engine_obj is None:
            self._engine_obj = Engine(self._devices)
        return self._engine_obj

    # -- recording interface (include/adept/StackStorageOrig.h:46-121) --------------------
    def register_gradients(self, n: int) -> int:
        """Hand out n consecutive gradient slots (Stack.h:219-231, simplified: no gap list)."""
        first = self._i_gradient
        self._i_gradient += int(n)
        self._max_gradient = max(self._max_gradient, self._i_gradient)
        return first

    def register_gradient(self) -> int:
        return self.register_gradients(1)

    def new_recording(self):
        """Stack.h:528-543: reset counters, max_gradient = i_gradient+1, push the {-1,0} sentinel."""
        self._n_stmt = 0
        self._n_ops = 0
        self.clear_independents()
        self.clear_dependents()
        self.clear_gradients()
        self._max_gradient = self._i_gradient + 1
        self.push_lhs(-1)
        self._epoch += 1

    def push_rhs(self, multiplier: float, gradient_index: int):
        if self._n_ops >= self._mult.shape[0]:
            self._mult = np.concatenate([self._mult, np.zeros_like(self._mult)])
            self._idx = np.concatenate([self._idx, np.zeros_like(self._idx)])
        self._mult[self._n_ops] = multiplier
        self._idx[self._n_ops] = gradient_index
        self._n_ops += 1
        self._epoch += 1

    def push_lhs(self, gradient_index: int):
        if self._n_stmt >= self._stmt.shape[0]:
            self._stmt = np.concatenate([self._stmt, np.zeros_like(self._stmt)])
        self._stmt[self._n_stmt] = (gradient_index, self._n_ops)
        self._n_stmt += 1
        self._epoch += 1

    def load_tape(self, stmt, mult, idx, max_gradient):
        """Adopt a whole recording (arrays as the reference holds them, sentinel included)."""
        self._stmt = np.ascontiguousarray(stmt, dtype=np.int32).reshape(-1, 2)
        self._mult = np.ascontiguousarray(mult, dtype=np.float64)
        self._idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._n_stmt = self._stmt.shape[0]
        self._n_ops = self._idx.shape[0]
        self._max_gradient = int(max_gradient)
        self._i_gradient = max(self._i_gradient, self._max_gradient - 1)
        self.clear_gradients()
        self._epoch += 1

    # -- bookkeeping ------------------------------------------------------------------------
    def n_statements(self):
        return self._n_stmt

    def n_operations(self):
        return self._n_ops

    def max_gradients(self):
        return self._max_gradient

    def independent(self, x):
        """Stack.h:431-442: x is a slot index or a sequence of them."""
        self._independent.extend(np.atleast_1d(np.asarray(x, dtype=np.int64)).tolist())

    def dependent(self, y):
        self._dependent.extend(np.atleast_1d(np.asarray(y, dtype=np.int64)).tolist())

    def n_independent(self):
        return len(self._independent)

    def n_dependent(self):
        return len(self._dependent)

    def clear_independents(self):
        self._independent = []

    def clear_dependents(self):
        self._dependent = []

    def clear_gradients(self):
        self._gradients_initialized = False

    def gradients_are_initialized(self):
        return self._gradients_initialized

    def set_max_jacobian_threads(self, n: int) -> int:
        """Reference: OpenMP thread control (adept/Stack.cpp:92-116).  Here the unit of Jacobian
        parallelism is a GPU; the call is accepted and reports the number of GPUs in use."""
        return len(self._devices)

    def max_jacobian_threads(self) -> int:
        return len(self._devices)

    # -- gradients (Stack.h:290-339, adept/Stack.cpp:515-531) ------------------------------
    def _initialize_gradients(self):
        self._gradient = np.zeros(max(self._max_gradient, 1), np.float64)
        self._gradients_initialized = True

    def set_gradients(self, start: int, end_plus_one: int, gradient):
        if not self._gradients_initialized:
            self._initialize_gradients()
        if end_plus_one > self._max_gradient:
            raise gradient_out_of_range("gradient index out of range")
        self._gradient[start:end_plus_one] = np.asarray(gradient, dtype=np.float64)[: end_plus_one - start]

    def get_gradients(self, start: int, end_plus_one: int):
        if not self._gradients_initialized:
            raise gradients_not_initialized("gradients not initialized")
        if end_plus_one > self._max_gradient:
            raise gradient_out_of_range("gradient index out of range")
        return self._gradient[start:end_plus_one].copy()

    # -- replay: everything below runs on the GPU --------------------------------------------
    def _sync_tape(self):
        """Upload when this Stack's tape changed -- or when the ENGINE no longer holds it: an Engine holds one tape,
        and another Stack sharing the engine (or an ensemble / chain call) may have replaced ours.  Every upload gets
        a process-unique tag that the engine hands back (adept_cuda_tape_epoch)."""
        global _next_tag
        key = (self._epoch, self._n_stmt, self._n_ops, self._max_gradient)
        if self._uploaded != key or self._engine.tape_tag() != self._tag:
            _next_tag += 1
            self._tag = _next_tag
            self._engine.upload_tape(self._stmt[: self._n_stmt], self._mult[: self._n_ops],
                                     self._idx[: self._n_ops], self._max_gradient, epoch=self._tag)
            self._uploaded = key

    def compute_adjoint(self, flags: int = 0):
        """adept/Stack.cpp:139-164."""
        if not self._gradients_initialized:
            raise gradients_not_initialized("gradients not initialized")
        self._sync_tape()
        self._gradient = self._engine.adjoint(self._gradient, True, flags)

    reverse = compute_adjoint

    def compute_tangent_linear(self, flags: int = 0):
        """adept/Stack.cpp:170-190."""
        if not self._gradients_initialized:
            raise gradients_not_initialized("gradients not initialized")
        self._sync_tape()
        self._gradient = self._engine.tangent_linear(self._gradient, True, flags)

    forward = compute_tangent_linear

    def _jac(self, mode, out, dep_offset, indep_offset):
        if not self._independent or not self._dependent:
            raise dependents_or_independents_not_identified(
                "Dependent or independent variables not identified before a Jacobian computation")
        self._sync_tape()
        return self._engine.jacobian(mode, self._independent, self._dependent, out, dep_offset, indep_offset)

    def jacobian_forward(self, out=None, dep_offset=1, indep_offset=0):
        """adept/jacobian.cpp:204-315.  Flat buffer; element (i,j) at i*dep_offset + j*indep_offset."""
        return self._jac(_abi.FORWARD, out, dep_offset, indep_offset)

    def jacobian_reverse(self, out=None, dep_offset=1, indep_offset=0):
        """adept/jacobian.cpp:462-647."""
        return self._jac(_abi.REVERSE, out, dep_offset, indep_offset)

    def jacobian(self, out=None, dep_offset=1, indep_offset=0):
        """Stack.h:374-385: forward when n_independent <= n_dependent, else reverse."""
        mode = _abi.FORWARD if self.n_independent() <= self.n_dependent() else _abi.REVERSE
        return self._jac(mode, out, dep_offset, indep_offset)

    def jacobian_matrix(self, mode=None):
        """The Array<2> overloads (adept/jacobian.cpp:651-707): an (n_dependent, n_independent) matrix."""
        if mode is None:
            mode = _abi.FORWARD if self.n_independent() <= self.n_dependent() else _abi.REVERSE
        flat = self._jac(mode, None, 1, 0)
        return flat.reshape(self.n_independent(), self.n_dependent()).T.copy()

    def jacobian_into(self, jac: np.ndarray, mode=None):
        """Array<2>-style call with the size check of adept/jacobian.cpp:652-655."""
        if jac.shape != (self.n_dependent(), self.n_independent()):
            raise size_mismatch("Jacobian matrix has wrong size")
        jac[...] = self.jacobian_matrix(mode)
        return jac

    def print_status(self):
        return (f"Automatic Differentiation Stack (CUDA replay)\n"
                f"   Number of statements: {self._n_stmt}\n   Number of operations: {self._n_ops}\n"
                f"   Max number of gradients: {self._max_gradient}\n"
                f"   Independents: {self.n_independent()}  Dependents: {self.n_dependent()}\n")

    # the tape as arrays (what the engine uploads)
    def arrays(self):
        return dict(stmt=self._stmt[: self._n_stmt].copy(), mult=self._mult[: self._n_ops].copy(),
                    idx=self._idx[: self._n_ops].copy(), max_gradient=self._max_gradient,
                    indep=np.array(self._independent, np.int32), dep=np.array(self._dependent, np.int32))
