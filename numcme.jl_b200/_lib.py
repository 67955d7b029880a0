"""ctypes binding of libncme.so (the C ABI declared in include/ncme.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
compute entry point is called, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libncme.so")


class NcmeError(RuntimeError):
    """A libncme call failed (CUDA / NCCL / solver failure)."""


class ArgumentError(ValueError):
    """Mirror of Julia's ArgumentError / DimensionMismatch raised by the reference."""


class SeparabilityError(NcmeError):
    """A joint propensity that was put on the separable path (rank-1 detection) turned out not to be a product
    c(t) g(x) at a time actually used.  ``solve`` catches it and repeats the segment on the exact joint path."""


class CallbackGuard:
    """ctypes prints and swallows exceptions raised inside callbacks, so the C integrator would carry on with stale
    data.  Callbacks run under ``guard.wrap``: the first exception is stored, libncme is told to stop
    (``ncme_request_abort``) and ``guard.reraise()`` raises it in the caller once the C call has returned."""

    def __init__(self):
        self.exc = None

    def wrap(self, fn):
        def guarded(*args):
            if self.exc is not None:
                return
            try:
                fn(*args)
            except BaseException as e:          # noqa: BLE001 -- must not unwind through C frames
                self.exc = e
                load().ncme_request_abort()
        return guarded

    def reraise(self):
        exc, self.exc = self.exc, None
        if exc is not None:
            raise exc


p_void = C.c_void_p
p_i64 = C.POINTER(C.c_int64)
p_i32 = C.POINTER(C.c_int32)
p_u32 = C.POINTER(C.c_uint32)
p_f64 = C.POINTER(C.c_double)
i64 = C.c_int64
f64 = C.c_double
cint = C.c_int

# name -> (restype, argtypes).  Every symbol include/ncme.h declares appears here.
SIGNATURES = {
    "ncme_version": (cint, []),
    "ncme_last_error": (C.c_char_p, []),
    "ncme_ctx_create": (cint, [cint, C.POINTER(p_void)]),
    "ncme_ctx_destroy": (cint, [p_void]),
    "ncme_ctx_set_stream": (cint, [p_void, p_void]),
    "ncme_ctx_sync": (cint, [p_void]),
    "ncme_ctx_device_info": (cint, [p_void, p_i64]),
    "ncme_ctx_launch_count": (cint, [p_void, p_i64]),
    "ncme_dmalloc": (cint, [p_void, C.c_size_t, C.POINTER(p_void)]),
    "ncme_dfree": (cint, [p_void, p_void]),
    "ncme_h2d": (cint, [p_void, p_void, p_void, C.c_size_t]),
    "ncme_d2h": (cint, [p_void, p_void, p_void, C.c_size_t]),
    "ncme_host_alloc": (cint, [C.c_size_t, C.POINTER(p_void)]),
    "ncme_host_free": (cint, [p_void]),
    "ncme_space_create": (cint, [p_void, cint, cint, p_i64, i64, p_i64, C.POINTER(p_void)]),
    "ncme_space_from_host": (cint, [p_void, cint, cint, p_i64, i64, p_i64, p_u32, p_u32, C.POINTER(p_void)]),
    "ncme_space_destroy": (cint, [p_void]),
    "ncme_space_expand": (cint, [p_void, cint, cint, p_i32]),
    "ncme_space_delete": (cint, [p_void, i64, p_i64]),
    "ncme_space_state_count": (cint, [p_void, p_i64]),
    "ncme_space_sink_count": (cint, [p_void, p_i64]),
    "ncme_space_download_states": (cint, [p_void, i64, i64, p_i64]),
    "ncme_space_download_state_columns": (cint, [p_void, i64, i64, p_f64]),
    "ncme_space_download_connectivity": (cint, [p_void, i64, i64, p_u32, p_u32]),
    "ncme_space_lookup": (cint, [p_void, i64, p_i64, p_u32]),
    "ncme_space_marginal": (cint, [p_void, p_void, cint, p_i32, i64, p_i64, p_i64, p_f64]),
    "ncme_matrix_create": (cint, [p_void, p_i32, p_f64, C.POINTER(p_void)]),
    "ncme_space_new_count": (cint, [p_void, p_i64, p_i64]),
    "ncme_matrix_create_incremental": (cint, [p_void, p_void, p_void, p_i32, p_f64, C.POINTER(p_void)]),
    "ncme_matrix_destroy": (cint, [p_void]),
    "ncme_matrix_size": (cint, [p_void, p_i64, p_i64]),
    "ncme_matrix_set_joint_values": (cint, [p_void, cint, p_f64]),
    "ncme_matrix_set_tuning": (cint, [p_void, cint]),
    "ncme_matvec": (cint, [p_void, p_f64, p_void, p_void, f64]),
    "ncme_matrix_set_pipe": (cint, [p_void, cint, cint]),
    "ncme_matrix_compression_info": (cint, [p_void, p_i64]),
    "ncme_matvec_local": (cint, [p_void, p_f64, p_void, p_void]),
    "ncme_matvec_host": (cint, [p_void, p_f64, p_void, p_void, f64]),
    "ncme_matrix_stats": (cint, [p_void, C.POINTER(cint), p_i64, p_i64, p_i64]),
    "ncme_sensmatrix_create": (cint, [p_void, cint, cint, p_i32, p_i32, p_f64, C.POINTER(p_void)]),
    "ncme_sensmatrix_create_incremental": (cint, [p_void, p_void, cint, cint, p_i32, p_i32, p_f64, C.POINTER(p_void)]),
    "ncme_sensmatrix_destroy": (cint, [p_void]),
    "ncme_sensmatrix_set_joint_values": (cint, [p_void, cint, p_f64]),
    "ncme_sens_matvec": (cint, [p_void, p_f64, p_f64, p_void, p_void]),
    "ncme_sensmatrix_set_tuning": (cint, [p_void, cint]),
    "ncme_sensmatrix_stats": (cint, [p_void, C.POINTER(cint), p_i64, p_i64, p_i64]),
    "ncme_comm_unique_id": (cint, [C.c_char_p]),
    "ncme_comm_create": (cint, [p_void, cint, cint, C.c_char_p, C.POINTER(p_void)]),
    "ncme_comm_destroy": (cint, [p_void]),
    "ncme_comm_rank": (cint, [p_void, C.POINTER(cint), C.POINTER(cint)]),
    "ncme_comm_allreduce_sum": (cint, [p_void, p_void, i64]),
    "ncme_comm_allgatherv": (cint, [p_void, p_void, p_void, p_i64, p_i64]),
    "ncme_matrix_register_buffer": (cint, [p_void, p_void, C.c_size_t, i64]),
    "ncme_matrix_unregister_buffer": (cint, [p_void, p_void]),
    "ncme_comm_info": (cint, [p_void, p_i64]),
    "ncme_matrix_create_sharded": (cint, [p_void, p_void, p_i32, p_f64, C.POINTER(p_void)]),
    "ncme_matrix_shard_info": (cint, [p_void, p_i64]),
    "ncme_matrix_shard_window": (cint, [p_void, p_void, p_i64]),
    "ncme_matrix_create_window": (cint, [p_void, p_void, p_i32, p_f64, i64, i64, C.POINTER(p_void)]),
    "ncme_space_prune_by_mass": (cint, [p_void, p_void, f64, cint, p_i64]),
    "ncme_space_compact_vector": (cint, [p_void, p_void, p_void]),
    "ncme_solve_segment": (cint, [p_void, p_void, p_void, p_void, f64, f64, p_void, p_void, p_void]),
    "ncme_sens_solve_segment": (cint, [p_void, p_void, p_void, p_void, f64, f64, p_void, p_void, p_void]),
    "ncme_request_abort": (None, []),
    "ncme_vec_fill": (cint, [p_void, i64, f64, p_void]),
    "ncme_vec_copy": (cint, [p_void, i64, p_void, p_void]),
    "ncme_vec_scale": (cint, [p_void, i64, f64, p_void]),
    "ncme_vec_axpy": (cint, [p_void, i64, f64, p_void, p_void]),
    "ncme_vec_lincomb": (cint, [p_void, i64, cint, p_f64, C.POINTER(p_void), p_void]),
    "ncme_vec_sum": (cint, [p_void, i64, p_void, p_f64]),
    "ncme_vec_dot": (cint, [p_void, i64, p_void, p_void, p_f64]),
    "ncme_vec_wrms": (cint, [p_void, i64, p_void, p_void, p_void, f64, f64, p_f64]),
    "ncme_vec_any_nonfinite": (cint, [p_void, i64, p_void, C.POINTER(cint)]),
    "ncme_vec_residuals": (cint, [p_void, i64, p_void, p_void, p_void, f64, f64, p_void]),
    "ncme_vec_shift": (cint, [p_void, i64, f64, p_void]),
}

COEF_FN = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.c_void_p)
SAVE_FN = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.c_void_p)


class SolveOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("event_slope", C.c_double), ("check_event", C.c_int),
                ("save_every_step", C.c_int), ("nsave", C.c_int), ("save_t", C.POINTER(C.c_double)),
                ("h_init", C.c_double), ("max_steps", C.c_int64), ("method", C.c_int)]


class SolveStats(C.Structure):
    _fields_ = [("t_final", C.c_double), ("h_last", C.c_double), ("event_hit", C.c_int), ("nsaved", C.c_int),
                ("steps", C.c_int64), ("rejected", C.c_int64), ("rhs_evals", C.c_int64), ("launches", C.c_int64)]


_lib = None


def load():
    """Load libncme.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NcmeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  libncme has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(lib, name):
            continue  # optional symbols are checked by tests/test_abi.py against include/ncme.h
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().ncme_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int):
    if status == 0:
        return
    msg = last_error()
    if status == -1:
        raise ArgumentError(msg)
    if status == -3:
        raise MemoryError(msg)
    raise NcmeError(f"libncme error {status}: {msg}")


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))
