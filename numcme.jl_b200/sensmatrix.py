"""ForwardSensFspMatrixSparse behind the reference's API (fused CUDA block matvec, K2).

Reference: src/forwardsensfspmatrix/forwardsensfspmatrixsparse/sensfspmatrixsparse.jl
(constructor :31-95, ``matvec!`` :97-142).  Vector layout ``[p; s_1; ...; s_P]``, each block of
length N = n + R, exactly as the reference.

Divergence (no reference test covers it): for joint time-varying reactions the
derivative matrix is filled with d f / d theta (the mathematically correct value); the reference
fills it with f itself.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .cmemodel import (CmeModelWithSensitivity, eval_over_states, get_parameters, get_propensities,
                       get_propensity_gradients, get_gradient_sparsity_patterns)
from .device import DeviceVector, device_ptr, is_device, vec_len
from .fspmatrix import FspMatrixSparse
from .statespace import StateSpaceSparse


class ForwardSensFspMatrixSparse:
    def __init__(self, model: CmeModelWithSensitivity, space: StateSpaceSparse,
                 previous: "ForwardSensFspMatrixSparse | None" = None):
        """``previous``: the sensitivity matrix this space was assembled into before its last prune / expand (the
        rebuild after an ``adapt!``, forwardsenscmesparse.jl:140).  Propensities AND their parameter derivatives are then
        evaluated on the appended states only; the rows of the surviving states are carried over on the device
        (``ncme_matrix_create_incremental`` + ``ncme_sensmatrix_create_incremental``).  Falls back to the full build
        whenever the plain matrix does (other model, too few states, a space assembled elsewhere in between)."""
        # the derivative entries below follow the user's classification of every reaction: no separability detection
        prevA = previous.fspmatrix if previous is not None and getattr(previous, "_h", None) else None
        self.fspmatrix = A = FspMatrixSparse(space, get_propensities(model), parameters=get_parameters(model),
                                             detect_separable=False, previous=prevA)
        self.ctx = A.ctx
        self.parameters = get_parameters(model)
        self.parameter_count = P = len(self.parameters)
        self.propensity_gradients = get_propensity_gradients(model)
        pattern = get_gradient_sparsity_patterns(model)
        # entries in the reference's order: parameter-major, reactions ascending (nzrange over CSC columns)
        ents = [(r, ip) for ip in range(P) for r in range(A.nr) if pattern[r, ip]]
        self.entries = ents
        n = A.n
        self.incremental = bool(prevA is not None and A.incremental and previous.entries == ents and
                                previous.propensity_gradients is self.propensity_gradients)
        if self.incremental:
            n_new = A.new_state_count
            states = space.get_states(n - n_new, n_new) if n_new else np.zeros((0, space.ns), dtype=np.int64)
        else:
            n_new, states = n, A.states
        dvals = np.zeros((max(len(ents), 1), max(n_new, 1)), dtype=np.float64)
        for e, (r, ip) in enumerate(ents):
            g = self.propensity_gradients[r]
            kind = A.propensities[r].kind
            if kind == "ti":
                dvals[e, :n_new] = eval_over_states(g.pardiffs[ip], states, self.parameters)
            elif kind == "sep":
                dvals[e, :n_new] = eval_over_states(g.statefactor_pardiffs[ip], states, self.parameters)
        dvals = np.ascontiguousarray(dvals[:, :n_new]) if n_new else dvals
        er = np.array([r + 1 for r, _ in ents], dtype=np.int32)
        ep = np.array([ip + 1 for _, ip in ents], dtype=np.int32)
        h = L.p_void()
        lib = L.load()
        if self.incremental:
            st = lib.ncme_sensmatrix_create_incremental(A.handle, previous._h, P, len(ents),
                                                        L.ptr(er, C.c_int32) if ents else None,
                                                        L.ptr(ep, C.c_int32) if ents else None,
                                                        L.ptr(dvals, C.c_double), C.byref(h))
            if st != 0:                      # should not happen once A was built incrementally; be safe, not wrong
                A.close()
                self.__init__(model, space)
                return
        else:
            L.check(lib.ncme_sensmatrix_create(A.handle, P, len(ents), L.ptr(er, C.c_int32) if ents else None,
                                               L.ptr(ep, C.c_int32) if ents else None,
                                               L.ptr(dvals, C.c_double), C.byref(h)))
        self._h = h
        self._dcoef = np.zeros(max(len(ents), 1), dtype=np.float64)
        self._joint_entries = [e for e, (r, _) in enumerate(ents) if A.propensities[r].kind == "joint"]
        self._stage = None

    def _prepare(self, t: float):
        A = self.fspmatrix
        th = self.parameters
        coef = A.coefficients(t)
        A._refresh_joint(t)
        for e, (r, ip) in enumerate(self.entries):
            if A.propensities[r].kind == "sep":
                self._dcoef[e] = float(self.propensity_gradients[r].tfactor_pardiffs[ip](t, th))
        for e in self._joint_entries:                 # re-evaluated every call, like the reference (:137)
            r, ip = self.entries[e]
            vals = eval_over_states(self.propensity_gradients[r].pardiffs[ip], A.states, th, t=t)
            L.check(L.load().ncme_sensmatrix_set_joint_values(self._h, e, L.ptr(vals, C.c_double)))
        return coef

    def matvec_(self, out, t, vs):
        A = self.fspmatrix
        total = A.rowcount * (self.parameter_count + 1)
        if vec_len(out) != total or vec_len(vs) != total:
            raise L.ArgumentError(f"DimensionMismatch: expected vectors of length {total}")
        coef = self._prepare(float(t))
        lib = L.load()
        if is_device(out) and is_device(vs):
            L.check(lib.ncme_sens_matvec(self._h, L.ptr(coef, C.c_double), L.ptr(self._dcoef, C.c_double),
                                         C.c_void_p(device_ptr(vs)), C.c_void_p(device_ptr(out))))
            return
        if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous):
            raise L.ArgumentError("out must be a contiguous float64 numpy array or a device vector")
        if self._stage is None or self._stage[0].n != total:
            self._stage = (DeviceVector(self.ctx, total), DeviceVector(self.ctx, total))
        dx, dy = self._stage
        dx.upload(np.ascontiguousarray(vs, dtype=np.float64))
        L.check(lib.ncme_sens_matvec(self._h, L.ptr(coef, C.c_double), L.ptr(self._dcoef, C.c_double),
                                     C.c_void_p(dx.ptr), C.c_void_p(dy.ptr)))
        out[:] = dy.to_host()

    def set_tuning(self, rows_per_thread: int):
        L.check(L.load().ncme_sensmatrix_set_tuning(self._h, int(rows_per_thread)))

    def stats(self) -> dict:
        """Stored entries per derivative term as the reference holds them, B_sens and the bytes the
        fused kernel streams per matvec."""
        nt, ab, db = C.c_int(), C.c_int64(), C.c_int64()
        nnz = (C.c_int64 * (2 * max(len(self.entries), 1) + 2))()
        L.check(L.load().ncme_sensmatrix_stats(self._h, C.byref(nt), nnz, C.byref(ab), C.byref(db)))
        return {"ndterms": nt.value, "nnz_per_dterm": [nnz[k] for k in range(nt.value)],
                "algorithmic_bytes": ab.value, "device_bytes": db.value}

    def close(self):
        if getattr(self, "_h", None):
            L.load().ncme_sensmatrix_destroy(self._h)
            self._h = None
        self.fspmatrix.close()

    def __del__(self):
        try:
            if self.ctx.handle:
                self.close()
        except Exception:
            pass


def sens_matvec_(out, t, SA: ForwardSensFspMatrixSparse, vs):
    """matvec!(out, t, SA::ForwardSensFspMatrixSparse, vs)"""
    SA.matvec_(out, t, vs)
