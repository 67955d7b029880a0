"""ctypes loader + build recipe for oracle/cpu_matvec.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_cpu.so")


def build(force=False):
    src = os.path.join(_HERE, "cpu_matvec.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ncme_oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


class CscTerms:
    """The reference's per-term CSC matrices (Int64 indices) ready for the serial C pass."""

    def __init__(self, terms):
        self.n = terms[0][1].shape[0]
        self.coef = np.array([c for c, _ in terms], dtype=np.float64)
        self.colptr = [np.ascontiguousarray(m.indptr, dtype=np.int64) for _, m in terms]
        self.rowval = [np.ascontiguousarray(m.indices, dtype=np.int64) for _, m in terms]
        self.nzval = [np.ascontiguousarray(m.data, dtype=np.float64) for _, m in terms]
        k = len(terms)
        self._cp = (ctypes.POINTER(ctypes.c_int64) * k)(*[_p(a, ctypes.c_int64) for a in self.colptr])
        self._rv = (ctypes.POINTER(ctypes.c_int64) * k)(*[_p(a, ctypes.c_int64) for a in self.rowval])
        self._nz = (ctypes.POINTER(ctypes.c_double) * k)(*[_p(a, ctypes.c_double) for a in self.nzval])

    def matvec(self, v, out):
        lib().ncme_oracle_fsp_matvec(ctypes.c_int(len(self.colptr)), ctypes.c_int64(self.n), self._cp, self._rv,
                                     self._nz, _p(self.coef, ctypes.c_double), _p(v, ctypes.c_double),
                                     _p(out, ctypes.c_double))


class CsrOmp:
    def __init__(self, csr):
        self.n = csr.shape[0]
        self.rowptr = np.ascontiguousarray(csr.indptr, dtype=np.int64)
        self.colind = np.ascontiguousarray(csr.indices, dtype=np.int32)
        self.val = np.ascontiguousarray(csr.data, dtype=np.float64)

    def matvec(self, v, out):
        lib().ncme_oracle_csr_omp(ctypes.c_int64(self.n), _p(self.rowptr, ctypes.c_int64),
                                  _p(self.colind, ctypes.c_int32), _p(self.val, ctypes.c_double),
                                  _p(v, ctypes.c_double), _p(out, ctypes.c_double))


def num_threads():
    return int(lib().ncme_oracle_num_threads())
