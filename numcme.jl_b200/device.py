"""Device context and the device-resident FSP vector.

``DeviceVector`` is the host-language wrapper the north star asks for: the FSP vector ``u`` stays in
HBM; an integrator only ever sees step-control scalars (norms, sink sums).  It wraps either memory
owned by libncme (``ncme_dmalloc``) or any CUDA buffer exposing ``data_ptr()`` (a torch tensor).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class Context:
    """One GPU, one stream (ncme_ctx).  Not thread-safe, like the reference's matvec!."""

    _default = None

    def __init__(self, device: int = 0):
        lib = L.load()
        h = L.p_void()
        L.check(lib.ncme_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._torch_stream = None

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = Context(0)
        return cls._default

    @property
    def handle(self):
        return self._h

    def use_torch_stream(self, stream=None):
        """Launch on torch's current CUDA stream so torch.cuda.Event timings see the kernels."""
        import torch

        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        self._torch_stream = s
        # torch's legacy default stream has handle 0 (== "restore own stream" in the ABI): use cudaStreamLegacy
        L.check(L.load().ncme_ctx_set_stream(self._h, C.c_void_p(s.cuda_stream or 1)))

    def sync(self):
        L.check(L.load().ncme_ctx_sync(self._h))

    def device_info(self) -> dict:
        info = (C.c_int64 * 4)()
        L.check(L.load().ncme_ctx_device_info(self._h, info))
        return {"sm_count": info[0], "l2_bytes": info[1], "total_mem": info[2], "cc": info[3]}

    def launch_count(self) -> int:
        v = C.c_int64()
        L.check(L.load().ncme_ctx_launch_count(self._h, C.byref(v)))
        return v.value

    def close(self):
        if self._h:
            L.load().ncme_ctx_destroy(self._h)
            self._h = None
            if Context._default is self:
                Context._default = None

    def __del__(self):  # best effort; explicit close() preferred
        try:
            self.close()
        except Exception:
            pass


def device_ptr(obj) -> int:
    """Raw device address of a DeviceVector or of anything with data_ptr() (torch CUDA tensor)."""
    if isinstance(obj, DeviceVector):
        return obj.ptr
    if hasattr(obj, "data_ptr"):
        if hasattr(obj, "is_cuda") and not obj.is_cuda:
            raise L.ArgumentError("expected a CUDA tensor")
        if hasattr(obj, "dtype") and str(obj.dtype) != "torch.float64":
            raise L.ArgumentError("expected a float64 tensor")
        if hasattr(obj, "is_contiguous") and not obj.is_contiguous():
            raise L.ArgumentError("expected a contiguous tensor")
        return int(obj.data_ptr())
    raise L.ArgumentError(f"not a device vector: {type(obj)!r}")


def is_device(obj) -> bool:
    return isinstance(obj, DeviceVector) or (hasattr(obj, "data_ptr") and getattr(obj, "is_cuda", False))


def vec_len(obj) -> int:
    if isinstance(obj, DeviceVector):
        return obj.n
    if hasattr(obj, "numel"):
        return int(obj.numel())
    return int(np.asarray(obj).size)


class DeviceVector:
    """float64 vector in HBM.  Supports views (offset slices) like Julia's ``view(v, a:b)``."""

    def __init__(self, ctx: Context, n: int, _ptr=None, _owner=None):
        self.ctx = ctx
        self.n = int(n)
        self._owner = _owner
        if _ptr is None:
            p = L.p_void()
            L.check(L.load().ncme_dmalloc(ctx.handle, max(self.n, 1) * 8, C.byref(p)))
            self.ptr = int(p.value)
            self._owned = True
        else:
            self.ptr = int(_ptr)
            self._owned = False

    # -- construction / transfer
    @classmethod
    def from_host(cls, ctx: Context, a) -> "DeviceVector":
        a = np.ascontiguousarray(a, dtype=np.float64)
        v = cls(ctx, a.size)
        v.upload(a)
        return v

    @classmethod
    def zeros(cls, ctx: Context, n: int) -> "DeviceVector":
        v = cls(ctx, n)
        v.fill(0.0)
        return v

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size != self.n:
            raise L.ArgumentError("size mismatch in upload")
        L.check(L.load().ncme_h2d(self.ctx.handle, C.c_void_p(self.ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))

    def to_host(self, start: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n - start if count is None else count
        out = np.empty(count, dtype=np.float64)
        if count:
            L.check(L.load().ncme_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr + 8 * start),
                                      out.nbytes))
        return out

    def download_into(self, out: np.ndarray, start: int = 0):
        """Device -> host into an existing contiguous float64 array (e.g. pinned memory): no allocation, no extra copy."""
        if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous):
            raise L.ArgumentError("download_into needs a contiguous float64 numpy array")
        if start < 0 or start + out.size > self.n:
            raise L.ArgumentError("size mismatch in download_into")
        if out.size:
            L.check(L.load().ncme_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr + 8 * start),
                                      out.nbytes))
        return out

    def view(self, start: int, count: int) -> "DeviceVector":
        if start < 0 or count < 0 or start + count > self.n:
            raise L.ArgumentError("view out of range")
        return DeviceVector(self.ctx, count, _ptr=self.ptr + 8 * start, _owner=self)

    def __len__(self):
        return self.n

    # -- K7 ops
    def fill(self, a: float):
        L.check(L.load().ncme_vec_fill(self.ctx.handle, self.n, float(a), C.c_void_p(self.ptr)))
        return self

    def copy_from(self, other):
        L.check(L.load().ncme_vec_copy(self.ctx.handle, self.n, C.c_void_p(device_ptr(other)), C.c_void_p(self.ptr)))
        return self

    def clone(self) -> "DeviceVector":
        return DeviceVector(self.ctx, self.n).copy_from(self)

    def scale(self, a: float):
        L.check(L.load().ncme_vec_scale(self.ctx.handle, self.n, float(a), C.c_void_p(self.ptr)))
        return self

    def axpy(self, a: float, x):
        """self += a * x"""
        L.check(L.load().ncme_vec_axpy(self.ctx.handle, self.n, float(a), C.c_void_p(device_ptr(x)), C.c_void_p(self.ptr)))
        return self

    def lincomb(self, coefs, xs):
        """self = sum_k coefs[k] * xs[k]   (k <= 8)"""
        k = len(coefs)
        cs = (C.c_double * k)(*[float(c) for c in coefs])
        ps = (C.c_void_p * k)(*[device_ptr(x) for x in xs])
        L.check(L.load().ncme_vec_lincomb(self.ctx.handle, self.n, k, cs, ps, C.c_void_p(self.ptr)))
        return self

    def sum(self, start: int = 0, count: int | None = None) -> float:
        count = self.n - start if count is None else count
        out = C.c_double()
        L.check(L.load().ncme_vec_sum(self.ctx.handle, count, C.c_void_p(self.ptr + 8 * start), C.byref(out)))
        return out.value

    def dot(self, other) -> float:
        out = C.c_double()
        L.check(L.load().ncme_vec_dot(self.ctx.handle, self.n, C.c_void_p(self.ptr), C.c_void_p(device_ptr(other)),
                                      C.byref(out)))
        return out.value

    def norm(self) -> float:
        return float(np.sqrt(self.dot(self)))

    def wrms(self, u0, u1, atol: float, rtol: float) -> float:
        out = C.c_double()
        L.check(L.load().ncme_vec_wrms(self.ctx.handle, self.n, C.c_void_p(self.ptr), C.c_void_p(device_ptr(u0)),
                                       C.c_void_p(device_ptr(u1)), float(atol), float(rtol), C.byref(out)))
        return out.value

    def residuals(self, x, u0, u1, atol: float, rtol: float):
        """self_i = x_i / (atol + rtol max(|u0_i|, |u1_i|))   (OrdinaryDiffEq's calculate_residuals!)"""
        L.check(L.load().ncme_vec_residuals(self.ctx.handle, self.n, C.c_void_p(device_ptr(x)), C.c_void_p(device_ptr(u0)),
                                            C.c_void_p(device_ptr(u1)), float(atol), float(rtol), C.c_void_p(self.ptr)))
        return self

    def shift(self, a: float):
        """self += a (elementwise)"""
        L.check(L.load().ncme_vec_shift(self.ctx.handle, self.n, float(a), C.c_void_p(self.ptr)))
        return self

    def any_nonfinite(self) -> bool:
        out = C.c_int()
        L.check(L.load().ncme_vec_any_nonfinite(self.ctx.handle, self.n, C.c_void_p(self.ptr), C.byref(out)))
        return bool(out.value)

    def free(self):
        if self._owned and self.ptr:
            L.load().ncme_dfree(self.ctx.handle, C.c_void_p(self.ptr))
            self.ptr = 0
            self._owned = False

    def __del__(self):
        try:
            if self.ctx.handle:
                self.free()
        except Exception:
            pass
