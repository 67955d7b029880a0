"""Row-sharded multi-GPU support (K8): one process per GPU, NCCL communicator owned by libncme.

torch.distributed is only the plumbing that distributes the NCCL unique id (and runs gloo in the CPU
tests); halo exchange and the scalar all-reduces are issued by the library on its own streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .device import Context, DeviceVector, device_ptr


def shard_bounds(n_global: int, nranks: int):
    """Row cuts used by ncme_matrix_create_sharded: contiguous blocks, boundaries on multiples of 64."""
    cuts = [0]
    for r in range(1, nranks):
        c = (n_global * r) // nranks
        cuts.append(min(n_global, (c + 63) // 64 * 64))
    cuts.append(n_global)
    return cuts


class Comm:
    """ncme_comm: NCCL communicator of this rank.  ``Comm.from_torch()`` bootstraps it through torch.distributed."""

    def __init__(self, ctx: Context, rank: int, nranks: int, unique_id: bytes | None):
        self.ctx, self.rank, self.nranks = ctx, int(rank), int(nranks)
        h = L.p_void()
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        L.check(L.load().ncme_comm_create(ctx.handle, self.rank, self.nranks, uid, C.byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        L.check(L.load().ncme_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch(cls, ctx: Context):
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        return cls(ctx, rank, world, box[0] if world > 1 else None)

    @property
    def handle(self):
        return self._h

    def allreduce_sum_(self, v: DeviceVector):
        L.check(L.load().ncme_comm_allreduce_sum(self._h, C.c_void_p(device_ptr(v)), v.n))

    def allgatherv(self, local: DeviceVector, full: DeviceVector, counts, displs):
        cn = np.ascontiguousarray(counts, dtype=np.int64)
        dp = np.ascontiguousarray(displs, dtype=np.int64)
        L.check(L.load().ncme_comm_allgatherv(self._h, C.c_void_p(device_ptr(local)), C.c_void_p(device_ptr(full)),
                                              L.ptr(cn, C.c_int64), L.ptr(dp, C.c_int64)))

    def info(self) -> dict:
        v = (C.c_int64 * 4)()
        L.check(L.load().ncme_comm_info(self._h, v))
        return {"p2p": bool(v[0]), "p2p_matvecs": int(v[1]), "nccl_matvecs": int(v[2]), "nccl_halo_bytes": int(v[3])}

    def close(self):
        if getattr(self, "_h", None):
            L.load().ncme_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.close()
        except Exception:
            pass


class ShardedVector:
    """Local slice [rows row_lo..row_hi | R sink entries] of an FSP vector, allocated with the halo margins a
    sharded matvec input needs (ncme_matrix_shard_info)."""

    def __init__(self, A, fill=None, register=False):
        """``register=True`` (collective over the ranks) exposes the buffer for the peer-memory halo."""
        self._A = A
        self._registered = False
        info = A.shard_info()
        self.hl, self.hh = info["halo_lo"], info["halo_hi"]
        self.nloc = info["row_hi"] - info["row_lo"]
        self.R = A.nr
        self._buf = DeviceVector(A.ctx, self.hl + self.nloc + self.R + self.hh)
        self._buf.fill(0.0)
        self.v = self._buf.view(self.hl, self.nloc + self.R)
        if fill is not None:
            self.v.upload(fill)
        if register and info["nranks"] > 1:
            L.check(L.load().ncme_matrix_register_buffer(A.handle, C.c_void_p(self._buf.ptr), self._buf.n * 8, self.hl))
            self._registered = True

    def unregister(self):
        """collective"""
        if self._registered:
            L.check(L.load().ncme_matrix_unregister_buffer(self._A.handle, C.c_void_p(self._buf.ptr)))
            self._registered = False
