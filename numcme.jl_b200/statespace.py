"""StateSpaceSparse behind the reference's API, backed by the device hash table / compaction kernels.

Reference: src/statespace/sparse/sparsestatespace.jl (struct :22-40, ctor :103-144, expand! :153-194,
deleteat! :276-331, getters :47-83).  Julia's ``expand!`` / ``deleteat!`` are ``expand_`` /
``deleteat_`` here (trailing underscore = mutating).  Indices are 1-based with 0 = none in everything
this class returns, as in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .device import Context


class StateSpaceSparse:
    def __init__(self, stoich_matrix, initstates, ctx: Context | None = None, _handle=None):
        self.ctx = ctx or Context.default()
        self.stoich_matrix = np.asarray(stoich_matrix, dtype=np.int64)
        if self.stoich_matrix.ndim != 2:
            raise L.ArgumentError("stoichiometry matrix must be 2-D (species x reactions)")
        self.ns, self.nr = self.stoich_matrix.shape
        self._states_cache = None
        self._states_cache_n = 0
        self.version = 0                         # bumped by every mutation (expand_, deleteat_, prune_by_mass_)
        if _handle is not None:
            self._h = _handle
            return
        init = np.asarray(initstates, dtype=np.int64)
        if init.ndim == 1:                       # single-state constructor (:142-144)
            init = init[None, :]
        init = np.ascontiguousarray(init.reshape(-1, self.ns))
        st = self._stoich_colmajor()
        h = L.p_void()
        L.check(L.load().ncme_space_create(self.ctx.handle, self.ns, self.nr, L.ptr(st, C.c_int64), init.shape[0],
                                           L.ptr(init, C.c_int64), C.byref(h)))
        self._h = h

    def _stoich_colmajor(self):
        # Julia Matrix layout: column (reaction) major
        return np.ascontiguousarray(self.stoich_matrix.T.reshape(-1), dtype=np.int64)

    @classmethod
    def from_host(cls, stoich_matrix, states, state_connectivity, sink_connectivity, ctx: Context | None = None):
        """Import a space (reference conventions: 1-based connectivity, 0 = none)."""
        self = cls.__new__(cls)
        ctx = ctx or Context.default()
        S = np.asarray(stoich_matrix, dtype=np.int64)
        ns, nr = S.shape
        states = np.ascontiguousarray(np.asarray(states, dtype=np.int64).reshape(-1, ns))
        sc = np.ascontiguousarray(state_connectivity, dtype=np.uint32).reshape(-1, nr)
        kc = np.ascontiguousarray(sink_connectivity, dtype=np.uint32).reshape(-1, nr)
        if not (states.shape[0] == sc.shape[0] == kc.shape[0]):
            raise L.ArgumentError("states and connectivity tables must have the same number of rows")
        st = np.ascontiguousarray(S.T.reshape(-1), dtype=np.int64)
        h = L.p_void()
        L.check(L.load().ncme_space_from_host(ctx.handle, ns, nr, L.ptr(st, C.c_int64), states.shape[0],
                                              L.ptr(states, C.c_int64), L.ptr(sc, C.c_uint32), L.ptr(kc, C.c_uint32),
                                              C.byref(h)))
        StateSpaceSparse.__init__(self, S, None, ctx=ctx, _handle=h)
        return self

    @property
    def handle(self):
        return self._h

    # -- getters
    def get_state_count(self) -> int:
        n = C.c_int64()
        L.check(L.load().ncme_space_state_count(self._h, C.byref(n)))
        return n.value

    def get_sink_count(self) -> int:
        return self.nr

    def get_stoich_matrix(self):
        return self.stoich_matrix

    def get_states(self, first: int = 0, count: int | None = None) -> np.ndarray:
        """States as an (n x NS) int64 array in index order (insertion order, as the reference)."""
        n = self.get_state_count()
        whole = first == 0 and (count is None or count == n)
        if whole and self._states_cache is not None and self._states_cache_n == n:
            return self._states_cache          # read-only view of the last download; invalidated by every mutation
        count = n - first if count is None else count
        out = np.empty((count, self.ns), dtype=np.int64)
        if count:
            L.check(L.load().ncme_space_download_states(self._h, first, count, L.ptr(out, C.c_int64)))
        if whole:
            out.setflags(write=False)
            self._states_cache, self._states_cache_n = out, n
        return out

    def get_state_columns(self, first: int = 0, count: int | None = None) -> list:
        """The states [first, first + count) species-major as float64 columns (one contiguous array per species): what
        a vectorised propensity evaluation over many states reads."""
        n = self.get_state_count()
        count = n - first if count is None else count
        buf = np.empty((self.ns, max(count, 0)), dtype=np.float64)
        if count:
            L.check(L.load().ncme_space_download_state_columns(self._h, first, count, L.ptr(buf, C.c_double)))
        return [buf[k] for k in range(self.ns)]

    @property
    def states(self):
        return self.get_states()

    def _connectivity(self):
        n = self.get_state_count()
        sc = np.zeros((n, self.nr), dtype=np.uint32)
        kc = np.zeros((n, self.nr), dtype=np.uint32)
        if n:
            L.check(L.load().ncme_space_download_connectivity(self._h, 0, n, L.ptr(sc, C.c_uint32), L.ptr(kc, C.c_uint32)))
        return sc, kc

    def get_state_connectivity(self) -> np.ndarray:
        return self._connectivity()[0]

    def get_sink_connectivity(self) -> np.ndarray:
        return self._connectivity()[1]

    def lookup(self, states) -> np.ndarray:
        """``get(state2idx, x, 0)`` for every row of ``states`` -> uint32 (1-based, 0 = absent)."""
        q = np.ascontiguousarray(np.asarray(states, dtype=np.int64).reshape(-1, self.ns))
        out = np.zeros(q.shape[0], dtype=np.uint32)
        if q.shape[0]:
            L.check(L.load().ncme_space_lookup(self._h, q.shape[0], L.ptr(q, C.c_int64), L.ptr(out, C.c_uint32)))
        return out

    def marginal(self, p_dev, dims):
        """``sum(p, dims)`` (fspvector.jl:66-99) for a DEVICE-resident probability vector over this space: the reduction
        runs on the GPU (temporary hash table + compaction + fp64 atomics), only the reduced states and values come
        back.  ``dims`` are the 1-based species summed out.  Returns an FspVectorSparse over the kept species."""
        from .device import device_ptr, vec_len
        from .fspvector import FspVectorSparse
        n = self.get_state_count()
        if vec_len(p_dev) < n:
            raise L.ArgumentError("State and value lists must have equal lengths.")
        dims = sorted(set(int(d) for d in dims))
        if not dims or not (dims[0] >= 1 and dims[-1] <= self.ns):
            raise L.ArgumentError(f"Input dimensions must be between 1 and {self.ns}.")
        d = np.ascontiguousarray(dims, dtype=np.int32)
        nkeep = self.ns - len(dims)
        nred = C.c_int64()
        ptr = C.c_void_p(device_ptr(p_dev))
        L.check(L.load().ncme_space_marginal(self._h, ptr, d.size, L.ptr(d, C.c_int32), 0, C.byref(nred), None, None))
        m = nred.value
        st = np.zeros((m, max(nkeep, 0)), dtype=np.int64)
        vals = np.zeros(m, dtype=np.float64)
        if m:
            L.check(L.load().ncme_space_marginal(self._h, ptr, d.size, L.ptr(d, C.c_int32), m, C.byref(nred),
                                                 L.ptr(st, C.c_int64) if nkeep else L.ptr(np.zeros(1, dtype=np.int64), C.c_int64),
                                                 L.ptr(vals, C.c_double)))
        return FspVectorSparse(st, vals)

    def get_statedict(self) -> dict:
        """The reference's ``state2idx`` Dict, materialised on the host (debug / small spaces)."""
        return {tuple(int(v) for v in s): i + 1 for i, s in enumerate(self.get_states())}

    # -- mutation
    def expand_(self, expansionlevel: int, onlyreactions=()):
        self._states_cache = None
        self.version += 1
        only = np.ascontiguousarray(list(onlyreactions), dtype=np.int32)
        L.check(L.load().ncme_space_expand(self._h, int(expansionlevel), only.size,
                                           L.ptr(only, C.c_int32) if only.size else None))

    def deleteat_(self, ids):
        self._states_cache = None
        self.version += 1
        ids = np.ascontiguousarray(np.asarray(ids, dtype=np.int64).reshape(-1))
        if ids.size:
            L.check(L.load().ncme_space_delete(self._h, ids.size, L.ptr(ids, C.c_int64)))

    def prune_by_mass_(self, p_dev, threshold: float, strict: bool) -> int:
        """The dropstates branch of adapt! on the device (rstepadapters.jl:40-46 / :93-99): deletes the
        ``dropcount`` least probable states; returns dropcount.  Follow with ``compact_vector``."""
        from .device import device_ptr
        self._states_cache = None
        self.version += 1
        dc = C.c_int64()
        L.check(L.load().ncme_space_prune_by_mass(self._h, C.c_void_p(device_ptr(p_dev)), float(threshold),
                                                  1 if strict else 0, C.byref(dc)))
        return dc.value

    def compact_vector(self, vin, vout):
        """deleteat!(p, dropids) for a device vector, using the map of the last delete/prune."""
        from .device import device_ptr
        L.check(L.load().ncme_space_compact_vector(self._h, C.c_void_p(device_ptr(vin)), C.c_void_p(device_ptr(vout))))

    def close(self):
        if getattr(self, "_h", None):
            L.load().ncme_space_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.close()
        except Exception:
            pass


def expand_(space: StateSpaceSparse, expansionlevel: int, onlyreactions=()):
    """expand!(statespace, expansionlevel; onlyreactions=[])"""
    space.expand_(expansionlevel, onlyreactions)


def deleteat_(space: StateSpaceSparse, ids):
    """deleteat!(statespace, ids)"""
    space.deleteat_(ids)


def get_state_count(space):
    return space.get_state_count()


def get_sink_count(space):
    return space.get_sink_count()


def get_states(obj):
    return obj.get_states()


def get_statedict(space):
    return space.get_statedict()


def get_state_connectivity(space):
    return space.get_state_connectivity()


def get_sink_connectivity(space):
    return space.get_sink_connectivity()
