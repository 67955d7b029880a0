"""FspVectorSparse / FspOutputSparse: host-side I/O containers of ``solve`` (reference:
src/fspvector/fspvector.jl:13-129, src/transientcme/sparse/fspoutput.jl:19-74).  Out of the kernels'
scope; the state list is shared between slices of one segment instead of
deep-copied per snapshot."""
from __future__ import annotations

import numpy as np

from ._lib import ArgumentError


class FspVectorSparse:
    def __init__(self, states, values, checksizes: bool = True):
        self.states = np.asarray(states, dtype=np.int64)
        if self.states.ndim == 1:          # a single state given as a flat list
            self.states = self.states.reshape(1, -1)
        self.values = np.asarray(values, dtype=np.float64)
        if checksizes and self.states.shape[0] != self.values.shape[0]:
            raise ArgumentError("State and value lists must have equal lengths.")
        self._dict = None

    @classmethod
    def from_pairs(cls, statespace, statevalpairs):
        """FspVectorSparse(statespace, [x => v, ...])  (fspvector.jl:42-55)"""
        st = statespace.get_states()
        vals = np.zeros(st.shape[0])
        xs = np.asarray([p[0] for p in statevalpairs], dtype=np.int64).reshape(len(statevalpairs), -1)
        idx = statespace.lookup(xs)
        for k, (_, v) in enumerate(statevalpairs):
            if idx[k]:
                vals[idx[k] - 1] = v
        return cls(st, vals)

    @property
    def state2idx(self):
        if self._dict is None:
            self._dict = {tuple(int(v) for v in s): i + 1 for i, s in enumerate(self.states)}
        return self._dict

    def get_states(self):
        return self.states

    def get_values(self):
        return self.values

    def nnz(self):
        return self.states.shape[0]

    def sum(self, dims=None):
        """sum(p) or the marginal sum(p, dims) over the 1-based species in ``dims`` (fspvector.jl:57-99)."""
        if dims is None:
            return float(self.values.sum())
        ns = self.states.shape[1]
        dims = sorted(set(int(d) for d in dims))
        if not (min(dims) >= 1 and max(dims) <= ns):
            raise ArgumentError(f"Input dimensions must be between 1 and {ns}.")
        keep = [k for k in range(ns) if (k + 1) not in dims]
        red = self.states[:, keep]
        m = red.shape[0]
        if m == 0 or not keep:
            tot = np.array([self.values.sum()]) if m else np.zeros(0)
            return FspVectorSparse(np.zeros((tot.size, len(keep)), dtype=np.int64), tot)
        # reduced states in order of first occurrence, values accumulated in state order (np.bincount adds its weights
        # sequentially by index: the same sums, in the same order, as the reference's loop)
        uniq, first, inv = np.unique(red, axis=0, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        vals = np.bincount(rank[np.asarray(inv).reshape(-1)], weights=self.values, minlength=order.size)
        return FspVectorSparse(uniq[order], vals)

    def to_array(self):
        """Array(p)  (fspvector.jl:107-129)"""
        if self.states.shape[0] == 0:
            raise ArgumentError("Cannot construct dense array from empty FspVector instance.")
        bounds = self.states.max(axis=0) + 1
        out = np.zeros(tuple(int(b) for b in bounds))
        out[tuple(self.states.T)] = self.values
        return out


def get_values(v: FspVectorSparse):
    """fspvector.jl:19"""
    return v.values


def nnz(v: FspVectorSparse) -> int:
    """fspvector.jl:21"""
    return v.nnz()


class FspOutputSliceSparse:
    def __init__(self, t, p, sinks):
        self.t, self.p, self.sinks = t, p, sinks


class FspOutputSparse:
    def __init__(self):
        self.t = []
        self.p = []
        self.sinks = []
        self.stats = {}

    def __len__(self):
        return len(self.t)

    def __getitem__(self, ind):
        if isinstance(ind, (list, tuple, np.ndarray)):
            return [self[int(i)] for i in ind]
        if ind < 0:
            ind += len(self.t)
        if ind >= len(self.t):
            raise ArgumentError("Requested index exceeds array limit.")
        return FspOutputSliceSparse(self.t[ind], self.p[ind], self.sinks[ind])
