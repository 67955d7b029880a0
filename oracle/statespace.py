"""Oracle restatement of ``StateSpaceSparse`` (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/statespace/sparse/sparsestatespace.jl:
  struct + invariants   :22-40
  constructor           :103-124  (one-state variant :142-144)
  expand!               :153-194
  _addstates!           :208-267
  deleteat!             :276-331

Indices are 1-based with 0 = "none", exactly as the reference stores them, so
that index-level comparisons with the worked example of SURVEY.md Appendix A are
literal.  Two implementations live here:

* ``StateSpaceOracle``      -- literal dict/list restatement (small cases).
* ``StateSpaceOracleFast``  -- numpy-vectorised, same ordering semantics, checked
                               against the literal one in tests; used for the
                               1e5..1e7-state parity fixtures.
"""
from __future__ import annotations

import numpy as np


def _as_stoich(stoich) -> np.ndarray:
    S = np.asarray(stoich, dtype=np.int64)
    if S.ndim != 2:
        raise ValueError("stoichiometry matrix must be 2-D (species x reactions)")
    return S


class StateSpaceOracle:
    """Literal restatement; ``states`` is a list of tuples in insertion order."""

    def __init__(self, stoich, initstates):
        self.stoich = _as_stoich(stoich)          # NS x R, column r = net change of reaction r
        self.ns, self.nr = self.stoich.shape
        self.sink_count = self.nr                  # sparsestatespace.jl:106
        self.states: list[tuple] = []
        self.state2idx: dict[tuple, int] = {}      # state -> 1-based index
        self.state_connectivity: list[list[int]] = []
        self.sink_connectivity: list[list[int]] = []
        initstates = np.asarray(initstates, dtype=np.int64)
        if initstates.ndim == 1:                   # single-state constructor, :142-144
            initstates = initstates[None, :]
        self._addstates([tuple(int(v) for v in s) for s in initstates])

    # -- getters (sparsestatespace.jl:47-83)
    def get_state_count(self):
        return len(self.states)

    def get_sink_count(self):
        return self.sink_count

    def states_array(self) -> np.ndarray:
        return np.asarray(self.states, dtype=np.int64).reshape(len(self.states), self.ns)

    def state_connectivity_array(self) -> np.ndarray:
        return np.asarray(self.state_connectivity, dtype=np.int64).reshape(len(self.states), self.nr)

    def sink_connectivity_array(self) -> np.ndarray:
        return np.asarray(self.sink_connectivity, dtype=np.int64).reshape(len(self.states), self.nr)

    # -- expand!  (sparsestatespace.jl:153-194)
    def expand(self, expansionlevel: int, onlyreactions=()):
        if expansionlevel <= 0:
            return
        S = self.stoich
        expandreactions = list(onlyreactions) if len(onlyreactions) else list(range(1, self.nr + 1))
        explorables = []                           # used as a LIFO stack (Deque push!/pop!)
        for idx in range(1, len(self.states) + 1):
            for ir in expandreactions:
                if self.sink_connectivity[idx - 1][ir - 1] != 0:
                    explorables.append(idx)
                    break
        for _level in range(expansionlevel):
            candidates = []
            while explorables:
                idx = explorables.pop()
                x = self.states[idx - 1]
                for ir in expandreactions:
                    candidates.append(tuple(x[k] + int(S[k, ir - 1]) for k in range(self.ns)))
            old = len(self.states)
            self._addstates(candidates)
            for idx in range(old + 1, len(self.states) + 1):
                explorables.append(idx)

    # -- _addstates!  (sparsestatespace.jl:208-267)
    def _addstates(self, newstates):
        S = self.stoich
        R = self.nr
        unique_new = []
        newidx = len(self.states)
        for st in newstates:
            if self.state2idx.get(st, 0) == 0 and all(v >= 0 for v in st):
                newidx += 1
                unique_new.append(st)
                self.state2idx[st] = newidx
        old = len(self.states)
        self.states.extend(unique_new)
        for _ in unique_new:
            self.state_connectivity.append([0] * R)
            self.sink_connectivity.append([0] * R)
        for new in range(old + 1, len(self.states) + 1):
            x = self.states[new - 1]
            for ir in range(1, R + 1):
                pred = tuple(x[k] - int(S[k, ir - 1]) for k in range(self.ns))
                ridx = self.state2idx.get(pred, 0)
                if ridx != 0:
                    self.state_connectivity[new - 1][ir - 1] = ridx
                    self.sink_connectivity[ridx - 1][ir - 1] = 0
            for ir in range(1, R + 1):
                succ = tuple(x[k] + int(S[k, ir - 1]) for k in range(self.ns))
                if all(v >= 0 for v in succ):
                    ridx = self.state2idx.get(succ, 0)
                    if ridx == 0:
                        self.sink_connectivity[new - 1][ir - 1] = ir
                    else:
                        self.state_connectivity[ridx - 1][ir - 1] = new

    # -- deleteat!  (sparsestatespace.jl:276-331)
    def deleteat(self, ids):
        """``ids`` are 1-based state indices (any order, duplicates tolerated)."""
        S = self.stoich
        R = self.nr
        nold = len(self.states)
        ids = sorted(set(int(i) for i in ids))
        newidxs = list(range(1, nold + 1))
        for i in ids:
            newidxs[i - 1] = 0
        nxt = 1
        for i in range(nold):
            if newidxs[i] != 0:
                newidxs[i] = nxt
                nxt += 1
        for i in ids:
            del self.state2idx[self.states[i - 1]]
        dead = set(ids)
        keep = [i for i in range(1, nold + 1) if i not in dead]
        self.states = [self.states[i - 1] for i in keep]
        self.state_connectivity = [self.state_connectivity[i - 1] for i in keep]
        self.sink_connectivity = [self.sink_connectivity[i - 1] for i in keep]
        if not self.states:
            return
        for i, x in enumerate(self.states):
            self.state2idx[x] = newidxs[self.state2idx[x] - 1]
            row = self.state_connectivity[i]
            for ir in range(R):
                row[ir] = newidxs[row[ir] - 1] if row[ir] != 0 else 0
        for i, x in enumerate(self.states):
            for ir in range(1, R + 1):
                if self.sink_connectivity[i][ir - 1] == 0:
                    succ = tuple(x[k] + int(S[k, ir - 1]) for k in range(self.ns))
                    if all(v >= 0 for v in succ) and self.state2idx.get(succ, 0) == 0:
                        self.sink_connectivity[i][ir - 1] = ir


# ----------------------------------------------------------------------------------------------
# Vectorised variant (same insertion order, same connectivity), for large parity fixtures.
# ----------------------------------------------------------------------------------------------
class _SortedKeyIndex:
    """Sorted int64 keys -> 1-based state index, with O(n) merge per insert batch."""

    def __init__(self):
        self.keys = np.empty(0, dtype=np.int64)
        self.idx = np.empty(0, dtype=np.int64)

    def lookup(self, q: np.ndarray) -> np.ndarray:
        """Return 1-based indices, 0 where absent."""
        if self.keys.size == 0:
            return np.zeros(q.shape, dtype=np.int64)
        pos = np.searchsorted(self.keys, q)
        pos_c = np.minimum(pos, self.keys.size - 1)
        hit = self.keys[pos_c] == q
        return np.where(hit, self.idx[pos_c], 0)

    def insert(self, k: np.ndarray, idx: np.ndarray):
        order = np.argsort(k, kind="stable")
        k, idx = k[order], idx[order]
        pos = np.searchsorted(self.keys, k)
        self.keys = np.insert(self.keys, pos, k)
        self.idx = np.insert(self.idx, pos, idx)

    def rebuild(self, k: np.ndarray, idx: np.ndarray):
        order = np.argsort(k, kind="stable")
        self.keys, self.idx = k[order], idx[order]


class StateSpaceOracleFast:
    """numpy restatement with identical ordering semantics (see module docstring)."""

    KEY_BITS = 62

    def __init__(self, stoich, initstates, bits_per_species=None):
        self.stoich = _as_stoich(stoich)
        self.ns, self.nr = self.stoich.shape
        self.sink_count = self.nr
        if bits_per_species is None:
            bits_per_species = [self.KEY_BITS // self.ns] * self.ns
        self.bits = np.asarray(bits_per_species, dtype=np.int64)
        assert self.bits.sum() <= 63
        self.shifts = np.concatenate(([0], np.cumsum(self.bits)[:-1])).astype(np.int64)
        self.states = np.empty((0, self.ns), dtype=np.int64)
        self.state_connectivity = np.empty((0, self.nr), dtype=np.int64)
        self.sink_connectivity = np.empty((0, self.nr), dtype=np.int64)
        self._index = _SortedKeyIndex()
        initstates = np.asarray(initstates, dtype=np.int64)
        if initstates.ndim == 1:
            initstates = initstates[None, :]
        self._addstates(initstates.reshape(-1, self.ns))

    def _pack(self, X: np.ndarray) -> np.ndarray:
        """Pack non-negative rows into int64 keys; rows with a negative entry get key -1."""
        neg = (X < 0).any(axis=1)
        if (X >= (np.int64(1) << self.bits)[None, :]).any():
            raise OverflowError("state component exceeds key width")
        k = (np.where(X < 0, 0, X) << self.shifts[None, :]).sum(axis=1)
        return np.where(neg, np.int64(-1), k)

    def get_state_count(self):
        return self.states.shape[0]

    def get_sink_count(self):
        return self.sink_count

    def states_array(self):
        return self.states

    def state_connectivity_array(self):
        return self.state_connectivity

    def sink_connectivity_array(self):
        return self.sink_connectivity

    def expand(self, expansionlevel: int, onlyreactions=()):
        if expansionlevel <= 0:
            return
        er = np.asarray(list(onlyreactions) if len(onlyreactions) else range(1, self.nr + 1), dtype=np.int64)
        Ssel = self.stoich[:, er - 1].T                      # nreact x NS
        explor = np.nonzero((self.sink_connectivity[:, er - 1] != 0).any(axis=1))[0] + 1
        for _level in range(expansionlevel):
            if explor.size == 0:
                break
            fr = explor[::-1]                                 # LIFO pop order
            cand = (self.states[fr - 1][:, None, :] + Ssel[None, :, :]).reshape(-1, self.ns)
            old = self.get_state_count()
            self._addstates(cand)
            explor = np.arange(old + 1, self.get_state_count() + 1, dtype=np.int64)

    def _addstates(self, cand: np.ndarray):
        R = self.nr
        old = self.get_state_count()
        keys = self._pack(cand)
        ok = keys >= 0
        ok &= self._index.lookup(np.where(ok, keys, 0)) == 0
        ck = keys[ok]
        _, first = np.unique(ck, return_index=True)
        first.sort()                                          # first occurrences in candidate order
        newstates = cand[ok][first]
        newkeys = ck[first]
        m = newstates.shape[0]
        newidx = np.arange(old + 1, old + m + 1, dtype=np.int64)
        self.states = np.concatenate([self.states, newstates])
        self.state_connectivity = np.concatenate([self.state_connectivity, np.zeros((m, R), np.int64)])
        self.sink_connectivity = np.concatenate([self.sink_connectivity, np.zeros((m, R), np.int64)])
        if m == 0:
            return
        self._index.insert(newkeys, newidx)
        for ir in range(R):
            s = self.stoich[:, ir][None, :]
            pk = self._pack(newstates - s)
            ridx = self._index.lookup(np.where(pk >= 0, pk, 0))
            ridx = np.where(pk >= 0, ridx, 0)
            hit = ridx != 0
            self.state_connectivity[old:, ir][hit] = ridx[hit]
            self.sink_connectivity[ridx[hit] - 1, ir] = 0
        for ir in range(R):
            s = self.stoich[:, ir][None, :]
            sk = self._pack(newstates + s)
            valid = sk >= 0
            ridx = np.where(valid, self._index.lookup(np.where(valid, sk, 0)), 0)
            sink = valid & (ridx == 0)
            self.sink_connectivity[old:, ir][sink] = ir + 1
            hit = valid & (ridx != 0)
            self.state_connectivity[ridx[hit] - 1, ir] = newidx[hit]

    def deleteat(self, ids):
        ids = np.unique(np.asarray(ids, dtype=np.int64))
        nold = self.get_state_count()
        keep = np.ones(nold, dtype=bool)
        keep[ids - 1] = False
        newidxs = np.where(keep, np.cumsum(keep), 0).astype(np.int64)
        self.states = self.states[keep]
        sc = self.state_connectivity[keep]
        self.sink_connectivity = self.sink_connectivity[keep]
        n = self.states.shape[0]
        if n == 0:
            self.state_connectivity = sc
            self._index.rebuild(np.empty(0, np.int64), np.empty(0, np.int64))
            return
        self.state_connectivity = np.where(sc != 0, newidxs[np.maximum(sc, 1) - 1], 0)
        self._index.rebuild(self._pack(self.states), np.arange(1, n + 1, dtype=np.int64))
        for ir in range(self.nr):
            s = self.stoich[:, ir][None, :]
            sk = self._pack(self.states + s)
            valid = sk >= 0
            ridx = np.where(valid, self._index.lookup(np.where(valid, sk, 0)), 0)
            newsink = (self.sink_connectivity[:, ir] == 0) & valid & (ridx == 0)
            self.sink_connectivity[newsink, ir] = ir + 1
