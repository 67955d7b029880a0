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
    """Sorted int64 keys -> 1-based state index.  Two sorted runs (a large one and a small recent one that is
    merged into it when it reaches 1/16 of its size), so that an insert batch costs O(batch) amortised."""

    def __init__(self):
        self.keys = np.empty(0, dtype=np.int64)
        self.idx = np.empty(0, dtype=np.int64)
        self.rkeys = np.empty(0, dtype=np.int64)
        self.ridx = np.empty(0, dtype=np.int64)

    @staticmethod
    def _find(keys, idx, q):
        if keys.size == 0:
            return np.zeros(q.shape, dtype=np.int64)
        pos_c = np.minimum(np.searchsorted(keys, q), keys.size - 1)
        return np.where(keys[pos_c] == q, idx[pos_c], 0)

    def lookup(self, q: np.ndarray) -> np.ndarray:
        """Return 1-based indices, 0 where absent."""
        out = self._find(self.keys, self.idx, q)
        if self.rkeys.size:
            out = np.maximum(out, self._find(self.rkeys, self.ridx, q))      # a key lives in one run only
        return out

    @staticmethod
    def _merge(ka, ia, kb, ib):
        pos = np.searchsorted(ka, kb)
        return np.insert(ka, pos, kb), np.insert(ia, pos, ib)

    def insert(self, k: np.ndarray, idx: np.ndarray):
        order = np.argsort(k, kind="stable")
        self.rkeys, self.ridx = self._merge(self.rkeys, self.ridx, k[order], idx[order])
        if self.rkeys.size > max(4096, self.keys.size // 16):
            self.keys, self.idx = self._merge(self.keys, self.idx, self.rkeys, self.ridx)
            self.rkeys, self.ridx = np.empty(0, dtype=np.int64), np.empty(0, dtype=np.int64)

    def rebuild(self, k: np.ndarray, idx: np.ndarray):
        order = np.argsort(k, kind="stable")
        self.keys, self.idx = k[order], idx[order]
        self.rkeys, self.ridx = np.empty(0, dtype=np.int64), np.empty(0, dtype=np.int64)


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
        self._n = 0                                             # growable buffers (capacity doubling): rows [0, _n) are live
        self._states = np.empty((0, self.ns), dtype=np.int64)
        self._sc = np.empty((0, self.nr), dtype=np.int64)
        self._kc = np.empty((0, self.nr), dtype=np.int64)
        self._index = _SortedKeyIndex()
        initstates = np.asarray(initstates, dtype=np.int64)
        if initstates.ndim == 1:
            initstates = initstates[None, :]
        self._addstates(initstates.reshape(-1, self.ns))

    states = property(lambda self: self._states[:self._n])
    state_connectivity = property(lambda self: self._sc[:self._n])
    sink_connectivity = property(lambda self: self._kc[:self._n])

    def _append(self, newstates: np.ndarray):
        m, n = newstates.shape[0], self._n
        if n + m > self._states.shape[0]:
            cap = max(1024, 2 * (n + m))
            for name in ("_states", "_sc", "_kc"):
                buf = getattr(self, name)
                grown = np.empty((cap, buf.shape[1]), dtype=np.int64)
                grown[:n] = buf[:n]
                setattr(self, name, grown)
        self._states[n:n + m] = newstates
        self._sc[n:n + m] = 0
        self._kc[n:n + m] = 0
        self._n = n + m

    def _pack(self, X: np.ndarray) -> np.ndarray:
        """Pack non-negative rows into int64 keys; rows with a negative entry get key -1."""
        neg = (X < 0).any(axis=1)
        if (X >= (np.int64(1) << self.bits)[None, :]).any():
            raise OverflowError("state component exceeds key width")
        k = (np.where(X < 0, 0, X) << self.shifts[None, :]).sum(axis=1)
        return np.where(neg, np.int64(-1), k)

    def get_state_count(self):
        return self._n

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
        self._append(newstates)
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
        sc = self.state_connectivity[keep]
        self._states, self._kc = self.states[keep], self.sink_connectivity[keep]
        n = self._n = self._states.shape[0]
        if n == 0:
            self._sc = sc
            self._index.rebuild(np.empty(0, np.int64), np.empty(0, np.int64))
            return
        self._sc = np.where(sc != 0, newidxs[np.maximum(sc, 1) - 1], 0)
        self._index.rebuild(self._pack(self.states), np.arange(1, n + 1, dtype=np.int64))
        for ir in range(self.nr):
            s = self.stoich[:, ir][None, :]
            sk = self._pack(self.states + s)
            valid = sk >= 0
            ridx = np.where(valid, self._index.lookup(np.where(valid, sk, 0)), 0)
            newsink = (self.sink_connectivity[:, ir] == 0) & valid & (ridx == 0)
            self.sink_connectivity[newsink, ir] = ir + 1
