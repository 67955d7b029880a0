"""Oracle restatement of ``FspMatrixSparse`` (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/fspmatrix/sparse/fspsparsematrix.jl:
  constructor (term classification, per-term CSC)   :47-108
  _generate_sparsematrix_entries (COO generator)    :120-152
  _update_sparsematrix! (joint-TV value refresh)    :154-166
  matvec! / matvecadd! / matvec / *                 :196-264
and the propensity calling conventions of src/cmemodel/propensity.jl:129-153.

``sparse(I,J,V,m,n)`` (Julia stdlib SparseArrays, not under /root/reference) sums
duplicate entries and keeps stored zeros; ``scipy.sparse.coo_matrix(...).tocsc()``
has the same two properties, which tests/test_oracle_fspmatrix.py pins through the
stored-entry counts quoted in SURVEY.md section 3A (3006 / 2002).

Propensities are duck-typed: any object with ``kind`` in {"ti","sep","joint"} and
the callables ``f`` (ti: f(x,p); joint: f(t,x,p)) or ``tfactor(t,p)`` +
``statefactor(x,p)`` (sep).  ``x`` is indexable by species (0-based here).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class OProp:
    """Minimal propensity record for oracle-only tests."""

    def __init__(self, kind, f=None, tfactor=None, statefactor=None):
        assert kind in ("ti", "sep", "joint")
        self.kind, self.f, self.tfactor, self.statefactor = kind, f, tfactor, statefactor


def eval_over_states(fn, states: np.ndarray, *lead_args_and_p, t=None) -> np.ndarray:
    """Evaluate ``fn(x,p)`` (or ``fn(t,x,p)`` if t is given) on every row of ``states``."""
    (p,) = lead_args_and_p
    n = states.shape[0]
    cols = [states[:, k].astype(np.float64) for k in range(states.shape[1])]
    try:
        v = fn(cols, p) if t is None else fn(t, cols, p)
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 0:
            v = np.full(n, float(v))
        if v.shape == (n,):
            return v
    except Exception:
        pass
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        x = [int(c) for c in states[i]]
        out[i] = fn(x, p) if t is None else fn(t, x, p)
    return out


def generate_sparsematrix_entries(states, state_conn, sink_conn, dvals, reactionidx):
    """COO triples (1-based) of one reaction's matrix.  fspsparsematrix.jl:120-152.

    ``dvals[j]`` is the state factor at state j (zeros when the reference passes
    ``statefactor === nothing`` for joint-TV terms, :129).
    """
    n = states.shape[0]
    r = reactionidx - 1
    rows = np.zeros(2 * n, dtype=np.int64)
    cols = np.zeros(2 * n, dtype=np.int64)
    vals = np.zeros(2 * n, dtype=np.float64)
    idx = np.arange(1, n + 1, dtype=np.int64)
    rows[:n] = idx
    cols[:n] = idx
    vals[:n] = -1.0 * dvals
    sink = sink_conn[:, r]
    hs = sink != 0
    cols[n:][hs] = idx[hs]
    rows[n:][hs] = n + sink[hs]
    vals[n:][hs] = dvals[hs]
    cidx = state_conn[:, r]
    hc = cidx != 0
    pos = cidx[hc] + n - 1
    rows[pos] = idx[hc]
    cols[pos] = cidx[hc]
    vals[pos] = -1.0 * vals[cidx[hc] - 1]
    keep = cols != 0
    return rows[keep], cols[keep], vals[keep]


def _csc(rows, cols, vals, N):
    return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(N, N)).tocsc()


class FspMatrixOracle:
    def __init__(self, space, propensities, parameters=()):
        self.parameters = parameters
        self.states = np.array(space.states_array(), copy=True)
        sc = space.state_connectivity_array()
        kc = space.sink_connectivity_array()
        n = self.states.shape[0]
        self.n = n
        self.rowcount = self.colcount = N = n + space.get_sink_count()
        self.propensities = list(propensities)
        self.ti_ids = [i + 1 for i, a in enumerate(propensities) if a.kind == "ti"]
        self.sep_ids = [i + 1 for i, a in enumerate(propensities) if a.kind == "sep"]
        self.joint_ids = [i + 1 for i, a in enumerate(propensities) if a.kind == "joint"]
        self._sc, self._kc = sc, kc

        R_, C_, V_ = [], [], []
        for r in self.ti_ids:
            d = eval_over_states(propensities[r - 1].f, self.states, parameters)
            a, b, c = generate_sparsematrix_entries(self.states, sc, kc, d, r)
            R_.append(a), C_.append(b), V_.append(c)
        self.timeinvariant_matrix = (
            _csc(np.concatenate(R_), np.concatenate(C_), np.concatenate(V_), N) if self.ti_ids else None
        )
        self.separabletv_factormatrices = []
        for r in self.sep_ids:
            d = eval_over_states(propensities[r - 1].statefactor, self.states, parameters)
            self.separabletv_factormatrices.append(_csc(*generate_sparsematrix_entries(self.states, sc, kc, d, r), N))
        self.jointtv_matrices = []
        for r in self.joint_ids:
            d = np.zeros(n)
            self.jointtv_matrices.append(_csc(*generate_sparsematrix_entries(self.states, sc, kc, d, r), N))
        self.t_cache = -np.inf

    def size(self, dim=None):
        if dim is None:
            return (self.rowcount, self.colcount)
        if dim not in (1, 2):
            raise ValueError("Second argument must be either 1 or 2.")
        return self.rowcount if dim == 1 else self.colcount

    @staticmethod
    def update_sparsematrix(M, states, fn, t, p):
        """fspsparsematrix.jl:154-166 -- diagonal <- -f, the other stored entry of the column <- +f."""
        n = states.shape[0]
        val = eval_over_states(fn, states, p, t=t)
        indptr, indices, data = M.indptr, M.indices, M.data
        colof = np.repeat(np.arange(M.shape[1]), np.diff(indptr))
        incol = colof < n
        v = np.where(incol, val[np.minimum(colof, n - 1)], 0.0)
        data[incol] = np.where(indices[incol] == colof[incol], -1.0 * v[incol], v[incol])

    def _tv_terms(self, out, t, v):
        p = self.parameters
        for i, r in enumerate(self.sep_ids):
            out += self.propensities[r - 1].tfactor(t, p) * (self.separabletv_factormatrices[i] @ v)
        needupdate = t != self.t_cache
        if needupdate:
            self.t_cache = t
        for i, r in enumerate(self.joint_ids):
            if needupdate:
                self.update_sparsematrix(self.jointtv_matrices[i], self.states, self.propensities[r - 1].f, t, p)
            out += self.jointtv_matrices[i] @ v

    def matvec_(self, out, t, v):
        if self.timeinvariant_matrix is None:
            out[:] = 0.0
        else:
            out[:] = self.timeinvariant_matrix @ v
        self._tv_terms(out, t, v)

    def matvecadd_(self, out, t, v):
        if self.timeinvariant_matrix is not None:
            out += self.timeinvariant_matrix @ v
        self._tv_terms(out, t, v)

    def matvec(self, t, v):
        w = np.empty_like(v)
        self.matvec_(w, t, v)
        return w

    def __matmul__(self, v):
        return self.matvec(0.0, v)

    # helpers for the timed C baseline: list of (coef-kind, CSC) exactly as the reference holds them
    def terms_at(self, t):
        p = self.parameters
        out = []
        if self.timeinvariant_matrix is not None:
            out.append((1.0, self.timeinvariant_matrix))
        for i, r in enumerate(self.sep_ids):
            out.append((float(self.propensities[r - 1].tfactor(t, p)), self.separabletv_factormatrices[i]))
        for i, r in enumerate(self.joint_ids):
            self.update_sparsematrix(self.jointtv_matrices[i], self.states, self.propensities[r - 1].f, t, p)
            out.append((1.0, self.jointtv_matrices[i]))
        self.t_cache = t
        return out

    def stored_entries(self):
        mats = ([self.timeinvariant_matrix] if self.timeinvariant_matrix is not None else []) \
            + self.separabletv_factormatrices + self.jointtv_matrices
        return [m.nnz for m in mats]

    def algorithmic_bytes(self):
        """sum_terms (8 nnz + 4 nnz_offdiag) + 16 N."""
        b = 16 * self.rowcount
        for nnz in self.stored_entries():
            b += 8 * nnz + 4 * (nnz - self.n)
        return b
