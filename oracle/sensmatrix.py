"""Oracle restatement of ``ForwardSensFspMatrixSparse`` (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/forwardsensfspmatrix/forwardsensfspmatrixsparse/sensfspmatrixsparse.jl:
  constructor  :31-95   (one dA matrix per (reaction, parameter) of the sparsity pattern)
  matvec!      :97-142  (block vector [p; s_1; ...; s_P], each block of length N = n + R)

Gradients are duck-typed per reaction, mirroring src/cmemodel/propensitygrad.jl:20-47:
  ti   : ``pardiffs[ip](x,p)``
  sep  : ``tfactor_pardiffs[ip](t,p)``, ``statefactor_pardiffs[ip](x,p)``
  joint: ``pardiffs[ip](t,x,p)``
``pattern`` is a dense bool array (reactions x parameters), the restatement of the
Bool CSC of src/cmemodel/senstools/sparsity_pattern.jl:19-30.

Divergence: for joint-TV reactions the reference refreshes the
dA matrix with the *propensity* instead of its derivative (sensfspmatrixsparse.jl:137
-> fspsparsematrix.jl:160, ``pdiff.f`` is the captured propensity).  No reference
test covers it => parity unpinned there.  This oracle computes the mathematically
correct derivative; ``reproduce_q6=True`` reproduces the reference's behaviour.
"""
from __future__ import annotations

import numpy as np

from .fspmatrix import FspMatrixOracle, _csc, eval_over_states, generate_sparsematrix_entries


class OGrad:
    def __init__(self, kind, pardiffs=None, tfactor_pardiffs=None, statefactor_pardiffs=None):
        self.kind = kind
        self.pardiffs = pardiffs
        self.tfactor_pardiffs = tfactor_pardiffs
        self.statefactor_pardiffs = statefactor_pardiffs


class SensFspMatrixOracle:
    def __init__(self, space, propensities, gradients, pattern, parameters, reproduce_q6=False):
        self.fspmatrix = A = FspMatrixOracle(space, propensities, parameters)
        self.gradients = list(gradients)
        self.pattern = np.asarray(pattern, dtype=bool)
        self.parameter_count = P = len(parameters)
        self.reproduce_q6 = reproduce_q6
        N = A.rowcount
        st, sc, kc = A.states, A._sc, A._kc
        self.timeinvariant_matdiffs = []
        for ip in range(P):
            R_, C_, V_ = [np.empty(0, np.int64)], [np.empty(0, np.int64)], [np.empty(0)]
            for r in A.ti_ids:
                if self.pattern[r - 1, ip]:
                    d = eval_over_states(self.gradients[r - 1].pardiffs[ip], st, parameters)
                    a, b, c = generate_sparsematrix_entries(st, sc, kc, d, r)
                    R_.append(a), C_.append(b), V_.append(c)
            self.timeinvariant_matdiffs.append(_csc(np.concatenate(R_), np.concatenate(C_), np.concatenate(V_), N))
        self.sep_entries = []      # (ip, r, j, dMatrix)
        for ip in range(P):
            for j, r in enumerate(A.sep_ids):
                if self.pattern[r - 1, ip]:
                    d = eval_over_states(self.gradients[r - 1].statefactor_pardiffs[ip], st, parameters)
                    self.sep_entries.append((ip, r, j, _csc(*generate_sparsematrix_entries(st, sc, kc, d, r), N)))
        self.joint_entries = []    # (ip, r, dMatrix)
        for ip in range(P):
            for r in A.joint_ids:
                if self.pattern[r - 1, ip]:
                    d = np.zeros(A.n)
                    self.joint_entries.append((ip, r, _csc(*generate_sparsematrix_entries(st, sc, kc, d, r), N)))

    def matvec_(self, out, t, vs):
        A = self.fspmatrix
        th = A.parameters
        n = A.rowcount
        P = self.parameter_count
        p = vs[:n]
        A.matvec_(out[:n], t, p)
        for ip in range(P):
            o = out[(ip + 1) * n:(ip + 2) * n]
            A.matvec_(o, t, vs[(ip + 1) * n:(ip + 2) * n])
            o += self.timeinvariant_matdiffs[ip] @ p
            for (jp, r, j, dM) in self.sep_entries:
                if jp != ip:
                    continue
                o += A.propensities[r - 1].tfactor(t, th) * (dM @ p)
                o += self.gradients[r - 1].tfactor_pardiffs[ip](t, th) * (A.separabletv_factormatrices[j] @ p)
            for (jp, r, dM) in self.joint_entries:
                if jp != ip:
                    continue
                fn = A.propensities[r - 1].f if self.reproduce_q6 else self.gradients[r - 1].pardiffs[ip]
                FspMatrixOracle.update_sparsematrix(dM, A.states, fn, t, th)
                o += dM @ p

    def matvec(self, t, vs):
        out = np.empty_like(vs)
        self.matvec_(out, t, vs)
        return out
