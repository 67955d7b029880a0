"""Pins the matrix / sensitivity-matrix oracle on the reference's tests: test/test_fspmat.jl:41-68,
test/sensmat/telegraph.jl:33-217 (analytic A(t) and dA/dtheta), test/sensmat/poisson.jl:6-21,
SURVEY.md Appendix A; and the timed C restatement against it."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from fixtures import (FSPMAT_THETA, RNACOUNT_MAX, SENS_THETA, TELEGRAPH_S, fspmat_propensities, sens_telegraph,
                      sens_tfactor)
from oracle.fspmatrix import FspMatrixOracle, OProp
from oracle.sensmatrix import OGrad, SensFspMatrixOracle
from oracle.statespace import StateSpaceOracle, StateSpaceOracleFast

EPS = np.finfo(float).eps


def _space2():
    s = StateSpaceOracle(TELEGRAPH_S, [1, 0, 0])
    s.expand(2)
    return s


def test_appendix_a_dense_matrix():
    A = FspMatrixOracle(_space2(), fspmat_propensities("ti"), FSPMAT_THETA)
    want = np.zeros((9, 9))
    want[:5, :5] = [[-0.05, 0.1, 1, 0, 0], [0.05, -5.1, 0, 1, 0], [0, 0, -1.05, 0.1, 2], [0, 5, 0.05, -6.1, 0], [0, 0, 0, 0, -2.05]]
    want[5, 4] = 0.05
    want[7, 3] = 5.0
    assert np.allclose(A.timeinvariant_matrix.toarray(), want, atol=0, rtol=1e-15)
    w = A.matvec(0.0, np.ones(9))
    assert np.allclose(w, [1.05, -4.05, 1.05, -1.05, -2.05, 0.05, 0, 5, 0], rtol=1e-14, atol=1e-15)


def test_fspmat_jl():  # test/test_fspmat.jl:41-68
    space = _space2()
    A = FspMatrixOracle(space, fspmat_propensities("ti"), FSPMAT_THETA)
    assert A.size(1) == space.get_state_count() + space.get_sink_count() == A.size(2)
    v = np.ones(A.size(1))
    for t in (1.0, 0.0):
        assert abs(A.matvec(t, v).sum()) <= 1e-14
    A1 = FspMatrixOracle(space, fspmat_propensities("tv"), FSPMAT_THETA)
    A2 = FspMatrixOracle(space, fspmat_propensities("tvj"), FSPMAT_THETA)
    for t in (1.0, 0.0):
        w1, w2 = A1.matvec(t, v), A2.matvec(t, v)
        assert abs(w1.sum()) <= 1e-14 and abs(w2.sum()) <= 1e-14
        assert np.linalg.norm(w1 - w2) == pytest.approx(0.0, abs=1e-15)
    with pytest.raises(ValueError):
        A.size(3)


def test_stored_entry_counts():
    """Julia's sparse() keeps stored zeros and sums duplicates: 3006 / 2002."""
    props, grads, pattern, states = sens_telegraph()
    A = FspMatrixOracle(StateSpaceOracleFast(TELEGRAPH_S, states), props, SENS_THETA)
    assert A.stored_entries() == [3006, 2002]


def test_matvecadd():
    space = _space2()
    A = FspMatrixOracle(space, fspmat_propensities("tv"), FSPMAT_THETA)
    rng = np.random.default_rng(0)
    v, o = rng.random(9), rng.random(9)
    out = o.copy()
    A.matvecadd_(out, 0.7, v)
    assert np.allclose(out, o + A.matvec(0.7, v), rtol=1e-15)


# ---- analytic telegraph generator, restated from test/sensmat/telegraph.jl:46-188 ------------
def _analytic(nmax, t, p):
    n = 2 * (nmax + 1)
    N = n + 4
    c4 = sens_tfactor(t, p)
    R, Cc, V = [], [], []
    dR = [([], [], []) for _ in range(5)]

    def add(lst, i, j, v):
        lst[0].append(i), lst[1].append(j), lst[2].append(v)
    for gene in (0, 1):
        for k in range(nmax + 1):
            x = (1 - gene, gene, k)
            j = gene * (nmax + 1) + k
            a1, a2, a3, a4 = p[0] * x[0], p[1] * x[1], p[2] * x[1], c4 * p[3] * x[2]
            add((R, Cc, V), j, j, -(a1 + a2 + a3 + a4))
            if gene == 0:
                add((R, Cc, V), nmax + 1 + k, j, a1)
                add(dR[0], j, j, -1.0), add(dR[0], nmax + 1 + k, j, 1.0)
            else:
                add((R, Cc, V), k, j, a2)
                add(dR[1], j, j, -1.0), add(dR[1], k, j, 1.0)
                tgt = nmax + 1 + k + 1 if k < nmax else n + 2      # sink of reaction 3 (0-based row n+2)
                add((R, Cc, V), tgt, j, a3)
                add(dR[2], j, j, -1.0), add(dR[2], tgt, j, 1.0)
            if k > 0:
                add((R, Cc, V), j - 1, j, a4)
                add(dR[3], j, j, -c4 * k), add(dR[3], j - 1, j, c4 * k)
                dl = p[3] * k * math.cos(math.pi * t / p[4]) * math.pi * t / p[4] ** 2 if 1.0 - math.sin(math.pi * t / p[4]) >= 0 else 0.0
                add(dR[4], j, j, -dl), add(dR[4], j - 1, j, dl)
    mk = lambda l: sp.coo_matrix((l[2], (l[0], l[1])), shape=(N, N)).tocsr()
    return mk((R, Cc, V)), [mk(d) for d in dR]


@pytest.mark.parametrize("t", [10.0, 20.0, 30.0, 100.0])
def test_sens_against_analytic(t):  # test/sensmat/telegraph.jl:190-217, tolerance n*eps
    props, grads, pattern, states = sens_telegraph()
    SA = SensFspMatrixOracle(StateSpaceOracleFast(TELEGRAPH_S, states), props, grads, pattern, SENS_THETA)
    n = SA.fspmatrix.rowcount
    P = 5
    v = np.ones(n * (P + 1))
    v /= v.sum()
    out = SA.matvec(t, v)
    A, dAs = _analytic(RNACOUNT_MAX, t, SENS_THETA)
    ref = np.empty_like(v)
    ref[:n] = A @ v[:n]
    for ip in range(P):
        ref[(ip + 1) * n:(ip + 2) * n] = dAs[ip] @ v[:n] + A @ v[(ip + 1) * n:(ip + 2) * n]
    assert np.abs(out - ref).max() <= n * EPS


def test_sens_zero_sum():  # test/sensmat/telegraph.jl:34-43
    props, grads, pattern, states = sens_telegraph()
    SA = SensFspMatrixOracle(StateSpaceOracleFast(TELEGRAPH_S, states), props, grads, pattern, SENS_THETA)
    n = SA.fspmatrix.rowcount
    v = np.ones(6 * n)
    out = SA.matvec(20.0, v)
    assert abs(out.sum() / v.sum()) <= n * 6 * EPS


def test_sens_poisson_zero_sum():  # test/sensmat/poisson.jl:6-21
    S = np.array([[1], [-1]]).T
    props = [OProp("ti", f=lambda x, p: p[0] + 0.0 * x[0]), OProp("ti", f=lambda x, p: p[1] + 0.0 * x[0])]
    one = lambda x, p: 1.0 + 0.0 * x[0]
    zero = lambda x, p: 0.0 * x[0]
    grads = [OGrad("ti", pardiffs=[one, zero]), OGrad("ti", pardiffs=[zero, one])]
    space = StateSpaceOracleFast(S, [[i] for i in range(1, 101)])
    SA = SensFspMatrixOracle(space, props, grads, np.eye(2, dtype=bool), [10.0, 5.0])
    out = SA.matvec(0.0, np.ones(3 * SA.fspmatrix.rowcount))
    assert out.sum() == pytest.approx(0.0, abs=1e-9)


def test_sens_joint_q6_divergence():
    """Joint-TV sensitivity: correct derivative by default, reference's behaviour behind a flag."""
    f = lambda t, x, p: (1.0 + 0.5 * math.sin(t)) * p[1] * x[1]
    df = lambda t, x, p: (1.0 + 0.5 * math.sin(t)) * x[1]
    zero3 = lambda t, x, p: 0.0 * x[0]
    props = fspmat_propensities("ti")
    props[1] = OProp("joint", f=f)
    zero = lambda x, p: 0.0 * x[0]
    grads = [OGrad("ti", pardiffs=[lambda x, p: 1.0 * x[0], zero, zero, zero]),
             OGrad("joint", pardiffs=[zero3, df, zero3, zero3]),
             OGrad("ti", pardiffs=[zero, zero, lambda x, p: 1.0 * x[1], zero]),
             OGrad("ti", pardiffs=[zero, zero, zero, lambda x, p: 1.0 * x[2]])]
    space = _space2()
    SA = SensFspMatrixOracle(space, props, grads, np.eye(4, dtype=bool), FSPMAT_THETA)
    n = SA.fspmatrix.rowcount
    rng = np.random.default_rng(3)
    v = rng.random(5 * n)
    out = SA.matvec(0.3, v)
    h = 1e-6
    th_p, th_m = list(FSPMAT_THETA), list(FSPMAT_THETA)
    th_p[1] += h
    th_m[1] -= h
    Ap = FspMatrixOracle(space, props, th_p).matvec(0.3, v[:n])
    Am = FspMatrixOracle(space, props, th_m).matvec(0.3, v[:n])
    fd = (Ap - Am) / (2 * h) + SA.fspmatrix.matvec(0.3, v[2 * n:3 * n])
    assert np.allclose(out[2 * n:3 * n], fd, rtol=1e-7, atol=1e-9)
    SB = SensFspMatrixOracle(space, props, grads, np.eye(4, dtype=bool), FSPMAT_THETA, reproduce_q6=True)
    assert not np.allclose(SB.matvec(0.3, v)[2 * n:3 * n], fd, rtol=1e-3)


def test_c_baseline_matches_scipy():
    from oracle import cbaseline
    space = StateSpaceOracleFast(TELEGRAPH_S, [1, 0, 0])
    space.expand(40)
    A = FspMatrixOracle(space, fspmat_propensities("tv"), FSPMAT_THETA)
    rng = np.random.default_rng(0)
    v = rng.random(A.rowcount)
    terms = cbaseline.CscTerms(A.terms_at(0.7))
    out = np.empty_like(v)
    terms.matvec(v, out)
    assert np.allclose(out, A.matvec(0.7, v), rtol=1e-13, atol=1e-15)
    fused = sum(c * m for c, m in A.terms_at(0.7)).tocsr()
    out2 = np.empty_like(v)
    cbaseline.CsrOmp(fused).matvec(v, out2)
    assert np.allclose(out2, out, rtol=1e-12, atol=1e-14)
    assert cbaseline.num_threads() >= 1
