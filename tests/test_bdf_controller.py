"""The step-size / order controller shared by the host loop and the multi-step kernel (numcme.jl_b200/csrc/bdf_ctl.h)
is plain host/device C++: here it is compiled with g++ and checked against a numpy restatement of the NDF rules
(Shampine & Reichelt; the formulation of scipy.integrate.BDF, whose `compute_R` / `change_D` these functions follow)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctl(tmp_path_factory):
    out = tmp_path_factory.mktemp("ctl") / "libctl.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "numcme.jl_b200", "csrc"),
                           os.path.join(ROOT, "tests", "harness", "ctl_harness.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.h_sink_sum_at.restype = C.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def np_compute_R(order, factor):
    """scipy.integrate._ivp.bdf.compute_R"""
    I = np.arange(1, order + 1)[:, None]
    J = np.arange(1, order + 1)
    M = np.zeros((order + 1, order + 1))
    M[1:, 1:] = (I - 1 - factor * J) / I
    M[0] = 1
    return np.cumprod(M, axis=0)


def np_change_D(D, order, factor):
    """scipy.integrate._ivp.bdf.change_D: D[:order+1] <- (R U)^T D[:order+1]"""
    R = np_compute_R(order, factor)
    U = np_compute_R(order, 1)
    RU = R.dot(U)
    D = D.copy()
    D[:order + 1] = RU.T.dot(D[:order + 1])
    return D


def test_constants(ctl):
    g, a, e = np.zeros(6), np.zeros(6), np.zeros(7)
    ctl.h_constants(_p(g), _p(a), _p(e))
    kappa = np.array([0, -0.1850, -1 / 9, -0.0823, -0.0415, 0])
    gamma = np.hstack((0, np.cumsum(1 / np.arange(1, 6))))
    assert np.allclose(g, gamma, rtol=0, atol=1e-15)
    assert np.allclose(a, (1 - kappa) * gamma, rtol=0, atol=1e-15)
    assert np.allclose(e[:6], kappa * gamma + 1 / np.arange(1, 7), rtol=0, atol=1e-15)


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_compute_R_and_composed_rescaling(ctl, order):
    rng = np.random.default_rng(order)
    for factor in (0.2, 0.5, 0.93, 1.7, 10.0):
        R = np.zeros(36)
        ctl.h_compute_R(order, C.c_double(factor), _p(R))
        assert np.allclose(R.reshape(6, 6)[:order + 1, :order + 1], np_compute_R(order, factor), rtol=1e-14, atol=1e-15)
    # a chain of pending rescalings composed on the controller == applying them one after the other (what bdf.cu does)
    factors = np.array([0.7, 1.3, 0.5])
    D = rng.standard_normal((8, 7))
    want = D.copy()
    for f in factors:
        want = np_change_D(want, order, f)
    P = np.zeros(36)
    ctl.h_queue_changes(order, len(factors), _p(factors), _p(P))
    P = P.reshape(6, 6)
    got = D.copy()
    got[:order + 1] = P[:order + 1, :order + 1].T.dot(D[:order + 1])     # D_r <- sum_j P[j][r] D_j, as in k_bdf_step
    assert np.allclose(got, want, rtol=1e-12, atol=1e-13)
    assert np.allclose(P[0, :order + 1], np.eye(order + 1)[0]) and np.allclose(P[1:, 0], 0)   # D_0 (the solution) is invariant
    assert np.all(P[order + 1:] == 0) and np.all(P[:, order + 1:] == 0)


def test_controller_rounds(ctl):
    out = np.zeros(9)
    g, a, e = np.zeros(6), np.zeros(6), np.zeros(7)
    ctl.h_constants(_p(g), _p(a), _p(e))
    d = C.c_double
    # plain step: t_new = t + h, c = h / alpha_k
    ctl.h_round(d(1.0), d(10.0), d(0.25), 3, 1, 0, d(0.3), d(1.0), d(1.0), d(100.0), _p(out))
    assert out[0] == 1.25 and out[1] == 0.25 and out[2] == pytest.approx(0.25 / a[3]) and out[3] == pytest.approx(1 / a[3])
    assert out[4] == pytest.approx(e[3]) and out[5] == 0.25 and out[6] == 3 and out[7] == 2 and out[8] == 0   # too few equal steps
    # landing on t1 queues a rescaling with factor (t1 - t) / h and resets the equal-step counter
    ctl.h_round(d(9.9), d(10.0), d(0.25), 2, 5, 1, d(4.0), d(0), d(0), d(100.0), _p(out))
    assert out[0] == 10.0 and out[1] == pytest.approx(0.1) and out[8] == 1
    # rejected: h *= max(0.2, 0.9 err^(-1/(k+1)))
    assert out[5] == pytest.approx(0.1 * max(0.2, 0.9 * 4.0 ** (-1 / 3))) and out[6] == 2 and out[7] == 0
    ctl.h_round(d(0.0), d(10.0), d(0.5), 4, 0, 1, d(1e300), d(0), d(0), d(100.0), _p(out))
    assert out[5] == pytest.approx(0.5 * 0.2)
    # linear solver failure: halve
    ctl.h_round(d(0.0), d(10.0), d(0.5), 4, 3, 2, d(0), d(0), d(0), d(100.0), _p(out))
    assert out[5] == 0.25 and out[7] == 0 and out[8] == 1
    # order selection after order+1 equal steps: the largest of err_m^(-1/k), err^(-1/(k+1)), err_p^(-1/(k+2)) wins
    N = 400.0
    order, err = 2, 0.05
    for sm, sp in ((1e-8, 1e2), (1e2, 1e-10), (1e2, 1e2)):
        ctl.h_round(d(0.0), d(10.0), d(0.1), order, order, 0, d(err), d(sm), d(sp), d(N), _p(out))
        em = e[order - 1] * np.sqrt(sm / N)
        ep = e[order + 1] * np.sqrt(sp / N)
        fac = [em ** (-1 / order), err ** (-1 / (order + 1)), ep ** (-1 / (order + 2))]
        best = int(np.argmax(fac))
        assert out[6] == order + best - 1
        assert out[5] == pytest.approx(0.1 * min(10.0, 0.9 * fac[best]))
        assert out[7] == 0 and out[8] == 1
    # order 1 cannot go down, order 5 cannot go up
    ctl.h_round(d(0.0), d(10.0), d(0.1), 1, 1, 0, d(0.5), d(1e-30), d(1e10), d(N), _p(out))
    assert out[6] == 1
    ctl.h_round(d(0.0), d(10.0), d(0.1), 5, 5, 0, d(0.5), d(1e10), d(1e-30), d(N), _p(out))
    assert out[6] == 5


def test_sink_dense_output(ctl):
    """Newton backward-difference interpolation of the sink entries: exact for polynomials of degree <= order."""
    rng = np.random.default_rng(0)
    R, order, h, t_new = 3, 4, 0.37, 5.0
    coef = rng.standard_normal((R, order + 1))
    f = lambda t: np.array([np.polyval(c, t) for c in coef])
    ys = np.array([f(t_new - j * h) for j in range(order + 1)])          # y_n, y_{n-1}, ...
    nd = np.zeros((order + 1, R))
    import math
    for j in range(order + 1):                                          # backward differences D_j = nabla^j y_n
        nd[j] = sum((-1) ** i * math.comb(j, i) * ys[i] for i in range(j + 1))
    nd = np.ascontiguousarray(nd)
    for tt in (t_new, t_new - 0.2 * h, t_new - h, t_new - 0.77 * h):
        got = ctl.h_sink_sum_at(R, order, _p(nd), C.c_double(t_new), C.c_double(h), C.c_double(tt))
        assert got == pytest.approx(f(tt).sum(), rel=1e-11, abs=1e-11)
    assert ctl.h_sizeof_ctl() % 8 == 0
