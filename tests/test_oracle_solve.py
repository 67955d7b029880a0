"""Pins the oracle's solve loop: the reference's own solver tests (test/test_solver.jl:60-99, conservation and
saveat length) and -- because the reference pins no solution values -- the analytic birth-death solution."""
import math

import numpy as np
import pytest
from scipy.stats import poisson

from fixtures import FSPMAT_THETA, TELEGRAPH_S, fspmat_propensities
from oracle.fspmatrix import OProp
from oracle.solve import RStepAdapterOracle, SelectiveRStepAdapterOracle, solve_adaptive, solve_fixed
from oracle.statespace import StateSpaceOracleFast


def test_fixed_conservation_and_saveat():  # test/test_solver.jl:55-77
    sp = StateSpaceOracleFast(TELEGRAPH_S, [1, 0, 0])
    sp.expand(20)
    st = sp.states_array()
    p0 = np.zeros(st.shape[0])
    p0[0] = 1.0
    touts = np.arange(0.0, 121.0, 20.0)
    sol = solve_fixed(TELEGRAPH_S, fspmat_propensities("tv"), FSPMAT_THETA, st, p0, (0.0, 120.0), saveat=touts,
                      odeatol=1e-14, odertol=1e-4)
    assert len(sol["t"]) == len(touts)
    for p, s in zip(sol["p"], sol["sinks"]):
        assert p.sum() + s.sum() == pytest.approx(1.0, abs=1e-6)


@pytest.mark.parametrize("adapter", [RStepAdapterOracle(5, 10, True), SelectiveRStepAdapterOracle(10, 10, True)])
def test_adaptive_conservation(adapter):  # test/test_solver.jl:80-96
    touts = np.arange(0.0, 121.0, 20.0)
    sol = solve_adaptive(TELEGRAPH_S, fspmat_propensities("tv"), FSPMAT_THETA, [[1, 0, 0]], [1.0], (0.0, 120.0), adapter,
                         saveat=touts, odeatol=1e-14, odertol=1e-4)
    assert sol["adapts"] >= 1
    for p, s in zip(sol["p"], sol["sinks"]):
        assert p.sum() + s.sum() == pytest.approx(1.0, abs=1e-6)
    assert sol["sinks"][-1].sum() <= 1e-6 * 1.01


def test_birth_death_poisson():
    """0 -> X at rate lam, X -> 0 at rate gam*x from x=0: X(t) ~ Poisson(lam/gam (1 - exp(-gam t)))."""
    S = np.array([[1], [-1]]).T
    lam, gam = 10.0, 0.5
    props = [OProp("ti", f=lambda x, p: p[0] + 0.0 * x[0]), OProp("ti", f=lambda x, p: p[1] * x[0])]
    sol = solve_adaptive(S, props, [lam, gam], [[0]], [1.0], (0.0, 4.0), RStepAdapterOracle(10, 10, False),
                         saveat=[1.0, 4.0], fsptol=1e-8, odeatol=1e-12, odertol=1e-9, method="LSODA")
    for k, t in enumerate([1.0, 4.0]):
        mu = lam / gam * (1 - math.exp(-gam * t))
        st = sol["states"][k][:, 0]
        assert np.abs(sol["p"][k] - poisson.pmf(st, mu)).max() < 1e-7


def test_prune_rule():  # rstepadapters.jl:41-43 vs :93-95 (>= vs >)
    p = np.array([0.5, 1e-9, 0.3, 2e-9, 0.2 - 3e-9])
    a, b = RStepAdapterOracle(1, 1, True), SelectiveRStepAdapterOracle(1, 1, True)
    assert a.drop_ids(p, 1.0, 1.0, 1e-8).tolist() == [2, 4]      # tails: 1-1e-9, 1-3e-9 >= 1-1e-8
    assert a.drop_ids(p, 1.0, 1.0, 2e-9).tolist() == [2]
    thr_exact = 1.0 - (p.sum() - np.cumsum(np.sort(p))[0])
    assert len(a.drop_ids(p, 1.0, 1.0, thr_exact)) >= len(b.drop_ids(p, 1.0, 1.0, thr_exact))


def test_telegraph_moments_analytic():
    """Second independent pin of transient VALUES (the reference's tests pin none): the telegraph model
    (examples/telegraph_cme.jl) has closed-form means.  P(gene on) = k01/(k01+k10) (1 - exp(-(k01+k10) t)) from the off
    state, and the mean mRNA count solves m' = lam * P_on(t) - gam * m."""
    k01, k10, lam, gam = 0.05, 0.1, 5.0, 0.5
    props = [OProp("ti", f=lambda x, p: p[0] * x[0]), OProp("ti", f=lambda x, p: p[1] * x[1]),
             OProp("ti", f=lambda x, p: p[2] * x[1]), OProp("ti", f=lambda x, p: p[3] * x[2])]
    touts = [5.0, 20.0, 60.0]
    sol = solve_adaptive(TELEGRAPH_S, props, [k01, k10, lam, gam], [[1, 0, 0]], [1.0], (0.0, 60.0),
                         RStepAdapterOracle(5, 10, True), saveat=touts, fsptol=1e-8, odeatol=1e-13, odertol=1e-9, method="LSODA")
    s = k01 + k10
    for k, t in enumerate(touts):
        st, p = sol["states"][k], sol["p"][k]
        pon = k01 / s * (1 - math.exp(-s * t))
        # m(t) = lam k01/s [ (1 - e^{-gam t})/gam - (e^{-s t} - e^{-gam t})/(gam - s) ]
        m = lam * k01 / s * ((1 - math.exp(-gam * t)) / gam - (math.exp(-s * t) - math.exp(-gam * t)) / (gam - s))
        assert float((p * st[:, 1]).sum()) == pytest.approx(pon, abs=2e-7)
        assert float((p * st[:, 2]).sum()) == pytest.approx(m, abs=2e-6)
