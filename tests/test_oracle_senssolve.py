"""Pins oracle/senssolve.py (restatement of forwardsenscmesparse.jl:99-215 + fsspaceadapterssparse.jl:21-60).

The reference's own test (test/test_sensfsp.jl) only checks that the solve runs; the values are pinned here by the
analytic sensitivities of the birth-death process: p(x,t) = Poisson(x; mu), mu = lam/gam (1 - exp(-gam t)),
d p/d theta = (Poisson(x-1; mu) - Poisson(x; mu)) d mu/d theta."""
import math

import numpy as np
from scipy.stats import poisson

from oracle.fspmatrix import OProp
from oracle.sensmatrix import OGrad
from oracle.senssolve import ForwardSensRStepAdapterOracle, solve_sens


def birth_death():
    S = np.array([[1], [-1]]).T
    props = [OProp("ti", f=lambda x, p: p[0] + 0.0 * x[0]), OProp("ti", f=lambda x, p: p[1] * x[0])]
    zero = lambda x, p: 0.0 * x[0]
    grads = [OGrad("ti", pardiffs=[lambda x, p: 1.0 + 0.0 * x[0], zero]), OGrad("ti", pardiffs=[zero, lambda x, p: 1.0 * x[0]])]
    return S, props, grads, np.eye(2, dtype=bool)


def analytic(states, t, lam, gam):
    x = states[:, 0]
    mu = lam / gam * (1 - math.exp(-gam * t))
    dmu = [(1 - math.exp(-gam * t)) / gam, lam * (t * math.exp(-gam * t) / gam - (1 - math.exp(-gam * t)) / gam ** 2)]
    dp = poisson.pmf(x - 1, mu) - poisson.pmf(x, mu)
    return poisson.pmf(x, mu), [dp * d for d in dmu]


def test_sens_solve_oracle_birth_death():
    S, props, grads, pattern = birth_death()
    lam, gam = 10.0, 0.5
    touts = [1.0, 4.0]
    out = solve_sens(S, props, grads, pattern, [lam, gam], [[0]], [1.0], [[0.0], [0.0]], (0.0, 4.0),
                     ForwardSensRStepAdapterOracle(10, 10, True), saveat=touts, fsptol=1e-8, odeatol=1e-13, odertol=1e-9)
    assert out["adapts"] >= 1
    assert len(out["t"]) == len(touts) + 1            # + the final slice, as the reference (Q3)
    for k, t in enumerate(touts):
        p, dps = analytic(out["states"][k], t, lam, gam)
        assert np.abs(out["p"][k] - p).max() < 5e-7
        for ip in range(2):
            assert np.abs(out["S"][k][ip] - dps[ip]).max() < 5e-6 * max(1.0, np.abs(dps[ip]).max())
            # d/dtheta of the total mass (states + sinks) is zero
            assert abs(out["S"][k][ip].sum() + out["dsinks"][k][ip].sum()) < 1e-9
        assert out["p"][k].sum() + out["sinks"][k].sum() == np.float64(1.0) or abs(out["p"][k].sum() + out["sinks"][k].sum() - 1) < 1e-9


def test_sens_adapter_oracle_prune_rule():
    """fsspaceadapterssparse.jl:45-52: `>=` rule, p and every sensitivity vector lose the same entries."""
    from oracle.statespace import StateSpaceOracleFast
    S, *_ = birth_death()
    sp = StateSpaceOracleFast(S, [[i] for i in range(6)])
    p = np.array([0.5, 0.3, 0.2 - 2e-8, 1e-9, 1e-8, 0.0])
    Sv = [np.arange(6.0), -np.arange(6.0)]
    ad = ForwardSensRStepAdapterOracle(1, 2, True)
    p2, S2 = ad.adapt(sp, p.copy(), [s.copy() for s in Sv], np.zeros(2), [np.zeros(2)] * 2, 1.0, 1.0, 1.5e-8)
    # tail mass stays >= 1 - 1.5e-8 after removing x = 5 (0) and x = 3 (1e-9), not after x = 4 (1e-8)
    assert sp.states_array()[:4, 0].tolist() == [0, 1, 2, 4]
    assert p2[:4].tolist() == [0.5, 0.3, 0.2 - 2e-8, 1e-8] and S2[0][:4].tolist() == [0.0, 1.0, 2.0, 4.0]
    assert sp.get_state_count() == p2.size == S2[1].size
