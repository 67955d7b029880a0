"""Models behind BASELINE.json's configs; shared by tests and bench.py."""
from __future__ import annotations

import math

import numpy as np

from .cmemodel import CmeModel, propensity


def telegraph_model(theta=(0.05, 0.1, 5.0, 0.5)) -> CmeModel:
    """examples/telegraph_cme.jl:15-42 (species: G0, G1, mRNA)."""
    S = np.array([[-1, 1, 0], [1, -1, 0], [0, 0, 1], [0, 0, -1]]).T
    props = [
        propensity(lambda x, p: p[0] * x[0]),
        propensity(lambda x, p: p[1] * x[1]),
        propensity(lambda x, p: p[2] * x[1]),
        propensity(lambda x, p: p[3] * x[2]),
    ]
    return CmeModel(S, props, list(theta))


TOGGLE_THETA = [2.2e-3, 6.8e-5, 1.7e-2, 1.6e-2, 2.6e-3, 6.1e-3, 3, 2.1, 3.8e-4, 3.8e-4, 10.0, 3600]


def _uv_rate(t, p):
    return p[9] + (1.0 if t <= p[11] else 0.0) * 0.002 * p[10] ** 2 / (1260 + p[10] ** 3)


def toggle_model(separable=True, theta=TOGGLE_THETA) -> CmeModel:
    """examples/toggleswitch_fsp_variants.jl:8-59 (separable beta_4 or joint alpha_4 formulation)."""
    S = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]]).T
    a1 = propensity(lambda x, p: p[0] + p[2] / (1.0 + p[4] * x[1] ** p[6]))
    a2 = propensity(lambda x, p: p[8] * x[0])
    a3 = propensity(lambda x, p: p[1] + p[3] / (1.0 + p[5] * x[0] ** p[7]))
    if separable:
        a4 = propensity(lambda x, p: 1.0 * x[1], _uv_rate)
    else:
        a4 = propensity(lambda t, x, p: _uv_rate(t, p) * x[1])
    return CmeModel(S, [a1, a2, a3, a4], list(theta))


def m2d_model() -> CmeModel:
    """M-2D: examples/2dstate_exploration.jl:6 stoichiometry with mass-action birth-death rates."""
    S = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]]).T
    props = [
        propensity(lambda x, p: p[0] + 0.0 * x[0]),
        propensity(lambda x, p: p[1] * x[0]),
        propensity(lambda x, p: p[2] + 0.0 * x[0]),
        propensity(lambda x, p: p[3] * x[1]),
    ]
    return CmeModel(S, props, [10.0, 1.0, 8.0, 0.5])


def m3d_model(time_varying=False) -> CmeModel:
    """M-3D: synthetic three-species birth-death network, R = 6 (BASELINE.json config 5).
    TV variant: death-3 is separable with c(t) = 1 + 0.5 sin(2 pi t / 10)."""
    S = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]).T
    props = [
        propensity(lambda x, p: p[0] + 0.0 * x[0]),
        propensity(lambda x, p: p[1] * x[0]),
        propensity(lambda x, p: p[2] + 0.0 * x[0]),
        propensity(lambda x, p: p[3] * x[1]),
        propensity(lambda x, p: p[4] + 0.0 * x[0]),
    ]
    if time_varying:
        props.append(propensity(lambda x, p: p[5] * x[2], lambda t, p: 1.0 + 0.5 * math.sin(2.0 * math.pi * t / 10.0)))
    else:
        props.append(propensity(lambda x, p: p[5] * x[2]))
    return CmeModel(S, props, [10.0, 1.0, 8.0, 0.5, 6.0, 0.3])


M2D_LEVELS = 1413   # n = 1 000 405
M3D_LEVELS = 390    # n = 10 039 316


def simplex_count(d: int, L: int) -> int:
    return math.comb(L + d, d)


# ---- Hog1p (examples/hog1p.jl:6-82): NS = 6 (G0..G3, RNAnuc, RNAcyt), R = 13, P = 14 -----------------------------
HOG_R1, HOG_R2, HOG_ETA, HOG_A, HOG_M = 6.1e-3, 6.9e-3, 5.9, 9.3e9, 2.2e-2


def hog1p_signal(t):
    u = (1.0 - math.exp(-HOG_R1 * t)) * math.exp(-HOG_R2 * t)
    return HOG_A * (u / (1.0 + u / HOG_M)) ** HOG_ETA


# parameter order: k01 k10 a k12 k21 k23 k32 l0 l1 l2 l3 gnuc ktrans gcyt
HOG1P_THETA = [2.6e-3, 1.9e01, 0.0, 7.63e-3, 1.2e-2, 4e-3, 3.1e-3, 5.9e-4, 1.7e-1, 1.0, 3e-2, 2.2e-6, 2.6e-1, 8.3e-3]


def hog1p_model(theta=HOG1P_THETA, separable=True) -> CmeModel:
    """The 13-reaction Hog1p-driven gene model of examples/hog1p.jl.  Reaction 2 (G1 -> G0) has the time-varying rate
    max(0, k10 - a*Hog1p(t)): Catalyst import makes it a joint propensity in the reference (catalyst_interface.jl:21-27);
    it is rank-1, so the separable form is offered as well."""
    S = np.zeros((6, 13), dtype=np.int64)
    def rx(r, frm=None, to=None, plus=None, minus=None):
        if frm is not None:
            S[frm, r] -= 1
            S[to, r] += 1
        if plus is not None:
            S[plus, r] += 1
        if minus is not None:
            S[minus, r] -= 1
    rx(0, 0, 1); rx(1, 1, 0); rx(2, 1, 2); rx(3, 2, 1); rx(4, 2, 3); rx(5, 3, 2)
    rx(6, plus=4); rx(7, plus=4); rx(8, plus=4); rx(9, plus=4)
    rx(10, minus=4); rx(11, minus=4, plus=5); rx(12, minus=5)
    k10t = lambda t, p: max(0.0, p[1] - p[2] * hog1p_signal(t))
    props = [propensity(lambda x, p: p[0] * x[0])]
    if separable:
        props.append(propensity(lambda x, p: 1.0 * x[1], k10t))
    else:
        props.append(propensity(lambda t, x, p: k10t(t, p) * x[1]))
    props += [
        propensity(lambda x, p: p[3] * x[1]), propensity(lambda x, p: p[4] * x[2]),
        propensity(lambda x, p: p[5] * x[2]), propensity(lambda x, p: p[6] * x[3]),
        propensity(lambda x, p: p[7] * x[0]), propensity(lambda x, p: p[8] * x[1]),
        propensity(lambda x, p: p[9] * x[2]), propensity(lambda x, p: p[10] * x[3]),
        propensity(lambda x, p: p[11] * x[4]), propensity(lambda x, p: p[12] * x[4]),
        propensity(lambda x, p: p[13] * x[5]),
    ]
    return CmeModel(S, props, list(theta))


def _zero_x(x, p):
    return 0.0 * x[0]


def _zero_t(t, p):
    return 0.0


def hog1p_sens_model(theta=HOG1P_THETA):
    """BASELINE.json config 3: the Hog1p model of examples/hog1p.jl:33-82 as a ``CmeModelWithSensitivity`` (NS = 6,
    R = 13, P = 14).  Gradient sparsity as src/cmemodel/senstools/sparsity_pattern.jl:19-30 derives it from the rate
    laws: one parameter per reaction, plus (k10, a) both on reaction 2 (G1 -> G0, rate max(0, k10 - a Hog1p(t))) --
    14 (reaction, parameter) entries (12 + 2).  Reaction 2 is separable: its time factor carries both parameters."""
    from .cmemodel import CmeModelWithSensitivity, propensitygrad, propensitygrad_timevarying
    model = hog1p_model(theta, separable=True)
    P = 14
    par_of_reaction = {0: 0, 2: 3, 3: 4, 4: 5, 5: 6, 6: 7, 7: 8, 8: 9, 9: 10, 10: 11, 11: 12, 12: 13}
    species_of_reaction = {0: 0, 2: 1, 3: 2, 4: 2, 5: 3, 6: 0, 7: 1, 8: 2, 9: 3, 10: 4, 11: 4, 12: 5}
    pattern = np.zeros((13, P), dtype=bool)
    grads = []
    for r in range(13):
        if r == 1:
            pattern[r, 1] = pattern[r, 2] = True
            active = lambda t, p: 1.0 if p[1] - p[2] * hog1p_signal(t) > 0.0 else 0.0
            dt = [_zero_t] * P
            dt[1] = active
            dt[2] = lambda t, p: -hog1p_signal(t) * active(t, p)
            a = model.propensities[r]
            grads.append(propensitygrad_timevarying(a.tfactor, a.statefactor, dt, [_zero_x] * P))
            continue
        ip, k = par_of_reaction[r], species_of_reaction[r]
        pattern[r, ip] = True
        d = [_zero_x] * P
        d[ip] = (lambda x, p, k=k: 1.0 * x[k])
        grads.append(propensitygrad(d))
    return CmeModelWithSensitivity(model, pattern, grads)


def m3d_sens_model(time_varying=True):
    """M-3D (BASELINE.json config 5) with its P = 6 rate constants as sensitivity parameters: reaction r depends on
    parameter r only (6 derivative matrices); the separable time factor of death-3 carries no parameter."""
    from .cmemodel import CmeModelWithSensitivity, propensitygrad, propensitygrad_timevarying
    model = m3d_model(time_varying=time_varying)
    P = 6
    dstate = [lambda x, p: 1.0 + 0.0 * x[0], lambda x, p: 1.0 * x[0], lambda x, p: 1.0 + 0.0 * x[0],
              lambda x, p: 1.0 * x[1], lambda x, p: 1.0 + 0.0 * x[0], lambda x, p: 1.0 * x[2]]
    grads = []
    for r in range(6):
        d = [_zero_x] * P
        d[r] = dstate[r]
        a = model.propensities[r]
        if a.kind == "sep":
            grads.append(propensitygrad_timevarying(a.tfactor, a.statefactor, [_zero_t] * P, d))
        else:
            grads.append(propensitygrad(d))
    return CmeModelWithSensitivity(model, np.eye(6, dtype=bool), grads)
