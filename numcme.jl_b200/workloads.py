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
