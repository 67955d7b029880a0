"""Model fixtures restated from the reference's tests (shared by oracle and GPU parity tests)."""
import math

import numpy as np

from oracle.fspmatrix import OProp
from oracle.sensmatrix import OGrad

TELEGRAPH_S = np.array([[-1, 1, 0], [1, -1, 0], [0, 0, 1], [0, 0, -1]]).T   # test/test_statespace.jl:9
TOGGLE_S = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]]).T                   # test/test_statespace.jl:39


def tv(t, p):
    return max(0.0, 1.0 - math.sin(math.pi * t / 2))


def fspmat_propensities(kind):
    """test/test_fspmat.jl:15-38 (0-based parameter/species indices)."""
    a1 = OProp("ti", f=lambda x, p: p[0] * x[0])
    a2 = OProp("ti", f=lambda x, p: p[1] * x[1])
    a2tv = OProp("sep", tfactor=tv, statefactor=lambda x, p: p[1] * x[1])
    a2tvj = OProp("joint", f=lambda t, x, p: tv(t, p) * p[1] * x[1])
    a3 = OProp("ti", f=lambda x, p: p[2] * x[1])
    a4 = OProp("ti", f=lambda x, p: p[3] * x[2])
    return {"ti": [a1, a2, a3, a4], "tv": [a1, a2tv, a3, a4], "tvj": [a1, a2tvj, a3, a4]}[kind]


FSPMAT_THETA = [0.05, 0.1, 5.0, 1.0]

# ---- test/sensmat/telegraph.jl:7-31
SENS_THETA = [0.05, 0.1, 5.0, 0.5, 20.0]
RNACOUNT_MAX = 500


def sens_tfactor(t, p):
    return max(0.0, 1.0 - math.sin(math.pi * t / p[4]))


def sens_dtfactor_dL(t, p):
    # d/dL max(0, 1 - sin(pi t / L)) = cos(pi t/L) * pi t / L^2 where the max is active
    return math.cos(math.pi * t / p[4]) * math.pi * t / p[4] ** 2 if 1.0 - math.sin(math.pi * t / p[4]) >= 0 else 0.0


def sens_telegraph():
    props = [
        OProp("ti", f=lambda x, p: p[0] * x[0]),
        OProp("ti", f=lambda x, p: p[1] * x[1]),
        OProp("ti", f=lambda x, p: p[2] * x[1]),
        OProp("sep", tfactor=sens_tfactor, statefactor=lambda x, p: p[3] * x[2]),
    ]
    zero = lambda x, p: 0.0 * x[0]
    zt = lambda t, p: 0.0
    grads = [
        OGrad("ti", pardiffs=[lambda x, p: 1.0 * x[0], zero, zero, zero, zero]),
        OGrad("ti", pardiffs=[zero, lambda x, p: 1.0 * x[1], zero, zero, zero]),
        OGrad("ti", pardiffs=[zero, zero, lambda x, p: 1.0 * x[1], zero, zero]),
        OGrad("sep", tfactor_pardiffs=[zt, zt, zt, zt, sens_dtfactor_dL],
              statefactor_pardiffs=[zero, zero, zero, lambda x, p: 1.0 * x[2], zero]),
    ]
    pattern = np.zeros((4, 5), dtype=bool)   # test/test_autodiff.jl:56-58: sparse([1,2,3,4,4],[1,2,3,4,5])
    for r, ip in [(0, 0), (1, 1), (2, 2), (3, 3), (3, 4)]:
        pattern[r, ip] = True
    states = [[1, 0, i] for i in range(RNACOUNT_MAX + 1)] + [[0, 1, i] for i in range(RNACOUNT_MAX + 1)]
    return props, grads, pattern, np.array(states, dtype=np.int64)
