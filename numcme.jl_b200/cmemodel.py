"""Host-side model description, mirroring src/cmemodel/*.jl of the reference.

Propensities are opaque host callables, exactly as in the reference: they are evaluated on the host
once per (state, reaction) at matrix assembly (fspsparsematrix.jl:129) and the separable time
factors once per right-hand-side call (:204).  Only their *values* cross the C ABI.

Python conventions (0-based): a state factor is ``f(x, p)`` where ``x[k]`` is the count of species k
and ``p`` the parameter vector; it is called with ``x[k]`` being numpy arrays over all states (one
vectorised call) and falls back to one call per state if the callable is not array-friendly.
"""
from __future__ import annotations

import inspect

import numpy as np

from ._lib import ArgumentError

TIME_INVARIANT, SEPARABLE_TV, JOINT_TV = 0, 1, 2


class Propensity:
    """Base type (propensity.jl:8)."""
    kind = "ti"

    @property
    def kind_code(self) -> int:
        return {"ti": TIME_INVARIANT, "sep": SEPARABLE_TV, "joint": JOINT_TV}[self.kind]


class StandardTimeInvariantPropensity(Propensity):
    """propensity.jl:49-51 -- ``f(x, p)``."""
    kind = "ti"

    def __init__(self, f):
        self.f = f

    def __call__(self, x, p=()):
        return self.f(x, p)


class SeparableTimeVaryingPropensity(Propensity):
    """propensity.jl:93-96 -- ``tfactor(t, p) * statefactor(x, p)``."""
    kind = "sep"

    def __init__(self, tfactor, statefactor):
        self.tfactor = tfactor
        self.statefactor = statefactor

    def __call__(self, t, x, p=()):
        return self.tfactor(t, p) * self.statefactor(x, p)


class JointTimeVaryingPropensity(Propensity):
    """propensity.jl:124-126 -- ``f(t, x, p)``."""
    kind = "joint"

    def __init__(self, f):
        self.f = f

    def __call__(self, t, x, p=()):
        return self.f(t, x, p)


def istimevarying(a: Propensity) -> bool:
    return a.kind != "ti"


def istimeseparable(a: Propensity) -> bool:
    return a.kind == "sep"


def propensity(f, tfactor=None) -> Propensity:
    """propensity.jl:131-144.  ``propensity(f)`` with 2 arguments (x,p) -> time-invariant, with 3
    arguments (t,x,p) -> joint time-varying; ``propensity(xfactor, tfactor)`` -> separable."""
    if tfactor is not None:
        return SeparableTimeVaryingPropensity(tfactor, f)
    nargs = len(inspect.signature(f).parameters)
    if nargs == 2:
        return StandardTimeInvariantPropensity(f)
    if nargs == 3:
        return JointTimeVaryingPropensity(f)
    raise ArgumentError("The callable passed to `propensity()` must have either two arguments (x,p) "
                        "or three arguments (t,x,p).")


# -- gradients (propensitygrad.jl:20-47)
class PropensityGradient:
    kind = "ti"


class StandardTimeInvariantPropensityGradient(PropensityGradient):
    kind = "ti"

    def __init__(self, pardiffs):
        self.pardiffs = list(pardiffs)


class SeparableTimeVaryingPropensityGradient(PropensityGradient):
    kind = "sep"

    def __init__(self, tfactor, statefactor, tfactor_pardiffs, statefactor_pardiffs):
        self.tfactor = tfactor
        self.statefactor = statefactor
        self.tfactor_pardiffs = list(tfactor_pardiffs)
        self.statefactor_pardiffs = list(statefactor_pardiffs)


class JointTimeVaryingPropensityGradient(PropensityGradient):
    kind = "joint"

    def __init__(self, pardiffs):
        self.pardiffs = list(pardiffs)


def propensitygrad(pardiffs):
    return StandardTimeInvariantPropensityGradient(pardiffs)


def propensitygrad_timevarying(*args):
    if len(args) == 1:
        return JointTimeVaryingPropensityGradient(args[0])
    if len(args) == 4:
        return SeparableTimeVaryingPropensityGradient(*args)
    raise ArgumentError("propensitygrad_timevarying takes (pardiffs) or (tfactor, statefactor, dtfactor, dstatefactor)")


class CmeModel:
    """cmemodel.jl:49-53.  ``stoich_matrix`` is species x reactions."""

    def __init__(self, stoich_matrix, propensities, parameters=()):
        self.stoich_matrix = np.asarray(stoich_matrix, dtype=np.int64)
        if self.stoich_matrix.ndim != 2:
            raise ArgumentError("stoichiometry matrix must be 2-D (species x reactions)")
        self.propensities = list(propensities)
        if len(self.propensities) != self.stoich_matrix.shape[1]:
            raise ArgumentError("one propensity per reaction (column of the stoichiometry matrix) is required")
        self.parameters = parameters

    def __repr__(self):
        return (f"Stochastic reaction network with {get_species_count(self)} species, "
                f"{get_reaction_count(self)} reactions and {get_parameter_count(self)} parameters.")


class CmeModelWithSensitivity:
    """cmemodel.jl:104-108: a CmeModel + gradient sparsity pattern (reactions x parameters, bool) +
    one PropensityGradient per reaction.  The reference derives both with ForwardDiff/ModelingToolkit
    (host-side model preparation, out of this path's scope); here they are supplied explicitly, or
    derived by central finite differences with ``CmeModelWithSensitivity.from_finite_differences``."""

    def __init__(self, cmemodel: CmeModel, gradient_sparsity_patterns, propensity_gradients):
        self.cmemodel = cmemodel
        self.gradient_sparsity_patterns = np.asarray(gradient_sparsity_patterns, dtype=bool)
        self.propensity_gradients = list(propensity_gradients)
        R, P = get_reaction_count(cmemodel), get_parameter_count(cmemodel)
        if self.gradient_sparsity_patterns.shape != (R, P):
            raise ArgumentError(f"gradient sparsity pattern must be {R} x {P}")
        if len(self.propensity_gradients) != R:
            raise ArgumentError("one PropensityGradient per reaction is required")


def get_parameters(m):
    return m.cmemodel.parameters if isinstance(m, CmeModelWithSensitivity) else m.parameters


def get_stoich_matrix(m):
    return m.cmemodel.stoich_matrix if isinstance(m, CmeModelWithSensitivity) else m.stoich_matrix


def get_propensities(m):
    return m.cmemodel.propensities if isinstance(m, CmeModelWithSensitivity) else m.propensities


def get_species_count(m):
    return get_stoich_matrix(m).shape[0]


def get_reaction_count(m):
    return get_stoich_matrix(m).shape[1]


def get_parameter_count(m):
    return len(get_parameters(m))


def get_propensity_gradients(m: CmeModelWithSensitivity):
    return m.propensity_gradients


def get_gradient_sparsity_patterns(m: CmeModelWithSensitivity):
    return m.gradient_sparsity_patterns


def eval_over_columns(fn, cols, p, t=None) -> np.ndarray:
    """Evaluate ``fn(x,p)`` (or ``fn(t,x,p)``) over n states given species-major: ``cols[k]`` = float64[n] of species k.

    One vectorised call with ``x[k] = cols[k]`` is tried first.  Its result is only trusted after a spot check against
    scalar calls on a few states: a callable that reduces over species with numpy (``np.sum(x)``, ``np.max(x)`` ...)
    returns a 0-d or wrongly shaped value for column arrays, and a broadcast of it would silently assemble a wrong
    generator.  On any mismatch the per-state loop (the reference's calling convention, fspsparsematrix.jl:129) is used."""
    n = cols[0].shape[0] if len(cols) else 0

    def scalar(i):
        x = [int(c[i]) for c in cols]
        return float(fn(x, p) if t is None else fn(t, x, p))

    try:
        v = fn(cols, p) if t is None else fn(t, cols, p)
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 0:
            v = np.full(n, float(v))
        if v.shape == (n,):
            ok = True
            for i in sorted({0, n // 2, n - 1} if n else ()):
                ref = scalar(i)
                if not (v[i] == ref or abs(v[i] - ref) <= 1e-12 * max(abs(ref), abs(v[i])) or
                        (np.isnan(ref) and np.isnan(v[i]))):
                    ok = False
                    break
            if ok:
                return np.ascontiguousarray(v)
    except Exception:
        pass
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        out[i] = scalar(i)
    return out


def eval_over_states(fn, states: np.ndarray, p, t=None) -> np.ndarray:
    """``eval_over_columns`` for states given row-major (n x NS integers) -> float64[n]."""
    return eval_over_columns(fn, [states[:, k].astype(np.float64) for k in range(states.shape[1])], p, t)
