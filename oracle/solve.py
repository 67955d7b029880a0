"""Oracle restatement of the space adapters and the FSP ``solve`` loops (TEST INFRASTRUCTURE ONLY).

Follows
  /root/reference/src/transientcme/sparse/rstepadapters.jl:23-52   (RStepAdapter init!/adapt!)
  /root/reference/src/transientcme/sparse/rstepadapters.jl:74-110  (SelectiveRStepAdapter)
  /root/reference/src/transientcme/sparse/fspsolve.jl:10-41        (fixed-space solve)
  /root/reference/src/transientcme/sparse/fspsolve.jl:105-197      (adaptive solve)

The reference integrates with DifferentialEquations.jl / Sundials CVODE_BDF(GMRES) and a
ContinuousCallback (third-party, not under /root/reference).  The stand-in here is
``scipy.integrate.solve_ivp`` (BDF or LSODA) with a terminal event on
``sum(sinks) - fsptol*t/tend``.  Transient solution values are therefore *parity unpinned by
the reference* (its tests only pin conservation and self-consistency, test/test_solver.jl:60-99);
tests pin them through analytic solutions instead.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp

from .fspmatrix import FspMatrixOracle
from .statespace import StateSpaceOracleFast

EPS = np.finfo(np.float64).eps


def _sparse_jac(A):
    """Exact Jacobian A(t) = sum of the reference's per-term CSC matrices (the ODE is linear): lets scipy's implicit
    methods factorise a sparse matrix instead of estimating a dense one column by column."""
    def jac(t, u):
        J = None
        for c, M in A.terms_at(t):
            J = c * M if J is None else J + c * M
        return J.tocsc()
    return jac


class RStepAdapterOracle:
    selective = False

    def __init__(self, initial_step_count, max_step_count, dropstates):
        self.initial_step_count = initial_step_count
        self.max_step_count = max_step_count
        self.dropstates = dropstates

    def init(self, space, p):
        nold = space.get_state_count()
        space.expand(self.initial_step_count)
        return np.concatenate([p, np.zeros(space.get_state_count() - nold)])

    def drop_ids(self, p, t, tend, fsptol):
        """rstepadapters.jl:41-43 / :93-96 -- returns sorted 1-based ids to delete."""
        pids = np.argsort(p, kind="stable")
        tail = p.sum() - np.cumsum(p[pids])
        thr = 1.0 - t * fsptol / tend
        dropcount = int(np.sum(tail > thr if self.selective else tail >= thr))
        return np.sort(pids[:dropcount]) + 1

    def adapt(self, space, p, sinks, t, tend, fsptol, dsinks=None):
        if self.selective and p.size == 0:
            raise ValueError("Empty `p` input in `adapt!`.")
        if self.dropstates:
            ids = self.drop_ids(p, t, tend, fsptol)
            if ids.size:
                space.deleteat(ids)
                p = np.delete(p, ids - 1)
        nold = space.get_state_count()
        if self.selective:
            only = [int(r) + 1 for r in np.nonzero(dsinks > 0)[0]]
            # reference: an empty `onlyreactions` means "all reactions" (sparsestatespace.jl:163)
            space.expand(self.max_step_count, onlyreactions=only)
        else:
            space.expand(self.max_step_count)
        return np.concatenate([p, np.zeros(space.get_state_count() - nold)])


class SelectiveRStepAdapterOracle(RStepAdapterOracle):
    selective = True


def solve_fixed(stoich, propensities, parameters, states, p0, tspan, saveat=None,
                odeatol=1e-6, odertol=1e-4, method="BDF", sparse_jac=False):
    """fspsolve.jl:10-41 on a fixed state space given by ``states`` (rows)."""
    space = StateSpaceOracleFast(stoich, states)
    A = FspMatrixOracle(space, propensities, parameters)
    R = space.get_sink_count()
    u0 = np.concatenate([p0, np.zeros(R)])
    kw = {"jac": _sparse_jac(A)} if sparse_jac and method in ("BDF", "Radau") else {}
    sol = solve_ivp(lambda t, u: A.matvec(t, u), tspan, u0, method=method, atol=odeatol, rtol=odertol,
                    t_eval=saveat, **kw)
    n = space.get_state_count()
    return {"t": sol.t, "states": space.states_array(), "p": [sol.y[:n, k] for k in range(sol.t.size)],
            "sinks": [sol.y[n:, k] for k in range(sol.t.size)]}


def solve_adaptive(stoich, propensities, parameters, states0, p0, tspan, adapter, saveat=None,
                   fsptol=1e-6, odeatol=1e-6, odertol=1e-4, method="BDF", verbose=False, sparse_jac=False):
    """fspsolve.jl:105-197.  Returns dict(t, states[k], p[k], sinks[k])."""
    tstart, tend = min(tspan), max(tspan)
    saveat = None if saveat is None else np.asarray(saveat, dtype=np.float64)
    space = StateSpaceOracleFast(stoich, states0)
    R = space.get_sink_count()
    p = adapter.init(space, np.array(p0, dtype=np.float64))
    tnow = tstart
    unow = np.concatenate([p, np.zeros(R)])
    A = FspMatrixOracle(space, propensities, parameters)
    out = {"t": [], "states": [], "p": [], "sinks": [], "rhs_calls": 0, "adapts": 0}

    while tnow < tend:
        n = space.get_state_count()

        def rhs(t, u, A=A):
            out["rhs_calls"] += 1
            return A.matvec(t, u)

        def event(t, u, n=n):
            return u[n:].sum() - fsptol * t / tend
        event.terminal = True
        event.direction = 1     # rising crossings only (g(t0) = 0 at the very start must not fire)

        te = None
        if saveat is not None:
            te = saveat[(saveat >= tnow) & (saveat <= tend)]
        kw = {"jac": _sparse_jac(A)} if sparse_jac and method in ("BDF", "Radau") else {}
        sol = solve_ivp(rhs, (tnow, tend), unow, method=method, atol=odeatol, rtol=odertol,
                        events=event, t_eval=te, dense_output=False, **kw)
        hit = sol.status == 1
        sol.t = np.asarray(sol.t, dtype=np.float64)
        sol.y = np.asarray(sol.y, dtype=np.float64).reshape(unow.size, -1)
        t_stop = float(sol.t_events[0][0]) if hit else tend
        u_stop = sol.y_events[0][0] if hit else (sol.y[:, -1] if te is None else None)
        st = space.states_array().copy()
        for k in range(sol.t.size):
            if sol.t[k] <= t_stop:
                out["t"].append(float(sol.t[k]))
                out["states"].append(st)
                out["p"].append(sol.y[:n, k].copy())
                out["sinks"].append(sol.y[n:, k].copy())
        if u_stop is None:                       # reached tend with t_eval: integrate state at tend
            s2 = solve_ivp(rhs, (tnow, tend), unow, method=method, atol=odeatol, rtol=odertol, **kw)
            u_stop = s2.y[:, -1]
        tnow = t_stop
        if tnow < tend:
            p = u_stop[:n].copy()
            sinks = u_stop[n:].copy()
            dsinks = A.matvec(tnow, u_stop)[n:]
            p = adapter.adapt(space, p, sinks, tnow, tend, fsptol, dsinks=dsinks)
            A = FspMatrixOracle(space, propensities, parameters)
            out["adapts"] += 1
            if sinks.sum() >= tnow * fsptol / tend:
                sinks -= EPS
            unow = np.concatenate([p, sinks])
            if verbose:
                print(f"t = {tnow:.2f}. Update state space. New size: {space.get_state_count()}.")
        else:
            out["t"].append(tnow)
            out["states"].append(st)
            out["p"].append(u_stop[:n].copy())
            out["sinks"].append(u_stop[n:].copy())
    return out
