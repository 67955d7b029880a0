"""Oracle restatement of the forward-sensitivity FSP ``solve`` loop (TEST INFRASTRUCTURE ONLY).

Follows
  /root/reference/src/forwardsenscme/sparse/forwardsenscmesparse.jl:99-215   (solve)
  /root/reference/src/forwardsenscme/sparse/fsspaceadapterssparse.jl:21-60   (ForwardSensRStepAdapter init!/adapt!)

Block vector ``u = [p; s_1; ...; s_P]``, every block of length ``N = n + R`` (states then sinks), right-hand side
``SensFspMatrixOracle.matvec`` (sensfspmatrixsparse.jl:97-142), terminal event on the sinks of the probability block
(``:153-155``).  The reference integrates with DifferentialEquations.jl / CVODE (third-party); the stand-in is
``scipy.integrate.solve_ivp`` with the exact block Jacobian, so transient values are *parity unpinned by the reference*
(test/test_sensfsp.jl only checks that the solve runs and conserves mass); tests pin them through the analytic
sensitivities of the birth-death process instead.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.integrate import solve_ivp

from .sensmatrix import SensFspMatrixOracle
from .statespace import StateSpaceOracleFast

EPS = np.finfo(np.float64).eps


class ForwardSensRStepAdapterOracle:
    """fsspaceadapterssparse.jl:10-14"""

    def __init__(self, initial_step_count, max_step_count, dropstates):
        self.initial_step_count = initial_step_count
        self.max_step_count = max_step_count
        self.dropstates = dropstates

    def init(self, space, p, S):
        """:21-30"""
        nold = space.get_state_count()
        space.expand(self.initial_step_count)
        z = np.zeros(space.get_state_count() - nold)
        return np.concatenate([p, z]), [np.concatenate([s, z]) for s in S]

    def adapt(self, space, p, S, sinks, dsinks, t, tend, fsptol):
        """:37-60"""
        if self.dropstates:
            pids = np.argsort(p, kind="stable")                              # sortperm(p)
            dropcount = int(np.sum((p.sum() - np.cumsum(p[pids])) >= (1.0 - t * fsptol / tend)))
            ids = np.sort(pids[:dropcount]) + 1
            if ids.size:
                space.deleteat(ids)
                p = np.delete(p, ids - 1)
                S = [np.delete(s, ids - 1) for s in S]
        nold = space.get_state_count()
        space.expand(self.max_step_count)
        z = np.zeros(space.get_state_count() - nold)
        return np.concatenate([p, z]), [np.concatenate([s, z]) for s in S]


def _block_jac(SA: SensFspMatrixOracle):
    """d(rhs)/du of the block system at time t: A(t) on the diagonal blocks, dA/dtheta_ip(t) in block (ip, 0)."""
    A = SA.fspmatrix
    P = SA.parameter_count

    def jac(t, u):
        th = A.parameters
        At = None
        for c, M in A.terms_at(t):
            At = c * M if At is None else At + c * M
        rows = [[At] + [None] * P]
        for ip in range(P):
            D = SA.timeinvariant_matdiffs[ip].copy()
            for (jp, r, j, dM) in SA.sep_entries:
                if jp == ip:
                    D = D + A.propensities[r - 1].tfactor(t, th) * dM \
                        + SA.gradients[r - 1].tfactor_pardiffs[ip](t, th) * A.separabletv_factormatrices[j]
            for (jp, r, dM) in SA.joint_entries:
                if jp == ip:
                    from .fspmatrix import FspMatrixOracle
                    FspMatrixOracle.update_sparsematrix(dM, A.states, SA.gradients[r - 1].pardiffs[ip], t, th)
                    D = D + dM
            row = [D] + [None] * P
            row[ip + 1] = At
            rows.append(row)
        return sp.bmat(rows, format="csc")
    return jac


def solve_sens(stoich, propensities, gradients, pattern, parameters, states0, p0, S0, tspan, adapter, saveat=None,
               fsptol=1e-6, odeatol=1e-10, odertol=1e-4, method="BDF"):
    """forwardsenscmesparse.jl:99-215.  Returns dict(t, states[k], p[k], sinks[k], S[k][ip], dsinks[k][ip])."""
    tstart, tend = min(tspan), max(tspan)
    P = len(parameters)
    if len(S0) != P:
        raise ValueError("Initial condition does not match CME model.")
    saveat = None if saveat is None else np.asarray(saveat, dtype=np.float64)
    space = StateSpaceOracleFast(stoich, states0)
    R = space.get_sink_count()
    p, S = adapter.init(space, np.array(p0, dtype=np.float64), [np.array(s, dtype=np.float64) for s in S0])
    tnow = tstart
    N = p.size + R
    unow = np.zeros(N * (P + 1))
    unow[:p.size] = p
    for ip in range(P):
        unow[(ip + 1) * N:(ip + 1) * N + p.size] = S[ip]
    out = {"t": [], "states": [], "p": [], "sinks": [], "S": [], "dsinks": [], "adapts": 0, "rhs_calls": 0}

    def push(t, st, u, n, N):
        out["t"].append(float(t))
        out["states"].append(st)
        out["p"].append(u[:n].copy())
        out["sinks"].append(u[n:N].copy())
        out["S"].append([u[(ip + 1) * N:(ip + 1) * N + n].copy() for ip in range(P)])
        out["dsinks"].append([u[(ip + 1) * N + n:(ip + 2) * N].copy() for ip in range(P)])

    while tnow < tend:
        SA = SensFspMatrixOracle(space, propensities, gradients, pattern, parameters)
        N = SA.fspmatrix.rowcount
        n = N - R

        def rhs(t, u, SA=SA):
            out["rhs_calls"] += 1
            return SA.matvec(t, u)

        def event(t, u, n=n, N=N):
            return u[n:N].sum() - fsptol * t / tend
        event.terminal = True
        event.direction = 1
        te = None if saveat is None else saveat[(saveat >= tnow) & (saveat <= tend)]
        kw = {"jac": _block_jac(SA)} if method in ("BDF", "Radau") else {}
        sol = solve_ivp(rhs, (tnow, tend), unow, method=method, atol=odeatol, rtol=odertol, events=event, t_eval=te, **kw)
        hit = sol.status == 1
        y = np.asarray(sol.y, dtype=np.float64).reshape(unow.size, -1)
        t_stop = float(sol.t_events[0][0]) if hit else tend
        u_stop = sol.y_events[0][0] if hit else (y[:, -1] if te is None else None)
        st = space.states_array().copy()
        for k in range(np.asarray(sol.t).size):
            if sol.t[k] <= t_stop:
                push(sol.t[k], st, y[:, k], n, N)
        if u_stop is None:
            u_stop = solve_ivp(rhs, (tnow, tend), unow, method=method, atol=odeatol, rtol=odertol, **kw).y[:, -1]
        tnow = t_stop
        if tnow < tend:
            p = u_stop[:n].copy()
            sinks = u_stop[n:N].copy()
            S = [u_stop[(ip + 1) * N:(ip + 1) * N + n].copy() for ip in range(P)]
            dsinks = [u_stop[(ip + 1) * N + n:(ip + 2) * N].copy() for ip in range(P)]
            p, S = adapter.adapt(space, p, S, sinks, dsinks, tnow, tend, fsptol)
            out["adapts"] += 1
            if sinks.sum() >= tnow * fsptol / tend:            # :187-189
                sinks -= EPS
            n2 = p.size
            N2 = n2 + R
            unow = np.zeros(N2 * (P + 1))
            unow[:n2] = p
            unow[n2:N2] = sinks
            for ip in range(P):
                unow[(ip + 1) * N2:(ip + 1) * N2 + n2] = S[ip]
                unow[(ip + 1) * N2 + n2:(ip + 2) * N2] = dsinks[ip]
        else:
            push(tnow, st, u_stop, n, N)
    return out
