"""Forward-sensitivity FSP solve behind the reference's API.

Reference: src/forwardsenscme/sparse/forwardsenscmesparse.jl (initial condition :37-75, ``solve`` :99-215),
src/forwardsenscme/sparse/fsspaceadapterssparse.jl (``ForwardSensRStepAdapter`` :10-60),
src/forwardsenscme/sparse/sensoutputsparse.jl (outputs :20-85).

The block vector ``[p; s_1; ...; s_P]`` (each block n + R long) stays in HBM; the right-hand side is the fused block
matvec (K2), the integrator is the native one (``ncme_sens_solve_segment``), pruning/expansion run on the device.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import _lib as L
from .cmemodel import CmeModelWithSensitivity, get_parameter_count, get_stoich_matrix
from .device import DeviceVector
from .fspvector import FspVectorSparse
from .sensmatrix import ForwardSensFspMatrixSparse
from .statespace import StateSpaceSparse
from .transientcme import EPS, _method_code, _saveat_array


class ForwardSensFspInitialConditionSparse:
    """forwardsenscmesparse.jl:37-52"""

    def __init__(self, states, p, S):
        self.states = np.asarray(states, dtype=np.int64)
        if self.states.ndim == 1:
            self.states = self.states.reshape(1, -1)
        self.p = np.asarray(p, dtype=np.float64)
        self.S = [np.asarray(s, dtype=np.float64) for s in S]


def forwardsens_initial_condition(states, probabilities, sensitivity):
    """forwardsenscmesparse.jl:61-75"""
    if len(states) == 0:
        raise L.ArgumentError("Empty state list in input.")
    return ForwardSensFspInitialConditionSparse(states, probabilities, sensitivity)


def get_probability(ic: ForwardSensFspInitialConditionSparse):
    """forwardsenscmesparse.jl:52"""
    return ic.p


def get_sensitivity(ic: ForwardSensFspInitialConditionSparse):
    """forwardsenscmesparse.jl:53"""
    return ic.S


class ForwardSensRStepAdapter:
    """fsspaceadapterssparse.jl:10-14"""

    def __init__(self, initial_step_count: int, max_step_count: int, dropstates: bool):
        self.initial_step_count = int(initial_step_count)
        self.max_step_count = int(max_step_count)
        self.dropstates = bool(dropstates)


class AdaptiveForwardSensFspSparse:
    """forwardsenscmesparse.jl:15-24"""

    def __init__(self, space_adapter, ode_method=None):
        self.space_adapter = space_adapter
        self.ode_method = ode_method


class ForwardSensFspOutputSliceSparse:
    def __init__(self, t, p, sinks, S, dsinks):
        self.t, self.p, self.sinks, self.S, self.dsinks = t, p, sinks, S, dsinks


class ForwardSensFspOutputSparse:
    """sensoutputsparse.jl:20-26"""

    def __init__(self):
        self.t, self.p, self.sinks, self.S, self.dsinks = [], [], [], [], []
        self.stats = {}

    def __len__(self):
        return len(self.t)

    def __getitem__(self, ind):
        if isinstance(ind, (list, tuple, np.ndarray)):
            return [self[int(i)] for i in ind]
        if ind < 0:
            ind += len(self.t)
        if ind >= len(self.t):
            raise L.ArgumentError("Requested index exceeds array limit.")
        return ForwardSensFspOutputSliceSparse(self.t[ind], self.p[ind], self.sinks[ind], self.S[ind], self.dsinks[ind])

    def _push(self, t, states, uu, n, R, P):
        N = n + R
        self.t.append(float(t))
        self.p.append(FspVectorSparse(states, uu[:n]))
        self.sinks.append(uu[n:N].copy())
        self.S.append([FspVectorSparse(states, uu[(ip + 1) * N:(ip + 1) * N + n]) for ip in range(P)])
        self.dsinks.append([uu[(ip + 1) * N + n:(ip + 2) * N].copy() for ip in range(P)])


def _grow_all(ctx, vecs, n_new):
    out = []
    for v in vecs:
        if v.n == n_new:
            out.append(v)
            continue
        q = DeviceVector.zeros(ctx, n_new)
        if v.n:
            q.view(0, v.n).copy_from(v)
        out.append(q)
    return out


def sens_init_(space, adapter, vecs):
    """init!(statespace, adapter, p, S, t, fsptol)   fsspaceadapterssparse.jl:21-30  (vecs = [p, S_1..S_P])"""
    space.expand_(adapter.initial_step_count)
    return _grow_all(space.ctx, vecs, space.get_state_count())


def sens_adapt_(space, adapter, vecs, t, tend, fsptol):
    """adapt!(statespace, adapter, p, S, sinks, dsinks, t, tend, fsptol)   fsspaceadapterssparse.jl:37-60"""
    ctx = space.ctx
    if adapter.dropstates and vecs[0].n:
        dropped = space.prune_by_mass_(vecs[0], 1.0 - t * fsptol / tend, strict=False)
        if dropped:
            new = []
            for v in vecs:                       # deleteat!(p, dropids); deleteat!(svec, dropids)
                q = DeviceVector(ctx, space.get_state_count())
                space.compact_vector(v, q)
                new.append(q)
            vecs = new
    space.expand_(adapter.max_step_count)
    return _grow_all(ctx, vecs, space.get_state_count())


def solve_sens(model: CmeModelWithSensitivity, initial_condition: ForwardSensFspInitialConditionSparse, tspan,
               sensfspalgorithm: AdaptiveForwardSensFspSparse, saveat=None, fsptol=1.0e-6, odeatol=1.0e-10,
               odertol=1.0e-4, verbose=False, ctx=None) -> ForwardSensFspOutputSparse:
    """solve(model::CmeModelWithSensitivity, ic, tspan, alg; saveat, fsptol, odeatol, odertol, verbose)
    forwardsenscmesparse.jl:99-215"""
    tstart, tend = min(tspan), max(tspan)
    P = get_parameter_count(model)
    if len(initial_condition.S) != P:
        raise L.ArgumentError("Initial condition does not match CME model. Initial condition must contain `np` "
                              "sensitivity vectors where `np` is the number of CME model parameters.")
    adapter = sensfspalgorithm.space_adapter
    method = _method_code(sensfspalgorithm.ode_method)
    sv = _saveat_array(saveat, tspan)
    space = StateSpaceSparse(get_stoich_matrix(model), initial_condition.states, ctx=ctx)
    ctx = space.ctx
    R = space.get_sink_count()
    idx = space.lookup(initial_condition.states)
    n0 = space.get_state_count()

    def place(vals):
        a = np.zeros(n0)
        a[idx[idx > 0] - 1] = np.asarray(vals)[idx > 0]
        return DeviceVector.from_host(ctx, a)
    t_wall = time.perf_counter()
    vecs = sens_init_(space, adapter, [place(initial_condition.p)] + [place(s) for s in initial_condition.S])
    sinks = np.zeros(R)
    dsinks = [np.zeros(R) for _ in range(P)]
    out = ForwardSensFspOutputSparse()
    tot = {"steps": 0, "rejected": 0, "rhs_evals": 0, "launches": 0, "adapts": 0, "incremental_builds": 0}
    tnow = tstart
    prev_SA = None
    while tnow < tend:
        # after an adapt! only the appended states are evaluated (propensities and their parameter derivatives); the
        # rows of the survivors are carried over on the device (the reference rebuilds from scratch, :140)
        SA = ForwardSensFspMatrixSparse(model, space, previous=prev_SA)
        if prev_SA is not None:
            prev_SA.close()
            prev_SA = None
        tot["incremental_builds"] += int(SA.incremental)
        n = space.get_state_count()
        N = n + R
        U = DeviceVector.zeros(ctx, N * (P + 1))
        for b, v in enumerate(vecs):
            U.view(b * N, n).copy_from(v)
        U.view(n, R).upload(sinks)
        for ip in range(P):
            U.view((ip + 1) * N + n, R).upload(dsinks[ip])
        saved_t, saved_u = [], []
        nr, nent = SA.fspmatrix.nr, len(SA.entries)

        def coef_cb(t, ptr, _user, SA=SA, nr=nr, nent=nent):
            coef = SA._prepare(t)
            for r in range(nr):
                ptr[r] = coef[r]
            for e in range(nent):
                ptr[nr + e] = SA._dcoef[e]

        def save_cb(t, ptr, _user, L_=N * (P + 1)):
            saved_t.append(float(t))
            saved_u.append(np.ctypeslib.as_array(ptr, shape=(L_,)).copy())
        guard = L.CallbackGuard()
        ccb, scb = L.COEF_FN(guard.wrap(coef_cb)), L.SAVE_FN(guard.wrap(save_cb))
        opts = L.SolveOpts()
        opts.rtol, opts.atol = float(odertol), float(odeatol)
        opts.check_event, opts.event_slope = 1, float(fsptol / tend)
        opts.save_every_step = 1 if sv is None else 0
        sva = np.ascontiguousarray(sv if sv is not None else [], dtype=np.float64)
        opts.nsave, opts.save_t = int(sva.size), sva.ctypes.data_as(C.POINTER(C.c_double))
        opts.h_init, opts.max_steps, opts.method = 0.0, 0, method
        stats = L.SolveStats()
        status = L.load().ncme_sens_solve_segment(SA._h, C.cast(ccb, C.c_void_p), C.cast(scb, C.c_void_p), None,
                                                  float(tnow), float(tend), C.c_void_p(U.ptr), C.byref(opts),
                                                  C.byref(stats))
        guard.reraise()                  # exceptions of user closures raised inside the callbacks
        L.check(status)
        for k in ("steps", "rejected", "rhs_evals", "launches"):
            tot[k] += getattr(stats, k)
        states = space.get_states()
        for t, uu in zip(saved_t, saved_u):
            out._push(t, states, uu, n, R, P)
        tnow = stats.t_final
        if stats.event_hit and tnow < tend:
            vecs = [U.view(b * N, n).clone() for b in range(P + 1)]
            sinks = U.to_host(n, R)
            dsinks = [U.to_host((ip + 1) * N + n, R) for ip in range(P)]
            prev_SA = SA
            vecs = sens_adapt_(space, adapter, vecs, tnow, tend, fsptol)
            tot["adapts"] += 1
            if sinks.sum() >= tnow * fsptol / tend:      # forwardsenscmesparse.jl:187-189
                sinks = sinks - EPS
            if verbose:
                print(f"At t = {tnow:.2f}: update sate space. New size: {space.get_state_count()} states.")
        else:
            out._push(tnow, states, U.to_host(), n, R, P)
            SA.close()
            tnow = tend
    tot["wall_s"] = time.perf_counter() - t_wall
    tot["final_states"] = space.get_state_count()
    out.stats = tot
    return out
