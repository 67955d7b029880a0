"""Transient FSP solves behind the reference's API, with the FSP vector resident in HBM.

Reference: src/transientcme/sparse/fspsolve.jl (fixed-space ``solve`` :10-41, ``AdaptiveFspSparse``
:59-62, adaptive ``solve`` :105-197) and src/transientcme/sparse/rstepadapters.jl (``RStepAdapter``
:12-52, ``SelectiveRStepAdapter`` :61-110).

The reference hands the right-hand side to DifferentialEquations.jl / Sundials (host vectors).  Here
``ode_method=None`` (a legal value of the reference's ``Union{Nothing,AbstractODEAlgorithm}`` field)
selects the native device-resident integrator of libncme (``ncme_solve_segment``): u, the stage
vectors, the error norm and the sink event all stay on the GPU; per step only a norm and R sink
entries cross PCIe.  Space adaptation (prune by mass, expand) runs on the device as well.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import _lib as L
from .cmemodel import CmeModel
from .device import DeviceVector
from .fspmatrix import FspMatrixSparse, matvec_
from .fspvector import FspOutputSparse, FspVectorSparse
from .statespace import StateSpaceSparse

EPS = float(np.finfo(np.float64).eps)


class NativeRK45:
    """Device-resident explicit Dormand-Prince 5(4) (libncme method 0): exact linear invariants, best for non-stiff
    problems and tight tolerances."""
    method = 0


class NativeBDF:
    """Device-resident variable-order BDF (NDF) + matrix-free Jacobi-preconditioned GMRES (libncme method 1): the
    counterpart of the ``CVODE_BDF(linear_solver=:GMRES)`` every example of the reference uses.  ``ode_method=None``
    selects it."""
    method = 1


class NativeBDFClassic(NativeBDF):
    """method 2: force the launch-per-operation BDF of csrc/bdf.cu (what sharded and very large problems use)."""
    method = 2


class NativeBDFFused(NativeBDF):
    """method 3: force the one-kernel-per-step BDF of csrc/bdf_fused.cu (the default below ~2e6 states on one GPU)."""
    method = 3


class RStepAdapter:
    """rstepadapters.jl:12-16"""
    selective = False

    def __init__(self, initial_step_count: int, max_step_count: int, dropstates: bool):
        self.initial_step_count = int(initial_step_count)
        self.max_step_count = int(max_step_count)
        self.dropstates = bool(dropstates)


class SelectiveRStepAdapter(RStepAdapter):
    """rstepadapters.jl:61-65: only explores through reactions whose sink is still gaining mass."""
    selective = True


class AdaptiveFspSparse:
    """fspsolve.jl:59-62"""

    def __init__(self, ode_method=None, space_adapter=None):
        if space_adapter is None:
            raise L.ArgumentError("space_adapter is required")
        self.ode_method = ode_method
        self.space_adapter = space_adapter


def _grow(ctx, p: DeviceVector, n_new: int) -> DeviceVector:
    """append!(p, zeros(n_new - length(p)))"""
    if n_new == p.n:
        return p
    q = DeviceVector.zeros(ctx, n_new)
    if p.n:
        q.view(0, p.n).copy_from(p)
    return q


def init_(space: StateSpaceSparse, adapter: RStepAdapter, p: DeviceVector, t: float, fsptol: float) -> DeviceVector:
    """init!(statespace, adapter, p, t, fsptol)   rstepadapters.jl:23-28 / :74-79"""
    space.expand_(adapter.initial_step_count)
    return _grow(space.ctx, p, space.get_state_count())


def adapt_(space: StateSpaceSparse, adapter: RStepAdapter, p: DeviceVector, sinks: np.ndarray, t: float, tend: float,
           fsptol: float, dsinks: np.ndarray | None = None) -> DeviceVector:
    """adapt!(statespace, adapter, p, sinks, t, tend, fsptol; integrator)   rstepadapters.jl:35-52 / :86-110.
    ``dsinks`` (d sinks / dt at t, what the reference reads through get_du!) is required by the selective adapter."""
    ctx = space.ctx
    if adapter.selective and p.n == 0:
        raise L.ArgumentError("Empty `p` input in `adapt!`.")
    if adapter.dropstates and p.n:
        thr = 1.0 - t * fsptol / tend
        dropped = space.prune_by_mass_(p, thr, strict=adapter.selective)
        if dropped:
            q = DeviceVector(ctx, space.get_state_count())
            space.compact_vector(p, q)
            p = q
    if adapter.selective:
        if dsinks is None:
            raise L.ArgumentError("SelectiveRStepAdapter needs the sink derivatives")
        only = [int(r) + 1 for r in np.nonzero(np.asarray(dsinks) > 0)[0]]
        space.expand_(adapter.max_step_count, onlyreactions=only)   # empty => all, like the reference (:163)
    else:
        space.expand_(adapter.max_step_count)
    return _grow(ctx, p, space.get_state_count())


class _Segment:
    """One call of ncme_solve_segment with host callbacks for time factors and output slices."""

    def __init__(self, A: FspMatrixSparse, rtol, atol, method=0):
        self.A = A
        self.rtol, self.atol, self.method = rtol, atol, method
        self.saved_t, self.saved_u = [], []
        N = A.rowcount

        def coef_cb(t, coef_ptr, _user):
            c = A.coefficients(t)
            A._refresh_joint(t)
            for r in A.device_separable_ids:
                coef_ptr[r - 1] = c[r - 1]

        def save_cb(t, u_ptr, _user):
            self.saved_t.append(float(t))
            self.saved_u.append(np.ctypeslib.as_array(u_ptr, shape=(N,)).copy())

        self._guard = L.CallbackGuard()
        self._coef_cb = L.COEF_FN(self._guard.wrap(coef_cb))
        self._save_cb = L.SAVE_FN(self._guard.wrap(save_cb))
        self.needs_coef = bool(A.device_separable_ids or A.device_joint_ids)

    def run(self, u: DeviceVector, t0, t1, saveat=None, save_every_step=False, event_slope=None):
        opts = L.SolveOpts()
        opts.rtol, opts.atol = float(self.rtol), float(self.atol)
        opts.check_event = 0 if event_slope is None else 1
        opts.event_slope = 0.0 if event_slope is None else float(event_slope)
        opts.save_every_step = 1 if save_every_step else 0
        sv = np.ascontiguousarray(saveat if saveat is not None else [], dtype=np.float64)
        opts.nsave = int(sv.size)
        opts.save_t = sv.ctypes.data_as(C.POINTER(C.c_double))
        opts.h_init = 0.0
        opts.max_steps = 0
        opts.method = int(self.method)
        stats = L.SolveStats()
        self.saved_t, self.saved_u = [], []
        cb = C.cast(self._coef_cb, C.c_void_p) if self.needs_coef else None
        status = L.load().ncme_solve_segment(self.A.handle, cb, C.cast(self._save_cb, C.c_void_p), None, float(t0),
                                             float(t1), C.c_void_p(u.ptr), C.byref(opts), C.byref(stats))
        self._guard.reraise()            # an exception raised inside a callback (the C side stopped: NCME_ERR_ABORTED)
        L.check(status)
        return stats


def _method_code(ode_method):
    if ode_method is None:
        return 1          # the native counterpart of the reference's CVODE_BDF(linear_solver=:GMRES)
    if hasattr(ode_method, "method"):
        return int(ode_method.method)
    raise L.ArgumentError("ode_method must be None (native device integrator) or a Native* method object; "
                          "DifferentialEquations.jl algorithms only exist on the Julia side")


def _saveat_array(saveat, tspan):
    if saveat is None:
        return None
    if np.isscalar(saveat):
        return np.arange(tspan[0], tspan[1] + 0.5 * saveat, saveat, dtype=np.float64)
    sv = np.asarray(list(saveat), dtype=np.float64)
    return sv if sv.size else None


def _initial(model, initial_distribution):
    if not isinstance(initial_distribution, FspVectorSparse):
        raise L.ArgumentError("initial_distribution must be a FspVectorSparse")
    return initial_distribution.states, np.array(initial_distribution.values, dtype=np.float64)


def solve(model: CmeModel, initial_distribution: FspVectorSparse, tspan, algorithm=None, saveat=None, fsptol=1.0e-6,
          odeatol=None, odertol=1.0e-4, verbose=False, ctx=None, comm=None, detect_separable=True) -> FspOutputSparse:
    """``solve(model, p0, tspan, ode_method; saveat, odeatol, odertol)``   (fixed space, fspsolve.jl:10-41) when
    ``algorithm`` is None or an ODE method, and
    ``solve(model, p0, tspan, fspalgorithm::AdaptiveFspSparse; saveat, fsptol, odeatol, odertol, verbose)``
    (adaptive, fspsolve.jl:105-197) when it is an ``AdaptiveFspSparse``.

    ``comm`` (parallel.Comm, one process per GPU) row-shards the matrix and every FSP vector over the ranks; the
    state space and the adaptation decisions are replicated, every rank returns the same (gathered) output.

    ``detect_separable`` (not in the reference): joint time-varying propensities that are numerically a product
    c(t) g(x) are integrated on the separable path (see ``FspMatrixSparse``).  The classification is re-checked on
    sentinel states at every time the integrator uses; if it ever fails, the running segment is discarded and
    repeated with the exact joint path (``_update_sparsematrix!``) -- never a silently wrong generator."""
    if comm is not None and ctx is None:
        ctx = comm.ctx
    from .cmemodel import CmeModelWithSensitivity
    sens = isinstance(model, CmeModelWithSensitivity)
    if odeatol is None:                                   # code defaults: fspsolve.jl:109 / forwardsenscmesparse.jl:106
        odeatol = 1.0e-10 if sens else 1.0e-6
    if sens:                                              # forwardsenscmesparse.jl:99
        from .forwardsenscme import solve_sens
        return solve_sens(model, initial_distribution, tspan, algorithm, saveat=saveat, fsptol=fsptol, odeatol=odeatol,
                          odertol=odertol, verbose=verbose, ctx=ctx)
    if isinstance(algorithm, AdaptiveFspSparse):
        return _solve_adaptive(model, initial_distribution, tspan, algorithm, saveat, fsptol, odeatol, odertol, verbose,
                               ctx, comm, detect_separable)
    return _solve_fixed(model, initial_distribution, tspan, algorithm, saveat, odeatol, odertol, ctx, comm,
                        detect_separable)


class _Dist:
    """Distribution of one FSP vector over the ranks for a given (sharded or not) matrix."""

    def __init__(self, A: FspMatrixSparse, comm):
        from .parallel import ShardedVector, shard_bounds
        self.A, self.comm = A, comm
        info = A.shard_info()
        self.lo, self.hi, self.ng = info["row_lo"], info["row_hi"], info["n_global"]
        self.nloc = self.hi - self.lo
        self.R = A.nr
        self.u = ShardedVector(A)
        P = comm.nranks if comm is not None else 1
        cuts = shard_bounds(self.ng, P)
        self.counts = [cuts[r + 1] - cuts[r] for r in range(P)]
        self.displs = cuts[:-1]

    def load(self, p_full: DeviceVector, sinks: np.ndarray):
        if self.nloc:
            self.u.v.view(0, self.nloc).copy_from(p_full.view(self.lo, self.nloc))
        self.u.v.view(self.nloc, self.R).upload(sinks)

    def gather(self) -> DeviceVector:
        """p (all states) on every rank"""
        full = DeviceVector(self.A.ctx, self.ng)
        if self.comm is not None and self.comm.nranks > 1:
            self.comm.allgatherv(self.u.v, full, self.counts, self.displs)
        elif self.ng:
            full.copy_from(self.u.v.view(0, self.ng))
        return full

    def sinks(self) -> np.ndarray:
        return self.u.v.to_host(self.nloc, self.R)


def _solve_fixed(model, p0, tspan, ode_method, saveat, odeatol, odertol, ctx, comm=None, detect_separable=True):
    states0, vals0 = _initial(model, p0)
    space = StateSpaceSparse(model.stoich_matrix, states0, ctx=ctx)
    R = space.get_sink_count()
    # duplicates/negatives are dropped by the space; place p0 by lookup
    idx = space.lookup(states0)
    n = space.get_state_count()
    pv = np.zeros(n)
    pv[idx[idx > 0] - 1] = vals0[idx > 0]
    sv = _saveat_array(saveat, tspan)
    t_wall = time.perf_counter()
    p_dev = DeviceVector.from_host(space.ctx, pv)
    while True:
        A = FspMatrixSparse(space, model.propensities, parameters=model.parameters, comm=comm,
                            detect_separable=detect_separable)
        dist = _Dist(A, comm)
        dist.load(p_dev, np.zeros(R))
        seg = _Segment(A, odertol, odeatol, _method_code(ode_method))
        try:
            stats = seg.run(dist.u.v, tspan[0], tspan[1], saveat=sv, save_every_step=sv is None)
            break
        except L.SeparabilityError:          # a detected c(t) g(x) form broke down: repeat on the exact joint path
            if not detect_separable:
                raise
            detect_separable = False
            A.close()
    out = FspOutputSparse()
    states = space.get_states()
    for t, uu in zip(seg.saved_t, seg.saved_u):
        out.t.append(t)
        out.p.append(FspVectorSparse(states, uu[:n]))
        out.sinks.append(uu[n:].copy())
    out.stats = {"steps": stats.steps, "rejected": stats.rejected, "rhs_evals": stats.rhs_evals,
                 "launches": stats.launches, "adapts": 0, "wall_s": time.perf_counter() - t_wall, "final_states": n}
    return out


def _solve_adaptive(model, p0, tspan, alg, saveat, fsptol, odeatol, odertol, verbose, ctx, comm=None,
                    detect_separable=True):
    tstart, tend = min(tspan), max(tspan)
    sv = _saveat_array(saveat, tspan)
    adapter = alg.space_adapter
    method = _method_code(alg.ode_method)
    states0, vals0 = _initial(model, p0)
    space = StateSpaceSparse(model.stoich_matrix, states0, ctx=ctx)
    ctx = space.ctx
    R = space.get_sink_count()
    idx = space.lookup(states0)
    pv = np.zeros(space.get_state_count())
    pv[idx[idx > 0] - 1] = vals0[idx > 0]
    t_wall = time.perf_counter()
    br = {"expand": 0.0, "matrix": 0.0, "integrate": 0.0, "output": 0.0}      # wall-time breakdown (seconds)
    p = init_(space, adapter, DeviceVector.from_host(ctx, pv), tstart, fsptol)   # all states, replicated
    br["expand"] += time.perf_counter() - t_wall
    tnow = tstart
    sinks = np.zeros(R)
    tq = time.perf_counter()
    A = FspMatrixSparse(space, model.propensities, parameters=model.parameters, comm=comm,
                        detect_separable=detect_separable)
    br["matrix"] += time.perf_counter() - tq
    out = FspOutputSparse()
    tot = {"steps": 0, "rejected": 0, "rhs_evals": 0, "launches": 0, "adapts": 0, "matrix_builds": 1}
    while tnow < tend:
        n = space.get_state_count()
        dist = _Dist(A, comm)
        dist.load(p, sinks)
        seg = _Segment(A, odertol, odeatol, method)
        tq = time.perf_counter()
        try:
            stats = seg.run(dist.u.v, tnow, tend, saveat=sv, save_every_step=sv is None, event_slope=fsptol / tend)
            br["integrate"] += time.perf_counter() - tq
        except L.SeparabilityError:
            # a joint propensity detected as c(t) g(x) broke the product form at a time the integrator used: discard
            # this segment (p, sinks still hold its initial state) and repeat it on the exact joint path
            if not detect_separable:
                raise
            detect_separable = False
            A.close()
            A = FspMatrixSparse(space, model.propensities, parameters=model.parameters, comm=comm, detect_separable=False)
            tot["matrix_builds"] += 1
            tot["separability_fallbacks"] = tot.get("separability_fallbacks", 0) + 1
            continue
        for k in ("steps", "rejected", "rhs_evals", "launches"):
            tot[k] += getattr(stats, k)
        tq = time.perf_counter()
        states = space.get_states() if seg.saved_t else None
        for t, uu in zip(seg.saved_t, seg.saved_u):
            out.t.append(t)
            out.p.append(FspVectorSparse(states, uu[:n]))
            out.sinks.append(uu[n:].copy())
        br["output"] += time.perf_counter() - tq
        tnow = stats.t_final
        sinks = dist.sinks()
        if stats.event_hit and tnow < tend:
            dsinks = None
            if adapter.selective:                        # get_du!(du, integrator) (rstepadapters.jl:100-103)
                du = DeviceVector(ctx, dist.nloc + R)
                matvec_(du, tnow, A, dist.u.v)
                dsinks = du.to_host(dist.nloc, R)
            tq = time.perf_counter()
            p = adapt_(space, adapter, dist.gather(), sinks, tnow, tend, fsptol, dsinks=dsinks)
            br["expand"] += time.perf_counter() - tq
            tq = time.perf_counter()
            A_old = A                                    # incremental rebuild: only the new states are evaluated (H8)
            A = FspMatrixSparse(space, model.propensities, parameters=model.parameters, comm=comm, previous=A_old,
                                detect_separable=detect_separable)
            A_old.close()
            br["matrix"] += time.perf_counter() - tq
            tot["incremental_builds"] = tot.get("incremental_builds", 0) + (1 if A.incremental else 0)
            tot["adapts"] += 1
            tot["matrix_builds"] += 1
            if sinks.sum() >= tnow * fsptol / tend:      # re-arm the event (fspsolve.jl:179-181)
                sinks = sinks - EPS
            if verbose:
                print(f"t = {tnow:.2f}. Update state space. New size: {space.get_state_count()}.")
        else:
            tq = time.perf_counter()
            out.t.append(tnow)                           # final slice (duplicates the last saveat point when
            out.p.append(FspVectorSparse(space.get_states(), dist.gather().to_host()))   # tend is in saveat, Q3)
            out.sinks.append(sinks.copy())
            br["output"] += time.perf_counter() - tq
            tnow = tend
    tot["wall_s"] = time.perf_counter() - t_wall
    tot["breakdown_s"] = {k: round(v, 4) for k, v in br.items()}
    tot["final_states"] = space.get_state_count()
    out.stats = tot
    return out
