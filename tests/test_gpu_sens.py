"""GPU parity: fused sensitivity block matvec (K2) vs the oracle pinned on test/sensmat/telegraph.jl."""
import math

import numpy as np
import pytest

from fixtures import FSPMAT_THETA, SENS_THETA, TELEGRAPH_S, fspmat_propensities, sens_telegraph
from oracle.fspmatrix import OProp
from oracle.sensmatrix import OGrad, SensFspMatrixOracle
from oracle.statespace import StateSpaceOracle, StateSpaceOracleFast
from test_gpu_matvec import _relerr, _to_pkg_props

pytestmark = pytest.mark.gpu


def _to_pkg_grads(pkg, props, grads):
    out = []
    for a, g in zip(props, grads):
        if g.kind == "ti":
            out.append(pkg.propensitygrad(g.pardiffs))
        elif g.kind == "sep":
            out.append(pkg.propensitygrad_timevarying(a.tfactor, a.statefactor, g.tfactor_pardiffs, g.statefactor_pardiffs))
        else:
            out.append(pkg.propensitygrad_timevarying(g.pardiffs))
    return out


def _sensmodel(pkg, S, props, grads, pattern, theta):
    return pkg.CmeModelWithSensitivity(pkg.CmeModel(S, _to_pkg_props(pkg, props), theta), pattern,
                                       _to_pkg_grads(pkg, props, grads))


@pytest.mark.parametrize("t", [10.0, 20.0, 30.0, 100.0])
def test_sens_telegraph(pkg, ctx, t):  # test/sensmat/telegraph.jl:190-217
    props, grads, pattern, states = sens_telegraph()
    model = _sensmodel(pkg, TELEGRAPH_S, props, grads, pattern, SENS_THETA)
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, states)
    SA = pkg.ForwardSensFspMatrixSparse(model, sp)
    OS = SensFspMatrixOracle(StateSpaceOracleFast(TELEGRAPH_S, states), props, grads, pattern, SENS_THETA)
    n = SA.fspmatrix.rowcount
    v = np.ones(6 * n)
    v /= v.sum()
    out = np.empty_like(v)
    pkg.matvec_(out, t, SA, v)
    ref = OS.matvec(t, v)
    assert np.abs(out - ref).max() <= n * np.finfo(float).eps          # the reference test's own tolerance
    rng = np.random.default_rng(7)
    v = rng.random(6 * n)
    dv, do = pkg.DeviceVector.from_host(ctx, v), pkg.DeviceVector(ctx, 6 * n)
    pkg.matvec_(do, t, SA, dv)
    assert _relerr(do.to_host(), OS.matvec(t, v)) <= 1e-12
    ones = np.ones(6 * n)
    pkg.matvec_(out, t, SA, ones)
    assert abs(out.sum() / ones.sum()) <= n * 6 * np.finfo(float).eps  # telegraph.jl:34-43


def test_sens_poisson(pkg):  # test/sensmat/poisson.jl
    S = np.array([[1], [-1]]).T
    props = [OProp("ti", f=lambda x, p: p[0] + 0.0 * x[0]), OProp("ti", f=lambda x, p: p[1] + 0.0 * x[0])]
    one = lambda x, p: 1.0 + 0.0 * x[0]
    zero = lambda x, p: 0.0 * x[0]
    grads = [OGrad("ti", pardiffs=[one, zero]), OGrad("ti", pardiffs=[zero, one])]
    states = [[i] for i in range(1, 101)]
    model = _sensmodel(pkg, S, props, grads, np.eye(2, dtype=bool), [10.0, 5.0])
    SA = pkg.ForwardSensFspMatrixSparse(model, pkg.StateSpaceSparse(S, states))
    OS = SensFspMatrixOracle(StateSpaceOracleFast(S, states), props, grads, np.eye(2, dtype=bool), [10.0, 5.0])
    v = np.ones(3 * SA.fspmatrix.rowcount)
    out = np.empty_like(v)
    pkg.matvec_(out, 0.0, SA, v)
    assert out.sum() == pytest.approx(0.0, abs=1e-9)
    assert _relerr(out, OS.matvec(0.0, v)) <= 1e-12


def test_sens_joint(pkg):
    f = lambda t, x, p: (1.0 + 0.5 * math.sin(t)) * p[1] * x[1]
    df = lambda t, x, p: (1.0 + 0.5 * math.sin(t)) * x[1]
    zero3 = lambda t, x, p: 0.0 * x[0]
    zero = lambda x, p: 0.0 * x[0]
    props = fspmat_propensities("ti")
    props[1] = OProp("joint", f=f)
    grads = [OGrad("ti", pardiffs=[lambda x, p: 1.0 * x[0], zero, zero, zero]),
             OGrad("joint", pardiffs=[zero3, df, zero3, zero3]),
             OGrad("ti", pardiffs=[zero, zero, lambda x, p: 1.0 * x[1], zero]),
             OGrad("ti", pardiffs=[zero, zero, zero, lambda x, p: 1.0 * x[2]])]
    osp = StateSpaceOracle(TELEGRAPH_S, [1, 0, 0])
    osp.expand(6)
    OS = SensFspMatrixOracle(osp, props, grads, np.eye(4, dtype=bool), FSPMAT_THETA)
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    sp.expand_(6)
    SA = pkg.ForwardSensFspMatrixSparse(_sensmodel(pkg, TELEGRAPH_S, props, grads, np.eye(4, dtype=bool), FSPMAT_THETA), sp)
    rng = np.random.default_rng(3)
    v = rng.random(5 * SA.fspmatrix.rowcount)
    out = np.empty_like(v)
    for t in (0.3, 1.7):
        pkg.matvec_(out, t, SA, v)
        assert _relerr(out, OS.matvec(t, v)) <= 1e-12


def test_sens_solve(pkg):
    """test/test_sensfsp.jl (a smoke test in the reference) + the checks the reference lacks: the probability block
    equals the plain FSP solve, d/dtheta of total mass is zero, sensitivities match central finite differences."""
    import math
    props, grads, pattern, _ = sens_telegraph()
    model = _sensmodel(pkg, TELEGRAPH_S, props, grads, pattern, SENS_THETA)
    ic = pkg.forwardsens_initial_condition([[1, 0, 0]], [1.0], [[0.0] for _ in range(5)])
    alg = pkg.AdaptiveForwardSensFspSparse(ode_method=pkg.NativeRK45(), space_adapter=pkg.ForwardSensRStepAdapter(10, 10, True))
    touts = [10.0, 40.0]
    sol = pkg.solve(model, ic, (0.0, 40.0), alg, saveat=touts, fsptol=1e-8, odeatol=1e-12, odertol=1e-8)
    assert sol.stats["adapts"] >= 1
    assert isinstance(sol[0], pkg.ForwardSensFspOutputSliceSparse) and len(sol[0].S) == 5
    for k in range(2):
        assert sol.p[k].sum() + sol.sinks[k].sum() == pytest.approx(1.0, abs=1e-9)
        for ip in range(5):
            assert abs(sol.S[k][ip].values.sum() + sol.dsinks[k][ip].sum()) <= 1e-8
    # probability block == plain solve
    plain_alg = pkg.AdaptiveFspSparse(pkg.NativeRK45(), pkg.RStepAdapter(10, 10, True))

    def plain(theta):
        m = pkg.CmeModel(TELEGRAPH_S, _to_pkg_props(pkg, props), theta)
        return pkg.solve(m, pkg.FspVectorSparse([[1, 0, 0]], [1.0]), (0.0, 40.0), plain_alg, saveat=touts, fsptol=1e-8,
                         odeatol=1e-12, odertol=1e-8)

    def on_states(fv, states):
        d = fv.state2idx
        return np.array([fv.values[d[tuple(s)] - 1] if tuple(s) in d else 0.0 for s in states.tolist()])
    base = plain(SENS_THETA)
    st = sol.p[1].states
    assert np.abs(on_states(base.p[1], st) - sol.p[1].values).max() < 1e-7
    for ip, h in [(2, 1e-4), (3, 1e-5), (4, 1e-3)]:                  # lambda, gamma, L (time factor parameter)
        tp, tm = list(SENS_THETA), list(SENS_THETA)
        tp[ip] += h
        tm[ip] -= h
        fd = (on_states(plain(tp).p[1], st) - on_states(plain(tm).p[1], st)) / (2 * h)
        s = sol.S[1][ip].values
        assert np.abs(fd - s).max() <= 2e-4 * max(np.abs(s).max(), 1e-3), (ip, np.abs(fd - s).max(), np.abs(s).max())


def test_sens_solve_bdf_matches_explicit(pkg):
    """test/test_sensfsp.jl passes CVODE_BDF(GMRES); the native BDF on the block system agrees with the explicit one."""
    props, grads, pattern, _ = sens_telegraph()
    model = _sensmodel(pkg, TELEGRAPH_S, props, grads, pattern, SENS_THETA)
    ic = pkg.forwardsens_initial_condition([[1, 0, 0]], [1.0], [[0.0] for _ in range(5)])
    out = {}
    for name, m, rt in (("bdf", None, 1e-7), ("rk", pkg.NativeRK45(), 1e-8)):
        alg = pkg.AdaptiveForwardSensFspSparse(ode_method=m, space_adapter=pkg.ForwardSensRStepAdapter(10, 10, False))
        out[name] = pkg.solve(model, ic, (0.0, 30.0), alg, saveat=[30.0], fsptol=1e-8, odeatol=1e-12, odertol=rt)
    a, b = out["bdf"], out["rk"]
    assert a.stats["adapts"] >= 1

    def on(fv, states):
        d = fv.state2idx
        return np.array([fv.values[d[tuple(s)] - 1] if tuple(s) in d else 0.0 for s in states.tolist()])
    st = b.p[0].states
    assert np.abs(on(a.p[0], st) - b.p[0].values).max() < 2e-6
    for ip in range(5):
        sb = b.S[0][ip].values
        assert np.abs(on(a.S[0][ip], st) - sb).max() <= 2e-5 * max(np.abs(sb).max(), 1e-2), ip


def test_sens_hog1p_matrix_vs_oracle(pkg, ctx):
    """BASELINE.json config 3: Hog1p (examples/hog1p.jl:33-82, NS=6, R=13, P=14) with the forward-sensitivity matrix
    (sensfspmatrixsparse.jl:31-142): 14 (reaction, parameter) entries, separable reaction 2 carrying two parameters in
    its time factor.  >= 10^4 states, 1e-12 relative, device-resident and host-buffer entry points."""
    th = list(pkg.workloads.HOG1P_THETA)
    th[2] = 3.2e4                                     # a: signal on, so d c / d(k10, a) are non-trivial
    model = pkg.workloads.hog1p_sens_model(th)
    cm = model.cmemodel
    sp = pkg.StateSpaceSparse(cm.stoich_matrix, [1, 0, 0, 0, 0, 0], ctx=ctx)
    sp.expand_(105)                                   # gene states x (nuclear, cytoplasmic) RNA counts: ~4 L^2 / 2 states
    osp = StateSpaceOracleFast(cm.stoich_matrix, [1, 0, 0, 0, 0, 0])
    osp.expand(105)
    n = sp.get_state_count()
    assert n >= 10000 and np.array_equal(sp.get_states(), osp.states_array())
    SA = pkg.ForwardSensFspMatrixSparse(model, sp)
    assert len(SA.entries) == 14 and SA.parameter_count == 14
    OS = SensFspMatrixOracle(osp, cm.propensities, model.propensity_gradients, model.gradient_sparsity_patterns, th)
    N = SA.fspmatrix.rowcount
    rng = np.random.default_rng(5)
    v = rng.random(15 * N)
    dv, do = pkg.DeviceVector.from_host(ctx, v), pkg.DeviceVector(ctx, 15 * N)
    out = np.empty_like(v)
    for t in (0.0, 45.0, 120.0, 900.0):               # k10 - a Hog1p(t) changes sign between these times
        ref = OS.matvec(t, v)
        pkg.matvec_(do, t, SA, dv)
        got = do.to_host()
        for b in range(15):                           # per block: every sensitivity block to 1e-12 of its own scale
            sl = slice(b * N, (b + 1) * N)
            assert _relerr(got[sl], ref[sl]) <= 1e-12, (t, b)
        pkg.matvec_(out, t, SA, v)                    # host-buffer entry point
        assert np.array_equal(out, got)
    ones = np.ones(15 * N)
    pkg.matvec_(out, 120.0, SA, ones)                 # column sums of A and of every dA vanish (telegraph.jl:34-43)
    assert abs(out.sum()) <= 1e-9 * np.abs(out).sum()


def test_sens_m3d_matrix_vs_oracle(pkg, ctx):
    """The bench's sensitivity workload (M-3D, P = 6, separable death-3) at 1.8e5 states vs the oracle."""
    model = pkg.workloads.m3d_sens_model()
    cm = model.cmemodel
    sp = pkg.StateSpaceSparse(cm.stoich_matrix, [0, 0, 0], ctx=ctx)
    sp.expand_(100)
    osp = StateSpaceOracleFast(cm.stoich_matrix, [0, 0, 0])
    osp.expand(100)
    SA = pkg.ForwardSensFspMatrixSparse(model, sp)
    OS = SensFspMatrixOracle(osp, cm.propensities, model.propensity_gradients, model.gradient_sparsity_patterns, cm.parameters)
    N = SA.fspmatrix.rowcount
    v = np.random.default_rng(9).random(7 * N)
    dv, do = pkg.DeviceVector.from_host(ctx, v), pkg.DeviceVector(ctx, 7 * N)
    for t in (0.0, 2.5):
        pkg.matvec_(do, t, SA, dv)
        got, ref = do.to_host(), OS.matvec(t, v)
        for b in range(7):
            assert _relerr(got[b * N:(b + 1) * N], ref[b * N:(b + 1) * N]) <= 1e-12, (t, b)


def _on_states(states_from, values, states_to):
    d = {tuple(s): v for s, v in zip(states_from.tolist(), values)}
    return np.array([d.get(tuple(s), 0.0) for s in states_to.tolist()])


@pytest.mark.parametrize("method", ["bdf", "rk45"])
def test_sens_solve_vs_oracle(pkg, method):
    """The forward-sensitivity solve loop against its ORACLE (oracle/senssolve.py, restating
    forwardsenscmesparse.jl:99-215 and pinned by analytic birth-death sensitivities) on the reference's own telegraph
    sensitivity fixture (test/test_sensfsp.jl, test/sensmat/telegraph.jl): probabilities, all five sensitivity blocks,
    sinks and d(sinks)/d(theta)."""
    from oracle.senssolve import ForwardSensRStepAdapterOracle, solve_sens
    props, grads, pattern, _ = sens_telegraph()
    model = _sensmodel(pkg, TELEGRAPH_S, props, grads, pattern, SENS_THETA)
    ic = pkg.forwardsens_initial_condition([[1, 0, 0]], [1.0], [[0.0] for _ in range(5)])
    ode = None if method == "bdf" else pkg.NativeRK45()
    alg = pkg.AdaptiveForwardSensFspSparse(ode_method=ode, space_adapter=pkg.ForwardSensRStepAdapter(10, 10, True))
    touts = [5.0, 20.0, 40.0]
    rt = 1e-7 if method == "bdf" else 1e-8
    sol = pkg.solve(model, ic, (0.0, 40.0), alg, saveat=touts, fsptol=1e-8, odeatol=1e-13, odertol=rt)
    ref = solve_sens(TELEGRAPH_S, props, grads, pattern, SENS_THETA, [[1, 0, 0]], [1.0], [[0.0]] * 5, (0.0, 40.0),
                     ForwardSensRStepAdapterOracle(10, 10, True), saveat=touts, fsptol=1e-8, odeatol=1e-13, odertol=1e-9)
    assert len(sol) == len(ref["t"]) == len(touts) + 1
    tol = 2e-5 if method == "bdf" else 2e-6
    for k in range(len(touts)):
        st = ref["states"][k]
        assert np.abs(_on_states(sol.p[k].states, sol.p[k].values, st) - ref["p"][k]).max() < tol
        assert np.abs(sol.sinks[k] - ref["sinks"][k]).max() < tol
        for ip in range(5):
            s_ref = ref["S"][k][ip]
            scale = max(np.abs(s_ref).max(), 1e-2)
            assert np.abs(_on_states(sol.S[k][ip].states, sol.S[k][ip].values, st) - s_ref).max() <= tol * scale, (k, ip)
            assert np.abs(sol.dsinks[k][ip] - ref["dsinks"][k][ip]).max() <= tol * scale, (k, ip)


def test_sens_incremental_rebuild_after_adapt(pkg, ctx):
    """The sensitivity matrix after a prune + expand built from its predecessor -- propensities and parameter derivatives
    evaluated on the appended states only, the survivors' rows carried over on the device -- is bit-for-bit the matrix
    built from scratch (Hog1p, 14 entries, two adapt rounds), falls back when the pattern's model changes, and the
    adaptive sensitivity solve uses it (forwardsenscmesparse.jl:140 rebuilds from scratch)."""
    rng = np.random.default_rng(12)
    th = list(pkg.workloads.HOG1P_THETA)
    th[2] = 3.2e4
    model = pkg.workloads.hog1p_sens_model(th)
    cm = model.cmemodel
    sp = pkg.StateSpaceSparse(cm.stoich_matrix, [1, 0, 0, 0, 0, 0], ctx=ctx)
    sp.expand_(60)
    assert sp.get_state_count() >= 2048                      # (below INCREMENTAL_MIN_STATES matrices are rebuilt from scratch)
    SA = pkg.ForwardSensFspMatrixSparse(model, sp)
    assert not SA.incremental
    for rnd in range(2):
        n = sp.get_state_count()
        drop = np.sort(rng.choice(np.arange(2, n + 1), size=n // 6, replace=False))
        sp.deleteat_(drop)
        sp.expand_(3 + rnd)
        SB = pkg.ForwardSensFspMatrixSparse(model, sp, previous=SA)
        assert SB.incremental and SB.fspmatrix.incremental and 0 < SB.fspmatrix.new_state_count < sp.get_state_count()
        SF = pkg.ForwardSensFspMatrixSparse(model, sp)       # from scratch (re-marks the space)
        assert not SF.incremental and SB.stats() == SF.stats()
        N = SB.fspmatrix.rowcount
        v = rng.random(15 * N)
        a, b = np.empty_like(v), np.empty_like(v)
        for t in (0.0, 45.0, 900.0):
            pkg.matvec_(a, t, SB, v)
            pkg.matvec_(b, t, SF, v)
            assert np.array_equal(a, b), (rnd, t)
        SA.close()
        SB.close()
        SA = SF
    # another model object (other gradient closures) -> full build, never a stale carry-over
    other = pkg.workloads.hog1p_sens_model(th)
    sp.expand_(1)
    assert not pkg.ForwardSensFspMatrixSparse(other, sp, previous=SA).incremental
    # the adaptive sensitivity solve goes through it: identical results with and without the carry-over
    import numcme_jl_b200.fspmatrix as FM
    props, grads, pattern, _ = sens_telegraph()
    tm = _sensmodel(pkg, TELEGRAPH_S, props, grads, pattern, SENS_THETA)
    ic = pkg.forwardsens_initial_condition([[1, 0, 0]], [1.0], [[0.0] for _ in range(5)])
    alg = pkg.AdaptiveForwardSensFspSparse(ode_method=pkg.NativeRK45(), space_adapter=pkg.ForwardSensRStepAdapter(10, 10, True))
    saved = FM.INCREMENTAL_MIN_STATES
    try:
        FM.INCREMENTAL_MIN_STATES = 0
        s1 = pkg.solve(tm, ic, (0.0, 40.0), alg, saveat=[40.0], fsptol=1e-8, odeatol=1e-12, odertol=1e-8)
        assert s1.stats["adapts"] >= 1 and s1.stats["incremental_builds"] >= 1
        FM.INCREMENTAL_MIN_STATES = 1 << 60
        s0 = pkg.solve(tm, ic, (0.0, 40.0), alg, saveat=[40.0], fsptol=1e-8, odeatol=1e-12, odertol=1e-8)
        assert s0.stats["incremental_builds"] == 0 and s0.stats["adapts"] == s1.stats["adapts"]
        assert np.array_equal(s0.p[0].values, s1.p[0].values)
        for x, y in zip(s0.S[0], s1.S[0]):
            assert np.array_equal(x.values, y.values)
    finally:
        FM.INCREMENTAL_MIN_STATES = saved
