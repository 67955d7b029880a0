"""GPU parity for the transient solve path: native device integrator, sink event, prune/expand adapters,
saveat -- against the oracle loop (scipy, tight tolerances), analytic solutions, and the reference's own
solver tests (test/test_solver.jl)."""
import math

import numpy as np
import pytest
from scipy.stats import poisson

from fixtures import FSPMAT_THETA, TELEGRAPH_S, TOGGLE_S, fspmat_propensities
from oracle.solve import RStepAdapterOracle, SelectiveRStepAdapterOracle, solve_adaptive, solve_fixed
from oracle.statespace import StateSpaceOracleFast
from test_gpu_matvec import _to_pkg_props

pytestmark = pytest.mark.gpu


def _align(states_a, vals_a, states_b, vals_b):
    """values of two sparse vectors on the union of their supports"""
    d = {}
    for s, v in zip(map(tuple, states_a.tolist()), vals_a):
        d[s] = [v, 0.0]
    for s, v in zip(map(tuple, states_b.tolist()), vals_b):
        d.setdefault(s, [0.0, 0.0])[1] = v
    arr = np.array(list(d.values()))
    return arr[:, 0], arr[:, 1]


def test_fixed_space_solve(pkg):  # test/test_solver.jl:55-77
    model = pkg.CmeModel(TELEGRAPH_S, _to_pkg_props(pkg, fspmat_propensities("tv")), FSPMAT_THETA)
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    sp.expand_(20)
    p0 = pkg.FspVectorSparse.from_pairs(sp, [([1, 0, 0], 1.0)])
    touts = np.arange(0.0, 121.0, 20.0)
    sol = pkg.solve(model, p0, (0.0, 120.0), None, odertol=1e-4, odeatol=1e-14, saveat=touts)   # None = native BDF
    assert len(sol) == len(touts)
    assert isinstance(sol[0], pkg.FspOutputSliceSparse)
    for p, s in zip(sol.p, sol.sinks):
        assert p.sum() + s.sum() == pytest.approx(1.0, abs=1.5e-8)      # the reference test's own `isapprox` tolerance
    dense = pkg.solve(model, p0, (0.0, 120.0), None, odertol=1e-4, odeatol=1e-14)   # saveat = []: every step
    assert len(dense) == dense.stats["steps"] + 1
    # values vs the oracle at tight tolerance
    tight = pkg.solve(model, p0, (0.0, 120.0), pkg.NativeRK45(), odertol=1e-9, odeatol=1e-13, saveat=touts)
    ref = solve_fixed(TELEGRAPH_S, fspmat_propensities("tv"), FSPMAT_THETA, sp.get_states(), p0.values, (0.0, 120.0),
                      saveat=touts, odeatol=1e-13, odertol=1e-10, method="LSODA")
    for k in range(len(touts)):
        assert np.abs(tight.p[k].values - ref["p"][k]).max() < 1e-8
        assert np.abs(tight.sinks[k] - ref["sinks"][k]).max() < 1e-8


@pytest.mark.parametrize("selective", [False, True])
def test_adaptive_solve_reference_tests(pkg, selective):  # test/test_solver.jl:80-99
    ada = (pkg.SelectiveRStepAdapter if selective else pkg.RStepAdapter)(10, 10, True)
    alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=ada)
    k01, k10, lam, gam = FSPMAT_THETA
    tv = lambda t, p: max(0.0, 1.0 - math.sin(math.pi * t / 2))
    m1 = pkg.CmeModel(TELEGRAPH_S, [pkg.propensity(lambda x, p: k01 * x[0]), pkg.propensity(lambda x, p: k10 * x[1], tv),
                                    pkg.propensity(lambda x, p: lam * x[1]), pkg.propensity(lambda x, p: gam * x[2])], [])
    m2 = pkg.CmeModel(TELEGRAPH_S, _to_pkg_props(pkg, fspmat_propensities("tv")), FSPMAT_THETA)
    p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
    touts = np.arange(0.0, 121.0, 20.0)
    s1 = pkg.solve(m1, p0, (0.0, 120.0), alg, odertol=1e-4, odeatol=1e-14, saveat=touts)
    s2 = pkg.solve(m2, p0, (0.0, 120.0), alg, odertol=1e-4, odeatol=1e-14, saveat=touts)
    assert s1.stats["adapts"] >= 1
    for sol in (s1, s2):
        for p, s in zip(sol.p, sol.sinks):
            assert p.sum() + s.sum() == pytest.approx(1.0, abs=1.5e-8)  # the reference test's own `isapprox` tolerance
        assert sol.sinks[-1].sum() <= 1e-6 * 1.001
    assert len(s1) == len(s2)
    for a, b in zip(s1.p, s2.p):
        assert np.array_equal(a.states, b.states)
        assert np.abs(a.values - b.values).sum() <= 1e-14


def test_adaptive_values_vs_oracle(pkg):
    """Knife-edge pruning may give different fringe sets: compare values on the union."""
    alg = pkg.AdaptiveFspSparse(ode_method=pkg.NativeRK45(), space_adapter=pkg.RStepAdapter(5, 10, True))
    model = pkg.workloads.telegraph_model()
    p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
    touts = [50.0, 150.0, 300.0]
    sol = pkg.solve(model, p0, (0.0, 300.0), alg, saveat=touts, fsptol=1e-6, odeatol=1e-12, odertol=1e-8)
    ref = solve_adaptive(TELEGRAPH_S, model.propensities, model.parameters, [[1, 0, 0]], [1.0], (0.0, 300.0),
                         RStepAdapterOracle(5, 10, True), saveat=touts, fsptol=1e-6, odeatol=1e-13, odertol=1e-10,
                         method="LSODA")
    for k in range(len(touts)):
        assert sol.t[k] == pytest.approx(touts[k])
        a, b = _align(sol.p[k].states, sol.p[k].values, ref["states"][k], ref["p"][k])
        assert np.abs(a - b).max() < 2e-6          # both are within fsptol of the exact CME solution
        assert sol.p[k].sum() + sol.sinks[k].sum() == pytest.approx(1.0, abs=1e-9)


def test_birth_death_poisson(pkg):
    S = np.array([[1], [-1]]).T
    lam, gam = 10.0, 0.5
    model = pkg.CmeModel(S, [pkg.propensity(lambda x, p: p[0] + 0.0 * x[0]), pkg.propensity(lambda x, p: p[1] * x[0])],
                         [lam, gam])
    alg = pkg.AdaptiveFspSparse(ode_method=pkg.NativeRK45(), space_adapter=pkg.RStepAdapter(10, 10, False))
    sol = pkg.solve(model, pkg.FspVectorSparse([[0]], [1.0]), (0.0, 4.0), alg, saveat=[1.0, 4.0], fsptol=1e-8,
                    odeatol=1e-13, odertol=1e-9)
    for k, t in enumerate([1.0, 4.0]):
        mu = lam / gam * (1 - math.exp(-gam * t))
        assert np.abs(sol.p[k].values - poisson.pmf(sol.p[k].states[:, 0], mu)).max() < 1e-7


def test_toggle_variants_agree(pkg):  # examples/toggleswitch_fsp_variants.jl (shortened horizon)
    touts = np.arange(0.0, 7201.0, 600.0)
    p0 = pkg.FspVectorSparse([[0, 0]], [1.0])
    res = {}
    for name, sep, ada in [("full_sep", True, pkg.RStepAdapter(20, 5, True)), ("sel_sep", True, pkg.SelectiveRStepAdapter(20, 5, True)),
                           ("full_joint", False, pkg.RStepAdapter(20, 5, True))]:
        model = pkg.workloads.toggle_model(separable=sep)
        res[name] = pkg.solve(model, p0, (0.0, 7200.0), pkg.AdaptiveFspSparse(pkg.NativeRK45(), ada), saveat=touts,
                              odertol=1e-6, odeatol=1e-14)
    for name, sol in res.items():
        assert len(sol) == len(touts) + 1                      # + the final slice (duplicate of tend, as the reference)
        for p, s in zip(sol.p, sol.sinks):
            assert p.sum() + s.sum() == pytest.approx(1.0, abs=1e-9)
    for k in range(len(touts)):
        a, b = _align(res["full_sep"].p[k].states, res["full_sep"].p[k].values, res["full_joint"].p[k].states,
                      res["full_joint"].p[k].values)
        assert np.abs(a - b).max() < 1e-7
        a, b = _align(res["full_sep"].p[k].states, res["full_sep"].p[k].values, res["sel_sep"].p[k].states,
                      res["sel_sep"].p[k].values)
        assert np.abs(a - b).max() < 5e-6


def test_prune_by_mass_matches_oracle(pkg, ctx):
    rng = np.random.default_rng(11)
    for n, strict in [(10, False), (1000, False), (5000, True), (70001, False)]:
        osp = StateSpaceOracleFast(TOGGLE_S, [0, 0])
        L = int(math.sqrt(2 * n)) + 2
        osp.expand(L)
        m = osp.get_state_count()
        p = rng.random(m) ** 8
        p[rng.integers(0, m, size=m // 10)] = 0.0               # ties
        p[rng.integers(0, m, size=3)] = -1e-18                  # tiny negatives, as an integrator leaves them
        p /= p.sum() / (1 - 3e-7)
        ada = (SelectiveRStepAdapterOracle if strict else RStepAdapterOracle)(1, 1, True)
        ids = ada.drop_ids(p, 0.5, 1.0, 1e-6)
        sp = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
        sp.expand_(L)
        dp = pkg.DeviceVector.from_host(ctx, p)
        dropped = sp.prune_by_mass_(dp, 1.0 - 0.5 * 1e-6, strict)
        assert abs(dropped - len(ids)) <= 2                     # knife-edge: cumulative rounding may shift the count
        if dropped == len(ids):
            osp.deleteat(ids)
            assert np.array_equal(sp.get_states(), osp.states_array())
            assert np.array_equal(sp.get_sink_connectivity().astype(np.int64), osp.sink_connectivity_array())
            q = pkg.DeviceVector(ctx, sp.get_state_count())
            sp.compact_vector(dp, q)
            assert np.array_equal(q.to_host(), np.delete(p, ids - 1))
        assert p.sum() - np.sort(p)[:dropped].sum() >= 1.0 - 0.5e-6 - 1e-12


def test_vector_ops(pkg, ctx):
    rng = np.random.default_rng(0)
    n = 100003
    a, b, c = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    da, db, dc = (pkg.DeviceVector.from_host(ctx, v) for v in (a, b, c))
    assert da.sum() == pytest.approx(a.sum(), rel=1e-12, abs=1e-9)
    assert da.dot(db) == pytest.approx(a @ b, rel=1e-12, abs=1e-9)
    assert da.sum(7, 1000) == pytest.approx(a[7:1007].sum(), rel=1e-12, abs=1e-12)
    out = pkg.DeviceVector(ctx, n)
    out.lincomb([2.0, -1.0, 0.5], [da, db, dc])
    assert np.allclose(out.to_host(), 2 * a - b + 0.5 * c, rtol=1e-15, atol=1e-15)
    out.axpy(3.0, da)
    assert np.allclose(out.to_host(), 5 * a - b + 0.5 * c, rtol=1e-14, atol=1e-14)
    out.scale(0.5)
    assert np.allclose(out.to_host(), 0.5 * (5 * a - b + 0.5 * c), rtol=1e-14, atol=1e-14)
    w = da.wrms(db, dc, 1e-6, 1e-3)
    assert w == pytest.approx(np.sqrt(np.mean((a / (1e-6 + 1e-3 * np.maximum(np.abs(b), np.abs(c)))) ** 2)), rel=1e-12)
    out.residuals(da, db, dc, 1e-6, 1e-3)                         # calculate_residuals! (Julia Broadcast surface)
    assert np.allclose(out.to_host(), a / (1e-6 + 1e-3 * np.maximum(np.abs(b), np.abs(c))), rtol=1e-14, atol=0)    # the kernel contracts atol + rtol*max into an fma
    out.shift(0.25)
    assert np.allclose(out.to_host(), a / (1e-6 + 1e-3 * np.maximum(np.abs(b), np.abs(c))) + 0.25, rtol=1e-14, atol=0)
    assert not da.any_nonfinite()
    a2 = a.copy()
    a2[5] = np.nan
    assert pkg.DeviceVector.from_host(ctx, a2).any_nonfinite()
    assert np.array_equal(da.view(10, 20).to_host(), a[10:30])
    assert da.sum() == da.sum()                                  # deterministic


# ---- the stiff integrator (BDF/NDF + GMRES, method 1) against the same references ------------------------------
def test_bdf_fixed_space(pkg):
    model = pkg.CmeModel(TELEGRAPH_S, _to_pkg_props(pkg, fspmat_propensities("tv")), FSPMAT_THETA)
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    sp.expand_(20)
    p0 = pkg.FspVectorSparse.from_pairs(sp, [([1, 0, 0], 1.0)])
    touts = np.arange(0.0, 121.0, 20.0)
    sol = pkg.solve(model, p0, (0.0, 120.0), pkg.NativeBDF(), odertol=1e-4, odeatol=1e-14, saveat=touts)
    assert len(sol) == len(touts)
    print("bdf stats", sol.stats, [float(p.sum() + s.sum() - 1.0) for p, s in zip(sol.p, sol.sinks)])
    for p, s in zip(sol.p, sol.sinks):
        # the inexact (Krylov) linear solve would drift; the invariant projection of the BDF keeps total mass
        assert p.sum() + s.sum() == pytest.approx(1.0, abs=1e-9)
    tight = pkg.solve(model, p0, (0.0, 120.0), pkg.NativeBDF(), odertol=1e-8, odeatol=1e-13, saveat=touts)
    ref = solve_fixed(TELEGRAPH_S, fspmat_propensities("tv"), FSPMAT_THETA, sp.get_states(), p0.values, (0.0, 120.0),
                      saveat=touts, odeatol=1e-13, odertol=1e-10, method="LSODA")
    for k in range(len(touts)):
        assert np.abs(tight.p[k].values - ref["p"][k]).max() < 5e-7
        assert np.abs(tight.sinks[k] - ref["sinks"][k]).max() < 5e-7
    loose = pkg.solve(model, p0, (0.0, 120.0), pkg.NativeBDF(), odertol=1e-4, odeatol=1e-8, saveat=touts)
    for k in range(len(touts)):
        assert np.abs(loose.p[k].values - ref["p"][k]).max() < 2e-3


def test_bdf_adaptive_and_poisson(pkg):
    import math
    S = np.array([[1], [-1]]).T
    lam, gam = 10.0, 0.5
    model = pkg.CmeModel(S, [pkg.propensity(lambda x, p: p[0] + 0.0 * x[0]), pkg.propensity(lambda x, p: p[1] * x[0])],
                         [lam, gam])
    alg = pkg.AdaptiveFspSparse(ode_method=pkg.NativeBDF(), space_adapter=pkg.RStepAdapter(10, 10, False))
    sol = pkg.solve(model, pkg.FspVectorSparse([[0]], [1.0]), (0.0, 4.0), alg, saveat=[1.0, 4.0], fsptol=1e-8,
                    odeatol=1e-13, odertol=1e-8)
    assert sol.stats["adapts"] >= 1
    for k, t in enumerate([1.0, 4.0]):
        mu = lam / gam * (1 - math.exp(-gam * t))
        assert np.abs(sol.p[k].values - poisson.pmf(sol.p[k].states[:, 0], mu)).max() < 1e-6
    # telegraph adaptive (examples/telegraph_cme.jl) : BDF vs explicit integrator
    tm = pkg.workloads.telegraph_model()
    p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
    a = pkg.solve(tm, p0, (0.0, 300.0), pkg.AdaptiveFspSparse(pkg.NativeBDF(), pkg.RStepAdapter(5, 10, True)),
                  saveat=[150.0, 300.0], odeatol=1e-12, odertol=1e-7)
    b = pkg.solve(tm, p0, (0.0, 300.0), pkg.AdaptiveFspSparse(pkg.NativeRK45(), pkg.RStepAdapter(5, 10, True)),
                  saveat=[150.0, 300.0], odeatol=1e-12, odertol=1e-8)
    for k in range(2):
        x, y = _align(a.p[k].states, a.p[k].values, b.p[k].states, b.p[k].values)
        assert np.abs(x - y).max() < 3e-6
        assert a.p[k].sum() + a.sinks[k].sum() == pytest.approx(1.0, abs=1e-7)     # pruning drops <= fsptol of mass


def test_bdf_stiff_is_cheaper_than_explicit(pkg):
    """A stiff birth-death chain (death rate up to 2000): the BDF needs far fewer RHS evaluations than DP5."""
    S = np.array([[1], [-1]]).T
    model = pkg.CmeModel(S, [pkg.propensity(lambda x, p: p[0] + 0.0 * x[0]), pkg.propensity(lambda x, p: p[1] * x[0])],
                         [20.0, 1.0])
    sp = pkg.StateSpaceSparse(S, [[0]])
    sp.expand_(2000)
    p0 = pkg.FspVectorSparse.from_pairs(sp, [([0], 1.0)])
    r = {}
    for name, m in (("bdf", pkg.NativeBDF()), ("rk", pkg.NativeRK45())):
        r[name] = pkg.solve(model, p0, (0.0, 5.0), m, saveat=[5.0], odertol=1e-5, odeatol=1e-10)
    mu = 20.0 * (1 - np.exp(-5.0))
    for name in r:
        st = r[name].p[0].states[:, 0]
        assert np.abs(r[name].p[0].values - poisson.pmf(st, mu)).max() < 1e-4, name
    assert r["bdf"].stats["rhs_evals"] < 0.25 * r["rk"].stats["rhs_evals"], (r["bdf"].stats, r["rk"].stats)


# ---- fused-step BDF (csrc/bdf_fused.cu, one kernel per step attempt) against the launch-per-operation BDF ---------
# 41 / 801 states: one CTA, work vectors in shared memory; 1001 states: one CTA, vectors in global memory
@pytest.mark.parametrize("levels", [20, 400, 500])
def test_bdf_fused_matches_classic_fixed(pkg, levels):
    model = pkg.CmeModel(TELEGRAPH_S, _to_pkg_props(pkg, fspmat_propensities("tv")), FSPMAT_THETA)
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    sp.expand_(levels)
    p0 = pkg.FspVectorSparse.from_pairs(sp, [([1, 0, 0], 1.0)])
    touts = np.arange(0.0, 121.0, 20.0)
    ref = solve_fixed(TELEGRAPH_S, fspmat_propensities("tv"), FSPMAT_THETA, sp.get_states(), p0.values, (0.0, 120.0),
                      saveat=touts, odeatol=1e-13, odertol=1e-10, method="LSODA")
    out = {}
    for name, m in (("classic", pkg.NativeBDFClassic()), ("fused", pkg.NativeBDFFused())):
        out[name] = pkg.solve(model, p0, (0.0, 120.0), m, odertol=1e-8, odeatol=1e-13, saveat=touts)
        for k in range(len(touts)):
            assert np.abs(out[name].p[k].values - ref["p"][k]).max() < 5e-7, name
            assert np.abs(out[name].sinks[k] - ref["sinks"][k]).max() < 5e-7, name
            assert out[name].p[k].sum() + out[name].sinks[k].sum() == pytest.approx(1.0, abs=1e-9)
    f, c = out["fused"].stats, out["classic"].stats
    print("fused", f, "classic", c)
    assert f["launches"] < 0.2 * c["launches"]                         # one launch per step attempt (+ dense output)
    assert abs(f["steps"] - c["steps"]) <= 0.1 * c["steps"] + 2        # same controller, same algorithm
    # every-step output and bitwise reproducibility of the fused path
    a = pkg.solve(model, p0, (0.0, 120.0), pkg.NativeBDFFused(), odertol=1e-4, odeatol=1e-14)
    b = pkg.solve(model, p0, (0.0, 120.0), pkg.NativeBDFFused(), odertol=1e-4, odeatol=1e-14)
    assert len(a) == a.stats["steps"] + 1 and a.t == b.t
    assert all(np.array_equal(x.values, y.values) for x, y in zip(a.p, b.p))


def test_bdf_fused_cooperative_grid(pkg):
    """20 301 states (examples/2dstate_exploration.jl as shipped): 20 CTAs, grid-wide barriers; and a 2-D model large
    enough for the full co-resident grid.  Checked against the explicit integrator and the classic BDF."""
    model = pkg.workloads.m2d_model()
    for levels, tol in ((200, 2e-6), (700, 2e-6)):
        sp = pkg.StateSpaceSparse(model.stoich_matrix, [0, 0])
        sp.expand_(levels)
        p0 = pkg.FspVectorSparse.from_pairs(sp, [([0, 0], 1.0)])
        res = {}
        for name, m, rt in (("fused", pkg.NativeBDFFused(), 1e-7), ("classic", pkg.NativeBDFClassic(), 1e-7),
                            ("rk", pkg.NativeRK45(), 1e-9)):
            res[name] = pkg.solve(model, p0, (0.0, 2.0), m, saveat=[0.5, 2.0], odertol=rt, odeatol=1e-13)
        print(levels, {k: (v.stats["steps"], v.stats["rhs_evals"], v.stats["launches"], round(v.stats["wall_s"], 4))
                       for k, v in res.items()})
        for k in range(2):
            assert np.abs(res["fused"].p[k].values - res["rk"].p[k].values).max() < tol
            assert np.abs(res["classic"].p[k].values - res["rk"].p[k].values).max() < tol
            assert res["fused"].p[k].sum() + res["fused"].sinks[k].sum() == pytest.approx(1.0, abs=1e-9)
        # product-form Poisson marginals: mean x1 = 10 (1 - e^-t), mean x2 = 16 (1 - e^-t/2)
        st = res["fused"].p[1].states
        v = res["fused"].p[1].values
        assert (v * st[:, 0]).sum() == pytest.approx(10.0 * (1 - math.exp(-2.0)), rel=1e-5)
        assert (v * st[:, 1]).sum() == pytest.approx(16.0 * (1 - math.exp(-1.0)), rel=1e-5)


def test_bdf_fused_adaptive_event_and_timevarying(pkg):
    """Adaptive solves (sink event inside a fused step, dense output at the event, prune/expand in between) with a
    separable and a joint time-varying reaction: fused == classic within the solver tolerance."""
    touts = np.arange(0.0, 3601.0, 600.0)
    p0 = pkg.FspVectorSparse([[0, 0]], [1.0])
    for sep in (True, False):
        model = pkg.workloads.toggle_model(separable=sep)
        out = {}
        for name, m in (("fused", pkg.NativeBDFFused()), ("classic", pkg.NativeBDFClassic())):
            out[name] = pkg.solve(model, p0, (0.0, 3600.0), pkg.AdaptiveFspSparse(m, pkg.RStepAdapter(20, 5, True)),
                                  saveat=touts, odertol=1e-7, odeatol=1e-14)
        assert out["fused"].stats["adapts"] >= 1
        for k in range(len(touts)):
            a, b = _align(out["fused"].p[k].states, out["fused"].p[k].values, out["classic"].p[k].states,
                          out["classic"].p[k].values)
            assert np.abs(a - b).max() < 5e-6
            assert out["fused"].p[k].sum() + out["fused"].sinks[k].sum() == pytest.approx(1.0, abs=1e-7)


def test_bdf_fused_multistep_launches(pkg):
    """Time-invariant matrices on one CTA: up to 64 step attempts per launch with the controller on the device
    (every-step output ring, sink-event sign test, order selection in the kernel).  Same answers as one launch per step
    (NCME_BDF_SINGLE_STEP=1) and as the launch-per-operation BDF; far fewer launches."""
    import os
    tm = pkg.workloads.telegraph_model()
    p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
    # (a) adaptive, every-step output, events + prune/expand in between (examples/telegraph_cme.jl)
    alg = pkg.AdaptiveFspSparse(pkg.NativeBDFFused(), pkg.RStepAdapter(5, 10, True))
    multi = pkg.solve(tm, p0, (0.0, 300.0), alg)
    os.environ["NCME_BDF_SINGLE_STEP"] = "1"
    try:
        single = pkg.solve(tm, p0, (0.0, 300.0), alg)
    finally:
        del os.environ["NCME_BDF_SINGLE_STEP"]
    classic = pkg.solve(tm, p0, (0.0, 300.0), pkg.AdaptiveFspSparse(pkg.NativeBDFClassic(), pkg.RStepAdapter(5, 10, True)))
    print("multi", multi.stats, "single", single.stats, "classic", classic.stats)
    assert multi.stats["launches"] < 0.5 * single.stats["launches"]
    assert abs(multi.stats["steps"] - single.stats["steps"]) <= 3
    for sol in (multi, single, classic):
        assert sol.t[0] == 0.0 and sol.t[-1] == 300.0
        assert all(b >= a for a, b in zip(sol.t, sol.t[1:]))
        assert len(sol) >= sol.stats["steps"] + 1            # every accepted step (+ t0, + the final slice) is in the output
        for p, s in zip(sol.p, sol.sinks):
            assert p.sum() + s.sum() == pytest.approx(1.0, abs=1e-6)
    a, b = _align(multi.p[-1].states, multi.p[-1].values, single.p[-1].states, single.p[-1].values)
    assert np.abs(a - b).max() < 2e-6
    a, b = _align(multi.p[-1].states, multi.p[-1].values, classic.p[-1].states, classic.p[-1].values)
    assert np.abs(a - b).max() < 2e-6
    # every-step slices of the multi-step run are consistent with its own dense trajectory: interpolate the mean
    mean = [float((p.values * p.states[:, 2]).sum()) for p in multi.p]
    ms = [float((p.values * p.states[:, 2]).sum()) for p in single.p]
    assert np.interp(150.0, multi.t, mean) == pytest.approx(np.interp(150.0, single.t, ms), rel=1e-3)
    # (b) fixed space, every step, no event; tight tolerance against the explicit integrator; reproducible bitwise
    sp = pkg.StateSpaceSparse(tm.stoich_matrix, [1, 0, 0])
    sp.expand_(60)
    pf = pkg.FspVectorSparse.from_pairs(sp, [([1, 0, 0], 1.0)])
    f1 = pkg.solve(tm, pf, (0.0, 50.0), pkg.NativeBDFFused(), odertol=1e-8, odeatol=1e-13)
    f2 = pkg.solve(tm, pf, (0.0, 50.0), pkg.NativeBDFFused(), odertol=1e-8, odeatol=1e-13)
    rk = pkg.solve(tm, pf, (0.0, 50.0), pkg.NativeRK45(), saveat=[50.0], odertol=1e-9, odeatol=1e-13)
    assert len(f1) == f1.stats["steps"] + 1 and f1.t == f2.t and f1.t[-1] == 50.0
    assert all(np.array_equal(x.values, y.values) for x, y in zip(f1.p, f2.p))
    assert f1.stats["launches"] < 0.1 * f1.stats["steps"] + 10
    assert np.abs(f1.p[-1].values - rk.p[0].values).max() < 5e-7
    assert np.abs(f1.sinks[-1] - rk.sinks[0]).max() < 5e-7


def test_callback_exceptions_propagate(pkg):
    """an exception raised by a user time factor / propensity inside the integrator's host callback
    must stop the C integrator (NCME_ERR_ABORTED) and surface in the caller."""
    class Boom(RuntimeError):
        pass

    def tfac(t, p):
        if t > 3.0:
            raise Boom(f"time factor failed at t = {t}")
        return 1.0
    props = [pkg.propensity(lambda x, p: 0.05 * x[0]), pkg.propensity(lambda x, p: 0.1 * x[1], tfac),
             pkg.propensity(lambda x, p: 5.0 * x[1]), pkg.propensity(lambda x, p: 1.0 * x[2])]
    model = pkg.CmeModel(TELEGRAPH_S, props, [])
    p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
    for ode in (None, pkg.NativeRK45(), pkg.NativeBDFClassic()):
        alg = pkg.AdaptiveFspSparse(ode_method=ode, space_adapter=pkg.RStepAdapter(10, 10, True))
        with pytest.raises(Boom):
            pkg.solve(model, p0, (0.0, 50.0), alg)
    # the library is usable afterwards (the abort flag is cleared at the next segment)
    ok = pkg.CmeModel(TELEGRAPH_S, _to_pkg_props(pkg, fspmat_propensities("tv")), FSPMAT_THETA)
    alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(10, 10, True))
    sol = pkg.solve(ok, p0, (0.0, 10.0), alg)
    assert sol.p[-1].sum() + sol.sinks[-1].sum() == pytest.approx(1.0, abs=1e-6)


def test_separability_fallback_in_solve(pkg):
    """f = c1 x0 + 1{5<t<10} c2 x1 (c2 below the decay rate of x1) is classified as separable from the probe times; the run-time
    sentinels must catch it and solve() must repeat the segment on the exact joint path -- same answer as
    detect_separable=False, and different from the (wrong) separable generator."""
    S = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]]).T
    f = lambda t, x, p: 0.3 * x[0] + (0.4 * x[1] if 5.0 < t < 10.0 else 0.0 * x[1])   # 0.4 < the decay rate 0.5: no blow-up
    props = [pkg.propensity(lambda x, p: 4.0 + 0.0 * x[0]), pkg.propensity(lambda x, p: 0.2 * x[0]),
             pkg.propensity(f), pkg.propensity(lambda x, p: 0.5 * x[1])]
    model = pkg.CmeModel(S, props, [])
    p0 = pkg.FspVectorSparse([[3, 2]], [1.0])
    alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(15, 10, True))
    kw = dict(saveat=[4.0, 8.0, 12.0], fsptol=1e-6, odertol=1e-7, odeatol=1e-12)
    auto = pkg.solve(model, p0, (0.0, 12.0), alg, **kw)
    exact = pkg.solve(model, p0, (0.0, 12.0), alg, detect_separable=False, **kw)
    assert auto.stats.get("separability_fallbacks", 0) == 1
    for k in range(len(exact)):
        a, b = _align(auto.p[k].states, auto.p[k].values, exact.p[k].states, exact.p[k].values)
        assert np.abs(a - b).max() < 1e-6
    # sanity: the window matters (mean of species 2 at t = 12 differs visibly from a run that ignores it)
    g = lambda t, x, p: 0.3 * x[0]
    m2 = pkg.CmeModel(S, props[:2] + [pkg.propensity(g), props[3]], [])
    wrong = pkg.solve(m2, p0, (0.0, 12.0), alg, **kw)
    mean = lambda sol: float((sol.p[-1].values * sol.p[-1].states[:, 1]).sum())
    assert abs(mean(exact) - mean(wrong)) > 0.1


@pytest.mark.parametrize("variant", ["full_sep", "sel_sep", "full_joint"])
def test_toggle_full_horizon_vs_oracle(pkg, variant):
    """BASELINE.json config 2 at the example's FULL horizon (examples/toggleswitch_fsp_variants.jl:62-75: t in
    [0, 8 h], saveat every 60 s = 481 slices + the final one, RStepAdapter / SelectiveRStepAdapter (20, 5, true),
    odertol 1e-4, odeatol 1e-14) against oracle.solve.solve_adaptive (scipy BDF with the exact sparse Jacobian at
    odertol 1e-8).  Tolerance = the solver tolerance: |p_gpu - p_oracle| <= 1e-4 * max p per slice at the example's
    odertol = 1e-4, and 2e-6 at odertol = 1e-7."""
    sep = variant != "full_joint"
    selective = variant == "sel_sep"
    model = pkg.workloads.toggle_model(separable=sep)
    tend = 8 * 3600.0
    touts = np.arange(0.0, tend + 1.0, 60.0)
    ada = (pkg.SelectiveRStepAdapter if selective else pkg.RStepAdapter)(20, 5, True)
    oada = (SelectiveRStepAdapterOracle if selective else RStepAdapterOracle)(20, 5, True)
    ref = solve_adaptive(model.stoich_matrix, model.propensities, model.parameters, [[0, 0]], [1.0], (0.0, tend), oada,
                         saveat=touts, fsptol=1e-6, odeatol=1e-14, odertol=1e-8, method="BDF", sparse_jac=True)
    assert len(ref["t"]) == 482
    p0 = pkg.FspVectorSparse([[0, 0]], [1.0])
    for rt, tol in ((1e-4, 1e-4), (1e-7, 2e-6)):
        sol = pkg.solve(model, p0, (0.0, tend), pkg.AdaptiveFspSparse(None, ada), saveat=touts, odertol=rt, odeatol=1e-14)
        assert len(sol) == 482 and np.allclose(sol.t, ref["t"])
        worst = 0.0
        for k in range(0, 482, 13):
            a, b = _align(sol.p[k].states, sol.p[k].values, ref["states"][k], ref["p"][k])
            worst = max(worst, np.abs(a - b).max() / b.max())
            assert sol.p[k].sum() + sol.sinks[k].sum() == pytest.approx(1.0, abs=1e-6)
        print(f"toggle {variant} odertol={rt:g}: worst slice error {worst:.2e} of max p; {sol.stats}")
        assert worst <= tol
