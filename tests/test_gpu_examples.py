"""BASELINE.json configs 1-4 as parity cases (the reference's example scripts), end to end through the product path."""
import time

import numpy as np
import pytest

from oracle.fspmatrix import FspMatrixOracle
from oracle.statespace import StateSpaceOracleFast

pytestmark = pytest.mark.gpu


def test_telegraph_example(pkg):  # examples/telegraph_cme.jl: solve(model, p0, (0,300), RStepAdapter(5,10,true)), defaults
    model = pkg.workloads.telegraph_model()
    p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
    alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(5, 10, True))
    sol = pkg.solve(model, p0, (0.0, 300.0), alg)                       # saveat = []: every step, default tolerances
    t0 = time.perf_counter()
    sol = pkg.solve(model, p0, (0.0, 300.0), alg)
    wall = time.perf_counter() - t0
    print(f"telegraph adaptive solve: {wall*1e3:.1f} ms, {sol.stats} (reference: 5.45 ms on an Apple M1, docs/src/examples/telegraph.md:89)")
    assert sol.t[-1] == 300.0
    assert sol.p[-1].sum() + sol.sinks[-1].sum() == pytest.approx(1.0, abs=1e-6)
    assert sol.sinks[-1].sum() <= 1e-6 * 1.01
    # stationary mean mRNA = lambda/gamma * k01/(k01+k10) = 10/3
    mean = float((sol.p[-1].values * sol.p[-1].states[:, 2]).sum())
    assert mean == pytest.approx(10.0 / 3.0, rel=2e-3)


def test_hog1p_example(pkg):  # examples/hog1p.jl: two phases, RStepAdapter(10,20,true), NS=6, R=13
    th = list(pkg.workloads.HOG1P_THETA)
    alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(10, 20, True))
    m0 = pkg.workloads.hog1p_model(th)
    p0 = pkg.FspVectorSparse([[1, 0, 0, 0, 0, 0]], [1.0])
    t0 = time.perf_counter()
    s0 = pkg.solve(m0, p0, (0.0, 8 * 3600.0), alg, saveat=[8 * 3600.0], fsptol=1e-6, odeatol=1e-14, odertol=1e-6)
    pend = s0[-1].p
    assert pend.sum() + s0[-1].sinks.sum() == pytest.approx(1.0, abs=1e-6)
    th[2] = 3.2e4                                                      # a: MAPK signal switched on
    res = {}
    for sep in (True, False):
        m1 = pkg.workloads.hog1p_model(th, separable=sep)
        res[sep] = pkg.solve(m1, pend, (0.0, 600.0), alg, saveat=np.arange(0.0, 601.0, 60.0), fsptol=1e-4,
                             odeatol=1e-14, odertol=1e-6)
    print(f"hog1p two-phase solve: {time.perf_counter()-t0:.2f} s, phase-1 {s0.stats}, phase-2 {res[True].stats}")
    for sol in res.values():
        for p, s in zip(sol.p, sol.sinks):
            assert p.sum() + s.sum() == pytest.approx(1.0, abs=1e-5)
    a, b = res[True], res[False]                                       # separable == joint formulation
    for k in range(len(a)):
        da = {tuple(s): v for s, v in zip(a.p[k].states.tolist(), a.p[k].values)}
        db = {tuple(s): v for s, v in zip(b.p[k].states.tolist(), b.p[k].values)}
        keys = set(da) | set(db)
        assert max(abs(da.get(q, 0.0) - db.get(q, 0.0)) for q in keys) < 2e-5
    # marginals as in the example's plotting code: sum(p, [1,2,3,4,6]) -> nuclear mRNA
    nuc = a.p[-1].sum([1, 2, 3, 4, 6]).to_array()
    assert nuc.sum() == pytest.approx(a.p[-1].sum(), rel=1e-12)


def test_hog1p_matrix_vs_oracle(pkg):
    th = list(pkg.workloads.HOG1P_THETA)
    th[2] = 3.2e4
    m = pkg.workloads.hog1p_model(th)
    sp = pkg.StateSpaceSparse(m.stoich_matrix, [1, 0, 0, 0, 0, 0])
    sp.expand_(12)
    osp = StateSpaceOracleFast(m.stoich_matrix, [1, 0, 0, 0, 0, 0])
    osp.expand(12)
    assert np.array_equal(sp.get_states(), osp.states_array())
    A = pkg.FspMatrixSparse(sp, m.propensities, parameters=m.parameters)
    OA = FspMatrixOracle(osp, m.propensities, m.parameters)
    st = A.stats()
    assert st["nnz_per_term"] == OA.stored_entries()      # 4 RNA-production reactions share one stoichiometry -> merged
    rng = np.random.default_rng(0)
    v = rng.random(A.size(1))
    for t in (0.0, 120.0, 900.0):
        w, wr = pkg.matvec(t, A, v), OA.matvec(t, v)
        assert np.abs(w - wr).max() <= 1e-12 * np.abs(wr).max()


def test_2d_exploration_million(pkg):  # examples/2dstate_exploration.jl scaled: expand!(., 1413) -> 1 000 405 states
    S = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]]).T
    sp = pkg.StateSpaceSparse(S, [0, 0])
    t0 = time.perf_counter()
    sp.expand_(1413)
    wall = time.perf_counter() - t0
    n = sp.get_state_count()
    assert n == 1000405
    st = sp.get_states()
    assert (st.sum(axis=1) <= 1413).all() and np.unique(st[:, 0] * 2000 + st[:, 1]).size == n
    sp2 = pkg.StateSpaceSparse(S, [0, 0])
    t0 = time.perf_counter()
    sp2.expand_(200)
    wall200 = time.perf_counter() - t0
    assert sp2.get_state_count() == 20301
    print(f"expand!(L=1413) -> {n} states in {wall:.3f} s; expand!(L=200) -> 20301 states in {wall200*1e3:.1f} ms")
