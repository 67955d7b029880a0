"""GPU parity: fused FSP matvec (K1) + assembly (K6) vs the oracle; tolerance 1e-12 relative
(BASELINE.json north_star), through the C ABI (device and host-buffer entry points)."""
import numpy as np
import pytest

from fixtures import FSPMAT_THETA, SENS_THETA, TELEGRAPH_S, TOGGLE_S, fspmat_propensities, sens_telegraph
from oracle.fspmatrix import FspMatrixOracle
from oracle.statespace import StateSpaceOracle, StateSpaceOracleFast

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _to_pkg_props(pkg, oprops):
    out = []
    for a in oprops:
        if a.kind == "ti":
            out.append(pkg.StandardTimeInvariantPropensity(a.f))
        elif a.kind == "sep":
            out.append(pkg.SeparableTimeVaryingPropensity(a.tfactor, a.statefactor))
        else:
            out.append(pkg.JointTimeVaryingPropensity(a.f))
    return out


@pytest.mark.parametrize("kind", ["ti", "tv", "tvj", "tvj-exact"])
def test_fspmat_jl_on_gpu(pkg, kind):  # test/test_fspmat.jl:41-68
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    sp.expand_(2)
    detect = kind != "tvj-exact"      # "tvj": the joint propensity is found to be c(t) g(x); "-exact": _update_sparsematrix! path
    kind = kind.split("-")[0]
    props = fspmat_propensities(kind)
    A = pkg.FspMatrixSparse(sp, _to_pkg_props(pkg, props), parameters=FSPMAT_THETA, detect_separable=detect)
    if kind == "tvj":
        assert (A.device_joint_ids == []) == detect and A.jointtv_propensity_ids == [2]
    assert A.size(1) == sp.get_state_count() + sp.get_sink_count() == A.size(2)
    osp = StateSpaceOracle(TELEGRAPH_S, [1, 0, 0])
    osp.expand(2)
    OA = FspMatrixOracle(osp, props, FSPMAT_THETA)
    v = np.ones(A.size(1))
    for t in (1.0, 0.0, 0.37):
        w = pkg.matvec(t, A, v)
        assert abs(w.sum()) <= 1e-14
        assert _relerr(w, OA.matvec(t, v)) <= RTOL
    with pytest.raises(pkg.ArgumentError):
        A.size(3)
    with pytest.raises(pkg.ArgumentError):
        pkg.matvec(0.0, A, np.ones(3))
    assert _relerr(A @ v, OA.matvec(0.0, v)) <= RTOL


def test_separable_equals_joint(pkg):  # test/test_fspmat.jl:68
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    sp.expand_(2)
    A1 = pkg.FspMatrixSparse(sp, _to_pkg_props(pkg, fspmat_propensities("tv")), parameters=FSPMAT_THETA)
    v = np.ones(A1.size(1))
    for detect in (True, False):
        A2 = pkg.FspMatrixSparse(sp, _to_pkg_props(pkg, fspmat_propensities("tvj")), parameters=FSPMAT_THETA,
                                 detect_separable=detect)
        for t in (1.0, 0.0):
            assert np.linalg.norm(pkg.matvec(t, A1, v) - pkg.matvec(t, A2, v)) <= 1e-15
        assert np.linalg.norm(pkg.matvec(0.0, A1, v) - pkg.matvec(0.0, A2, v)) == 0.0    # the reference asserts `≈ 0`


@pytest.mark.parametrize("rows", [0, 1, 2, 4, 16, 17, 18, 20, 32, 64])   # +16: byte-compressed indices; +32/+64: smem-pipelined kernel
def test_rectangular_telegraph_all_kernel_variants(pkg, rows):
    props, grads, pattern, states = sens_telegraph()
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, states)
    A = pkg.FspMatrixSparse(sp, _to_pkg_props(pkg, props), parameters=SENS_THETA)
    A.set_tuning(rows)
    OA = FspMatrixOracle(StateSpaceOracleFast(TELEGRAPH_S, states), props, SENS_THETA)
    st = A.stats()
    assert st["nnz_per_term"] == OA.stored_entries() == [3006, 2002]
    assert st["algorithmic_bytes"] == OA.algorithmic_bytes()
    rng = np.random.default_rng(0)
    v = rng.random(A.size(1))
    for t in (10.0, 20.0, 30.0, 100.0):
        assert _relerr(pkg.matvec(t, A, v), OA.matvec(t, v)) <= RTOL


def test_device_resident_and_matvecadd(pkg, ctx):
    sp = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
    sp.expand_(60)
    model = pkg.workloads.toggle_model(separable=True)
    A = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters)
    osp = StateSpaceOracleFast(TOGGLE_S, [0, 0])
    osp.expand(60)
    OA = FspMatrixOracle(osp, model.propensities, model.parameters)
    rng = np.random.default_rng(5)
    v, o = rng.random(A.size(1)), rng.random(A.size(1))
    dv, do = pkg.DeviceVector.from_host(ctx, v), pkg.DeviceVector.from_host(ctx, o)
    pkg.matvecadd_(do, 100.0, A, dv)
    want = o.copy()
    OA.matvecadd_(want, 100.0, v)
    assert _relerr(do.to_host(), want) <= RTOL
    pkg.matvec_(do, 5000.0, A, dv)
    assert _relerr(do.to_host(), OA.matvec(5000.0, v)) <= RTOL
    out = o.copy()
    pkg.matvecadd_(out, 100.0, A, v)        # host-buffer path
    assert _relerr(out, want) <= RTOL
    with pytest.raises(pkg.ArgumentError):
        pkg.matvec_(dv, 0.0, A, dv)          # aliasing


def test_joint_toggle(pkg):
    sp = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
    sp.expand_(40)
    mj, ms = pkg.workloads.toggle_model(separable=False), pkg.workloads.toggle_model(separable=True)
    As = pkg.FspMatrixSparse(sp, ms.propensities, parameters=ms.parameters)
    rng = np.random.default_rng(2)
    v = rng.random(As.size(1))
    for detect in (False, True):
        Aj = pkg.FspMatrixSparse(sp, mj.propensities, parameters=mj.parameters, detect_separable=detect)
        assert (Aj.device_joint_ids == []) == detect
        for t in (0.0, 3600.0, 3601.0, 7000.0):
            assert _relerr(pkg.matvec(t, Aj, v), pkg.matvec(t, As, v)) <= RTOL


def test_rank1_detection(pkg):
    """joint propensities that are numerically c(t) g(x) go to the separable path; genuinely
    joint ones stay exact; a propensity that only looks separable on the probe times is caught by the sentinel states."""
    import math
    sp = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
    sp.expand_(30)
    base = pkg.workloads.toggle_model(separable=True)
    nonsep = pkg.propensity(lambda t, x, p: 1e-3 * x[1] * (1.0 + np.sin(0.01 * t * (1.0 + x[0]))))
    fake = pkg.propensity(lambda t, x, p: 1e-3 * x[1] * (1.0 + (1.0 if t == 5.0 else 0.0) * x[0]))
    A = pkg.FspMatrixSparse(sp, base.propensities[:3] + [nonsep], parameters=base.parameters)
    assert A.device_joint_ids == [4]
    Ax = pkg.FspMatrixSparse(sp, base.propensities[:3] + [nonsep], parameters=base.parameters, detect_separable=False)
    v = np.random.default_rng(4).random(A.size(1))
    assert np.array_equal(pkg.matvec(12.5, A, v), pkg.matvec(12.5, Ax, v))
    F = pkg.FspMatrixSparse(sp, base.propensities[:3] + [fake], parameters=base.parameters)
    assert F.device_joint_ids == []                     # looks separable on the probe times ...
    pkg.matvec(4.0, F, v)
    with pytest.raises(pkg.NcmeError):                  # ... but the sentinels notice at the time where it is not
        pkg.matvec(5.0, F, v)


def test_m2d_100k_vs_oracle(pkg):
    """M-2D at L=446 (100 128 states): oracle finishes in seconds."""
    model = pkg.workloads.m2d_model()
    sp = pkg.StateSpaceSparse(model.stoich_matrix, [0, 0])
    sp.expand_(446)
    osp = StateSpaceOracleFast(model.stoich_matrix, [0, 0])
    osp.expand(446)
    assert np.array_equal(sp.get_states(), osp.states_array())
    A = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters)
    OA = FspMatrixOracle(osp, model.propensities, model.parameters)
    assert A.stats()["algorithmic_bytes"] == OA.algorithmic_bytes()
    rng = np.random.default_rng(0)
    v = rng.random(A.size(1))
    v /= v.sum()
    ref = OA.matvec(0.0, v)
    outs = []
    for rows in (1, 2, 4, 17, 18, 20, 32, 64):
        A.set_tuning(rows)
        outs.append(pkg.matvec(0.0, A, v))
        assert _relerr(outs[-1], ref) <= RTOL
    assert all(np.array_equal(outs[0], o) for o in outs[1:])      # all variants are bitwise identical
    assert abs(pkg.matvec(0.0, A, np.ones(A.size(1))).sum()) <= 1e-9


def test_m3d_properties_at_scale(pkg, ctx):
    """M-3D TV variant at L=180 (1.0 M states): zero column sums, linearity, coefficient linearity,
    determinism -- size-independent properties (the oracle is too slow to assemble at full size)."""
    model = pkg.workloads.m3d_model(time_varying=True)
    sp = pkg.StateSpaceSparse(model.stoich_matrix, [0, 0, 0])
    sp.expand_(180)
    A = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters)
    N = A.size(1)
    rng = np.random.default_rng(0)
    x1, x2 = rng.random(N), rng.random(N)
    ones = pkg.matvec(2.5, A, np.ones(N))
    assert abs(ones.sum()) <= 1e-12 * np.abs(ones).sum()
    y1, y2, y12 = pkg.matvec(2.5, A, x1), pkg.matvec(2.5, A, x2), pkg.matvec(2.5, A, 2.0 * x1 - 3.0 * x2)
    assert _relerr(y12, 2.0 * y1 - 3.0 * y2) <= 1e-12
    assert np.array_equal(y1, pkg.matvec(2.5, A, x1))          # bitwise deterministic
    # A(t) = A_ti + c(t) A_sep  =>  A(t2) x - A(t1) x is proportional to c(t2) - c(t1)
    ya, yb, yc = pkg.matvec(0.0, A, x1), pkg.matvec(2.5, A, x1), pkg.matvec(7.5, A, x1)
    assert _relerr(yb - ya, -(yc - ya)) <= 1e-10


def test_incremental_rebuild_after_adapt(pkg):
    """the matrix after a prune + expand built from its predecessor -- only the appended states are
    evaluated on the host -- is bit-for-bit the matrix built from scratch (toggle switch with a separable and, through
    the rank-1 detection, a joint time factor; two adapt rounds; fallbacks)."""
    import numcme_jl_b200.fspmatrix as FM
    rng = np.random.default_rng(8)
    FM.INCREMENTAL_MIN_STATES, saved = 0, FM.INCREMENTAL_MIN_STATES      # (small spaces rebuild from scratch by default)
    for sep in (True, False):
        model = pkg.workloads.toggle_model(separable=sep)
        sp = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
        sp.expand_(25)
        A = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters)
        assert not A.incremental
        for rnd in range(2):
            n = sp.get_state_count()
            drop = np.sort(rng.choice(np.arange(1, n + 1), size=n // 5, replace=False))
            sp.deleteat_(drop)
            sp.expand_(4 + rnd)
            B = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters, previous=A)
            assert B.incremental and 0 < B.new_state_count < sp.get_state_count()
            F = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters)      # from scratch (re-marks the space)
            assert B.stats() == F.stats()
            v = rng.random(B.size(1))
            for t in (0.0, 1234.5, 5000.0):
                assert np.array_equal(pkg.matvec(t, B, v), pkg.matvec(t, F, v))
            # F was assembled in between: B is no longer the space's latest matrix -> silent fallback to a full build
            sp.expand_(1)
            C2 = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters, previous=B)
            assert not C2.incremental
            A = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters, previous=C2)
            assert A.incremental
            G = pkg.FspMatrixSparse(sp, model.propensities, parameters=model.parameters)
            v = rng.random(A.size(1))
            assert np.array_equal(pkg.matvec(77.0, A, v), pkg.matvec(77.0, G, v))
            A = G
    # other propensity objects -> full build
    other = pkg.workloads.toggle_model(separable=True)
    sp.expand_(1)
    assert not pkg.FspMatrixSparse(sp, other.propensities, parameters=other.parameters, previous=A).incremental
    # the adaptive solve uses it
    sol = pkg.solve(pkg.workloads.telegraph_model(), pkg.FspVectorSparse([[1, 0, 0]], [1.0]), (0.0, 300.0),
                    pkg.AdaptiveFspSparse(None, pkg.RStepAdapter(5, 10, True)), saveat=[300.0])
    assert sol.stats["incremental_builds"] == sol.stats["adapts"] >= 1
    FM.INCREMENTAL_MIN_STATES = saved
