"""GPU parity: StateSpaceSparse on the device vs the oracle, index-exact (not only after sorting)."""
import numpy as np
import pytest

from fixtures import TELEGRAPH_S, TOGGLE_S
from oracle.statespace import StateSpaceOracle, StateSpaceOracleFast

pytestmark = pytest.mark.gpu


def _same(space, osp):
    assert space.get_state_count() == osp.get_state_count()
    assert np.array_equal(space.get_states(), osp.states_array())
    assert np.array_equal(space.get_state_connectivity().astype(np.int64), osp.state_connectivity_array())
    assert np.array_equal(space.get_sink_connectivity().astype(np.int64), osp.sink_connectivity_array())


def test_reference_kats(pkg):  # test/test_statespace.jl
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    assert sp.get_state_count() == 1
    sp.expand_(0)
    assert sp.get_state_count() == 1
    sp.expand_(1)
    assert sp.get_state_count() == 3
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [[1, 0, 0]])
    sp.expand_(3)
    assert sp.get_state_count() == 7
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    sp.expand_(4)
    expected = sorted([[1, 0, 0], [0, 1, 0], [1, 0, 1], [0, 1, 1], [1, 0, 2], [0, 1, 2], [1, 0, 3], [0, 1, 3], [1, 0, 4]])
    assert sorted(sp.get_states().tolist()) == expected
    sp = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
    sp.expand_(1)
    assert sp.get_state_count() == 3
    sp = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
    sp.expand_(3)
    assert sorted(sp.get_states().tolist()) == sorted(
        [[0, 0], [1, 0], [2, 0], [3, 0], [0, 1], [1, 1], [2, 1], [0, 2], [1, 2], [0, 3]])


def test_appendix_a(pkg):
    sp = pkg.StateSpaceSparse(TELEGRAPH_S, [1, 0, 0])
    sp.expand_(2)
    assert sp.get_states().tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 1], [0, 1, 1], [1, 0, 2]]
    assert sp.get_state_connectivity().tolist() == [[0, 2, 0, 3], [1, 0, 0, 4], [0, 4, 1, 5], [3, 0, 2, 0], [0, 0, 3, 0]]
    assert sp.get_sink_connectivity().tolist() == [[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 3, 0], [1, 0, 3, 0]]


def test_initial_list_semantics(pkg):  # sparsestatespace.jl:221 duplicates / negatives dropped
    init = [[0, 1], [0, 1], [-1, 2], [10, 1], [0, 10]]
    sp = pkg.StateSpaceSparse(TOGGLE_S, init)
    _same(sp, StateSpaceOracle(TOGGLE_S, init))
    assert sp.lookup([[10, 1], [5, 5], [-1, 0], [0, 1]]).tolist() == [2, 0, 0, 1]


@pytest.mark.parametrize("S,x0,L", [(TELEGRAPH_S, [1, 0, 0], 12), (TOGGLE_S, [0, 0], 15), (TOGGLE_S, [[3, 4], [0, 0]], 6)])
def test_expand_delete_sequence_index_exact(pkg, S, x0, L):
    sp, osp = pkg.StateSpaceSparse(S, x0), StateSpaceOracle(S, x0)
    for step in (1, 2, L):
        sp.expand_(step)
        osp.expand(step)
        _same(sp, osp)
    sp.expand_(3, onlyreactions=[1, 3])
    osp.expand(3, onlyreactions=[1, 3])
    _same(sp, osp)
    rng = np.random.default_rng(1)
    ids = rng.choice(osp.get_state_count(), size=osp.get_state_count() // 3, replace=False) + 1
    sp.deleteat_(ids)
    osp.deleteat(ids)
    _same(sp, osp)
    sp.expand_(2)
    osp.expand(2)
    _same(sp, osp)
    sp.deleteat_(np.arange(1, osp.get_state_count() + 1))
    assert sp.get_state_count() == 0


def test_rectangular_fixture(pkg):  # test/sensmat/telegraph.jl:29-31
    st = [[1, 0, i] for i in range(501)] + [[0, 1, i] for i in range(501)]
    _same(pkg.StateSpaceSparse(TELEGRAPH_S, st), StateSpaceOracleFast(TELEGRAPH_S, st))


def test_2d_exploration_200(pkg):  # examples/2dstate_exploration.jl: 20 301 states
    sp = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
    sp.expand_(200)
    osp = StateSpaceOracleFast(TOGGLE_S, [0, 0])
    osp.expand(200)
    assert sp.get_state_count() == 20301
    _same(sp, osp)


def test_from_host_roundtrip(pkg):
    osp = StateSpaceOracleFast(TOGGLE_S, [0, 0])
    osp.expand(30)
    sp = pkg.StateSpaceSparse.from_host(TOGGLE_S, osp.states_array(), osp.state_connectivity_array(),
                                        osp.sink_connectivity_array())
    _same(sp, osp)
    sp.expand_(5)
    osp.expand(5)
    _same(sp, osp)


def test_three_species_million(pkg):
    """Size-independent properties at scale: simplex count, sorted set, every interior row has all
    predecessors (M-3D at L=180: 1 004 731 states)."""
    S = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]).T
    sp = pkg.StateSpaceSparse(S, [0, 0, 0])
    L = 180
    sp.expand_(L)
    n = (L + 1) * (L + 2) * (L + 3) // 6
    assert sp.get_state_count() == n
    st = sp.get_states()
    assert (st.sum(axis=1) <= L).all() and (st >= 0).all()
    assert np.unique(st, axis=0).shape[0] == n
    kc = sp.get_sink_connectivity()
    on_boundary = st.sum(axis=1) == L
    assert ((kc[:, 0] != 0) == on_boundary).all() and (kc[:, 1] == 0).all()
    sc = sp.get_state_connectivity()
    assert ((sc[:, 0] != 0) == (st[:, 0] > 0)).all()          # predecessor through +e1 exists iff x1 > 0
    assert ((sc[:, 1] != 0) == (~on_boundary)).all()          # predecessor through -e1 is x+e1
    j = sc[:, 0][st[:, 0] > 0].astype(np.int64) - 1
    assert np.array_equal(st[j] + S[:, 0], st[st[:, 0] > 0])


def test_key_relayout_when_a_species_outgrows_its_field(pkg):
    """6 species start with 10 key bits each (max count 1023); a species that grows past that triggers a re-division
    of the 63-bit budget instead of an error -- ordering and connectivity stay index-exact."""
    S = np.zeros((6, 3), dtype=np.int64)
    S[5, 0] = 1          # birth of species 6
    S[5, 1] = -1         # death of species 6
    S[0, 2], S[1, 2] = -1, 1   # G0 -> G1
    sp = pkg.StateSpaceSparse(S, [[1, 0, 0, 0, 0, 0]])
    sp.expand_(1500)
    osp = StateSpaceOracleFast(S, [[1, 0, 0, 0, 0, 0]], bits_per_species=[4, 4, 4, 4, 4, 40])
    osp.expand(1500)
    assert sp.get_states()[:, 5].max() == 1500 > 1023
    _same(sp, osp)
    assert sp.lookup([[0, 1, 0, 0, 0, 1499], [1, 0, 0, 0, 0, 1501]]).tolist() == [int(osp._index.lookup(osp._pack(np.array([[0, 1, 0, 0, 0, 1499]])))[0]), 0]
    sp.deleteat_([3, 4, 5])
    osp.deleteat([3, 4, 5])
    sp.expand_(3)
    osp.expand(3)
    _same(sp, osp)
    # an initial state that does not fit the equal split
    sp2 = pkg.StateSpaceSparse(S, [[0, 1, 0, 0, 0, 5000]])
    assert sp2.get_states().tolist() == [[0, 1, 0, 0, 0, 5000]]
    sp2.expand_(2)
    osp2 = StateSpaceOracleFast(S, [[0, 1, 0, 0, 0, 5000]], bits_per_species=[4, 4, 4, 4, 4, 40])
    osp2.expand(2)
    _same(sp2, osp2)


def test_small_to_general_handover_index_exact(pkg):
    """An expansion that starts in the single-launch path (k_expand_small), outgrows its pre-reserved capacity several
    times and finally the 65 536-state limit, where the per-level pipeline takes over from the last batch of new
    states: states and connectivity stay index-exact against the oracle (3 species, L = 75: 76 076 states)."""
    S = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]).T
    sp, osp = pkg.StateSpaceSparse(S, [0, 0, 0]), StateSpaceOracleFast(S, [0, 0, 0])
    sp.expand_(75)
    osp.expand(75)
    assert sp.get_state_count() == 76 * 77 * 78 // 6 > 65536
    _same(sp, osp)
    # ... and restricted expansions (SelectiveRStepAdapter) from there and from a small space
    sp.expand_(2, onlyreactions=[1, 4])
    osp.expand(2, onlyreactions=[1, 4])
    _same(sp, osp)
    sp2, osp2 = pkg.StateSpaceSparse(S, [[2, 1, 0], [0, 0, 3]]), StateSpaceOracleFast(S, [[2, 1, 0], [0, 0, 3]])
    for only in ([2], [1, 3, 5], []):
        sp2.expand_(4, onlyreactions=only)
        osp2.expand(4, onlyreactions=only)
        _same(sp2, osp2)


@pytest.mark.parametrize("R", [33, 48, 64])
def test_more_than_32_reactions(pkg, R):
    """33..64 reactions (NCME_MAX_REACTIONS = 64, 64-bit sink masks): state space index-exact incl. expand with
    onlyreactions above bit 31 and deleteat!, matvec <= 1e-12 (generic > 16-slot kernel), sink rows of reactions >= 32,
    and a BDF solve on the launch-per-operation path (the fused step kernel handles R <= 32) against DP5."""
    from oracle.fspmatrix import FspMatrixOracle, OProp
    from test_gpu_matvec import _to_pkg_props
    rng = np.random.default_rng(100 + R)
    S = rng.integers(-1, 2, size=(6, R))
    S[:, R - 1] = [1, 0, 0, 0, 0, -1]          # make sure the highest reaction has a non-trivial stoichiometry
    S[:, 36 % R] = S[:, 2]                     # duplicate stoichiometry across the 32-bit boundary
    x0 = [2, 2, 2, 2, 2, 2]
    sp, osp = pkg.StateSpaceSparse(S, x0), StateSpaceOracleFast(S, x0)
    sp.expand_(1)
    osp.expand(1)
    _same(sp, osp)
    hi = [r for r in (R, R - 1, 34, 2) if r <= R]
    sp.expand_(1, onlyreactions=hi)
    osp.expand(1, onlyreactions=hi)
    _same(sp, osp)
    n = sp.get_state_count()
    drop = sorted(rng.choice(np.arange(2, n + 1), size=n // 5, replace=False).tolist())
    sp.deleteat_(drop)
    osp.deleteat(drop)
    _same(sp, osp)
    props = [OProp("ti", f=(lambda x, p, r=r: (0.05 + 0.003 * r) * (1.0 + x[r % 6]))) for r in range(R)]
    props[R - 2] = OProp("sep", tfactor=lambda t, p: 1.0 + 0.5 * np.sin(t), statefactor=lambda x, p: 0.3 * x[2])
    A = pkg.FspMatrixSparse(sp, _to_pkg_props(pkg, props), parameters=[])
    OA = FspMatrixOracle(osp, props, [])
    v = rng.random(A.size(1))
    for t in (0.0, 1.3):
        w, wr = pkg.matvec(t, A, v), OA.matvec(t, v)
        assert np.abs(w - wr).max() <= 1e-12 * np.abs(wr).max()
        assert np.abs(wr[-R:]).min() >= 0.0 and np.abs(wr[-(R - 32):]).max() > 0.0   # sink rows above bit 31 are populated
    p0 = pkg.FspVectorSparse.from_pairs(sp, [(x0, 1.0)])
    model = pkg.CmeModel(S, _to_pkg_props(pkg, props), [])
    a = pkg.solve(model, p0, (0.0, 0.3), None, saveat=[0.3], odertol=1e-8, odeatol=1e-13)
    b = pkg.solve(model, p0, (0.0, 0.3), pkg.NativeRK45(), saveat=[0.3], odertol=1e-9, odeatol=1e-13)
    assert np.abs(a.p[0].values - b.p[0].values).max() < 1e-7 and np.abs(a.sinks[0] - b.sinks[0]).max() < 1e-7
    with pytest.raises(pkg.ArgumentError):      # 65 reactions: a clean error, not a truncated mask
        pkg.StateSpaceSparse(rng.integers(-1, 2, size=(3, 65)), [1, 1, 1])


def test_maximum_reaction_count_and_degenerate_spaces(pkg):
    """R = 32 (the most the fused BDF step kernel handles) over 8 species incl. a zero-stoichiometry reaction and two reactions with identical
    stoichiometry; a one-state space; a space emptied by deleteat!."""
    from oracle.fspmatrix import FspMatrixOracle, OProp
    rng = np.random.default_rng(17)
    S = rng.integers(-1, 2, size=(8, 32))
    S[:, 5] = 0                      # x -> x
    S[:, 9] = S[:, 3]                # duplicate stoichiometry (merged into one slot, like Julia's sparse())
    x0 = [2, 2, 2, 2, 2, 2, 2, 2]
    sp, osp = pkg.StateSpaceSparse(S, x0), StateSpaceOracleFast(S, x0)
    sp.expand_(2)
    osp.expand(2)
    _same(sp, osp)
    props = [OProp("ti", f=(lambda x, p, r=r: (0.1 + 0.01 * r) * (1.0 + x[r % 8]))) for r in range(32)]
    props[7] = OProp("sep", tfactor=lambda t, p: 1.0 + 0.5 * np.sin(t), statefactor=lambda x, p: 0.3 * x[2])
    from test_gpu_matvec import _to_pkg_props
    A = pkg.FspMatrixSparse(sp, _to_pkg_props(pkg, props), parameters=[])
    OA = FspMatrixOracle(osp, props, [])
    v = rng.random(A.size(1))
    for t in (0.0, 1.3):
        w, wr = pkg.matvec(t, A, v), OA.matvec(t, v)
        assert np.abs(w - wr).max() <= 1e-12 * np.abs(wr).max()
    # fused-step BDF with 32 sink rows
    p0 = pkg.FspVectorSparse.from_pairs(sp, [(x0, 1.0)])
    model = pkg.CmeModel(S, _to_pkg_props(pkg, props), [])
    a = pkg.solve(model, p0, (0.0, 0.5), pkg.NativeBDFFused(), saveat=[0.5], odertol=1e-8, odeatol=1e-13)
    b = pkg.solve(model, p0, (0.0, 0.5), pkg.NativeRK45(), saveat=[0.5], odertol=1e-9, odeatol=1e-13)
    assert np.abs(a.p[0].values - b.p[0].values).max() < 1e-7 and np.abs(a.sinks[0] - b.sinks[0]).max() < 1e-7
    # (this random network is not mass-action: reactions with a negative successor keep a positive propensity and leak
    #  mass, SURVEY.md 3A -- both integrators must lose the same amount)
    assert a.p[0].sum() + a.sinks[0].sum() == pytest.approx(b.p[0].sum() + b.sinks[0].sum(), abs=2e-5)   # 493 + 32 entries
    # one state, no expansion: everything leaks into the sinks
    one = pkg.StateSpaceSparse(TOGGLE_S, [0, 0])
    m1 = pkg.workloads.m2d_model()
    s1 = pkg.solve(m1, pkg.FspVectorSparse([[0, 0]], [1.0]), (0.0, 0.1), None, saveat=[0.1])
    assert s1.p[0].values[0] == pytest.approx(np.exp(-1.8), rel=2e-3)       # exp(-(10 + 8) t) at the default odertol = 1e-4
    assert s1.p[0].sum() + s1.sinks[0].sum() == pytest.approx(1.0, abs=1e-9)
    # delete everything, then the space is empty and expansion is a no-op (nothing to explore)
    one.expand_(2)
    one.deleteat_(list(range(1, one.get_state_count() + 1)))
    assert one.get_state_count() == 0 and one.get_states().shape == (0, 2)
    one.expand_(3)
    assert one.get_state_count() == 0
