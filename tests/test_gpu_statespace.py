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
