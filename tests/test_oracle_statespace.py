"""Pins the state-space oracle on the reference's own tests (test/test_statespace.jl) and on the
worked example of SURVEY.md Appendix A; checks the vectorised oracle against the literal one."""
import numpy as np
import pytest

from fixtures import TELEGRAPH_S, TOGGLE_S
from oracle.statespace import StateSpaceOracle, StateSpaceOracleFast

IMPLS = [StateSpaceOracle, StateSpaceOracleFast]


@pytest.mark.parametrize("cls", IMPLS)
def test_telegraph_counts_and_sets(cls):  # test/test_statespace.jl:14-34
    sp = cls(TELEGRAPH_S, [1, 0, 0])
    assert sp.get_state_count() == 1
    sp.expand(0)
    assert sp.get_state_count() == 1
    sp.expand(1)
    assert sp.get_state_count() == 3
    sp = cls(TELEGRAPH_S, [[1, 0, 0]])
    sp.expand(3)
    assert sp.get_state_count() == 7
    sp = cls(TELEGRAPH_S, [1, 0, 0])
    sp.expand(4)
    expected = sorted([[1, 0, 0], [0, 1, 0], [1, 0, 1], [0, 1, 1], [1, 0, 2], [0, 1, 2], [1, 0, 3], [0, 1, 3], [1, 0, 4]])
    assert sorted(sp.states_array().tolist()) == expected


@pytest.mark.parametrize("cls", IMPLS)
def test_toggle_counts_and_sets(cls):  # test/test_statespace.jl:46-63
    sp = cls(TOGGLE_S, [0, 0])
    assert sp.get_state_count() == 1
    sp.expand(1)
    assert sp.get_state_count() == 3
    sp = cls(TOGGLE_S, [0, 0])
    sp.expand(3)
    assert sp.get_state_count() == 10
    expected = sorted([[0, 0], [1, 0], [2, 0], [3, 0], [0, 1], [1, 1], [2, 1], [0, 2], [1, 2], [0, 3]])
    assert sorted(sp.states_array().tolist()) == expected
    # insertion order quoted in SURVEY.md section 3A
    assert sp.states_array().tolist() == [[0, 0], [1, 0], [0, 1], [1, 1], [0, 2], [2, 0], [3, 0], [2, 1], [1, 2], [0, 3]]


@pytest.mark.parametrize("cls", IMPLS)
def test_appendix_a_connectivity(cls):
    sp = cls(TELEGRAPH_S, [1, 0, 0])
    sp.expand(2)
    assert sp.states_array().tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 1], [0, 1, 1], [1, 0, 2]]
    assert sp.state_connectivity_array().tolist() == [[0, 2, 0, 3], [1, 0, 0, 4], [0, 4, 1, 5], [3, 0, 2, 0], [0, 0, 3, 0]]
    assert sp.sink_connectivity_array().tolist() == [[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 3, 0], [1, 0, 3, 0]]


@pytest.mark.parametrize("cls", IMPLS)
def test_duplicates_and_negatives_dropped(cls):  # sparsestatespace.jl:221
    sp = cls(TOGGLE_S, [[0, 1], [0, 1], [-1, 2], [10, 1], [0, 10]])
    assert sp.states_array().tolist() == [[0, 1], [10, 1], [0, 10]]


def _same(a, b):
    assert np.array_equal(a.states_array(), b.states_array())
    assert np.array_equal(a.state_connectivity_array(), b.state_connectivity_array())
    assert np.array_equal(a.sink_connectivity_array(), b.sink_connectivity_array())


@pytest.mark.parametrize("S,x0,L", [(TELEGRAPH_S, [1, 0, 0], 12), (TOGGLE_S, [0, 0], 15), (TOGGLE_S, [[3, 4], [0, 0]], 6)])
def test_fast_equals_literal(S, x0, L):
    a, b = StateSpaceOracle(S, x0), StateSpaceOracleFast(S, x0)
    for step in (1, 2, L):
        a.expand(step)
        b.expand(step)
        _same(a, b)
    a.expand(3, onlyreactions=[1, 3])
    b.expand(3, onlyreactions=[1, 3])
    _same(a, b)
    rng = np.random.default_rng(1)
    ids = rng.choice(a.get_state_count(), size=a.get_state_count() // 3, replace=False) + 1
    a.deleteat(ids)
    b.deleteat(ids)
    _same(a, b)
    a.expand(2)
    b.expand(2)
    _same(a, b)


def test_invariants_after_delete():
    """struct invariants of sparsestatespace.jl:35-39 hold after deleteat! (no reference test covers it)."""
    sp = StateSpaceOracle(TOGGLE_S, [0, 0])
    sp.expand(8)
    sp.deleteat([1, 5, 9, 20, 21])
    st = sp.states_array()
    d = {tuple(s): i + 1 for i, s in enumerate(st.tolist())}
    assert d == sp.state2idx
    for i, x in enumerate(st):
        for r in range(4):
            pred = tuple(x - TOGGLE_S[:, r])
            assert sp.state_connectivity[i][r] == d.get(pred, 0)
            succ = x + TOGGLE_S[:, r]
            want = r + 1 if (succ >= 0).all() and tuple(succ) not in d else 0
            assert sp.sink_connectivity[i][r] == want


def test_simplex_count():
    sp = StateSpaceOracleFast(TOGGLE_S, [0, 0])
    sp.expand(200)                       # examples/2dstate_exploration.jl:8 -> 20 301 states
    assert sp.get_state_count() == 20301
