"""FspVectorSparse marginals (reference: src/fspvector/fspvector.jl:66-129, test/test_fspvec.jl:5-17): the host
container against the oracle's literal restatement, and the device-side reduction (ncme_space_marginal) against both."""
import numpy as np
import pytest

from oracle.fspvector import marginal_sum
from oracle.statespace import StateSpaceOracleFast


def _hog1p_like_states(L=6):
    S = np.array([[-1, 1, 0, 0, 0, 0], [1, -1, 0, 0, 0, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, -1, 1], [0, 0, 0, 0, 0, -1],
                  [0, -1, 1, 0, 0, 0], [0, 1, -1, 0, 0, 0], [0, 0, -1, 1, 0, 0], [0, 0, 1, -1, 0, 0]]).T
    sp = StateSpaceOracleFast(S, [1, 0, 0, 0, 0, 0])
    sp.expand(L)
    return S, sp.states_array()


def test_reference_fspvec_test(pkg):  # test/test_fspvec.jl:5-17
    states = np.array([[i, j] for i in range(1, 4) for j in range(1, 5)])
    p = pkg.FspVectorSparse(states, np.ones(12) / 12.0)
    assert p.nnz() == 12
    m1 = p.sum([2])
    assert m1.nnz() == 3 and np.allclose(m1.values, 4 / 12.0)
    m2 = p.sum([1])
    assert m2.nnz() == 4 and np.allclose(m2.values, 3 / 12.0)
    assert p.sum() == pytest.approx(1.0)
    with pytest.raises(pkg.ArgumentError):
        pkg.FspVectorSparse(states, np.ones(11))
    with pytest.raises(pkg.ArgumentError):
        p.sum([3])


def test_host_marginal_equals_oracle_bitwise(pkg):
    _, states = _hog1p_like_states()
    rng = np.random.default_rng(3)
    vals = rng.random(states.shape[0]) ** 4
    p = pkg.FspVectorSparse(states, vals)
    for dims in ([1, 2, 3, 4, 6], [1, 2, 3, 4, 5], [5, 6], [1], [2, 3, 4, 5, 6]):
        rs, rv = marginal_sum(states, vals, dims)
        m = p.sum(dims)
        assert np.array_equal(m.states, rs)
        assert np.array_equal(m.values, rv)          # same sums in the same order
    assert p.sum([1, 2, 3, 4, 5, 6]).values[0] == pytest.approx(vals.sum())


@pytest.mark.gpu
def test_device_marginal(pkg):
    S, states = _hog1p_like_states(7)
    sp = pkg.StateSpaceSparse(S, [1, 0, 0, 0, 0, 0])
    sp.expand_(7)
    assert np.array_equal(sp.get_states(), states)
    rng = np.random.default_rng(5)
    vals = rng.random(states.shape[0]) ** 4
    dv = pkg.DeviceVector.from_host(sp.ctx, vals)
    for dims in ([1, 2, 3, 4, 6], [1, 2, 3, 4, 5], [5, 6], [1]):
        rs, rv = marginal_sum(states, vals, dims)
        m = sp.marginal(dv, dims)
        assert np.array_equal(m.states, rs)          # reduced states in the reference's first-occurrence order
        assert np.abs(m.values - rv).max() <= 1e-13 * np.abs(rv).max()
    with pytest.raises(pkg.ArgumentError):
        sp.marginal(dv, [7])
    # a million-state 2-D space: marginal of species 1 = row sums
    S2 = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]]).T
    sp2 = pkg.StateSpaceSparse(S2, [0, 0])
    sp2.expand_(1000)
    st = sp2.get_states()
    v2 = rng.random(st.shape[0])
    v2 /= v2.sum()
    m = sp2.marginal(pkg.DeviceVector.from_host(sp2.ctx, v2), [2])
    ref = np.bincount(st[:, 0], weights=v2)
    assert m.nnz() == 1001 and np.abs(m.values[np.argsort(m.states[:, 0])] - ref).max() < 1e-15
