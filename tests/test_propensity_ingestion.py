"""Host-side propensity ingestion: rank-1 separability test of joint time-varying
propensities (the form the reference's Catalyst import produces, src/cmemodel/catalyst_interface.jl:19-27)."""
import numpy as np

from oracle.statespace import StateSpaceOracleFast


def _states():
    S = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]]).T
    sp = StateSpaceOracleFast(S, [0, 0])
    sp.expand(25)
    return sp.states_array()


def test_detect_rank1_host(pkg):
    import numcme_jl_b200.fspmatrix as FM
    st = _states()
    mj = pkg.workloads.toggle_model(separable=False)
    ms = pkg.workloads.toggle_model(separable=True)
    found = FM.detect_rank1(mj.propensities[3].f, st, mj.parameters)
    assert found is not None
    g, sent = found
    # g is the joint propensity at the first probe time where it does not vanish: a multiple of the state factor
    sf = np.array([ms.propensities[3].statefactor(list(x), ms.parameters) for x in st], dtype=float)
    nz = sf != 0
    assert np.array_equal(g != 0, nz)
    r = g[nz] / sf[nz]
    assert np.abs(r - r[0]).max() <= 1e-13 * abs(r[0])
    assert 1 <= len(sent) <= 4 and all(g[i] != 0 for i in sent)
    # time factor recovered from one state == the separable model's time factor (up to the constant r[0])
    for t in (0.0, 100.0, 3600.0, 3600.5, 9000.0):
        c = mj.propensities[3].f(t, [int(v) for v in st[sent[0]]], mj.parameters) / g[sent[0]]
        assert abs(c * r[0] - ms.propensities[3].tfactor(t, ms.parameters)) <= 1e-15
    # genuinely joint, and identically zero propensities are left alone
    assert FM.detect_rank1(lambda t, x, p: x[1] * (1.0 + np.sin(0.01 * t * (1.0 + x[0]))), st, []) is None
    assert FM.detect_rank1(lambda t, x, p: 0.0 * x[0], st, []) is None
    # support that moves in time is not separable
    assert FM.detect_rank1(lambda t, x, p: np.where(x[0] > t, 1.0, 0.0) * (1.0 + x[1]), st, []) is None


def test_eval_over_states_rejects_species_reductions(pkg):
    """a propensity that reduces over species with numpy must not be broadcast as a constant."""
    from numcme_jl_b200.cmemodel import eval_over_states
    st = np.array([[0, 1], [2, 3], [4, 5]], dtype=np.int64)
    assert eval_over_states(lambda x, p: p[0] * np.sum(x), st, [2.0]).tolist() == [2.0, 10.0, 18.0]
    assert eval_over_states(lambda x, p: float(np.max(x)), st, []).tolist() == [1.0, 3.0, 5.0]
    assert eval_over_states(lambda x, p: np.prod(x), st, []).tolist() == [0.0, 6.0, 20.0]
    # genuinely constant and properly vectorised callables keep the one-call path
    assert eval_over_states(lambda x, p: 3.0, st, []).tolist() == [3.0, 3.0, 3.0]
    assert eval_over_states(lambda x, p: p[0] * x[0] * x[1], st, [0.5]).tolist() == [0.0, 3.0, 10.0]
    assert eval_over_states(lambda t, x, p: t * x[1], st, [], t=2.0).tolist() == [2.0, 6.0, 10.0]
    assert eval_over_states(lambda x, p: 1.0, st[:0], []).shape == (0,)


def test_rank1_probe_miss_is_caught_by_sentinels(pkg):
    """f = x0 + 1{5<t<10} 2 x1 passes the six probe times; the run-time sentinels must refuse it."""
    import numcme_jl_b200.fspmatrix as FM
    import numcme_jl_b200._lib as L
    st = _states()
    f = lambda t, x, p: 1.0 * x[0] + (2.0 * x[1] if 5.0 < t < 10.0 else 0.0 * x[1])
    info = FM._rank1_info(f, st, [])
    assert info is not None                      # the probe times miss the window: classified as separable
    g = info["g"]
    info["sent_states"] = [[int(v) for v in st[i]] for i in info["sent"]]
    info["sent_g"] = [float(g[i]) for i in info["sent"]]
    info["zero_states"] = [[int(v) for v in st[i]] for i in info["zero_sent"]]
    assert info["zero_states"] and info["zero_states"][0][0] == 0       # a state outside the support of g = x0

    class Holder:
        parameters = []
    tf = FM.FspMatrixSparse._make_rank1_tfactor(Holder(), 1, f, info)
    assert tf(1.0) == 1.0 and tf(12.0) == 1.0
    import pytest
    with pytest.raises(L.SeparabilityError):
        tf(7.0)


def test_callback_guard_stores_first_exception(pkg):
    """exceptions inside ctypes callbacks are stored and re-raised, never swallowed."""
    import numcme_jl_b200._lib as L
    import pytest
    calls = []
    guard = L.CallbackGuard()

    def boom(t):
        calls.append(t)
        raise KeyError("user tfactor failed")
    cb = guard.wrap(boom)
    cb(1.0)              # must not raise here (we would be inside a C frame)
    cb(2.0)              # after a failure the callback is not entered again
    assert calls == [1.0]
    with pytest.raises(KeyError):
        guard.reraise()
    guard.reraise()      # cleared
