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
