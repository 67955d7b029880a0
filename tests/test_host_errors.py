"""Host-side error behaviour of the mirror (no device needed): same exception classes / messages as the reference's
ArgumentErrors (fspvector.jl:29,68,110; fspoutput.jl:46; rstepadapters.jl:91; fspsolve.jl signature)."""
import numpy as np
import pytest


def test_fspvector_and_output_errors(pkg):
    with pytest.raises(pkg.ArgumentError, match="equal lengths"):
        pkg.FspVectorSparse([[0, 0], [1, 0]], [1.0])
    p = pkg.FspVectorSparse([[0, 0], [1, 0], [1, 1]], [0.2, 0.3, 0.5])
    with pytest.raises(pkg.ArgumentError, match="between 1 and 2"):
        p.sum([0])
    with pytest.raises(pkg.ArgumentError, match="empty"):
        pkg.FspVectorSparse(np.zeros((0, 2), dtype=np.int64), []).to_array()
    assert p.to_array().shape == (2, 2) and p.to_array()[1, 1] == 0.5
    assert pkg.nnz(p) == 3 and pkg.get_values(p) is p.values and p.state2idx[(1, 1)] == 3
    out = pkg.FspOutputSparse()
    out.t.append(0.0)
    out.p.append(p)
    out.sinks.append(np.zeros(2))
    assert len(out) == 1 and out[0].t == 0.0 and out[-1].p is p and len(out[[0, 0]]) == 2
    with pytest.raises(pkg.ArgumentError, match="exceeds array limit"):
        out[1]


def test_solver_argument_errors(pkg):
    model = pkg.workloads.telegraph_model()
    with pytest.raises(pkg.ArgumentError):
        pkg.AdaptiveFspSparse(ode_method=None, space_adapter=None)
    with pytest.raises(pkg.ArgumentError, match="FspVectorSparse"):
        pkg.solve(model, [[1, 0, 0]], (0.0, 1.0), None)
    from numcme_jl_b200.transientcme import _method_code, _saveat_array
    assert _method_code(None) == 1 and _method_code(pkg.NativeRK45()) == 0
    assert _method_code(pkg.NativeBDFClassic()) == 2 and _method_code(pkg.NativeBDFFused()) == 3
    with pytest.raises(pkg.ArgumentError, match="DifferentialEquations"):
        _method_code("CVODE_BDF")
    assert _saveat_array(None, (0, 1)) is None and _saveat_array([], (0, 1)) is None
    assert _saveat_array(0.5, (0.0, 2.0)).tolist() == [0.0, 0.5, 1.0, 1.5, 2.0]
    ic = pkg.forwardsens_initial_condition([[1, 0, 0]], [1.0], [[0.0]] * 4)
    assert pkg.get_probability(ic) is ic.p and len(pkg.get_sensitivity(ic)) == 4
    with pytest.raises(pkg.ArgumentError, match="Empty state list"):
        pkg.forwardsens_initial_condition([], [], [])


def test_propensity_classification(pkg):  # test/test_propensity.jl
    a = pkg.propensity(lambda x, p: p[0] * x[0])
    b = pkg.propensity(lambda t, x, p: t * x[0])
    c = pkg.propensity(lambda x, p: x[0], lambda t, p: 1.0 + t)
    assert not pkg.istimevarying(a) and pkg.istimevarying(b) and pkg.istimevarying(c)
    assert pkg.istimeseparable(c) and not pkg.istimeseparable(b)
    assert (a.kind, b.kind, c.kind) == ("ti", "joint", "sep")
