"""Measured CPU baseline of the fixed-space FSP solve (TEST / BENCH INFRASTRUCTURE ONLY).

The reference integrates `du/dt = A(t) u` (fspsolve.jl:10-41) with Sundials' CVODE_BDF(linear_solver=:GMRES)
(examples/telegraph_cme.jl:9, examples/hog1p.jl:69): variable-order BDF, matrix-free Newton-Krylov, every
right-hand side one serial `matvec!` (fspsparsematrix.jl:196-217).  Sundials is third-party and absent here
(Project.toml:35 `Sundials = "4"`), so this module times the same *structure* on the host:

  * time stepping / order selection / error control: scipy.integrate.BDF (the NDF formulation of the same family),
  * the linear systems (I - c J) d = b: restarted GMRES with a Jacobi preconditioner instead of scipy's sparse LU
    (scipy's BDF keeps its factorisation behind two replaceable callables, `lu` and `solve_lu`),
  * right-hand sides: the C restatement of the reference's serial per-term CSC passes (oracle/cpu_matvec.c).

It is a reported baseline, not a parity oracle: results are checked only for conservation and against the GPU solve's
moments by the caller.
"""
from __future__ import annotations

import time

import numpy as np
from scipy.integrate import BDF
from scipy.sparse.linalg import LinearOperator, gmres

from . import cbaseline


def bdf_gmres_fixed(OA, u0, tspan, rtol=1e-4, atol=1e-8, lin_rtol=1e-3, restart=24):
    """Integrate du/dt = A(t) u on the fixed space of the oracle matrix `OA` (FspMatrixOracle).
    Returns (u_end, stats) with stats = wall_s, steps, rhs_evals, krylov_matvecs, lin_solves, jac_evals."""
    n_rows = OA.rowcount
    stats = {"rhs_evals": 0, "krylov_matvecs": 0, "lin_solves": 0, "jac_evals": 0}
    terms = cbaseline.CscTerms(OA.terms_at(tspan[0]))
    out = np.empty(n_rows)

    def coefs(t):
        return np.array([c for c, _ in OA.terms_at(t)], dtype=np.float64) if (OA.sep_ids or OA.joint_ids) else terms.coef

    def rhs(t, u):
        stats["rhs_evals"] += 1
        terms.coef[:] = coefs(t)
        terms.matvec(np.ascontiguousarray(u), out)
        return out.copy()

    def jac(t, u):
        stats["jac_evals"] += 1
        J = None
        for c, M in OA.terms_at(t):
            J = c * M if J is None else J + c * M
        return J.tocsr()

    t_wall = time.perf_counter()
    solver = BDF(rhs, tspan[0], np.asarray(u0, dtype=np.float64), tspan[1], rtol=rtol, atol=atol, jac=jac)

    class _Sys:                                   # what `lu(I - c J)` returns: the operator and its Jacobi preconditioner
        def __init__(self, M):
            self.M = M.tocsr()
            d = self.M.diagonal()
            self.dinv = 1.0 / np.where(d != 0.0, d, 1.0)

    def lu(M):
        solver.nlu += 1
        return _Sys(M)

    def solve_lu(S, b):
        stats["lin_solves"] += 1

        def mv(v):
            stats["krylov_matvecs"] += 1
            return S.M @ v
        x, info = gmres(LinearOperator(S.M.shape, matvec=mv, dtype=np.float64), b, rtol=lin_rtol, atol=0.0, restart=restart,
                        maxiter=20, M=LinearOperator(S.M.shape, matvec=lambda v: S.dinv * v, dtype=np.float64))
        return x

    solver.lu, solver.solve_lu, solver.LU = lu, solve_lu, None
    steps = 0
    while solver.status == "running":
        msg = solver.step()
        steps += 1
        if solver.status == "failed":
            raise RuntimeError(f"cpu BDF/GMRES baseline failed: {msg}")
    stats.update({"wall_s": time.perf_counter() - t_wall, "steps": steps})
    return solver.y.copy(), stats
