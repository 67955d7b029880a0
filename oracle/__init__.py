"""CPU oracle for the NumCME.jl FSP right-hand-side hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy / scipy /
plain C) of the reference's algorithm for the path named in BASELINE.json.  It
may only be imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; it is the checker,
never the thing shipped or measured as the product.

The reference (pure Julia) cannot be executed in this environment (no ``julia``
binary here or on the GPU boxes), so the oracle is pinned by the reference's own
known-answer tests, restated in ``tests/test_oracle_*.py``:

* ``test/test_statespace.jl:23-34,55-63``  (state counts + exact sorted sets)
* ``test/test_fspmat.jl:41-68``            (zero column sums, separable == joint)
* ``test/sensmat/telegraph.jl:46-217``     (analytic A(t) and dA/dtheta, 4 times)
* ``test/sensmat/poisson.jl:6-21``         (sens zero-sum)
* ``test/test_solver.jl:60-99``            (conservation, saveat length)
* SURVEY.md Appendix A worked example      (index-level connectivity + dense A)

Third-party arithmetic the reference delegates to and which is NOT under
/root/reference (restated from its published semantics):

* Julia stdlib SparseArrays (``SparseArrays = "1"``, ``julia = "1.9"``,
  Project.toml:24,36): ``sparse(I,J,V,m,n)`` sums duplicates and keeps stored
  zeros; ``mul!(C,A,B,alpha,beta)`` is a serial column-oriented CSC pass.
  Pinned at the matvec boundary by the KATs above.
* DifferentialEquations 7/8 + Sundials 4 (CVODE_BDF/GMRES): step control and
  event root finding.  The reference's tests only pin conservation and
  self-consistency for transient solutions => *transient solution values:
  parity unpinned by the reference*; the oracle pins them independently with
  analytic solutions (birth-death Poisson) and tight-tolerance SciPy BDF/LSODA.
"""
