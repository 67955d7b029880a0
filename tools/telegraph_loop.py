"""A few telegraph adaptive solves (for ncu captures of the fused BDF step kernel)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
model = pkg.workloads.telegraph_model()
p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(5, 10, True))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    sol = pkg.solve(model, p0, (0.0, 300.0), alg)
print(sol.stats)
m2 = pkg.workloads.m2d_model()
sp = pkg.StateSpaceSparse(m2.stoich_matrix, [0, 0])
sp.expand_(446)
pf = pkg.FspVectorSparse.from_pairs(sp, [([0, 0], 1.0)])
print(pkg.solve(m2, pf, (0.0, 2.0), None, saveat=[2.0]).stats)
