"""Fused-step BDF (bdf_fused.cu) vs launch-per-operation BDF (bdf.cu) vs DP5 as a function of the state count
(2-D birth-death model, fixed space, t in [0, 2]).  Run on a GPU box."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
model = pkg.workloads.m2d_model()
for levels in ([int(v) for v in sys.argv[1:]] or [44, 140, 446, 1000, 1413, 2000, 2800]):
    sp = pkg.StateSpaceSparse(model.stoich_matrix, [0, 0])
    sp.expand_(levels)
    p0 = pkg.FspVectorSparse.from_pairs(sp, [([0, 0], 1.0)])
    row = {}
    for name, m in (("fused", pkg.NativeBDFFused()), ("classic", pkg.NativeBDFClassic()), ("dp5", pkg.NativeRK45())):
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            sol = pkg.solve(model, p0, (0.0, 2.0), m, saveat=[2.0], odertol=1e-4, odeatol=1e-8)
            best = min(best, sol.stats["wall_s"])
        row[name] = (round(best * 1e3, 2), sol.stats["steps"], sol.stats["rhs_evals"])
    print(sp.get_state_count(), row, flush=True)
