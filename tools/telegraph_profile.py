"""cProfile of the telegraph example (BASELINE config 1) through the Python mirror: where the ~5.8 ms per solve go."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
tm = pkg.workloads.telegraph_model()
p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(5, 10, True))
ctx = pkg.Context(0)
for _ in range(5):
    pkg.solve(tm, p0, (0.0, 300.0), alg, ctx=ctx)
tt = []
for _ in range(20):
    t0 = time.perf_counter()
    sol = pkg.solve(tm, p0, (0.0, 300.0), alg, ctx=ctx)
    tt.append(time.perf_counter() - t0)
print("best %.3f ms  median %.3f ms" % (min(tt) * 1e3, sorted(tt)[10] * 1e3), sol.stats)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    pkg.solve(tm, p0, (0.0, 300.0), alg, ctx=ctx)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
