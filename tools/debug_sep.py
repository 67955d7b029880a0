import faulthandler, sys, os, time
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "40")), exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
S = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]]).T
f = lambda t, x, p: 0.3 * x[0] + (0.4 * x[1] if 5.0 < t < 10.0 else 0.0 * x[1])   # 0.4 < the decay rate 0.5: no blow-up
props = [pkg.propensity(lambda x, p: 4.0 + 0.0 * x[0]), pkg.propensity(lambda x, p: 0.2 * x[0]),
         pkg.propensity(f), pkg.propensity(lambda x, p: 0.5 * x[1])]
model = pkg.CmeModel(S, props, [])
p0 = pkg.FspVectorSparse([[3, 2]], [1.0])
alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(15, 10, True))
kw = dict(saveat=[4.0, 8.0, 12.0], fsptol=1e-6, odertol=1e-7, odeatol=1e-12)
which = sys.argv[1]
t0 = time.time()
if which == "exact":
    sol = pkg.solve(model, p0, (0.0, 12.0), alg, detect_separable=False, verbose=True, **kw)
else:
    sol = pkg.solve(model, p0, (0.0, 12.0), alg, verbose=True, **kw)
print(which, "done", time.time() - t0, sol.stats)
