"""Where does the time of the small parity configs go?  (telegraph example).  Run on a GPU box."""
import os
import sys
import time
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
import numcme_jl_b200.transientcme as T
import numcme_jl_b200.statespace as SS
import numcme_jl_b200.fspmatrix as FM
import numcme_jl_b200.fspvector as FV

acc = defaultdict(float)
cnt = defaultdict(int)


def wrap(obj, name, label):
    f = getattr(obj, name)

    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[label] += time.perf_counter() - t0
            cnt[label] += 1
    setattr(obj, name, w)


wrap(T._Segment, "run", "segment.run")
wrap(T._Segment, "__init__", "segment.init")
wrap(SS.StateSpaceSparse, "expand_", "expand_")
wrap(SS.StateSpaceSparse, "prune_by_mass_", "prune")
wrap(SS.StateSpaceSparse, "get_states", "get_states")
wrap(SS.StateSpaceSparse, "__init__", "space.init")
wrap(SS.StateSpaceSparse, "lookup", "space.lookup")
wrap(FM.FspMatrixSparse, "__init__", "matrix.init")
wrap(FM.FspMatrixSparse, "close", "matrix.close")
wrap(T._Dist, "__init__", "dist.init")
wrap(T._Dist, "load", "dist.load")
wrap(T._Dist, "gather", "dist.gather")
wrap(T._Dist, "sinks", "dist.sinks")
wrap(FV.FspVectorSparse, "__init__", "FspVectorSparse")
T.FspVectorSparse = FV.FspVectorSparse

model = pkg.workloads.telegraph_model()
p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
for meth, name in ((None, "bdf"),):
    alg = pkg.AdaptiveFspSparse(ode_method=meth, space_adapter=pkg.RStepAdapter(5, 10, True))
    for _ in range(3):
        sol = pkg.solve(model, p0, (0.0, 300.0), alg)
    acc.clear()
    cnt.clear()
    ts = []
    NREP = 10
    for _ in range(NREP):
        t0 = time.perf_counter()
        sol = pkg.solve(model, p0, (0.0, 300.0), alg)
        ts.append(time.perf_counter() - t0)
    print(name, "telegraph wall ms", [round(t * 1e3, 2) for t in ts], sol.stats)
    tot = sum(ts) / NREP * 1e3
    print(f"mean {tot:.2f} ms; breakdown (ms per solve, calls per solve):")
    for k in sorted(acc, key=lambda k: -acc[k]):
        print(f"  {k:20s} {acc[k]/NREP*1e3:7.3f}  x{cnt[k]/NREP:.0f}")

# ---- variants: no per-step output; fixed space (one segment)
alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(5, 10, True))
for label, kw in (("every step", {}), ("saveat=[300]", {"saveat": [300.0]})):
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        sol = pkg.solve(model, p0, (0.0, 300.0), alg, **kw)
        ts.append(time.perf_counter() - t0)
    print(f"telegraph adaptive, {label}: best {min(ts)*1e3:.2f} ms median {sorted(ts)[len(ts)//2]*1e3:.2f} ms", sol.stats)
sp = pkg.StateSpaceSparse(model.stoich_matrix, [1, 0, 0])
sp.expand_(60)
pf = pkg.FspVectorSparse.from_pairs(sp, [([1, 0, 0], 1.0)])
for label, kw in (("every step", {}), ("saveat=[300]", {"saveat": [300.0]})):
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        sol = pkg.solve(model, pf, (0.0, 300.0), None, **kw)
        ts.append(time.perf_counter() - t0)
    print(f"telegraph fixed space n={sp.get_state_count()}, {label}: best {min(ts)*1e3:.2f} ms", sol.stats)
