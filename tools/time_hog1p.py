"""Hog1p phase 1 (adaptive, 15 adapts, ~1e5 states) with and without the incremental matrix rebuild.  GPU box."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
import numcme_jl_b200.fspmatrix as FM
th = list(pkg.workloads.HOG1P_THETA)
alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(10, 20, True))
m0 = pkg.workloads.hog1p_model(th)
p0 = pkg.FspVectorSparse([[1, 0, 0, 0, 0, 0]], [1.0])
for label, thr in (("warm-up", 2048), ("incremental", 2048), ("full rebuilds", 1 << 60), ("incremental", 2048), ("full rebuilds", 1 << 60)):
    FM.INCREMENTAL_MIN_STATES = thr
    t0 = time.perf_counter()
    s0 = pkg.solve(m0, p0, (0.0, 8 * 3600.0), alg, saveat=[8 * 3600.0], fsptol=1e-6, odeatol=1e-14, odertol=1e-6)
    print(f"{label:14s} {time.perf_counter()-t0:.3f} s", {k: s0.stats[k] for k in ("steps", "adapts", "incremental_builds", "final_states")}, flush=True)
