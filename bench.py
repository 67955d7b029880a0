#!/usr/bin/env python
"""bench.py -- FSP matvec throughput on the M-3D workload (BASELINE.json config 5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--levels L]

A "step" is one fused FSP matvec y = A(t) x over the whole state space (n = 10 039 316 states,
R = 6, time-varying variant: 1 041 MB algorithmic bytes per step, far larger than the 126 MB L2).
`value` = algorithmic GB/s with x, y and A resident in HBM; `e2e` = the same metric through the
host-buffer C-ABI entry point (pinned host x -> H2D -> kernel -> D2H y inside the timed region).
At N > 1 (torchrun, one rank per GPU) the SAME global operator is row-sharded over the ranks (strong scaling):
the boundary rows pull their halo of x from the neighbours' HBM over NVLink (CUDA IPC + flag epochs; grouped
ncclSend/ncclRecv as fallback) concurrently with the halo-free rows, value = global bytes / max-over-ranks time.  The JSON line also carries `solve`: wall time of a fixed-space FSP solve on the same operator with the
native device-resident integrator (the second half of BASELINE.json's metric).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION/INFO; the contract is ONE JSON line on stdout
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and "NCCL_DEBUG_FILE" not in os.environ:
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"

METRIC = "fsp_matvec_hbm_gbs"
UNIT = "GB/s"
CPU_SAMPLE_LEVELS = 180          # M-3D at L=180: 1 004 731 states (~104 MB) -- bounded CPU sample


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(levels, steps, warmup, time_varying=True):
    """The reference's CPU matvec (serial CSC passes, one per term) restated in C (oracle/cpu_matvec.c),
    on a bounded sample of the workload.  Returns (GB/s, info dict)."""
    import __graft_entry__ as g
    pkg = g.load_package()
    from oracle import cbaseline
    from oracle.fspmatrix import FspMatrixOracle
    from oracle.statespace import StateSpaceOracleFast

    model = pkg.workloads.m3d_model(time_varying=time_varying)
    osp = StateSpaceOracleFast(model.stoich_matrix, [0, 0, 0])
    osp.expand(levels)
    OA = FspMatrixOracle(osp, model.propensities, model.parameters)
    terms = cbaseline.CscTerms(OA.terms_at(2.5))
    nbytes = OA.algorithmic_bytes()
    rng = np.random.default_rng(0)
    v = rng.random(OA.rowcount)
    v /= v.sum()
    out = np.empty_like(v)
    for _ in range(warmup):
        terms.matvec(v, out)
    t0 = time.perf_counter()
    for _ in range(steps):
        terms.matvec(v, out)
    dt = (time.perf_counter() - t0) / steps
    # additional "best CPU" line (not how the reference computes): one fused CSR matrix, row-parallel over all host
    # threads OpenMP gives this process (torchrun pins OMP_NUM_THREADS=1; the plain N=1 run gets every core)
    best_omp = None
    try:
        if levels > 200:       # the extra line is for the bounded sample only (summing 70 M-entry scipy matrices is slow)
            raise RuntimeError("skipped at full size")
        fused = None
        for c, m in OA.terms_at(2.5):
            fused = c * m if fused is None else fused + c * m
        omp = cbaseline.CsrOmp(fused.tocsr())
        for _ in range(2):
            omp.matvec(v, out)
        t1 = time.perf_counter()
        for _ in range(max(steps, 5)):
            omp.matvec(v, out)
        dto = (time.perf_counter() - t1) / max(steps, 5)
        best_omp = {"value": nbytes / dto / 1e9, "unit": UNIT, "cores": cbaseline.num_threads(), "ms_per_matvec": dto * 1e3,
                    "what": "fused CSR (all terms summed), OpenMP row-parallel, 32-bit column indices"}
        terms.matvec(v, out)
    except Exception as exc:   # the baseline arm must not fail the bench
        best_omp = {"error": repr(exc)}
    info = {"kind": "port", "cores": 1, "unit": UNIT, "value": nbytes / dt / 1e9, "best_omp": best_omp,
            "states": osp.get_state_count(), "algorithmic_bytes": nbytes,
            "sample": f"M-3D TV at L={levels} (n={osp.get_state_count()}, {nbytes/1e6:.1f} MB algorithmic bytes per matvec), "
                      f"{steps} serial CSC matvecs (one pass per term, Int64 indices) = SparseArrays.mul! restated in C; "
                      f"{dt*1e3:.2f} ms per matvec; host has {os.cpu_count()} cores, the reference path uses 1"}
    return nbytes / dt / 1e9, dt, info, (OA, v, out)


CPU_SOLVE_LEVELS = 120           # M-3D at L=120: 302 621 states; holds all but ~1e-8 of the mass at t = 10


def cpu_solve_arm(pkg, t_end, n_full):
    """MEASURED CPU baseline of the fixed-space solve (BASELINE.md section 2 `cpu_solve`): the structure of the
    reference's CVODE_BDF(linear_solver=:GMRES) run (variable-order BDF, matrix-free Jacobi-GMRES, one serial per-term
    CSC matvec per right-hand side) on a bounded sample of the workload -- oracle/cpu_solve.py."""
    from oracle.cpu_solve import bdf_gmres_fixed
    from oracle.fspmatrix import FspMatrixOracle
    from oracle.statespace import StateSpaceOracleFast
    model = pkg.workloads.m3d_model(time_varying=True)
    osp = StateSpaceOracleFast(model.stoich_matrix, [0, 0, 0])
    osp.expand(CPU_SOLVE_LEVELS)
    OA = FspMatrixOracle(osp, model.propensities, model.parameters)
    ns = osp.get_state_count()
    u0 = np.zeros(OA.rowcount)
    u0[0] = 1.0
    u, st = bdf_gmres_fixed(OA, u0, (0.0, t_end), rtol=1e-4, atol=1e-8)
    X = osp.states_array()
    return {"kind": "port", "cores": 1, "unit": "s", "value": st["wall_s"], "steps": st["steps"],
            "rhs_evals": st["rhs_evals"], "krylov_matvecs": st["krylov_matvecs"], "mass": float(u.sum()),
            "mean_x": [float((u[:ns] * X[:, k]).sum()) for k in range(3)],
            "scaled_to_full_size_s": st["wall_s"] * n_full / ns,
            "sample": f"M-3D TV at L={CPU_SOLVE_LEVELS} (n={ns} of the {n_full} states; the same initial condition, horizon and "
                      f"tolerances; the truncated tail holds < 1e-7 of the mass, see mean_x), scipy BDF (NDF) stepping with "
                      f"matrix-free Jacobi-GMRES(24) linear solves and the C restatement of the serial per-term CSC matvec "
                      f"as right-hand side = the structure of the reference's CVODE_BDF(GMRES) path; measured wall time on "
                      f"1 of {os.cpu_count()} host cores; scaled_to_full_size_s = value x n_full / n (cost per step is linear in n)"}


def cpu_sens_arm(model, x0, levels, steps, warmup, t=2.5):
    """The reference's CPU sensitivity matvec (sensfspmatrixsparse.jl:97-142) restated with the C port of
    SparseArrays.mul!: (P+1) full matvec! passes over A's terms, one pass of the summed time-invariant derivative
    matrix per parameter, two passes per separable (reaction, parameter) entry -- all serial, Int64 indices."""
    import ctypes
    from oracle import cbaseline
    from oracle.sensmatrix import SensFspMatrixOracle
    from oracle.statespace import StateSpaceOracleFast
    cm = model.cmemodel
    osp = StateSpaceOracleFast(cm.stoich_matrix, x0)
    osp.expand(levels)
    OS = SensFspMatrixOracle(osp, cm.propensities, model.propensity_gradients, model.gradient_sparsity_patterns, cm.parameters)
    A = OS.fspmatrix
    N, P, th = A.rowcount, OS.parameter_count, cm.parameters
    terms = cbaseline.CscTerms(A.terms_at(t))
    lib = cbaseline.lib()
    f64p, i64p = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)

    def csc(M):
        M = M.tocsc()
        return (np.ascontiguousarray(M.indptr, dtype=np.int64), np.ascontiguousarray(M.indices, dtype=np.int64),
                np.ascontiguousarray(M.data, dtype=np.float64))
    dti = [csc(M) for M in OS.timeinvariant_matdiffs]
    dsep = [(ip, float(A.propensities[r - 1].tfactor(t, th)), float(OS.gradients[r - 1].tfactor_pardiffs[ip](t, th)),
             csc(dM), csc(A.separabletv_factormatrices[j])) for (ip, r, j, dM) in OS.sep_entries]

    def mul(M, x, alpha, y):
        lib.ncme_oracle_csc_mul(ctypes.c_int64(N), ctypes.c_int64(N), M[0].ctypes.data_as(i64p), M[1].ctypes.data_as(i64p),
                                M[2].ctypes.data_as(f64p), x.ctypes.data_as(f64p), ctypes.c_double(alpha),
                                ctypes.c_double(1.0), y.ctypes.data_as(f64p))
    rng = np.random.default_rng(0)
    v = rng.random(N * (P + 1))
    out = np.empty_like(v)
    blocks = [v[b * N:(b + 1) * N] for b in range(P + 1)]
    oblocks = [out[b * N:(b + 1) * N] for b in range(P + 1)]

    def step():
        terms.matvec(blocks[0], oblocks[0])
        for ip in range(P):
            terms.matvec(blocks[ip + 1], oblocks[ip + 1])
            if dti[ip][2].size:
                mul(dti[ip], blocks[0], 1.0, oblocks[ip + 1])
            for (jp, c, dc, dM, FM) in dsep:
                if jp == ip:
                    mul(dM, blocks[0], c, oblocks[ip + 1])
                    mul(FM, blocks[0], dc, oblocks[ip + 1])
    for _ in range(warmup):
        step()
    ref = OS.matvec(t, v)
    err = float(np.abs(out - ref).max() / np.abs(ref).max())
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    # B_sens of this sample from the oracle's own structures
    nb = A.algorithmic_bytes() - 16 * N + 16 * N * (P + 1)
    for M in OS.timeinvariant_matdiffs:
        if M.nnz:
            nb += 8 * M.nnz + 4 * (M.nnz - A.n)
    for (_, _, _, dM) in OS.sep_entries:
        nb += 8 * dM.nnz + 4 * (dM.nnz - A.n)
    return {"kind": "port", "cores": 1, "unit": UNIT, "value": nb / dt / 1e9, "ms_per_matvec": dt * 1e3,
            "check_vs_oracle_relerr": err,
            "sample": f"L={levels} (n={A.n}, P={P}, {nb/1e6:.1f} MB B_sens), {steps} sensitivity matvecs as the reference "
                      f"computes them: (P+1) x (1+n_sep) serial CSC passes over A + one pass per derivative matrix; "
                      f"host has {os.cpu_count()} cores, the reference path uses 1"}


def telegraph_cpu_ms(pkg):
    from oracle.solve import RStepAdapterOracle, solve_adaptive
    tm = pkg.workloads.telegraph_model()
    res = {}
    for m in ("BDF", "LSODA"):
        tt = []
        for _ in range(3):
            tq = time.perf_counter()
            solve_adaptive(tm.stoich_matrix, tm.propensities, tm.parameters, [[1, 0, 0]], [1.0], (0.0, 300.0),
                           RStepAdapterOracle(5, 10, True), method=m)
            tt.append(time.perf_counter() - tq)
        res[m] = min(tt) * 1e3
    return res


def sens_leg(pkg, ctx, which, levels, steps, warmup, cpu=True, rows=0):
    """K2 measurement (SURVEY 8(d) "Algorithmic bytes, sensitivity matvec"): one step = one fused block matvec
    Y = [A p; A s_ip + dA_ip p] over all P + 1 blocks."""
    import torch
    if which == "hog1p":
        th = list(pkg.workloads.HOG1P_THETA)
        th[2] = 3.2e4
        model, x0, tt = pkg.workloads.hog1p_sens_model(th), [1, 0, 0, 0, 0, 0], 120.0
        cpu_levels = min(levels, 400)
    else:
        model, x0, tt = pkg.workloads.m3d_sens_model(), [0, 0, 0], 2.5
        cpu_levels = min(levels, 100)
    cm = model.cmemodel
    t0 = time.perf_counter()
    space = pkg.StateSpaceSparse(cm.stoich_matrix, x0, ctx=ctx)
    space.expand_(levels)
    t1 = time.perf_counter()
    SA = pkg.ForwardSensFspMatrixSparse(model, space)
    t2 = time.perf_counter()
    if rows:
        SA.set_tuning(rows)
    st = SA.stats()
    n, N, P = SA.fspmatrix.n, SA.fspmatrix.rowcount, SA.parameter_count
    rng = np.random.default_rng(0)
    X = pkg.DeviceVector.from_host(ctx, rng.random(N * (P + 1)))
    Y = pkg.DeviceVector(ctx, N * (P + 1))
    for _ in range(max(warmup, 3)):
        pkg.matvec_(Y, tt, SA, X)
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pkg.matvec_(Y, tt, SA, X)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = ctx.launch_count() - l0
    peak, peak_src = measured_peak()
    gbs = st["algorithmic_bytes"] / (ms * 1e-3) / 1e9
    out = {"metric": "fsp_sens_matvec_hbm_gbs", "value": gbs, "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "gpu_launches": int(launches),
           "config": {"workload": f"{which} forward-sensitivity block matvec, L={levels}", "states": n, "parameters": P,
                      "entries": len(SA.entries), "slots": SA.fspmatrix.stats()["nterms"], "blocks": P + 1,
                      "algorithmic_bytes_per_step": st["algorithmic_bytes"], "streamed_bytes_per_step": st["device_bytes"],
                      "expand_s": round(t1 - t0, 3), "assemble_s": round(t2 - t1, 3),
                      "l2_policy": "operands larger than the 126 MB L2" if st["device_bytes"] > 4e8 else
                                   "operands fit the 126 MB L2 (small reference configuration)"},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": UNIT, "frac": gbs / peak, "traffic": None,
                        "peak_source": peak_src, "frac_of_nominal_8TBs": gbs / 8000.0, "kernel": "k_sens_matvec",
                        "streamed_gbs": st["device_bytes"] / (ms * 1e-3) / 1e9,
                        "note": "achieved = B_sens (SURVEY 8(d)) / CUDA-event time per launch"}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr) and which == "m3d":
        try:
            out["roofline"]["traffic"] = json.load(open(tr)).get("k_sens_matvec_dram_bytes_per_launch")
        except Exception:
            pass
    SA.close()
    if cpu:
        out["cpu_baseline"] = cpu_sens_arm(model, x0, cpu_levels, 3 if which == "m3d" else 10, 1, tt)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--levels", type=int, default=None, help="expansion depth of the M-3D simplex (default 390)")
    ap.add_argument("--rows", type=int, default=0, help="matvec kernel variant: rows per thread (0 = auto)")
    ap.add_argument("--pipe", default="", help="experiments: ROWS,STAGES of the shared-memory pipelined matvec kernel")
    ap.add_argument("--workload", default="matvec", choices=["matvec", "sens", "sens-hog1p"],
                    help="matvec: the headline K1 line (default); sens / sens-hog1p: the K2 line (M-3D P=6 / Hog1p P=14)")
    ap.add_argument("--sens-rows", type=int, default=0, help="K2 kernel variant: rows per thread (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-solve", action="store_true", help="skip the fixed-space solve leg")
    ap.add_argument("--solve-t", type=float, default=10.0, help="horizon of the solve leg")
    ap.add_argument("--solve-method", default="both", choices=["dp5", "bdf", "both"], help="native integrator(s) timed in the solve leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    import __graft_entry__ as g
    pkg = g.load_package()
    levels = args.levels or pkg.workloads.M3D_LEVELS
    config = {"workload": f"M-3D three-species birth-death FSP, time-varying variant, simplex L={levels}",
              "levels": levels, "reactions": 6, "l2_policy": "inputs (matrix ~1 GB) larger than the 126 MB L2",
              "parallelism": "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        # the SAME configuration as our arm (M-3D TV at the full L): building it through the oracle's numpy expand!
        # takes ~1.5 min and ~9 GB of host memory; every timed step is one full-size serial matvec (~0.1-0.2 s)
        gbs, dt, info, _ = cpu_reference_arm(levels, K, W)
        config.update({"states": info.pop("states"), "algorithmic_bytes_per_step": info.pop("algorithmic_bytes")})
        line = {"impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": info,
                "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if args.workload != "matvec":
        if rank != 0:
            return
        ctx = pkg.Context(local_rank)
        ctx.use_torch_stream()
        which = "hog1p" if args.workload == "sens-hog1p" else "m3d"
        lv = args.levels or (600 if which == "hog1p" else pkg.workloads.M3D_LEVELS)
        sampler = ClockSampler(local_rank)
        sampler.start()
        line = sens_leg(pkg, ctx, which, lv, K, W, cpu=not args.no_cpu, rows=args.sens_rows)
        line.update({"n_gpus": 1, "warmup": W, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                     "dtype": "f64", "data": "synthetic", "clocks": sampler.stop()})
        print(json.dumps(line))
        return
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ctx = pkg.Context(local_rank)
    ctx.use_torch_stream()
    comm = pkg.Comm.from_torch(ctx) if world > 1 else None
    config["parallelism"] = "single GPU" if world == 1 else (
        f"row-sharded over {world} GPUs (contiguous state blocks; halo of x pulled from the neighbours' HBM over NVLink "
        f"by the boundary rows (CUDA IPC + flag epochs), NCCL send/recv fallback; state space replicated)")

    # ---- build the workload through the product path: GPU expand, host propensities, GPU assembly
    model = pkg.workloads.m3d_model(time_varying=True)
    t_build = time.perf_counter()
    space = pkg.StateSpaceSparse(model.stoich_matrix, [0, 0, 0], ctx=ctx)
    space.expand_(levels)
    t_expand = time.perf_counter() - t_build
    A = pkg.FspMatrixSparse(space, model.propensities, parameters=model.parameters, comm=comm)
    t_assemble = time.perf_counter() - t_build - t_expand
    if args.rows:
        A.set_tuning(args.rows)
    if args.pipe:
        A.set_pipe(*[int(v) for v in args.pipe.split(",")])
    n = space.get_state_count()
    R = 6
    info = A.shard_info()
    nloc = info["row_hi"] - info["row_lo"]
    st = A.stats()
    # algorithmic bytes of the GLOBAL operator; per-rank stats cover the local rows only
    nb_t = torch.tensor([st["algorithmic_bytes"] - 16 * R], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(nb_t)
    nbytes = int(nb_t.item()) + 16 * R
    config.update({"states": n, "algorithmic_bytes_per_step": nbytes, "expand_s": round(t_expand, 3),
                   "assemble_s": round(t_assemble, 3), "halo_doubles": [info["halo_lo"], info["halo_hi"]]})
    rng = np.random.default_rng(0)
    xg = rng.random(n + R)
    xg /= xg.sum()
    xh = np.concatenate([xg[info["row_lo"]:info["row_hi"]], xg[n:]])
    x = pkg.ShardedVector(A, fill=xh, register=True)     # peer-memory halo when sharded (collective registration)
    y = pkg.DeviceVector(ctx, nloc + R)
    tt = 2.5

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: one step = one fused matvec over the whole (sharded) state space
    # The clock sampler (a fork + exec of nvidia-smi: milliseconds, different on every rank) starts BEFORE the barrier.
    # Ranks are then aligned on the DEVICE: an NCCL all-reduce on the compute stream followed by a few untimed pre-roll
    # matvecs, so that every GPU's queue already holds work when e0 is reached and host-side launch skew between the
    # ranks (which the boundary rows' flag wait would otherwise charge to the faster rank) stays outside the timed region.
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W):
        A.matvec_local_(y, tt, x.v)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    if world > 1:
        dist.all_reduce(torch.zeros(1, device=f"cuda:{local_rank}"))
    PRE = 8 if world > 1 else 0
    for _ in range(PRE):
        A.matvec_local_(y, tt, x.v)
    l0 = ctx.launch_count()
    ev[0].record()
    for i in range(K):
        A.matvec_local_(y, tt, x.v)
        ev[i + 1].record()
    barrier()
    launches = ctx.launch_count() - l0
    ms = ev[0].elapsed_time(ev[K]) / K
    per_step = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(K))
    step_stats = {"median_ms": per_step[K // 2], "min_ms": per_step[0], "max_ms": per_step[-1]}
    # nvidia-smi delivers a sample every ~100 ms and the timed region may be shorter: keep issuing the SAME kernel
    # (untimed) until at least 5 samples under this load are in, so that clocks / throttle reasons are observed
    ms_hold = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(ms_hold, op=dist.ReduceOp.MAX)
    n_hold = int(min(20000, max(0, (700.0 - K * float(ms_hold)) / float(ms_hold))))   # same count on every rank (lockstep)
    for _ in range(n_hold):
        A.matvec_local_(y, tt, x.v)
    barrier()
    clocks = sampler.stop()
    clocks["hold_steps_untimed"] = n_hold

    # ---- end to end through host buffers (pinned): H2D of x, kernel, D2H of y inside the timed region
    xp = torch.empty(nloc + R, dtype=torch.float64).pin_memory()
    yp = torch.empty(nloc + R, dtype=torch.float64).pin_memory()
    xp.numpy()[:] = xh
    Ke = max(3, min(K, 20))

    def e2e_step():
        if world == 1:
            pkg.matvec_(yp.numpy(), tt, A, xp.numpy())        # ncme_matvec_host: H2D + kernel + D2H
        else:
            x.v.upload(xp.numpy())                            # pinned host -> this rank's shard
            pkg.matvec_(y, tt, A, x.v)
            y.download_into(yp.numpy())                       # shard -> pinned host
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    ms_e2e = (time.perf_counter() - t0) / Ke * 1e3
    checksum = float(yp.numpy()[:nloc].sum())
    # what the host link of THIS box sustains for the same traffic (one pinned H2D and one D2H of the vector's size in
    # flight at once, plain copies): the floor of any host-buffer matvec, so the e2e number can be read against it
    link = None
    if world == 1:
        d_a = torch.empty(nloc + R, dtype=torch.float64, device=f"cuda:{local_rank}")
        d_b = torch.empty(nloc + R, dtype=torch.float64, device=f"cuda:{local_rank}")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def both():
            cur = torch.cuda.current_stream()
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                d_a.copy_(xp, non_blocking=True)
            with torch.cuda.stream(s2):
                yp.copy_(d_b, non_blocking=True)
            cur.wait_stream(s1)
            cur.wait_stream(s2)
        for _ in range(3):
            both()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            both()
        torch.cuda.synchronize()
        floor_ms = (time.perf_counter() - t0) / 10 * 1e3
        link = {"both_directions_floor_ms": floor_ms, "gbs_per_direction": 8 * (nloc + R) / (floor_ms * 1e-3) / 1e9,
                "e2e_frac_of_link_floor": floor_ms / ms_e2e,
                "note": "plain pinned copies of the vector's size, H2D and D2H at once on two streams, this box"}
        del d_a, d_b

    # ---- second half of the metric: full FSP solve wall time (fixed M-3D space, p0 = delta at the origin)
    solve_info = None
    if not args.no_solve:
        from numcme_jl_b200.transientcme import _Dist, _Segment
        runs = {}
        for name, code in (("dp5", 0), ("bdf", 1)):
            if args.solve_method not in (name, "both"):
                continue
            pfull = pkg.DeviceVector.zeros(ctx, n)
            pfull.view(0, 1).upload(np.ones(1))
            d = _Dist(A, comm)
            d.load(pfull, np.zeros(R))
            seg = _Segment(A, 1e-4, 1e-8, code)
            # untimed warm-up segment (workspace allocation, lazy kernel loading), like the matvec's warm-up steps
            seg.run(d.u.v, 0.0, min(0.02, args.solve_t), saveat=[min(0.02, args.solve_t)])
            # two full solves from p0, the faster one is reported (the first may still grow the Krylov workspace: the short
            # warm-up segment needs fewer basis vectors than the late steps of the full horizon); both are listed
            walls = []
            for _rep in range(2):
                d.load(pfull, np.zeros(R))
                seg.saved_u = []
                barrier()
                tw = time.perf_counter()
                sstats = seg.run(d.u.v, 0.0, args.solve_t, saveat=[args.solve_t])
                barrier()
                walls.append(time.perf_counter() - tw)
            wall = min(walls)
            uu = seg.saved_u[-1]
            runs[name] = {"wall_s": wall, "wall_s_runs": walls, "steps": int(sstats.steps), "rejected": int(sstats.rejected),
                          "rhs_evals": int(sstats.rhs_evals), "launches": int(sstats.launches),
                          "mass": float(uu.sum()), "sinks": float(uu[n:].sum()),
                          "mean_x": [float((uu[:n] * space.get_states()[:, k]).sum()) for k in range(3)] if rank == 0 else None}
            del d, seg
        best = min(runs, key=lambda k: runs[k]["wall_s"])
        solve_info = dict(runs[best])
        solve_info.update({"tspan": [0.0, args.solve_t], "odertol": 1e-4, "odeatol": 1e-8,
                           "warmup": "one untimed segment of horizon 0.02 per method, then two full solves from p0: the faster is wall_s, both in wall_s_runs",
                           "method": {"dp5": "native Dormand-Prince 5(4)", "bdf": "native BDF/NDF + Jacobi-GMRES"}[best] +
                                     ", device-resident", "all_methods": runs})

    # ---- the same solve through the PUBLIC API, end to end (SURVEY 8(d) "solve metric": wall time of `solve` incl. the
    # state-space build, propensity evaluation, matrix assembly, the adaptive loop and the output download; only
    # context / NCCL init excluded): solve(model, p0, tspan, AdaptiveFspSparse(RStepAdapter(L, 10, false)); saveat)
    if solve_info is not None:
        try:
            p0 = pkg.FspVectorSparse([[0, 0, 0]], [1.0])
            alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(levels, 10, False))
            barrier()
            tw = time.perf_counter()
            sol = pkg.solve(model, p0, (0.0, args.solve_t), alg, saveat=[args.solve_t], odertol=1e-4, odeatol=1e-8,
                            ctx=ctx, comm=comm)
            barrier()
            api_wall = time.perf_counter() - tw
            if world > 1:
                tw_t = torch.tensor([api_wall], dtype=torch.float64, device=f"cuda:{local_rank}")
                dist.all_reduce(tw_t, op=dist.ReduceOp.MAX)
                api_wall = float(tw_t)
            pl = sol.p[-1]
            solve_info["solve_api_wall_s"] = api_wall
            solve_info["solve_api"] = {"call": f"solve(model, p0, (0, {args.solve_t}), AdaptiveFspSparse(nothing, RStepAdapter({levels}, 10, false)); saveat=[{args.solve_t}], odertol=1e-4, odeatol=1e-8)",
                                       "steps": int(sol.stats["steps"]), "rhs_evals": int(sol.stats["rhs_evals"]),
                                       "adapts": int(sol.stats["adapts"]), "final_states": int(sol.stats["final_states"]),
                                       "breakdown_s": sol.stats.get("breakdown_s"),
                                       "mass": float(pl.values.sum() + sol.sinks[-1].sum()),
                                       "mean_x": [float((pl.values * pl.states[:, k]).sum()) for k in range(3)],
                                       "includes": "expand! of the simplex, host propensity evaluation, matrix assembly, "
                                                   "the integration, download of the final distribution and its states"}
            del sol, pl
        except Exception as exc:
            solve_info["solve_api"] = {"error": repr(exc)}

    per_rank_ms = [ms]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, {"ms": ms, **step_stats})
        per_rank_ms = [g_["ms"] for g_ in gathered]
        step_stats = {"median_ms": max(g_["median_ms"] for g_ in gathered), "min_ms": max(g_["min_ms"] for g_ in gathered),
                      "max_ms": max(g_["max_ms"] for g_ in gathered)}
        tms = torch.tensor([ms, ms_e2e, solve_info["wall_s"] if solve_info else 0.0], dtype=torch.float64,
                           device=f"cuda:{local_rank}")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tms[0]), float(tms[1])
        if solve_info:
            solve_info["wall_s"] = float(tms[2])
        cs = torch.tensor([checksum], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(cs)
        checksum = float(cs)

    value = nbytes / (ms * 1e-3) / 1e9                      # whole job: global operator bytes / max-over-ranks time
    e2e_value = nbytes / (ms_e2e * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    per_gpu = value / world
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
            "per_step": {**step_stats, "per_rank_ms_per_step": per_rank_ms,
                         "note": "ms_per_step = (event after step K - event before step 1) / K, max over ranks; "
                                 "median/min/max of the K per-step event intervals (max over ranks of each)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (nloc + R) * world,
                    "d2h_bytes_per_step": 8 * (nloc + R) * world, "ms_per_step": ms_e2e, "checksum_sum_y_states": checksum},
            "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": UNIT, "frac": per_gpu / peak,
                         "traffic": None, "peak_source": peak_src, "frac_of_nominal_8TBs": per_gpu / 8000.0,
                         "kernel": "k_fsp_matvec", "launch_us": ms * 1e3 / max(launches / K, 1),
                         "note": "achieved = algorithmic bytes (SURVEY 8(d)) / CUDA-event time per launch, per GPU"}}
    if abs(per_gpu / peak) > 1.5:
        line["roofline"]["timing_suspect"] = True
    if link:
        line["e2e"]["link"] = link
    if solve_info:
        line["solve"] = solve_info
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr) and world == 1:
        try:
            line["roofline"]["traffic"] = json.load(open(tr)).get("k_fsp_matvec_dram_bytes_per_launch")
        except Exception:
            pass
    if rank == 0 and world == 1 and not args.no_cpu:
        gbs, dt, info_cpu, _ = cpu_reference_arm(CPU_SAMPLE_LEVELS, 20, 2)
        line["cpu_baseline"] = info_cpu
        if solve_info:
            try:
                solve_info["cpu_baseline"] = cpu_solve_arm(pkg, args.solve_t, n)
            except Exception as exc:   # the baseline leg must not fail the bench
                solve_info["cpu_baseline"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_solve:
        # BASELINE.json configs[0] (the reference's own CPU-runnable case and its only published number):
        # examples/telegraph_cme.jl, adaptive FSP solve over t in [0, 300], defaults
        try:
            tm = pkg.workloads.telegraph_model()
            p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
            alg = pkg.AdaptiveFspSparse(ode_method=None, space_adapter=pkg.RStepAdapter(5, 10, True))
            for _ in range(3):
                pkg.solve(tm, p0, (0.0, 300.0), alg, ctx=ctx)
            tt_ = []
            for _ in range(10):
                tq = time.perf_counter()
                sol = pkg.solve(tm, p0, (0.0, 300.0), alg, ctx=ctx)
                tt_.append(time.perf_counter() - tq)
            line["parity_configs"] = {"telegraph_adaptive_solve_ms": {
                "best": min(tt_) * 1e3, "median": sorted(tt_)[len(tt_) // 2] * 1e3, "steps": sol.stats["steps"],
                "launches": sol.stats["launches"], "adapts": sol.stats["adapts"], "final_states": sol.stats["final_states"],
                "cpu_oracle_ms": telegraph_cpu_ms(pkg),
                "cpu_oracle_note": "oracle.solve.solve_adaptive (scipy BDF / LSODA restatement of fspsolve.jl:105-197), best of 3 each, on this host, 1 core",
                "reference_published_ms": 5.454, "reference_hardware": "Apple M1, Julia, docs/src/examples/telegraph.md:84-90",
                "note": "through the Python mirror of solve(); fused-step BDF, one kernel launch per step"}}
        except Exception as exc:
            line["parity_configs"] = {"error": repr(exc)}
    if comm is not None:
        line["config"]["halo_transport"] = comm.info()
        x.unregister()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
