"""Multi-rank parity check, launched by tests/test_gpu_multi.py through torchrun (one rank per GPU):
row-sharded matvec / solve against the single-GPU path computed on the same rank."""
import faulthandler
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    faulthandler.enable()
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = pkg.Context(local)
    comm = pkg.Comm.from_torch(ctx)
    levels = int(os.environ.get("NCME_DIST_LEVELS", "60"))

    # ---- matvec: M-3D TV variant, sharded vs full
    model = pkg.workloads.m3d_model(time_varying=True)
    space = pkg.StateSpaceSparse(model.stoich_matrix, [0, 0, 0], ctx=ctx)
    space.expand_(levels)
    n = space.get_state_count()
    A_full = pkg.FspMatrixSparse(space, model.propensities, parameters=model.parameters)
    A_sh = pkg.FspMatrixSparse(space, model.propensities, parameters=model.parameters, comm=comm)
    info = A_sh.shard_info()
    lo, hi = info["row_lo"], info["row_hi"]
    assert info["nranks"] == world and info["n_global"] == n
    cuts = pkg.shard_bounds(n, world)
    assert (lo, hi) == (cuts[rank], cuts[rank + 1]), (lo, hi, cuts)
    rng = np.random.default_rng(0)
    x = rng.random(n + 6)
    y_ref = pkg.matvec(2.5, A_full, x)
    xs = pkg.ShardedVector(A_sh, fill=np.concatenate([x[lo:hi], x[n:]]))
    ys = pkg.DeviceVector(ctx, hi - lo + 6)
    for _ in range(3):
        pkg.matvec_(ys, 2.5, A_sh, xs.v)
    y = ys.to_host()
    assert np.array_equal(y[:hi - lo], y_ref[lo:hi]), "state rows differ from the single-GPU result"
    # ... and against the ORACLE (numpy restatement of the reference's per-term CSC matvec), 1e-12 relative
    from oracle.fspmatrix import FspMatrixOracle
    from oracle.statespace import StateSpaceOracleFast
    osp = StateSpaceOracleFast(model.stoich_matrix, [0, 0, 0])
    osp.expand(levels)
    assert np.array_equal(space.get_states(), osp.states_array())
    y_or = FspMatrixOracle(osp, model.propensities, model.parameters).matvec(2.5, x)
    scale = np.abs(y_or).max()
    assert np.abs(y[:hi - lo] - y_or[lo:hi]).max() <= 1e-12 * scale, "sharded state rows differ from the oracle"
    assert np.abs(y[hi - lo:] - y_or[n:]).max() <= 1e-12 * scale, "sharded sink rows differ from the oracle"
    assert A_sh.build_window is not None and A_sh.build_window[0] <= lo and hi <= A_sh.build_window[1] and (
        world == 1 or A_sh.build_window[1] - A_sh.build_window[0] < n), "sharded build did not use a state-factor window"
    assert np.abs(y[hi - lo:] - y_ref[n:]).max() <= 1e-12 * np.abs(y_ref[n:]).max(), "sink rows differ"
    # same through the peer-memory halo (CUDA IPC): registered input buffer, no NCCL in the matvec
    before = comm.info()
    xr = pkg.ShardedVector(A_sh, fill=np.concatenate([x[lo:hi], x[n:]]), register=True)
    for rep in range(5):
        xrep = x * (1.0 + rep)
        yrep = pkg.matvec(2.5, A_full, xrep)
        xr.v.upload(np.concatenate([xrep[lo:hi], xrep[n:]]))     # overwrite between matvecs: exercises ready/done
        pkg.matvec_(ys, 2.5, A_sh, xr.v)
        y2 = ys.to_host()
        bad = np.nonzero(y2[:hi - lo] != yrep[lo:hi])[0]
        assert bad.size == 0, f"peer-memory halo: {bad.size} state rows differ (first {bad[:5]}, rep {rep})"
    after = comm.info()
    if after["p2p"]:
        assert after["p2p_matvecs"] - before["p2p_matvecs"] == 5 and after["nccl_matvecs"] == before["nccl_matvecs"]
    xr.unregister()
    gsum = torch.tensor([y[:hi - lo].sum()], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(gsum)
    ones = pkg.ShardedVector(A_sh, fill=np.ones(hi - lo + 6))
    pkg.matvec_(ys, 2.5, A_sh, ones.v)
    yo = ys.to_host()
    tot = torch.tensor([yo[:hi - lo].sum()], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(tot)
    assert abs(float(tot) + yo[hi - lo:].sum()) <= 1e-9, "column sums of the sharded operator are not zero"

    # ---- genuinely joint (non-separable) time-varying propensity on a row-sharded matrix: every rank refreshes the
    #      values over its rows + predecessor window (ncme_matrix_set_joint_values), vs single GPU and vs the oracle
    import math
    jprops = list(model.propensities[:5]) + [pkg.propensity(
        lambda t, x, p: p[5] * x[2] * (1.0 + 0.5 * math.sin(2.0 * math.pi * t / 10.0) * x[0] / (1.0 + x[0])))]
    jmodel = pkg.CmeModel(model.stoich_matrix, jprops, model.parameters)
    J_full = pkg.FspMatrixSparse(space, jprops, parameters=model.parameters)
    J_sh = pkg.FspMatrixSparse(space, jprops, parameters=model.parameters, comm=comm)
    assert J_sh.device_joint_ids == [6] and J_full.device_joint_ids == [6], "the propensity must stay on the joint path"
    assert J_sh.build_window is not None and (world == 1 or J_sh.build_window[1] - J_sh.build_window[0] < n)
    OJ = FspMatrixOracle(osp, jprops, model.parameters)
    for tj in (2.5, 7.25):
        yj_ref, yj_or = pkg.matvec(tj, J_full, x), OJ.matvec(tj, x)
        pkg.matvec_(ys, tj, J_sh, xs.v)
        yj = ys.to_host()
        assert np.array_equal(yj[:hi - lo], yj_ref[lo:hi]), "joint propensity: sharded state rows differ from single GPU"
        sj = np.abs(yj_or).max()
        assert np.abs(yj[:hi - lo] - yj_or[lo:hi]).max() <= 1e-12 * sj, "joint propensity: sharded rows differ from the oracle"
        assert np.abs(yj[hi - lo:] - yj_or[n:]).max() <= 1e-12 * sj, "joint propensity: sharded sink rows differ from the oracle"
    spj = pkg.StateSpaceSparse(model.stoich_matrix, [0, 0, 0], ctx=ctx)
    spj.expand_(30)
    pj = pkg.FspVectorSparse.from_pairs(spj, [([0, 0, 0], 1.0)])
    for meth in (pkg.NativeRK45(), pkg.NativeBDF()):
        j1 = pkg.solve(jmodel, pj, (0.0, 0.4), meth, saveat=[0.4], odeatol=1e-12, odertol=1e-7, ctx=ctx)
        j2 = pkg.solve(jmodel, pj, (0.0, 0.4), meth, saveat=[0.4], odeatol=1e-12, odertol=1e-7, comm=comm)
        assert np.abs(j1.p[0].values - j2.p[0].values).max() < 2e-7, "joint propensity: sharded solve differs"
        assert abs(j2.p[0].sum() + j2.sinks[0].sum() - 1.0) < 1e-9

    # ---- adaptive solve: telegraph, sharded vs single GPU (same adaptation decisions expected)
    tm = pkg.workloads.telegraph_model()
    p0 = pkg.FspVectorSparse([[1, 0, 0]], [1.0])
    alg = pkg.AdaptiveFspSparse(pkg.NativeRK45(), pkg.RStepAdapter(5, 10, True))
    touts = [50.0, 150.0, 300.0]
    s1 = pkg.solve(tm, p0, (0.0, 300.0), alg, saveat=touts, odeatol=1e-12, odertol=1e-8, ctx=ctx)
    s2 = pkg.solve(tm, p0, (0.0, 300.0), alg, saveat=touts, odeatol=1e-12, odertol=1e-8, comm=comm)
    assert len(s1) == len(s2)
    for a, b, sa, sb in zip(s1.p, s2.p, s1.sinks, s2.sinks):
        assert np.array_equal(a.states, b.states)
        assert np.abs(a.values - b.values).max() < 1e-10
        assert np.abs(sa - sb).max() < 1e-12
    assert s2.stats["adapts"] == s1.stats["adapts"] >= 1

    # ---- fixed-space solve on the 3-species model, sharded vs single
    sp0 = pkg.StateSpaceSparse(model.stoich_matrix, [0, 0, 0], ctx=ctx)
    sp0.expand_(40)
    pfull = pkg.FspVectorSparse.from_pairs(sp0, [([0, 0, 0], 1.0)])
    f1 = pkg.solve(model, pfull, (0.0, 0.5), pkg.NativeRK45(), saveat=[0.25, 0.5], odeatol=1e-12, odertol=1e-8, ctx=ctx)
    f2 = pkg.solve(model, pfull, (0.0, 0.5), pkg.NativeRK45(), saveat=[0.25, 0.5], odeatol=1e-12, odertol=1e-8, comm=comm)
    for a, b in zip(f1.p, f2.p):
        assert np.abs(a.values - b.values).max() < 1e-12
    assert f1.stats["steps"] == f2.stats["steps"]
    b1 = pkg.solve(model, pfull, (0.0, 0.5), pkg.NativeBDF(), saveat=[0.25, 0.5], odeatol=1e-12, odertol=1e-7, ctx=ctx)
    b2 = pkg.solve(model, pfull, (0.0, 0.5), pkg.NativeBDF(), saveat=[0.25, 0.5], odeatol=1e-12, odertol=1e-7, comm=comm)
    for a, b, c3 in zip(b1.p, b2.p, f1.p):
        d_ab, d_ac, d_bc = (float(np.abs(x.values - y.values).max()) for x, y in ((a, b), (a, c3), (b, c3)))
        if rank == 0:
            print(f"BDF single vs sharded {d_ab:.3e}; single vs explicit {d_ac:.3e}; sharded vs explicit {d_bc:.3e}; "
                  f"steps {b1.stats['steps']} / {b2.stats['steps']}, rhs {b1.stats['rhs_evals']} / {b2.stats['rhs_evals']}")
        # two BDF runs whose reductions are summed in different orders may take different step sequences: both must
        # sit within the solver tolerance (odertol = 1e-7) of the tight explicit solution, and of each other
        assert d_ab < 2e-7, "sharded BDF differs from single-GPU BDF"
        assert d_ac < 1e-6 and d_bc < 1e-6, "BDF differs from the explicit integrator"
    dist.barrier()
    if rank == 0:
        print(f"DIST_CHECK_OK world={world} n={n} comm={comm.info()} halo=({info['halo_lo']},{info['halo_hi']}) "
              f"interior=[{info['interior_begin']},{info['interior_end']})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
