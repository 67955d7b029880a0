// Native device-resident integrator for one FSP segment.
// Replaces DE.init(...)/DE.step!(integrator, tend - tnow, true) + the ContinuousCallback of
// src/transientcme/sparse/fspsolve.jl:145-161.  The FSP vector never leaves HBM: per step the host sees
// one weighted error norm and the R sink entries of the stage vectors (for the event function).
//
// method 0: Dormand-Prince 5(4) with FSAL, step-size control on the weighted RMS norm
//           sqrt(mean((err_i / (atol + rtol*max(|u_i|,|unew_i|)))^2)) and 4th-order dense output
//           (Hairer's dopri5 continuous extension) for saveat and for locating the sink event.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <functional>

#include "comm.cuh"
#include "matrix.cuh"
#include "ode.cuh"
#include "vec.cuh"

namespace ncme {

constexpr int ST = 256;

struct StepArgs {
    int64_t N;        // local vector length
    int64_t n;        // offset of the R event-sink entries ([n, n+R) is excluded from the kernel's sum)
    int R;
    const double* u;
    const double* unew;
    const double* k[7];
    double e[7];      // h * error coefficients
    double atol, rtol;
    double* partials;
    unsigned int* counter;
    double* result;   // [0] = sum of squares, [1 + v*R + r] = sink entry r of vector v (u, unew, k1..k7)
};

__global__ void __launch_bounds__(ST) k_rk_errnorm(const __grid_constant__ StepArgs a) {
    __shared__ double wsum[ST / 32];
    __shared__ bool is_last;
    double s = 0.0;
    // all entries but the R event sinks: those may be per-rank partial sums, the host adds their contribution
    for (int64_t i = (int64_t)blockIdx.x * ST + threadIdx.x; i < a.N; i += (int64_t)gridDim.x * ST) {
        if (i >= a.n && i < a.n + a.R) continue;
        double err = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (a.e[j] != 0.0) err = fma(a.e[j], a.k[j][i], err);
        const double w = a.atol + a.rtol * fmax(fabs(a.u[i]), fabs(a.unew[i]));
        const double q = err / w;
        s = fma(q, q, s);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < ST / 32; ++w) t += wsum[w];
        a.partials[blockIdx.x] = t;
        __threadfence();
        is_last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        const volatile double* p = a.partials;
        double t = 0.0;
        for (unsigned k = threadIdx.x; k < gridDim.x; k += 32) t += p[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (threadIdx.x == 0) {
            a.result[0] = t;
            *a.counter = 0u;
        }
    }
    // sink tails of the 9 vectors (the stage kernels that produced them finished before this launch)
    for (int q = threadIdx.x; q < 9 * a.R; q += ST) {
        const int v = q / a.R, r = q % a.R;
        const double* src = v == 0 ? a.u : (v == 1 ? a.unew : a.k[v - 2]);
        a.result[1 + q] = src[a.n + r];
    }
}

struct DenseArgs {
    int64_t N;
    const double* u;
    const double* unew;
    const double* k[7];
    double d[7];   // h * d_j
    double h, theta;
    double* out;
};

// Hairer's contd5: u + th*(r2 + (1-th)*(r3 + th*(r4 + (1-th)*r5)))
__global__ void __launch_bounds__(ST) k_rk_dense(const __grid_constant__ DenseArgs a) {
    const int64_t i = (int64_t)blockIdx.x * ST + threadIdx.x;
    if (i >= a.N) return;
    const double u0 = a.u[i], u1 = a.unew[i];
    const double r2 = u1 - u0;
    const double r3 = a.h * a.k[0][i] - r2;
    const double r4 = r2 - a.h * a.k[6][i] - r3;
    double r5 = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
        if (a.d[j] != 0.0) r5 = fma(a.d[j], a.k[j][i], r5);
    const double th = a.theta, th1 = 1.0 - a.theta;
    a.out[i] = u0 + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
}

// ---- Dormand-Prince tableau
static const double DP_C[7] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
static const double DP_A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
static const double DP_E[7] = {71.0 / 57600, 0, -71.0 / 16695, 71.0 / 1920, -17253.0 / 339200, 22.0 / 525, -1.0 / 40};
static const double DP_D[7] = {-12715105075.0 / 11282082432.0, 0, 87487479700.0 / 32700410799.0,
                               -10690763975.0 / 1880347072.0, 701980252875.0 / 199316789632.0,
                               -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0};

struct Workspace {       // views into the context's grow-only caches
    double* base = nullptr;
    double* k[7];
    double* ytmp;
    double* ytmp2;   // consecutive RHS inputs alternate buffers (lets the peer-memory halo skip its "done" handshake)
    double* ua;
    double* ub;
    double* full = nullptr;     // n_global + R, for gathered output slices (sharded runs)
    double* pinned = nullptr;
};

int cache_reserve(double** p, size_t* have, size_t want, bool pinned) {
    if (*have >= want) return NCME_OK;
    // grow-only with head-room: adaptive solves call this once per segment with a slightly larger state space each
    // time, and cudaMallocHost / cudaMalloc + their frees cost more than a whole small segment
    want = std::max<size_t>(want + std::min<size_t>(want / 2, (size_t)256 << 20), (size_t)4 << 20);
    if (*p) {
        if (pinned)
            cudaFreeHost(*p);
        else
            cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    cudaError_t e = pinned ? cudaMallocHost(p, want) : cudaMalloc(p, want);
    if (e != cudaSuccess) {
        set_error("integrator workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        return NCME_ERR_NOMEM;
    }
    *have = want;
    return NCME_OK;
}

// sum over sinks of the dense output at theta, from the (globally reduced) tails gathered by k_rk_errnorm
static double sink_dense_sum(const double* tails, int R, double h, double theta) {
    double s = 0.0;
    const double th = theta, th1 = 1.0 - theta;
    for (int r = 0; r < R; ++r) {
        const double u0 = tails[0 * R + r], u1 = tails[1 * R + r];
        const double r2 = u1 - u0;
        const double r3 = h * tails[2 * R + r] - r2;
        const double r4 = r2 - h * tails[8 * R + r] - r3;
        double r5 = 0.0;
        for (int j = 0; j < 7; ++j) r5 += h * DP_D[j] * tails[(2 + j) * R + r];
        s += u0 + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
    }
    return s;
}

__global__ void k_zero_tail(double* p, int R) {
    if ((int)threadIdx.x < R) p[threadIdx.x] = 0.0;
}

int SliceSaver::init(const OdeSystem& s, ncme_save_fn f, void* u, ncme_solve_stats* stats) {
    sys = &s;
    fn = f;
    user = u;
    st = stats;
    if (!fn) return NCME_OK;
    ncme_ctx* ctx = s.ctx;
    ncme_comm* comm = s.comm;
    cudaStream_t stream = ctx->stream;
    NCME_TRY(cache_reserve(&ctx->solve_pinned, &ctx->solve_pinned_bytes, (size_t)s.len_global * sizeof(double), true));
    pinned = ctx->solve_pinned;
    if (comm) {
        NCME_TRY(cache_reserve(&ctx->solve_full, &ctx->solve_full_bytes, (size_t)s.len_global * sizeof(double), false));
        full = ctx->solve_full;
        counts.resize(comm->nranks);
        displs.resize(comm->nranks);
        double mine = (double)s.sink_off;
        NCME_CUDA(cudaMemcpyAsync(comm->scratch + comm->rank, &mine, sizeof(double), cudaMemcpyHostToDevice, stream));
        NCME_NCCL(nccl_api()->AllGather(comm->scratch + comm->rank, comm->scratch, 1, ncclDouble, comm->nccl, stream));
        std::vector<double> hc(comm->nranks);
        NCME_CUDA(cudaMemcpyAsync(hc.data(), comm->scratch, sizeof(double) * comm->nranks, cudaMemcpyDeviceToHost, stream));
        NCME_CUDA(cudaStreamSynchronize(stream));
        int64_t off = 0;
        for (int r = 0; r < comm->nranks; ++r) {
            counts[r] = (int64_t)hc[r];
            displs[r] = off;
            off += counts[r];
        }
    }
    return NCME_OK;
}

// `v_dev` holds per-rank partial sink entries on sharded runs
int SliceSaver::save(double t, const double* v_dev) {
    if (!fn) return NCME_OK;
    cudaStream_t stream = sys->ctx->stream;
    ncme_comm* comm = sys->comm;
    const double* src = v_dev;
    if (comm) {
        const int64_t n = sys->sink_off;
        NCME_TRY(comm_allgatherv(comm, v_dev, full, counts.data(), displs.data(), stream));
        NCME_CUDA(cudaMemcpyAsync(full + sys->n_global, v_dev + n, (size_t)sys->R * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        NCME_TRY(comm_allreduce_sum(comm, full + sys->n_global, (size_t)sys->R, stream));
        src = full;
    }
    NCME_CUDA(cudaMemcpyAsync(pinned, src, (size_t)sys->len_global * sizeof(double), cudaMemcpyDeviceToHost, stream));
    NCME_CUDA(cudaStreamSynchronize(stream));
    fn(t, pinned, user);
    st->nsaved++;
    return NCME_OK;
}

int solve_dp5(const OdeSystem& sys, ncme_save_fn save_fn, void* user, double t0, double t1, double* u,
                     const ncme_solve_opts* o, ncme_solve_stats* st) {
    ncme_ctx* ctx = sys.ctx;
    ncme_comm* comm = sys.comm;          // nullptr on a single GPU
    cudaStream_t s = ctx->stream;
    const int64_t N = sys.len, n = sys.sink_off;   // local vector length / offset of the event sinks
    const int R = sys.R;
    const int64_t Nglob = sys.len_global;
    const int64_t launches0 = ctx->launches;
    Workspace ws;
    // every vector that can be a matvec input carries the halo margins: [hl | n rows | R sinks | hh]
    const size_t hl = round_up<size_t>((size_t)sys.hl, 32), hh = round_up<size_t>((size_t)sys.hh, 32);
    const size_t Npad = hl + round_up<size_t>((size_t)N, 32) + hh;
    // Sharded runs: the workspace is registered for peer access (CUDA IPC), so the neighbours can pull halo entries
    // of any stage vector straight from this GPU's HBM.
    if (comm) {
        NCME_TRY(comm_workspace(comm, Npad * 11 * sizeof(double), (int64_t)hl, (int64_t)Npad, 11, sys.peers, 2, &ws.base));
    } else {
        NCME_TRY(cache_reserve(&ctx->solve_ws, &ctx->solve_ws_bytes, Npad * 11 * sizeof(double), false));
        ws.base = ctx->solve_ws;
    }
    for (int j = 0; j < 7; ++j) ws.k[j] = ws.base + Npad * j + hl;
    ws.ytmp = ws.base + Npad * 7 + hl;
    ws.ua = ws.base + Npad * 8 + hl;
    ws.ub = ws.base + Npad * 9 + hl;
    ws.ytmp2 = ws.base + Npad * 10 + hl;
    auto rhs = [&](double t, const double* x, double* y) -> int {
        st->rhs_evals++;
        return sys.rhs(t, x, y);
    };
    SliceSaver saver;
    NCME_TRY(saver.init(sys, save_fn, user, st));
    auto save = [&](double t, const double* v_dev) -> int { return saver.save(t, v_dev); };
    auto lincomb = [&](int kterms, const double* cs, const double* const* xs, double* out) -> int {
        double c2[8];
        const double* x2[8];
        int m = 0;
        for (int j = 0; j < kterms; ++j)
            if (cs[j] != 0.0) {   // a_72 = 0
                c2[m] = cs[j];
                x2[m] = xs[j];
                ++m;
            }
        return ncme_vec_lincomb(ctx, N, m, c2, x2, out);
    };

    const double rtol = o->rtol > 0 ? o->rtol : 1e-4, atol = o->atol > 0 ? o->atol : 1e-6;
    const int64_t max_steps = o->max_steps > 0 ? o->max_steps : 100000000;
    const double tspan = t1 - t0;
    double t = t0;
    st->t_final = t0;
    st->event_hit = 0;

    // Inside the integrator the R sink entries of every vector are PER-RANK PARTIAL SUMS (their sum over ranks is
    // the true value): every operation on them is linear and they never feed back into A x, so no per-matvec
    // all-reduce is needed.  Rank 0 carries the incoming (replicated) sink values.
    double* ucur = ws.ua;
    double* unext = ws.ub;
    NCME_CUDA(cudaMemcpyAsync(ucur, u, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (comm && comm->rank != 0) {
        k_zero_tail<<<1, 32, 0, s>>>(ucur + n, R);
        ctx->launches++;
    }
    auto finish = [&](const double* src) -> int {   // hand the local slice back with reduced sinks
        if (src != u) NCME_CUDA(cudaMemcpyAsync(u, src, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s));
        NCME_TRY(comm_allreduce_sum(comm, u + n, (size_t)R, s));
        NCME_CUDA(cudaStreamSynchronize(s));
        if (comm && comm->my_flags) {
            unsigned int perr = 0;
            NCME_CUDA(cudaMemcpy(&perr, &comm->my_flags->error, sizeof(perr), cudaMemcpyDeviceToHost));
            if (perr) {
                set_error("peer-memory halo: a neighbouring rank did not signal within 2 s");
                return NCME_ERR_COMM;
            }
        }
        st->launches = ctx->launches - launches0;
        return NCME_OK;
    };

    int isave = 0;
    while (isave < o->nsave && o->save_t[isave] < t0) ++isave;
    if (o->save_every_step) NCME_TRY(save(t0, ucur));
    while (isave < o->nsave && o->save_t[isave] == t0) {
        NCME_TRY(save(t0, ucur));
        ++isave;
    }
    if (!(tspan > 0)) return finish(ucur);

    NCME_TRY(rhs(t, ucur, ws.k[0]));
    // initial step: h = 0.01 * ||u|| / ||f|| in the weighted norm (Hairer), clipped to the span
    double h = o->h_init;
    if (!(h > 0)) {
        double d0 = 0, d1 = 0;
        const int64_t nn = comm ? n : N;   // sharded: state rows only (the sink entries are partial sums)
        NCME_TRY(ncme_vec_wrms(ctx, nn > 0 ? nn : 1, ucur, ucur, ucur, atol, rtol, &d0));
        NCME_TRY(ncme_vec_wrms(ctx, nn > 0 ? nn : 1, ws.k[0], ucur, ucur, atol, rtol, &d1));
        double ss[2] = {d0 * d0 * (double)nn, d1 * d1 * (double)nn};
        if (comm) {
            NCME_CUDA(cudaMemcpyAsync(comm->scratch, ss, sizeof(ss), cudaMemcpyHostToDevice, s));
            NCME_TRY(comm_allreduce_sum(comm, comm->scratch, 2, s));
            NCME_CUDA(cudaMemcpyAsync(ss, comm->scratch, sizeof(ss), cudaMemcpyDeviceToHost, s));
            NCME_CUDA(cudaStreamSynchronize(s));
        }
        d0 = sqrt(ss[0] / (double)Nglob);
        d1 = sqrt(ss[1] / (double)Nglob);
        h = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h = std::min(h, tspan);
    }
    double g_prev = 0.0;
    bool have_g = false;
    bool last_rejected = false;
    StepArgs ea;
    ea.N = N;
    ea.n = n;
    ea.R = R;
    for (int j = 0; j < 7; ++j) ea.k[j] = ws.k[j];
    ea.atol = atol;
    ea.rtol = rtol;
    ea.partials = ctx->red_partials;
    ea.counter = ctx->red_counter;
    ea.result = ctx->red_result_dev;
    int64_t nb = (N + (int64_t)ST * 8 - 1) / ((int64_t)ST * 8);
    nb = std::max<int64_t>(1, std::min<int64_t>(nb, 4096));
    // smallest resolvable step: relative to where the segment STARTS (a long horizon must not forbid the tiny first
    // steps a tight absolute tolerance asks for: toggle example, odeatol = 1e-14 over 8 h)
    const double hmin = std::max(1e-14 * fabs(t0), 1e-20 * fabs(t1 - t0));
    const size_t nres = (size_t)(1 + 9 * R);

    while (t < t1) {
        if (abort_requested()) return abort_status();   // a save callback failed
        if (st->steps + st->rejected >= max_steps) {
            set_error("integrator: maximum number of steps (%lld) reached at t = %g", (long long)max_steps, t);
            return NCME_ERR_SOLVER;
        }
        bool last = false;
        if (t + h >= t1 || t1 - (t + h) < 1e-12 * tspan) {
            h = t1 - t;
            last = true;
        }
        for (int i = 1; i < 7; ++i) {   // stages 2..7
            double cs[8];
            const double* xs[8];
            cs[0] = 1.0;
            xs[0] = ucur;
            for (int j = 0; j < i; ++j) {
                cs[1 + j] = h * DP_A[i][j];
                xs[1 + j] = ws.k[j];
            }
            double* dst = (i == 6) ? unext : ((i & 1) ? ws.ytmp : ws.ytmp2);
            NCME_TRY(lincomb(1 + i, cs, xs, dst));
            NCME_TRY(rhs(t + DP_C[i] * h, dst, ws.k[i]));
        }
        for (int j = 0; j < 7; ++j) ea.e[j] = h * DP_E[j];
        ea.u = ucur;
        ea.unew = unext;
        k_rk_errnorm<<<(unsigned)nb, ST, 0, s>>>(ea);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        if (comm_hostreduce_available(comm) && nres <= NCME_HOSTREDUCE_MAX) {   // one small all-reduce per step:
            NCME_CUDA(cudaMemcpyAsync(ctx->red_result_host, ctx->red_result_dev, sizeof(double) * nres, cudaMemcpyDeviceToHost, s));
            NCME_CUDA(cudaStreamSynchronize(s));                               // through shared host memory ...
            NCME_TRY(comm_hostreduce_sum(comm, ctx->red_result_host, (size_t)nres));
        } else {                                                                // ... or NCCL on the device
            NCME_TRY(comm_allreduce_sum(comm, ctx->red_result_dev, nres, s));
            NCME_CUDA(cudaMemcpyAsync(ctx->red_result_host, ctx->red_result_dev, sizeof(double) * nres, cudaMemcpyDeviceToHost, s));
            NCME_CUDA(cudaStreamSynchronize(s));
        }
        const double* tails = ctx->red_result_host + 1;
        double sumsq = ctx->red_result_host[0];
        for (int r = 0; r < R; ++r) {   // sink rows, from the reduced tails
            double e = 0.0;
            for (int j = 0; j < 7; ++j) e += h * DP_E[j] * tails[(2 + j) * R + r];
            const double w = atol + rtol * std::max(fabs(tails[r]), fabs(tails[R + r]));
            sumsq += (e / w) * (e / w);
        }
        const double err = sqrt(sumsq / (double)Nglob);
        if (!(err <= 1.0)) {  // reject (also catches NaN)
            st->rejected++;
            const double fac = isfinite(err) ? std::max(0.2, 0.9 * pow(err, -0.2)) : 0.1;
            h *= fac;
            last_rejected = true;
            if (h < hmin) {
                set_error("integrator: step size underflow at t = %g (err = %g)", t, err);
                return NCME_ERR_SOLVER;
            }
            continue;
        }
        st->steps++;
        double theta_end = 1.0;
        bool event = false;
        if (o->check_event) {
            // g(theta) = sum_sinks(dense(theta)) - slope * (t + theta h), sampled like interp_points
            if (!have_g) {
                g_prev = sink_dense_sum(tails, R, h, 0.0) - o->event_slope * t;
                have_g = true;
            }
            const int NS_ = 16;
            double ga = g_prev, tha = 0.0;
            for (int q = 1; q <= NS_; ++q) {
                const double thb = (double)q / NS_;
                const double gb = sink_dense_sum(tails, R, h, thb) - o->event_slope * (t + thb * h);
                if (ga <= 0.0 && gb > 0.0) {
                    double lo = tha, hi = thb;
                    for (int it = 0; it < 60; ++it) {
                        const double mid = 0.5 * (lo + hi);
                        const double gm = sink_dense_sum(tails, R, h, mid) - o->event_slope * (t + mid * h);
                        if (gm > 0.0)
                            hi = mid;
                        else
                            lo = mid;
                    }
                    theta_end = hi;
                    event = true;
                    break;
                }
                ga = gb;
                tha = thb;
            }
            if (!event) g_prev = ga;
        }
        // dense-output saves inside (t, t + theta_end*h]
        DenseArgs da;
        da.N = N;
        da.u = ucur;
        da.unew = unext;
        for (int j = 0; j < 7; ++j) {
            da.k[j] = ws.k[j];
            da.d[j] = h * DP_D[j];
        }
        da.h = h;
        da.out = ws.ytmp;
        const double t_hi = t + theta_end * h;
        while (isave < o->nsave && o->save_t[isave] <= t_hi + 1e-14 * fabs(t_hi)) {
            const double ts = o->save_t[isave];
            const double th = std::min(1.0, std::max(0.0, (ts - t) / h));
            if (th >= 1.0 && !event) {
                NCME_TRY(save(ts, unext));
            } else {
                da.theta = th;
                k_rk_dense<<<(unsigned)((N + ST - 1) / ST), ST, 0, s>>>(da);
                ctx->launches++;
                NCME_TRY(save(ts, ws.ytmp));
            }
            ++isave;
        }
        if (event) {
            da.theta = theta_end;
            k_rk_dense<<<(unsigned)((N + ST - 1) / ST), ST, 0, s>>>(da);
            ctx->launches++;
            NCME_CUDA(cudaGetLastError());
            st->t_final = t_hi;
            st->event_hit = 1;
            st->h_last = h;
            return finish(ws.ytmp);
        }
        // accept: u <-> unew and k1 <-> k7 (FSAL) by pointer swaps, no copies
        std::swap(ucur, unext);
        std::swap(ws.k[0], ws.k[6]);
        ea.k[0] = ws.k[0];
        ea.k[6] = ws.k[6];
        t = last ? t1 : t + h;
        if (o->save_every_step) NCME_TRY(save(t, ucur));
        st->h_last = h;
        double fac = (err > 0) ? 0.9 * pow(err, -0.2) : 5.0;
        fac = std::min(last_rejected ? 1.0 : 5.0, std::max(0.2, fac));
        last_rejected = false;
        if (!last) h *= fac;
    }
    st->t_final = t;
    return finish(ucur);
}

}  // namespace ncme

using namespace ncme;

extern "C" int ncme_solve_segment(ncme_matrix* A, ncme_coef_fn coef_fn, ncme_save_fn save_fn, void* user, double t0,
                                  double t1, double* u_dev, const ncme_solve_opts* opts, ncme_solve_stats* stats) {
    NCME_RANGE("ncme_solve_segment");
    NCME_REQUIRE(A && u_dev && opts && stats, "null argument");
    NCME_REQUIRE(t1 >= t0, "solve_segment: t1 < t0");
    NCME_REQUIRE(opts->nsave == 0 || opts->save_t, "save_t is null");
    bool need_coef = false;
    for (int r = 0; r < A->nr; ++r) need_coef |= (A->kind[r] != NCME_TIME_INVARIANT);
    NCME_REQUIRE(coef_fn || !need_coef, "the matrix has time-varying reactions: a coefficient callback is required");
    memset(stats, 0, sizeof(*stats));
    clear_abort();
    double coef[NCME_MAX_REACTIONS];
    for (int r = 0; r < NCME_MAX_REACTIONS; ++r) coef[r] = 1.0;
    // time factors are pure functions of t (the reference evaluates them in every matvec!, fspsparsematrix.jl:204): one
    // host callback per DISTINCT time -- all right-hand sides of a BDF step share t_new
    double coef_t = NAN;
    auto refresh_coef = [&](double t) {
        if (coef_fn && !(t == coef_t)) {
            coef_fn(t, coef, user);
            coef_t = t;
        }
    };
    OdeSystem sys;
    sys.ctx = A->ctx;
    sys.comm = A->comm;
    sys.len = A->N;
    sys.sink_off = A->n;
    sys.R = A->nr;
    sys.hl = A->hl;
    sys.hh = A->hh;
    sys.len_global = A->n_global + A->nr;
    sys.n_global = A->n_global;
    sys.peers[0] = A->plo;
    sys.peers[1] = A->phi;
    sys.rhs = [&](double t, const double* x, double* y) -> int {
        refresh_coef(t);
        if (abort_requested()) return abort_status();
        return matvec_dist(A, coef, x, y, 0.0, /*no sink reduction, inputs alternate buffers*/ 2);
    };
    sys.rhs_safe = [&](double t, const double* x, double* y) -> int {
        refresh_coef(t);
        if (abort_requested()) return abort_status();
        return matvec_dist(A, coef, x, y, 0.0, 0);
    };
    sys.n_impl = A->n;
    sys.jac_diag = [&](double t, double* out) -> int {
        refresh_coef(t);
        if (abort_requested()) return abort_status();
        return matrix_diag(A, coef, out);
    };
    sys.rhs_sinks = [&](double t, const double* x, double* y) -> int {
        refresh_coef(t);
        if (abort_requested()) return abort_status();
        return matvec_sinks_only(A, coef, x, y);
    };
    if (opts->method == 0) return solve_dp5(sys, save_fn, user, t0, t1, u_dev, opts, stats);
    if (opts->method == 1 || opts->method == 2 || opts->method == 3) {
        // BDF: one fused kernel per step attempt (bdf_fused.cu) where launch latency dominates, the
        // launch-per-operation integrator (bdf.cu) for sharded matrices and for very large state spaces
        bool fused = false;
        if (opts->method == 3) {
            NCME_REQUIRE(bdf_fused_eligible(A), "method 3 (fused BDF step kernel) needs an unsharded matrix or the peer-memory transport");
            fused = true;
        } else if (opts->method == 1 && bdf_fused_eligible(A)) {
            static const long long max_rows = [] {
                const char* e = getenv("NCME_BDF_FUSED_MAX_ROWS");
                return e ? atoll(e) : 2000000LL;   // measured cross-over on B200 (tools/bdf_crossover.py)
            }();
            static const long long max_rows_sharded = [] {
                const char* e = getenv("NCME_BDF_FUSED_MAX_ROWS_SHARDED");
                return e ? atoll(e) : 500000LL;    // per rank.  Measured on 4 B200 (profiles/README.md): at 1.27e6 rows per
            }();                                   // rank the launch-per-operation path wins (0.085 vs 0.116 s), at 2.5e6
                                                   // 0.122 vs 0.201 s; the one-kernel step pays off for small shards only
            // decided from replicated facts only: every rank must take the same branch
            const int P = A->comm ? A->comm->nranks : 1;
            fused = (long long)(A->n_global / P) <= (A->comm ? max_rows_sharded : max_rows);
        }
        if (fused) return solve_bdf_fused(A, coef_fn, save_fn, user, t0, t1, u_dev, opts, stats);
        return solve_bdf(sys, save_fn, user, t0, t1, u_dev, opts, stats);
    }
    set_error("unknown integrator method %d", opts->method);
    return NCME_ERR_ARG;
}

// Forward-sensitivity segment: the same integrator on the block vector [p; s_1; ...; s_P] with the fused block
// matvec as right-hand side (reference: src/forwardsenscme/sparse/forwardsenscmesparse.jl:142-166).  The event
// watches the sinks of the probability block only (:153-155).  coef_fn fills coef[nr] then dcoef[nentries]
// (contiguous: coef at [0, nr), dcoef at [nr, nr + nentries)).
extern "C" int ncme_sens_solve_segment(ncme_sensmatrix* SA, ncme_coef_fn coef_fn, ncme_save_fn save_fn, void* user,
                                       double t0, double t1, double* U_dev, const ncme_solve_opts* opts,
                                       ncme_solve_stats* stats) {
    NCME_RANGE("ncme_sens_solve_segment");
    NCME_REQUIRE(SA && U_dev && opts && stats && coef_fn, "null argument");
    NCME_REQUIRE(t1 >= t0, "solve_segment: t1 < t0");
    NCME_REQUIRE(opts->nsave == 0 || opts->save_t, "save_t is null");
    memset(stats, 0, sizeof(*stats));
    clear_abort();
    ncme_matrix* A = nullptr;
    int npar = 0, nent = 0;
    NCME_TRY(sens_describe(SA, &A, &npar, &nent));
    std::vector<double> cf((size_t)A->nr + (size_t)nent + 1, 1.0);
    OdeSystem sys;
    sys.ctx = A->ctx;
    sys.comm = nullptr;
    sys.len = (int64_t)(npar + 1) * A->N;
    sys.sink_off = A->n;
    sys.R = A->nr;
    sys.len_global = sys.len;
    sys.n_global = A->n;
    sys.rhs = [&](double t, const double* x, double* y) -> int {
        coef_fn(t, cf.data(), user);
        if (abort_requested()) return abort_status();
        return ncme_sens_matvec(SA, cf.data(), cf.data() + A->nr, x, y);
    };
    // BDF: the whole block vector is implicit; Jacobian = [A 0; dA A] => block-Jacobi diagonal = diag(A) per block
    // (zero on the sink rows)
    sys.n_impl = sys.len;
    sys.jac_diag = [&](double t, double* out) -> int {
        coef_fn(t, cf.data(), user);
        if (abort_requested()) return abort_status();
        cudaStream_t st = A->ctx->stream;
        NCME_CUDA(cudaMemsetAsync(out, 0, (size_t)sys.len * sizeof(double), st));
        NCME_TRY(matrix_diag(A, cf.data(), out));
        for (int b = 1; b <= npar; ++b)
            NCME_CUDA(cudaMemcpyAsync(out + (size_t)b * A->N, out, (size_t)A->n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return NCME_OK;
    };
    if (opts->method == 0) return solve_dp5(sys, save_fn, user, t0, t1, U_dev, opts, stats);
    if (opts->method >= 1 && opts->method <= 3) return solve_bdf(sys, save_fn, user, t0, t1, U_dev, opts, stats);
    set_error("unknown integrator method %d", opts->method);
    return NCME_ERR_ARG;
}
