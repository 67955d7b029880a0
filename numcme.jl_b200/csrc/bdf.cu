// Native stiff integrator (method 1): variable-step variable-order BDF in the backward-difference (NDF) form of
// Shampine & Reichelt with a matrix-free, diagonally preconditioned GMRES linear solver -- the device-resident
// counterpart of `CVODE_BDF(linear_solver = :GMRES)`, the algorithm every example of the reference passes to
// `solve` (examples/telegraph_cme.jl:9, toggleswitch_fsp_variants.jl:68, hog1p.jl:69; test/test_solver.jl:60).
// The ODE is linear, du/dt = A(t) u, so the Newton iteration of a BDF step is ONE linear solve with the exact
// Jacobian:  (I - c A(t_new)) d = c A(t_new) y_pred - psi,   y_new = y_pred + d,   c = h / alpha_k.
//
// Everything lives in HBM.  The R sink rows never feed back (their columns are empty), so the Krylov solve runs on
// the state rows only and the sink entries of d are completed explicitly afterwards; on sharded runs they stay
// per-rank partial sums (all operations on them are linear), exactly as in the explicit integrator.
#include <math.h>

#include <algorithm>
#include <vector>

#include "comm.cuh"
#include "matrix.cuh"
#include "ode.cuh"
#include "vec.cuh"

namespace ncme {

constexpr int BT = 256;
constexpr int MAX_ORDER = 5;
constexpr int GM_M = 24;          // Krylov dimension before restart
constexpr int RED_SLOTS = GM_M + 2;
constexpr int RED_BLOCKS = 1184;  // 148 SMs x 8 CTAs: the reductions of a step (<= 2 values per CTA; the context's partials buffer
                                  // holds 16384 doubles) ran at 0.3-0.65 of the HBM rate with 512 CTAs (ncu launch list, round 2)

struct PtrList {
    const double* p[GM_M + 2];
};
struct PtrListRW {
    double* p[MAX_ORDER + 3];
};

// Ordered, PARALLEL combine of the per-CTA partials by the CTA that finished last (call with the whole CTA, after the
// fence that follows the completion counter).  Thread t sums the CTAs b = t / SP, t / SP + BT / SP, ... for value slot
// t % SP; the slot threads then add the BT / SP group sums in group order.  Fixed order => the same bits on every
// rank and in every run.  Round 2: this stage used one thread per slot walking all gridDim.x partials (1 184 dependent
// L2 round trips): k_bdf_errnorm ran at 0.30 and k_gm_apply_dots at 0.65-0.70 of the HBM rate because of that tail
// (profiles/README.md, per-kernel roofline of the BDF).
template <int NS>
__device__ __forceinline__ double final_combine(const double* partials, int nslot) {
    constexpr int SP = NS <= 1 ? 1 : NS <= 2 ? 2 : NS <= 4 ? 4 : NS <= 8 ? 8 : NS <= 16 ? 16 : 32;
    constexpr int NG = BT / SP;
    static_assert(NS <= 32, "final_combine handles up to 32 value slots");
    __shared__ double gsum[NG][SP];
    const int sl = threadIdx.x % SP, g = threadIdx.x / SP;
    double t = 0.0;
    if (sl < nslot)
        for (unsigned b = g; b < gridDim.x; b += NG) t += __ldcg(partials + (size_t)b * NS + sl);
    gsum[g][sl] = t;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x < nslot) {
#pragma unroll 8
        for (int q = 0; q < NG; ++q) tot += gsum[q][threadIdx.x];
    }
    return tot;   // valid in threads [0, nslot)
}

// ---- deterministic multi-value block reduction: value slots [0, nslot) ------------------------------------------
template <int NS>
__device__ __forceinline__ void block_reduce_store(double (&v)[NS], int nslot, double* partials, unsigned int* counter,
                                                   double* result) {
    __shared__ double sh[BT / 32][NS];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (s < nslot) {
            double x = v[s];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
            if (lane == 0) sh[wid][s] = x;
        }
    }
    __syncthreads();
    if (threadIdx.x < nslot) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < BT / 32; ++w) t += sh[w][threadIdx.x];
        partials[(size_t)blockIdx.x * NS + threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const double t = final_combine<NS>(partials, nslot);
    if (threadIdx.x < nslot) result[threadIdx.x] = t;
    if (threadIdx.x == 0) *counter = 0u;
}

// Same two-stage block reduction, followed -- inside the SAME kernel, by the CTA that finished last -- by the all-reduce
// over the ranks through peer memory (NVLink stores into every rank's PeerFlags::red_slot, one flag per source rank,
// summation in rank order => identical bits everywhere).  The reduced values land in device memory `result`, where
// the next kernel of the Krylov iteration reads them: no host round trip, no NCCL launch in the iteration.
__device__ __forceinline__ unsigned long long bdf_global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
template <int NS>
__device__ __forceinline__ void block_reduce_allreduce(double (&v)[NS], double* partials, unsigned int* counter,
                                                       double* result, const DevAllreduce& ar) {
    static_assert(NS <= NCME_RED_VALS, "too many values for the in-kernel all-reduce");
    __shared__ double sh[BT / 32][NS];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double x = v[s];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) sh[wid][s] = x;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < BT / 32; ++w) t += sh[w][threadIdx.x];
        partials[(size_t)blockIdx.x * NS + threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double t = final_combine<NS>(partials, NS);
    if (threadIdx.x == 0) *counter = 0u;
    if (ar.nranks > 1) {
        const int buf = (int)(ar.epoch % NCME_RED_BUFS);
        if (threadIdx.x < NS)
            for (int q = 0; q < ar.nranks; ++q) ar.flags[q]->red_slot[buf][ar.me][threadIdx.x] = t;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < ar.nranks)
            *((volatile unsigned int*)&ar.flags[threadIdx.x]->red_flag[buf][ar.me]) = ar.epoch;
        if (threadIdx.x < ar.nranks) {   // wait (bounded) for the partial sums of every rank
            const volatile unsigned int* f = &ar.flags[ar.me]->red_flag[buf][threadIdx.x];
            const unsigned long long t0 = bdf_global_ns();
            while ((int)(*f - ar.epoch) < 0) {
                if (bdf_global_ns() - t0 > 2000000000ull) {
                    atomicExch(&ar.flags[ar.me]->error, 1u);
                    break;
                }
            }
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < NS) {
            const volatile double* sl = &ar.flags[ar.me]->red_slot[buf][0][0];
            t = 0.0;
            for (int q = 0; q < ar.nranks; ++q) t += sl[q * NCME_RED_VALS + threadIdx.x];
        }
    }
    if (threadIdx.x < NS) result[threadIdx.x] = t;
}

// y_pred = sum_{j<=k} D_j ;  psi = (sum_{1<=j<=k} gamma_j D_j) / alpha_k
__global__ void __launch_bounds__(BT) k_bdf_predict(int64_t N, int order, PtrList D, double g1, double g2, double g3,
                                                     double g4, double g5, double inv_alpha, double* __restrict__ ypred,
                                                     double* __restrict__ psi) {
    const int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    if (i >= N) return;
    const double g[5] = {g1, g2, g3, g4, g5};
    double y = D.p[0][i], ps = 0.0;
    for (int j = 1; j <= order; ++j) {
        const double dj = D.p[j][i];
        y += dj;
        ps = fma(g[j - 1], dj, ps);
    }
    ypred[i] = y;
    psi[i] = ps * inv_alpha;
}

// scale = atol + rtol |y_pred| ; ps = 1 / ((1 - c diag) scale) ; w0 = (c A y_pred - psi) ps ; result[0] = |w0|^2
struct SetupArgs {
    int64_t n;
    const double* ypred;
    const double* Ay;
    const double* psi;
    const double* jdiag;
    double c, atol, rtol;
    double* scale;
    double* ps;
    double* w0;
    double* partials;
    unsigned int* counter;
    double* result;
};
__global__ void __launch_bounds__(BT) k_bdf_setup(const __grid_constant__ SetupArgs a) {
    double v[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * BT;
    int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    // four grid strides per trip, all loads first (the stores through the struct's plain pointers would otherwise keep
    // the compiler from hoisting the next element's loads): enough bytes in flight per SM for the HBM latency
    for (; i + 3 * stride < a.n; i += 4 * stride) {
        double yp[4], jd[4], ay[4], ps[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            yp[u] = a.ypred[i + u * stride];
            jd[u] = a.jdiag[i + u * stride];
            ay[u] = a.Ay[i + u * stride];
            ps[u] = a.psi[i + u * stride];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double sc = a.atol + a.rtol * fabs(yp[u]);
            const double p = 1.0 / ((1.0 - a.c * jd[u]) * sc);
            const double w = (a.c * ay[u] - ps[u]) * p;
            a.scale[i + u * stride] = sc;
            a.ps[i + u * stride] = p;
            a.w0[i + u * stride] = w;
            v[0] = fma(w, w, v[0]);
        }
    }
    for (; i < a.n; i += stride) {
        const double sc = a.atol + a.rtol * fabs(a.ypred[i]);
        const double p = 1.0 / ((1.0 - a.c * a.jdiag[i]) * sc);
        const double w = (a.c * a.Ay[i] - a.psi[i]) * p;
        a.scale[i] = sc;
        a.ps[i] = p;
        a.w0[i] = w;
        v[0] = fma(w, w, v[0]);
    }
    block_reduce_store<1>(v, 1, a.partials, a.counter, a.result);
}

// v = w * inv ; z = v * scale   (z is the next operator input: carries halo margins)
__global__ void __launch_bounds__(BT) k_gm_normalize(int64_t n, const double* __restrict__ w, double inv,
                                                      const double* __restrict__ scale, double* __restrict__ v,
                                                      double* __restrict__ z) {
    const int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    if (i >= n) return;
    const double x = w[i] * inv;
    v[i] = x;
    z[i] = x * scale[i];
}

// w = (z - c A z) ps ;  result[j] = <w, V_j> (j <= k) ; result[k+1] = <w, w>
struct ApplyArgs {
    int64_t n;
    int k;
    const double* z;
    const double* Az;
    const double* ps;
    double c;
    PtrList V;
    double* w;
    double* partials;
    unsigned int* counter;
    double* result;      // device: column k of the Hessenberg data, [0..k] dots, [RED_SLOTS-1] = <w,w>, summed over the ranks
    DevAllreduce ar;
};
// KB = number of basis vectors dotted (k+1 rounded up to a multiple of 4; the surplus pointers alias V_0 and their
// results are ignored): fully unrolled, no predication, all loads of an element independent.
template <int KB>
__global__ void __launch_bounds__(BT) k_gm_apply_dots(const __grid_constant__ ApplyArgs a) {
    double v[KB + 1];
#pragma unroll
    for (int s = 0; s <= KB; ++s) v[s] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * BT) {
        const double w = (a.z[i] - a.c * a.Az[i]) * a.ps[i];
        double vj[KB > 0 ? KB : 1];
#pragma unroll
        for (int j = 0; j < KB; ++j) vj[j] = a.V.p[j][i];
        a.w[i] = w;
#pragma unroll
        for (int j = 0; j < KB; ++j) v[j] = fma(w, vj[j], v[j]);
        v[KB] = fma(w, w, v[KB]);
    }
    // slot layout expected by the host: [0..k] dots, [RED_SLOTS-1] = <w,w>
    double full[RED_SLOTS];
#pragma unroll
    for (int s = 0; s < RED_SLOTS; ++s) full[s] = 0.0;
#pragma unroll
    for (int j = 0; j < KB; ++j) full[j] = v[j];
    full[RED_SLOTS - 1] = v[KB];
    block_reduce_allreduce<RED_SLOTS>(full, a.partials, a.counter, a.result, a.ar);
}

// One wave: the grid-stride kernel is launched with at most as many CTAs as are resident at once (SMs x occupancy of the
// instantiation).  With the former fixed 1 184 CTAs the instantiations that fit 3-5 CTAs per SM (48-128 registers) ran
// 1.6-2.7 waves, i.e. a last wave that left 30-40 % of the SMs idle while it drained.
template <int KB>
static unsigned apply_dots_wave(ncme_ctx* ctx) {
    static int per_sm = 0;
    if (per_sm == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gm_apply_dots<KB>, BT, 0) != cudaSuccess || nb < 1) {
            cudaGetLastError();
            nb = 1;
        }
        per_sm = nb;
    }
    return (unsigned)(ctx->sm_count * per_sm);
}
static int launch_apply_dots(ncme_ctx* ctx, const ApplyArgs& aa, unsigned grid) {
    cudaStream_t s = ctx->stream;
    const int kb = aa.k < 0 ? 0 : ((aa.k + 1 + 3) / 4) * 4;
#define NCME_AD(KB) k_gm_apply_dots<KB><<<std::min(grid, apply_dots_wave<KB>(ctx)), BT, 0, s>>>(aa)
    switch (kb) {
        case 0: NCME_AD(0); break;
        case 4: NCME_AD(4); break;
        case 8: NCME_AD(8); break;
        case 12: NCME_AD(12); break;
        case 16: NCME_AD(16); break;
        case 20: NCME_AD(20); break;
        default: NCME_AD(24); break;
    }
#undef NCME_AD
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

// v_{k+1} = (w - sum_j h_j V_j) * inv ;  z = v_{k+1} * scale
// The coefficients h_j = <w, V_j> and |w|^2 are read from DEVICE memory (column k written by k_gm_apply_dots, already
// summed over the ranks): the host is not needed between the inner products and the orthogonalisation, so several
// Krylov iterations can be enqueued back to back.  |w - sum h_j v_j| by Pythagoras, exactly as the host recomputes it.
struct OrthoArgs {
    int64_t n;
    int k;
    const double* w;
    PtrList V;
    const double* hcol;   // device: [0..k] = h_j, [RED_SLOTS-1] = <w,w>
    const double* scale;
    double* vout;
    double* z;
};
__host__ __device__ __forceinline__ double gm_hk1(const double* hh, int k) {
    const double ww = hh[RED_SLOTS - 1];
    double hsq = 0.0;
    for (int j = 0; j <= k; ++j) hsq += hh[j] * hh[j];
    double hk1sq = ww - hsq;
    if (!(hk1sq > 1e-10 * ww)) hk1sq = hk1sq > 0.0 ? hk1sq : 0.0;
    return sqrt(hk1sq > 0.0 ? hk1sq : 0.0);
}
// KB = basis vectors subtracted, k + 1 rounded up to a multiple of 4 (surplus pointers alias V_0 with a zero
// coefficient: fma(-0, v, x) == x): fully unrolled, every load of an element independent and issued before first use
template <int KB>
__global__ void __launch_bounds__(BT) k_gm_ortho(const __grid_constant__ OrthoArgs a) {
    __shared__ double sh[RED_SLOTS + 1];
    if (threadIdx.x < RED_SLOTS) sh[threadIdx.x] = a.hcol[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) sh[RED_SLOTS] = 1.0 / gm_hk1(sh, a.k);
    __syncthreads();
    if (threadIdx.x > a.k && threadIdx.x < RED_SLOTS) sh[threadIdx.x] = 0.0;   // (after gm_hk1 read <w,w> in the last slot)
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    if (i >= a.n) return;
    double vj[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) vj[j] = a.V.p[j][i];
    double x = a.w[i];
    const double sc = a.scale[i];
#pragma unroll
    for (int j = 0; j < KB; ++j) x = fma(-sh[j], vj[j], x);
    x *= sh[RED_SLOTS];
    a.vout[i] = x;
    a.z[i] = x * sc;
}
static void launch_ortho(const OrthoArgs& oa, unsigned grid, cudaStream_t s) {
    switch (((oa.k + 1 + 3) / 4) * 4) {
        case 4: k_gm_ortho<4><<<grid, BT, 0, s>>>(oa); break;
        case 8: k_gm_ortho<8><<<grid, BT, 0, s>>>(oa); break;
        case 12: k_gm_ortho<12><<<grid, BT, 0, s>>>(oa); break;
        case 16: k_gm_ortho<16><<<grid, BT, 0, s>>>(oa); break;
        case 20: k_gm_ortho<20><<<grid, BT, 0, s>>>(oa); break;
        default: k_gm_ortho<24><<<grid, BT, 0, s>>>(oa); break;
    }
}

// d = scale * sum_j y_j V_j ;  ynew = ypred + d     (state rows)
struct SolArgs {
    int64_t n;
    int k;
    PtrList V;
    double y[GM_M + 1];
    const double* scale;
    const double* ypred;
    double* d;
    double* ynew;
    int accumulate;   // restart: d += ...
};
__global__ void __launch_bounds__(BT) k_gm_solution(const __grid_constant__ SolArgs a) {
    const int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    if (i >= a.n) return;
    double x = 0.0;
    for (int j = 0; j < a.k; ++j) x = fma(a.y[j], a.V.p[j][i], x);
    x *= a.scale[i];
    if (a.accumulate) x += a.d[i];
    a.d[i] = x;
    a.ynew[i] = a.ypred[i] + x;
}

// sink entries: d_s = c (A ynew)_s - psi_s ; ynew_s = ypred_s + d_s      (linear: valid for per-rank partial sums)
__global__ void k_bdf_sinks(int R, int64_t off, double c, const double* __restrict__ Ay, const double* __restrict__ psi,
                            const double* __restrict__ ypred, double* __restrict__ d, double* __restrict__ ynew) {
    const int r = threadIdx.x;
    if (r >= R) return;
    const double x = c * Ay[off + r] - psi[off + r];
    d[off + r] = x;
    ynew[off + r] = ypred[off + r] + x;
}

// Linear invariant: columns of A sum to zero, so an exact BDF step gives sum_all(d + psi) = 0.  An inexact Krylov
// solve violates it by the sum of its residual; ms[0] = sum_states(d + psi), ms[1] = sum_states|ynew|,
// ms[2] = sum_sinks(d + psi) = c * sum_sinks(A ynew) measure the defect ...
struct MassArgs {
    int64_t n;
    const double* d;
    const double* psi;
    const double* ynew;
    int R;
    int64_t off;
    double* partials;
    unsigned int* counter;
    double* ms;
};
__global__ void __launch_bounds__(BT) k_bdf_massdefect(const __grid_constant__ MassArgs a) {
    double v[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * BT;
    int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    for (; i + 3 * stride < a.n; i += 4 * stride) {   // loads of four grid strides in flight, same summation order
        double dd[4], pp[4], yy[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dd[u] = a.d[i + u * stride];
            pp[u] = a.psi[i + u * stride];
            yy[u] = a.ynew[i + u * stride];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[0] += dd[u] + pp[u];
            v[1] += fabs(yy[u]);
        }
    }
    for (; i < a.n; i += stride) {
        v[0] += a.d[i] + a.psi[i];
        v[1] += fabs(a.ynew[i]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double t = 0.0;
        for (int r = 0; r < a.R; ++r) t += a.d[a.off + r] + a.psi[a.off + r];
        a.ms[2] = t;
    }
    block_reduce_store<2>(v, 2, a.partials, a.counter, a.ms);
}
// ... and the defect is removed by rescaling the solution, d_i -= defect * |ynew_i| / sum|ynew|: a relative
// perturbation of the size of the linear residual that puts nothing onto (near-)empty boundary states.
// Skipped when the defect is not small (a genuinely leaking model must not be "repaired").
__global__ void __launch_bounds__(BT) k_bdf_massfix(int64_t n, const double* __restrict__ ms, double limit,
                                                     double* __restrict__ d, double* __restrict__ ynew) {
    const int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    if (i >= n) return;
    const double defect = ms[0] + ms[2];
    if (!(fabs(defect) <= limit * ms[1]) || !(ms[1] > 0.0)) return;   // limit is relative to sum|ynew|
    const double delta = -defect / ms[1] * fabs(ynew[i]);
    d[i] += delta;
    ynew[i] += delta;
}

// result[0] = sum (d / (atol + rtol |ynew|))^2 over the implicit rows; result[1 + v*R + r] = sink entry r of vector v
struct ErrArgs {
    int64_t n;
    const double* d;
    const double* ynew;
    double atol, rtol;
    int ntail, R;
    int64_t off;
    PtrList tails;
    double* partials;
    unsigned int* counter;
    double* result;
};
__global__ void __launch_bounds__(BT) k_bdf_errnorm(const __grid_constant__ ErrArgs a) {
    double v[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * BT;
    int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    for (; i + 3 * stride < a.n; i += 4 * stride) {   // loads of four grid strides in flight, same summation order
        double dd[4], yy[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dd[u] = a.d[i + u * stride];
            yy[u] = a.ynew[i + u * stride];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double q = dd[u] / (a.atol + a.rtol * fabs(yy[u]));
            v[0] = fma(q, q, v[0]);
        }
    }
    for (; i < a.n; i += stride) {
        const double q = a.d[i] / (a.atol + a.rtol * fabs(a.ynew[i]));
        v[0] = fma(q, q, v[0]);
    }
    if (blockIdx.x == 0)
        for (int q = threadIdx.x; q < a.ntail * a.R; q += BT) a.result[1 + q] = a.tails.p[q / a.R][a.off + q % a.R];
    block_reduce_store<1>(v, 1, a.partials, a.counter, a.result);
}

// order selection: result[0] = sum (Dm / scale)^2 , result[1] = sum (Dp / scale)^2 with scale from y = D_0
struct OrdArgs {
    int64_t n;
    const double* y;
    const double* Dm;
    const double* Dp;
    double atol, rtol;
    double* partials;
    unsigned int* counter;
    double* result;
};
__global__ void __launch_bounds__(BT) k_bdf_ordnorms(const __grid_constant__ OrdArgs a) {
    double v[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * BT;
    int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    for (; i + 3 * stride < a.n; i += 4 * stride) {   // loads of four grid strides in flight, same summation order
        double yy[4], dm[4], dp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            yy[u] = a.y[i + u * stride];
            dm[u] = a.Dm ? a.Dm[i + u * stride] : 0.0;
            dp[u] = a.Dp ? a.Dp[i + u * stride] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double inv = 1.0 / (a.atol + a.rtol * fabs(yy[u]));
            const double qm = dm[u] * inv, qp = dp[u] * inv;
            v[0] = fma(qm, qm, v[0]);
            v[1] = fma(qp, qp, v[1]);
        }
    }
    for (; i < a.n; i += stride) {
        const double inv = 1.0 / (a.atol + a.rtol * fabs(a.y[i]));
        const double qm = a.Dm ? a.Dm[i] * inv : 0.0, qp = a.Dp ? a.Dp[i] * inv : 0.0;
        v[0] = fma(qm, qm, v[0]);
        v[1] = fma(qp, qp, v[1]);
    }
    block_reduce_store<2>(v, 2, a.partials, a.counter, a.result);
}

// accept: D_{k+2} = d - D_{k+1}; D_{k+1} = d; D_i += D_{i+1} (i = k..0)
__global__ void __launch_bounds__(BT) k_bdf_update(int64_t N, int order, PtrListRW D, const double* __restrict__ d) {
    const int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    if (i >= N) return;
    const double dd = d[i];
    D.p[order + 2][i] = dd - D.p[order + 1][i];
    D.p[order + 1][i] = dd;
    double run = dd;
    for (int j = order; j >= 0; --j) {
        run += D.p[j][i];
        D.p[j][i] = run;
    }
}

// step-size change: D[0..k] <- (R U)^T D[0..k]
struct ChangeArgs {
    int64_t N;
    int order;
    PtrListRW D;
    double RU[MAX_ORDER + 1][MAX_ORDER + 1];
};
__global__ void __launch_bounds__(BT) k_bdf_change(const __grid_constant__ ChangeArgs a) {
    const int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x;
    if (i >= a.N) return;
    double in[MAX_ORDER + 1], out[MAX_ORDER + 1];
#pragma unroll
    for (int j = 0; j <= MAX_ORDER; ++j) in[j] = j <= a.order ? a.D.p[j][i] : 0.0;
#pragma unroll
    for (int r = 0; r <= MAX_ORDER; ++r) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j <= MAX_ORDER; ++j) s = fma(a.RU[j][r], in[j], s);
        out[r] = s;
    }
#pragma unroll
    for (int r = 0; r <= MAX_ORDER; ++r)
        if (r <= a.order) a.D.p[r][i] = out[r];
}

__global__ void k_zero_range(double* p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

// ---- host-side pieces ---------------------------------------------------------------------------------------------
static void compute_R(int order, double factor, double R[MAX_ORDER + 1][MAX_ORDER + 1]) {
    double M[MAX_ORDER + 1][MAX_ORDER + 1] = {};
    for (int j = 0; j <= order; ++j) M[0][j] = 1.0;
    for (int i = 1; i <= order; ++i)
        for (int j = 1; j <= order; ++j) M[i][j] = ((double)i - 1.0 - factor * j) / (double)i;
    for (int j = 0; j <= order; ++j) {
        double run = 1.0;
        for (int i = 0; i <= order; ++i) {
            run *= (i == 0) ? M[0][j] : M[i][j];
            R[i][j] = (i == 0) ? M[0][j] : run;
        }
    }
    // column 0: M[i][0] = 0 for i >= 1 -> cumprod gives 1, 0, 0, ...
    R[0][0] = 1.0;
    for (int i = 1; i <= order; ++i) R[i][0] = 0.0;
}

static inline unsigned grid_for(int64_t n) { return (unsigned)std::max<int64_t>(1, (n + BT - 1) / BT); }
static inline unsigned red_grid(int64_t n) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(RED_BLOCKS, (n + (int64_t)BT * 4 - 1) / ((int64_t)BT * 4)));
}

int solve_bdf(const OdeSystem& sys, ncme_save_fn save_fn, void* user, double t0, double t1, double* u,
              const ncme_solve_opts* o, ncme_solve_stats* st) {
    ncme_ctx* ctx = sys.ctx;
    ncme_comm* comm = sys.comm;
    cudaStream_t s = ctx->stream;
    const int64_t N = sys.len, n = sys.n_impl, off = sys.sink_off;
    const int R = sys.R;
    const bool sinks_explicit = (n < N);
    const int64_t Nglob = sys.len_global;
    const int64_t launches0 = ctx->launches;
    NCME_REQUIRE(sys.jac_diag && sys.rhs && (!sinks_explicit || sys.rhs_sinks), "BDF: incomplete system description");
    NCME_REQUIRE(!sinks_explicit || (off == n && n + R == N), "BDF: unexpected sink layout");

    // ---- workspace: 8 difference vectors, ypred/ynew/z (operator inputs: halo margins), psi, Ay, scale, ps, jdiag,
    //      w, d, GM_M+1 Krylov vectors
    const size_t hl = round_up<size_t>((size_t)sys.hl, 32), hh = round_up<size_t>((size_t)sys.hh, 32);
    const size_t Npad = hl + round_up<size_t>((size_t)N, 32) + hh;
    const int NV = (MAX_ORDER + 3) + 3 + 7 + (GM_M + 1);
    // partial sums of the fused Krylov inner products: one slot row per CTA (deterministic two-stage reduction)
    const int GM_BLOCKS = 148 * 8;
    NCME_TRY(cache_reserve(&ctx->gm_partials, &ctx->gm_partials_bytes, (size_t)GM_BLOCKS * RED_SLOTS * sizeof(double), false));
    double* gm_partials = ctx->gm_partials;
    double* base = nullptr;
    if (comm) {
        NCME_TRY(comm_workspace(comm, Npad * NV * sizeof(double), (int64_t)hl, (int64_t)Npad, NV, sys.peers, 2, &base));
    } else {
        NCME_TRY(cache_reserve(&ctx->solve_ws, &ctx->solve_ws_bytes, Npad * NV * sizeof(double), false));
        base = ctx->solve_ws;
    }
    int slot = 0;
    auto vec = [&]() { return base + Npad * (slot++) + hl; };
    double* D[MAX_ORDER + 3];
    for (int j = 0; j < MAX_ORDER + 3; ++j) D[j] = vec();
    double* ypred = vec();
    double* ynew = vec();
    double* z = vec();
    double* psi = vec();
    double* Ay = vec();
    double* scale = vec();
    double* ps = vec();
    double* jdiag = vec();
    double* w = vec();
    double* d = vec();
    double* V[GM_M + 1];
    for (int j = 0; j <= GM_M; ++j) V[j] = vec();
    NCME_CUDA(cudaMemsetAsync(base, 0, Npad * NV * sizeof(double), s));

    SliceSaver saver;
    NCME_TRY(saver.init(sys, save_fn, user, st));
    const double rtol = o->rtol > 0 ? o->rtol : 1e-4, atol = o->atol > 0 ? o->atol : 1e-6;
    const int64_t max_steps = o->max_steps > 0 ? o->max_steps : 100000000;
    const double tspan = t1 - t0;
    const unsigned gm_grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(GM_BLOCKS, (n + (int64_t)BT * 2 - 1) / ((int64_t)BT * 2)));

    // NDF coefficients (Shampine & Reichelt; same constants as scipy's BDF)
    const double kappa[MAX_ORDER + 1] = {0, -0.1850, -1.0 / 9, -0.0823, -0.0415, 0};
    double gamma[MAX_ORDER + 1], alpha[MAX_ORDER + 1], error_const[MAX_ORDER + 2];
    gamma[0] = 0;
    for (int k = 1; k <= MAX_ORDER; ++k) gamma[k] = gamma[k - 1] + 1.0 / k;
    for (int k = 0; k <= MAX_ORDER; ++k) alpha[k] = (1 - kappa[k]) * gamma[k];
    for (int k = 0; k <= MAX_ORDER; ++k) error_const[k] = kappa[k] * gamma[k] + 1.0 / (k + 1);
    error_const[MAX_ORDER + 1] = 1.0 / (MAX_ORDER + 2);

    const bool host_reduce = comm_hostreduce_available(comm);
    // NCME_BDF_PROFILE=1: where the wall time of a segment goes (host blocked on the stream = the GPU was busy;
    // the rest = the GPU waited for the host) -- printed to stderr by rank 0 at the end of the segment
    static const bool profile = getenv("NCME_BDF_PROFILE") != nullptr;
    double prof_sync = 0.0, prof_hostred = 0.0;
    long long prof_nsync = 0;
    const double prof_t0 = profile ? wall_seconds() : 0.0;
    auto sync_stream = [&]() -> int {
        const double ta = profile ? wall_seconds() : 0.0;
        NCME_CUDA(cudaStreamSynchronize(s));
        if (profile) {
            prof_sync += wall_seconds() - ta;
            ++prof_nsync;
        }
        return NCME_OK;
    };
    auto fetch = [&](size_t count) -> int {   // reduced scalars -> pinned host
        if (host_reduce && count <= (size_t)NCME_HOSTREDUCE_MAX) {
            // sharded: this rank's partial sums go to the host, the ranks combine them through shared memory
            NCME_CUDA(cudaMemcpyAsync(ctx->red_result_host, ctx->red_result_dev, sizeof(double) * count, cudaMemcpyDeviceToHost, s));
            NCME_TRY(sync_stream());
            const double ta = profile ? wall_seconds() : 0.0;
            const int rc = comm_hostreduce_sum(comm, ctx->red_result_host, count);
            if (profile) prof_hostred += wall_seconds() - ta;
            return rc;
        }
        NCME_TRY(comm_allreduce_sum(comm, ctx->red_result_dev, count, s));
        NCME_CUDA(cudaMemcpyAsync(ctx->red_result_host, ctx->red_result_dev, sizeof(double) * count, cudaMemcpyDeviceToHost, s));
        NCME_TRY(sync_stream());
        return NCME_OK;
    };
    auto rhs = [&](double t, const double* x, double* y) -> int {
        st->rhs_evals++;
        return sys.rhs_safe ? sys.rhs_safe(t, x, y) : sys.rhs(t, x, y);
    };
    auto change_D = [&](int order, double factor) -> int {
        double Rm[MAX_ORDER + 1][MAX_ORDER + 1] = {}, Um[MAX_ORDER + 1][MAX_ORDER + 1] = {};
        compute_R(order, factor, Rm);
        compute_R(order, 1.0, Um);
        ChangeArgs ca;
        ca.N = N;
        ca.order = order;
        for (int j = 0; j < MAX_ORDER + 3; ++j) ca.D.p[j] = D[j];
        for (int i = 0; i <= MAX_ORDER; ++i)
            for (int j = 0; j <= MAX_ORDER; ++j) {
                double v = 0.0;
                if (i <= order && j <= order)
                    for (int q = 0; q <= order; ++q) v += Rm[i][q] * Um[q][j];
                ca.RU[i][j] = v;
            }
        k_bdf_change<<<grid_for(N), BT, 0, s>>>(ca);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        return NCME_OK;
    };
    auto finish = [&](const double* src) -> int {
        if (src != u) NCME_CUDA(cudaMemcpyAsync(u, src, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s));
        if (sinks_explicit) NCME_TRY(comm_allreduce_sum(comm, u + off, (size_t)R, s));
        NCME_CUDA(cudaStreamSynchronize(s));
        st->launches = ctx->launches - launches0;
        if (profile && (!comm || comm->rank == 0))
            fprintf(stderr, "[ncme bdf profile] wall %.4f s, blocked on the stream %.4f s in %lld syncs, host all-reduce %.4f s, "
                            "%lld launches, %lld rhs, %lld steps\n", wall_seconds() - prof_t0, prof_sync, prof_nsync, prof_hostred,
                    (long long)st->launches, (long long)st->rhs_evals, (long long)st->steps);
        if (comm && comm->my_flags) {
            unsigned int perr = 0;
            NCME_CUDA(cudaMemcpy(&perr, &comm->my_flags->error, sizeof(perr), cudaMemcpyDeviceToHost));
            if (perr) {
                set_error("peer-memory halo: a neighbouring rank did not signal within 2 s");
                return NCME_ERR_COMM;
            }
        }
        return NCME_OK;
    };

    // ---- initial state: D0 = u (sharded: only rank 0 keeps the incoming replicated sink values), D1 = h f(t0,u)
    NCME_CUDA(cudaMemcpyAsync(D[0], u, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (comm && comm->rank != 0 && sinks_explicit) {
        k_zero_range<<<1, 64, 0, s>>>(D[0] + off, R);
        ctx->launches++;
    }
    st->t_final = t0;
    st->event_hit = 0;
    int isave = 0;
    while (isave < o->nsave && o->save_t[isave] < t0) ++isave;
    if (o->save_every_step) NCME_TRY(saver.save(t0, D[0]));
    while (isave < o->nsave && o->save_t[isave] == t0) {
        NCME_TRY(saver.save(t0, D[0]));
        ++isave;
    }
    if (!(tspan > 0)) return finish(D[0]);

    NCME_TRY(rhs(t0, D[0], Ay));
    double h_abs = o->h_init;
    if (!(h_abs > 0)) {
        const int64_t nn = comm ? n : N;
        double d0 = 0, d1 = 0;
        NCME_TRY(ncme_vec_wrms(ctx, nn > 0 ? nn : 1, D[0], D[0], D[0], atol, rtol, &d0));
        NCME_TRY(ncme_vec_wrms(ctx, nn > 0 ? nn : 1, Ay, D[0], D[0], atol, rtol, &d1));
        double ss[2] = {d0 * d0 * (double)nn, d1 * d1 * (double)nn};
        if (comm) {
            NCME_CUDA(cudaMemcpyAsync(comm->scratch, ss, sizeof(ss), cudaMemcpyHostToDevice, s));
            NCME_TRY(comm_allreduce_sum(comm, comm->scratch, 2, s));
            NCME_CUDA(cudaMemcpyAsync(ss, comm->scratch, sizeof(ss), cudaMemcpyDeviceToHost, s));
            NCME_CUDA(cudaStreamSynchronize(s));
        }
        d0 = sqrt(ss[0] / (double)Nglob);
        d1 = sqrt(ss[1] / (double)Nglob);
        h_abs = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h_abs = std::min(h_abs, tspan);
    }
    {   // D1 = h * f0
        const double cs[1] = {h_abs};
        const double* xs[1] = {Ay};
        NCME_TRY(ncme_vec_lincomb(ctx, N, 1, cs, xs, D[1]));
    }
    int order = 1, n_equal_steps = 0;
    double t = t0;
    double g_prev = 0.0;
    bool have_g = false;
    // smallest resolvable step: relative to where the segment STARTS (a long horizon must not forbid the tiny first
    // steps a tight absolute tolerance asks for: toggle example, odeatol = 1e-14 over 8 h)
    const double hmin = std::max(1e-14 * fabs(t0), 1e-20 * fabs(t1 - t0));
    // weighted-RMS residual of the linear solve = CVODE's 0.05 x Newton tolerance 0.1.  The residual of the inexact
    // solve is the only source of total-mass drift (1^T A = 0); it is removed by the invariant projection below.
    const double lin_tol = 5e-3;
    std::vector<double> tails((size_t)(MAX_ORDER + 4) * std::max(R, 1));
    // Hessenberg columns of the running Krylov cycle: device scalars (written by k_gm_apply_dots) and their host image
    constexpr int HCOL_LD = 32, HCOL_OFF = 32;
    static_assert(HCOL_OFF + HCOL_LD * (GM_M + 1) <= 1000 && RED_SLOTS <= HCOL_LD, "Hessenberg columns must fit the scalar buffers");
    auto hcol_dev = [&](int k) { return ctx->red_result_dev + HCOL_OFF + (size_t)HCOL_LD * k; };
    auto hcol_host = [&](int k) { return ctx->red_result_host + HCOL_OFF + (size_t)HCOL_LD * k; };
    int k_pred = 2;   // Krylov iterations of the previous linear solve (batch size of the next one)

    while (t < t1) {
        if (abort_requested()) return abort_status();   // a save callback failed
        if (st->steps + st->rejected >= max_steps) {
            set_error("integrator: maximum number of steps (%lld) reached at t = %g", (long long)max_steps, t);
            return NCME_ERR_SOLVER;
        }
        if (h_abs < hmin) {
            set_error("integrator (BDF): step size underflow at t = %g", t);
            return NCME_ERR_SOLVER;
        }
        double t_new = t + h_abs;
        if (t_new > t1 || t1 - t_new < 1e-12 * tspan) {
            t_new = t1;
            NCME_TRY(change_D(order, fabs(t_new - t) / h_abs));
            n_equal_steps = 0;
        }
        const double h = t_new - t;
        h_abs = fabs(h);
        const double c = h / alpha[order];

        // ---- predictor and right-hand side of the linear system
        PtrList DL{};
        for (int j = 0; j < MAX_ORDER + 3; ++j) DL.p[j] = D[j];
        k_bdf_predict<<<grid_for(N), BT, 0, s>>>(N, order, DL, gamma[1], gamma[2], gamma[3], gamma[4], gamma[5],
                                                 1.0 / alpha[order], ypred, psi);
        ctx->launches++;
        NCME_TRY(rhs(t_new, ypred, Ay));
        NCME_TRY(sys.jac_diag(t_new, jdiag));
        SetupArgs sa{n, ypred, Ay, psi, jdiag, c, atol, rtol, scale, ps, w, ctx->red_partials, ctx->red_counter, ctx->red_result_dev};
        k_bdf_setup<<<red_grid(n), BT, 0, s>>>(sa);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        NCME_TRY(fetch(1));
        double beta = sqrt(std::max(ctx->red_result_host[0], 0.0));
        const double sqrtn = sqrt((double)std::max<int64_t>(1, Nglob - (sinks_explicit ? R : 0)));

        // ---- GMRES(m) on  W P^-1 (I - cA) W^-1 dt = W P^-1 b  (x0 = 0), restarted
        bool lin_ok = true;
        bool have_d = false;
        if (!(beta / sqrtn > lin_tol * 1e-3)) {
            // right-hand side already negligible: d = 0
            k_zero_range<<<grid_for(n), 256, 0, s>>>(d, n);
            NCME_CUDA(cudaMemcpyAsync(ynew, ypred, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
            ctx->launches++;
            if (!(beta == beta)) lin_ok = false;
        } else {
            int restarts = 0;
            while (true) {
                double H[GM_M + 1][GM_M] = {};
                double cs_[GM_M] = {}, sn_[GM_M] = {}, gvec[GM_M + 1] = {};
                gvec[0] = beta;
                k_gm_normalize<<<grid_for(n), BT, 0, s>>>(n, w, 1.0 / beta, scale, V[0], z);
                ctx->launches++;
                int k = 0;
                double resid = beta;
                // Krylov iterations are enqueued in BATCHES: the inner products stay on the device (summed over the
                // ranks inside k_gm_apply_dots), k_gm_ortho reads them there, and the host fetches the Hessenberg
                // columns of a whole batch with one copy.  The first batch of a cycle speculates on one iteration
                // fewer than the previous linear solve needed, then single iterations follow: (almost) no wasted
                // matvec, 2-3 host round trips per solve instead of one per iteration.
                bool cycle_done = false;
                bool first_batch = true;
                while (!cycle_done && lin_ok) {
                    const int nb = first_batch ? std::max(1, std::min(k_pred - 1, GM_M - k)) : 1;
                    first_batch = false;
                    for (int b = 0; b < nb; ++b) {
                        const int kk = k + b;
                        if (kk > 0) {   // v_kk, z from column kk-1 (device-resident coefficients)
                            OrthoArgs oa{};
                            oa.n = n;
                            oa.k = kk - 1;
                            oa.w = w;
                            for (int j = 0; j <= kk - 1; ++j) oa.V.p[j] = V[j];
                            for (int j = kk; j < GM_M + 2; ++j) oa.V.p[j] = V[0];
                            oa.hcol = hcol_dev(kk - 1);
                            oa.scale = scale;
                            oa.vout = V[kk];
                            oa.z = z;
                            launch_ortho(oa, grid_for(n), s);
                            ctx->launches++;
                        }
                        NCME_TRY(rhs(t_new, z, Ay));
                        ApplyArgs aa{};
                        aa.n = n;
                        aa.k = kk;
                        aa.z = z;
                        aa.Az = Ay;
                        aa.ps = ps;
                        aa.c = c;
                        for (int j = 0; j <= kk; ++j) aa.V.p[j] = V[j];
                        for (int j = kk + 1; j < GM_M + 2; ++j) aa.V.p[j] = V[0];
                        aa.w = w;
                        aa.partials = gm_partials;
                        aa.counter = ctx->red_counter;
                        aa.result = hcol_dev(kk);
                        const bool in_kernel = comm_dev_allreduce(comm, &aa.ar);
                        NCME_TRY(launch_apply_dots(ctx, aa, gm_grid));
                        if (!in_kernel) NCME_TRY(comm_allreduce_sum(comm, hcol_dev(kk), RED_SLOTS, s));
                    }
                    NCME_CUDA(cudaMemcpyAsync(hcol_host(k), hcol_dev(k), sizeof(double) * HCOL_LD * nb, cudaMemcpyDeviceToHost, s));
                    NCME_TRY(sync_stream());
                    if (abort_requested()) return abort_status();
                    for (int b = 0; b < nb && !cycle_done; ++b, ++k) {
                        const double* hh = hcol_host(k);
                        const double ww = hh[RED_SLOTS - 1];
                        for (int j = 0; j <= k; ++j) H[j][k] = hh[j];
                        const double hk1 = gm_hk1(hh, k);
                        H[k + 1][k] = hk1;
                        // Givens rotations on column k
                        for (int j = 0; j < k; ++j) {
                            const double tmp = cs_[j] * H[j][k] + sn_[j] * H[j + 1][k];
                            H[j + 1][k] = -sn_[j] * H[j][k] + cs_[j] * H[j + 1][k];
                            H[j][k] = tmp;
                        }
                        const double den = hypot(H[k][k], H[k + 1][k]);
                        if (!(den > 0.0) || !(ww == ww)) {
                            lin_ok = false;
                            break;
                        }
                        cs_[k] = H[k][k] / den;
                        sn_[k] = H[k + 1][k] / den;
                        H[k][k] = den;
                        H[k + 1][k] = 0.0;
                        gvec[k + 1] = -sn_[k] * gvec[k];
                        gvec[k] = cs_[k] * gvec[k];
                        resid = fabs(gvec[k + 1]);
                        const bool happy = hk1 <= 1e-14 * sqrt(std::max(ww, 1e-300));
                        if (resid / sqrtn <= lin_tol || happy || k + 1 == GM_M) cycle_done = true;   // k is advanced by the loop
                    }
                }
                if (lin_ok && restarts == 0) k_pred = k;
                if (!lin_ok) break;
                // back substitution H y = g
                double yk[GM_M + 1] = {};
                for (int i = k - 1; i >= 0; --i) {
                    double acc = gvec[i];
                    for (int j = i + 1; j < k; ++j) acc -= H[i][j] * yk[j];
                    yk[i] = acc / H[i][i];
                }
                SolArgs so{};
                so.n = n;
                so.k = k;
                for (int j = 0; j < k; ++j) {
                    so.V.p[j] = V[j];
                    so.y[j] = yk[j];
                }
                so.scale = scale;
                so.ypred = ypred;
                so.d = d;
                so.ynew = ynew;
                so.accumulate = have_d ? 1 : 0;
                k_gm_solution<<<grid_for(n), BT, 0, s>>>(so);
                ctx->launches++;
                have_d = true;
                if (resid / sqrtn <= lin_tol) break;
                if (++restarts > 3) {
                    lin_ok = false;
                    break;
                }
                // restart: residual of the current d:  r = (c A ypred - psi - (d - c A d)) ps = (c A ynew - psi - d) ps
                NCME_TRY(rhs(t_new, ynew, Ay));
                // reuse the setup kernel's algebra through a lincomb:  tmp = c*Ay - psi - d ; w = tmp * ps
                {
                    const double cs3[3] = {c, -1.0, -1.0};
                    const double* xs3[3] = {Ay, psi, d};
                    NCME_TRY(ncme_vec_lincomb(ctx, n, 3, cs3, xs3, w));
                    // w *= ps and norm: reuse apply kernel with z = w, Az = 0-vector trick is awkward; do it plainly
                    ApplyArgs aa{};
                    aa.n = n;
                    aa.k = -1;
                    aa.z = w;
                    aa.Az = w;
                    aa.ps = ps;
                    aa.c = 0.0;
                    for (int j = 0; j < GM_M + 2; ++j) aa.V.p[j] = V[0];
                    aa.w = w;
                    aa.partials = gm_partials;
                    aa.counter = ctx->red_counter;
                    aa.result = ctx->red_result_dev;
                    NCME_TRY(launch_apply_dots(ctx, aa, gm_grid));
                    NCME_TRY(fetch(RED_SLOTS));
                    beta = sqrt(std::max(ctx->red_result_host[RED_SLOTS - 1], 0.0));
                    if (!(beta > 0.0)) break;
                }
            }
        }
        if (!lin_ok) {   // linear solver failed: halve the step (CVODE's reaction to a convergence failure)
            st->rejected++;
            h_abs *= 0.5;
            NCME_TRY(change_D(order, 0.5));
            n_equal_steps = 0;
            continue;
        }
        // ---- sink entries of d (explicit, linear), with the total-mass defect of the inexact solve removed in between
        if (sinks_explicit) {
            NCME_TRY(sys.rhs_sinks(t_new, ynew, Ay));
            k_bdf_sinks<<<1, 64, 0, s>>>(R, off, c, Ay, psi, ypred, d, ynew);
            ctx->launches++;
            double* ms = ctx->red_result_dev + 1000;   // device scalars, untouched by the other reductions
            MassArgs ma{n, d, psi, ynew, R, off, ctx->red_partials, ctx->red_counter, ms};
            k_bdf_massdefect<<<red_grid(n), BT, 0, s>>>(ma);
            ctx->launches++;
            NCME_TRY(comm_allreduce_sum(comm, ms, 3, s));
            // "small" = a relative change below 10 rtol
            k_bdf_massfix<<<grid_for(n), BT, 0, s>>>(n, ms, 10.0 * rtol, d, ynew);
            ctx->launches++;
            NCME_TRY(sys.rhs_sinks(t_new, ynew, Ay));
            k_bdf_sinks<<<1, 64, 0, s>>>(R, off, c, Ay, psi, ypred, d, ynew);
            ctx->launches++;
            NCME_CUDA(cudaGetLastError());
        }
        // ---- local error test; the same D2H carries the sink tails of D_0..D_{k+1}, d (event) and ypred
        const int ntail = R > 0 ? order + 4 : 0;   // sink tails are always fetched (event); their error is added
                                                   // on the host only when they are not part of the implicit rows
        ErrArgs ea{};
        ea.n = n;
        ea.d = d;
        ea.ynew = ynew;
        ea.atol = atol;
        ea.rtol = rtol;
        ea.ntail = ntail;
        ea.R = R;
        ea.off = off;
        for (int j = 0; j <= order + 1; ++j) ea.tails.p[j] = D[j];
        ea.tails.p[order + 2] = d;
        ea.tails.p[order + 3] = ypred;
        ea.partials = ctx->red_partials;
        ea.counter = ctx->red_counter;
        ea.result = ctx->red_result_dev;
        k_bdf_errnorm<<<red_grid(n), BT, 0, s>>>(ea);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        NCME_TRY(fetch((size_t)(1 + ntail * R)));
        double sumsq = ctx->red_result_host[0];
        const double* tl = ctx->red_result_host + 1;
        for (int q = 0; q < ntail * R; ++q) tails[q] = tl[q];
        if (sinks_explicit)
            for (int r = 0; r < R; ++r) {
                const double ds = tails[(size_t)(order + 2) * R + r], yn = tails[(size_t)(order + 3) * R + r] + ds;
                const double q = ds / (atol + rtol * fabs(yn));
                sumsq += q * q;
            }
        const double safety = 0.9;   // one (exact) Newton iteration
        const double error_norm = error_const[order] * sqrt(sumsq / (double)Nglob);
        if (!(error_norm <= 1.0)) {
            st->rejected++;
            const double factor = (error_norm == error_norm) ? std::max(0.2, safety * pow(error_norm, -1.0 / (order + 1))) : 0.2;
            h_abs *= factor;
            NCME_TRY(change_D(order, factor));
            n_equal_steps = 0;
            continue;
        }
        // ---- accept
        st->steps++;
        n_equal_steps++;
        PtrListRW DW{};
        for (int j = 0; j < MAX_ORDER + 3; ++j) DW.p[j] = D[j];
        k_bdf_update<<<grid_for(N), BT, 0, s>>>(N, order, DW, d);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        // new difference tails of the sinks, on the host (the update is linear)
        double nd[MAX_ORDER + 3][NCME_MAX_REACTIONS] = {};
        if (R > 0) {
            for (int r = 0; r < R; ++r) {
                const double ds = tails[(size_t)(order + 2) * R + r];
                nd[order + 2][r] = ds - tails[(size_t)(order + 1) * R + r];
                nd[order + 1][r] = ds;
                double run = ds;
                for (int j = order; j >= 0; --j) {
                    run += tails[(size_t)j * R + r];
                    nd[j][r] = run;
                }
            }
        }
        // dense output of the sinks: y(tt) = D0 + sum_j D_j prod_{m<j} (tt - (t_new - m h)) / (h (m+1))
        auto sink_sum_at = [&](double tt) {
            double sum = 0.0;
            for (int r = 0; r < R; ++r) {
                double p = 1.0, y = nd[0][r];
                for (int j = 1; j <= order; ++j) {
                    p *= (tt - (t_new - (j - 1) * h)) / (h * j);
                    y += nd[j][r] * p;
                }
                sum += y;
            }
            return sum;
        };
        auto dense_to = [&](double tt, double* out) -> int {
            double cs[MAX_ORDER + 1];
            const double* xs[MAX_ORDER + 1];
            cs[0] = 1.0;
            xs[0] = D[0];
            double p = 1.0;
            for (int j = 1; j <= order; ++j) {
                p *= (tt - (t_new - (j - 1) * h)) / (h * j);
                cs[j] = p;
                xs[j] = D[j];
            }
            return ncme_vec_lincomb(ctx, N, order + 1, cs, xs, out);
        };
        double t_hi = t_new;
        bool event = false;
        if (o->check_event && R > 0) {
            if (!have_g) {
                double s0 = 0.0;
                for (int r = 0; r < R; ++r) s0 += tails[r];   // D_0 before the step = u(t)
                g_prev = s0 - o->event_slope * t;
                have_g = true;
            }
            const int NS_ = 16;
            double ga = g_prev, ta = t;
            for (int q = 1; q <= NS_; ++q) {
                const double tb = t + (t_new - t) * q / NS_;
                const double gb = sink_sum_at(tb) - o->event_slope * tb;
                if (ga <= 0.0 && gb > 0.0) {
                    double lo = ta, hi = tb;
                    for (int it = 0; it < 60; ++it) {
                        const double mid = 0.5 * (lo + hi);
                        if (sink_sum_at(mid) - o->event_slope * mid > 0.0)
                            hi = mid;
                        else
                            lo = mid;
                    }
                    t_hi = hi;
                    event = true;
                    break;
                }
                ga = gb;
                ta = tb;
            }
            if (!event) g_prev = ga;
        }
        while (isave < o->nsave && o->save_t[isave] <= t_hi + 1e-14 * fabs(t_hi)) {
            const double ts = o->save_t[isave];
            if (ts >= t_new && !event) {
                NCME_TRY(saver.save(ts, D[0]));
            } else {
                NCME_TRY(dense_to(std::min(ts, t_new), z));
                NCME_TRY(saver.save(ts, z));
            }
            ++isave;
        }
        if (event) {
            NCME_TRY(dense_to(t_hi, z));
            st->t_final = t_hi;
            st->event_hit = 1;
            st->h_last = h_abs;
            return finish(z);
        }
        t = t_new;
        st->h_last = h_abs;
        if (o->save_every_step) NCME_TRY(saver.save(t, D[0]));
        if (t >= t1) break;
        if (n_equal_steps < order + 1) continue;
        // ---- order / step-size selection
        OrdArgs oa{};
        oa.n = n;
        oa.y = D[0];
        oa.Dm = order > 1 ? D[order] : nullptr;
        oa.Dp = order < MAX_ORDER ? D[order + 2] : nullptr;
        oa.atol = atol;
        oa.rtol = rtol;
        oa.partials = ctx->red_partials;
        oa.counter = ctx->red_counter;
        oa.result = ctx->red_result_dev;
        k_bdf_ordnorms<<<red_grid(n), BT, 0, s>>>(oa);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        NCME_TRY(fetch(2));
        double sm = ctx->red_result_host[0], sp = ctx->red_result_host[1];
        if (sinks_explicit)
            for (int r = 0; r < R; ++r) {
                const double inv = 1.0 / (atol + rtol * fabs(nd[0][r]));
                sm += (nd[order][r] * inv) * (nd[order][r] * inv);
                sp += (nd[order + 2][r] * inv) * (nd[order + 2][r] * inv);
            }
        const double INF = 1e300;
        const double em = order > 1 ? error_const[order - 1] * sqrt(sm / (double)Nglob) : INF;
        const double ep = order < MAX_ORDER ? error_const[order + 1] * sqrt(sp / (double)Nglob) : INF;
        const double norms[3] = {em, error_norm, ep};
        double factors[3];
        for (int q = 0; q < 3; ++q)
            factors[q] = norms[q] >= INF ? 0.0 : (norms[q] > 0 ? pow(norms[q], -1.0 / (order + q)) : 1e9);
        int best = 1;
        for (int q = 0; q < 3; ++q)
            if (factors[q] > factors[best]) best = q;
        order += best - 1;
        const double factor = std::min(10.0, safety * factors[best]);
        h_abs *= factor;
        NCME_TRY(change_D(order, factor));
        n_equal_steps = 0;
    }
    st->t_final = t;
    return finish(D[0]);
}

}  // namespace ncme
