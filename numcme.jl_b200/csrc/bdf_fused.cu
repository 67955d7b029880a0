// Fused-step variant of the native BDF/NDF + Jacobi-GMRES integrator (bdf.cu): ONE kernel launch per step attempt.
//
// Why: the launch-per-operation integrator of bdf.cu needs ~34 launches and ~10 device->host scalar fetches per
// step (one per Krylov iteration).  That is irrelevant at 10^7 states (a matvec is 150 us) but it is all the cost of
// the reference's own example configurations (telegraph: 70 states, toggle switch: 10^3, Hog1p: 10^4..10^5 --
// SURVEY.md H4): 83 steps took 26 ms, against 5.45 ms for the reference's whole adaptive solve on a laptop CPU
// (docs/src/examples/telegraph.md:89).  Here the whole step -- pending rescaling of the difference array, predictor,
// A(t_new)*y_pred, Jacobi/weight setup, restarted GMRES(24) (matvec, Gram-Schmidt inner products, Givens rotations,
// stopping test), solution update, explicit sink rows, total-mass projection, local error test, update of the
// differences and the order-selection norms -- runs inside one persistent kernel: a single CTA (`__syncthreads`
// between phases) for up to ~1000 rows, a cooperative grid (grid-wide barriers) above.  All reductions are
// two-stage and ordered (bitwise reproducible); every CTA redundantly combines the partials and runs the small
// Hessenberg/Givens recurrences, so no broadcast step is needed.  The host only sees one result record per step
// (error norm, sink tails for the event function, iteration counts) and keeps the step-size / order controller,
// the sink event and the output slices, exactly as in bdf.cu.  All RHS evaluations of a BDF step are at t_new, so
// time-varying coefficients c_r(t_new) are passed by value with the launch (one host callback per step attempt).
//
// Multi-step mode (time-invariant matrices on a single CTA, no saveat list): the controller itself (bdf_ctl.h, shared
// host/device code) also runs inside the kernel, so one launch performs up to 64 step attempts -- rejected steps, order
// and step-size selection, the sign test of the sink event and the every-step output ring included -- and returns to the
// host only for what needs it: the step that contains the event or reaches t1, a full output ring, failures.  That
// removes the ~10 us launch + result round trip per step that is left once a step is a single kernel.
//
// Replaces, like bdf.cu: DE.init / DE.step! with CVODE_BDF(linear_solver=:GMRES) (src/transientcme/sparse/fspsolve.jl:158-161).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "bdf_ctl.h"
#include "comm.cuh"
#include "matrix.cuh"
#include "ode.cuh"
#include "vec.cuh"

namespace cg = cooperative_groups;

namespace ncme {

namespace {

constexpr int FBT = 512;              // threads per CTA
constexpr int FW = FBT / 32;          // warps per CTA
constexpr int MAXO = BDF_MAXO;        // maximum BDF order
constexpr int GM = 24;                // Krylov dimension before restart
constexpr int NSLOT = GM + 2;         // widest reduction (k+1 inner products + <w,w>)
constexpr int MAXR = 32;              // reactions the fused step kernel handles (its shared-memory staging is sized by it);
                                      // models with 33..NCME_MAX_REACTIONS reactions use the launch-per-operation path (bdf.cu)
static_assert(MAXR <= NCME_MAX_REACTIONS, "fused BDF reaction limit");
constexpr int SMEM_MAX_BYTES = 211 * 1024;   // dynamic shared memory of the SMEMV variant (+ ~15 KB static <= 227 KB)

struct StepResult {                   // written by CTA 0 straight into pinned (device-mapped) host memory
    double error_sumsq;               // sum over ALL entries (states + sinks) of (d / (atol + rtol |ynew|))^2
    double error_norm;
    double ord_sm, ord_sp;            // order-selection sums over the state rows (new differences)
    double beta0, resid;
    double ms0, ms1, ms2;             // mass defect diagnostics
    int lin_ok, accepted, kiters, rhs_evals, restarts, pad0;
    volatile unsigned int seq;        // written last (after a system-scope fence): the host polls it
    int pad1;
    double sink_old0[MAXR];           // sink entries of D_0 before the step (u(t))
    double nd[MAXO + 3][MAXR];        // sink entries of the NEW differences D_0..D_{order+2} (valid when accepted)
};

struct StepArgs {
    // matrix (single GPU: padded position == row index)
    const uint32_t* col;
    const double* val;
    const double* diag;
    int64_t n, ld, N;
    int nslots, ndiag, R, G;
    double slot_coef[MAXR];
    double diag_coef[MAXR + 1];
    double sink_coef[MAXR];
    const uint32_t* sink_row;
    const double* sink_val;
    int64_t sink_ptr[MAXR + 1];
    // vectors (each of stride `stride` inside one workspace)
    double* D;        // D_j = D + j*stride, j < MAXO+3
    double* V;        // V_j = V + j*stride, j <= GM
    double *ypred, *ynew, *z, *psi, *scale, *ps, *w, *d;
    int64_t stride;
    // step description (single-step launches; in multi-step mode the in-kernel controller fills its own StepDyn)
    StepDyn dyn;
    double gamma[MAXO + 1];
    double atol, rtol, lin_tol, sqrtn, massfix_limit;
    double Nglob;
    // multi-step mode (ctl != nullptr): controller state in pinned host memory, constants, output ring
    BdfCtl* ctl;
    BdfConst kc;
    double* ring;                     // [BDF_RING][stride] every-step output slices
    int max_attempts;
    // row sharding (P > 1, cooperative launches only): column indices are positions in the halo-padded window
    // [halo_lo | local rows | R sinks | halo_hi]; halo entries are read from the neighbours' HBM (NVLink), the ranks'
    // grids meet in barriers / all-reduces through PeerFlags::fz_flag / fz_slot
    int P, me;
    uint32_t self_off, lo_end, hi_begin;
    const double *ypred_lo, *ypred_hi, *z_lo, *z_hi, *ynew_lo, *ynew_hi;   // peer views, indexed by padded position
    PeerFlags* flags[NCME_RED_RANKS];
    // scratch
    double* partials;                 // [2][G][NSLOT]
    double* sinkbuf;                  // [2][MAXR]: sum val*x, sum val*|x|
    StepResult* res;                  // device alias of the pinned host record
    unsigned int seq;
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}

struct SyncState {          // identical in every thread of the grid
    int parity;             // which half of the partials buffer the next reduction uses
    unsigned int xe;        // sequence number of the cross-GPU exchanges (sharded mode)
};

struct Shared {
    double w[FW][NSLOT];
    double res[NSLOT];
    double H[GM + 1][GM];
    double cs[GM], sn[GM], g[GM + 1], y[GM + 1], hcol[GM + 1];
    double inv_hk1, resid;
    int stop, ok;
    double sink_d[MAXR];
    double sinkS[MAXR], sinkA[MAXR];   // sink rows of A x and of A |x| (every CTA keeps a copy)
    double xg[NCME_FZ_VALS];           // staging of a cross-GPU all-reduce
};

template <bool MULTI>
__device__ __forceinline__ void grid_barrier() {
    if (MULTI)
        cg::this_grid().sync();
    else
        __syncthreads();
}

// ---- sharded mode: exchanges between the ranks' cooperative grids (all CTAs of all ranks are resident) ------------
// Every rank executes the same sequence of exchanges (all control flow derives from all-reduced values), numbered by
// sy.xe.  Slot / flag buffers are reused every NCME_RED_BUFS exchanges; every exchange is preceded by a grid-wide
// barrier on each rank, so a rank can be at most one exchange ahead of any CTA of any other rank.  Waits are bounded
// (2 s): a lost peer raises PeerFlags::error instead of hanging the GPU.
__device__ __forceinline__ unsigned long long fz_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void xg_wait_all(const StepArgs& a, int buf, unsigned int e) {
    if ((int)threadIdx.x < a.P) {
        const volatile unsigned int* f = &a.flags[a.me]->fz_flag[buf][threadIdx.x];
        const volatile unsigned int* err = &a.flags[a.me]->error;
        const unsigned long long t0 = fz_ns();
        while ((int)(*f - e) < 0) {
            if (*err) break;   // an exchange already timed out: the step is lost, do not wait 2 s at every exchange
            if (fz_ns() - t0 > 2000000000ull) {
                atomicExch(&a.flags[a.me]->error, 1u);
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
}
// Barrier over all ranks.  Precondition: a grid-wide barrier on this rank since its last writes that peers will read
// (those writes are then in this GPU's L2, where peer loads are served).
__device__ __forceinline__ void xg_barrier(const StepArgs& a, SyncState& sy) {
    const unsigned int e = ++sy.xe;
    const int buf = (int)(e % NCME_RED_BUFS);
    if (blockIdx.x == 0 && (int)threadIdx.x < a.P) {
        __threadfence_system();
        *((volatile unsigned int*)&a.flags[threadIdx.x]->fz_flag[buf][a.me]) = e;
    }
    xg_wait_all(a, buf, e);
}
// In-place sum over the ranks of vals[0..nv) (shared memory, identical in every CTA of a rank on entry); summation in
// rank order => identical bits on every rank.  Precondition as xg_barrier.  Ends with __syncthreads.
__device__ __forceinline__ void xg_allreduce(const StepArgs& a, SyncState& sy, double* vals, int nv) {
    const unsigned int e = ++sy.xe;
    const int buf = (int)(e % NCME_RED_BUFS);
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < nv)
            for (int q = 0; q < a.P; ++q) a.flags[q]->fz_slot[buf][a.me][threadIdx.x] = vals[threadIdx.x];
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < a.P) *((volatile unsigned int*)&a.flags[threadIdx.x]->fz_flag[buf][a.me]) = e;
    }
    xg_wait_all(a, buf, e);
    if ((int)threadIdx.x < nv) {
        const volatile double* sl = &a.flags[a.me]->fz_slot[buf][0][0];
        double t = 0.0;
        for (int q = 0; q < a.P; ++q) t += sl[q * NCME_FZ_VALS + threadIdx.x];
        vals[threadIdx.x] = t;
    }
    __syncthreads();
}

// Ordered two-stage reduction of nv <= NV values; afterwards sh.res[0..nv) holds the totals in EVERY CTA (identical
// bits everywhere: same partials, same summation order).  Contains a grid-wide barrier when MULTI.
template <int NV, bool MULTI>
__device__ __forceinline__ void reduce_all(double (&v)[NV], int nv, const StepArgs& a, SyncState& sy, Shared& sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // single CTA: row i is handled by thread i % FBT, so warps beyond ceil(N / 32) only ever hold zeros -- they skip the
    // shuffles and the second stage skips their slots (a 100-state system keeps 4 of 16 warps busy)
    const int nw = MULTI ? FW : (int)((a.N + 31) / 32 < FW ? (a.N + 31) / 32 : FW);
    if (wid < nw) {
#pragma unroll
        for (int s = 0; s < NV; ++s) {
            if (s < nv) {
                const double x = warp_sum(v[s]);
                if (lane == 0) sh.w[wid][s] = x;
            }
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nv) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += sh.w[w][threadIdx.x];
        if (MULTI)
            a.partials[((size_t)sy.parity * a.G + blockIdx.x) * NSLOT + threadIdx.x] = t;
        else
            sh.res[threadIdx.x] = t;
    }
    if (MULTI) {
        cg::this_grid().sync();
        const double* P = a.partials + (size_t)sy.parity * a.G * NSLOT;
        for (int s = wid; s < nv; s += FW) {
            double t = 0.0;
            for (int b = lane; b < a.G; b += 32) t += __ldcg(P + (size_t)b * NSLOT + s);
            t = warp_sum(t);
            if (lane == 0) sh.res[s] = t;
        }
        sy.parity ^= 1;
    }
    __syncthreads();
    if (MULTI && a.P > 1) xg_allreduce(a, sy, sh.res, nv);
}

struct Vecs {   // where the per-step work vectors live (global workspace or shared memory)
    double *ypred, *ynew, *z, *psi, *scale, *ps, *w, *d, *V;
    int64_t vs;   // stride between Krylov vectors
};
struct XView {   // a matvec input: local rows + (sharded) the neighbours' copies, indexed by padded position
    const double *x, *lo, *hi;
};

// (A x)_i and diag(A)_i for one state row.  Sharded: the column index is a position in the halo-padded window; halo
// entries come from the neighbour's HBM over NVLink (L1 bypassed: the same addresses are rewritten every iteration).
__device__ __forceinline__ double row_apply(const StepArgs& a, const XView& xv, int64_t i, double& jd) {
    double acc = 0.0;
    const uint32_t* cp = a.col + i;
    const double* vp = a.val + i;
    const double* xb = xv.x - a.self_off;
#pragma unroll 4
    for (int s = 0; s < a.nslots; ++s) {
        const uint32_t c = __ldg(cp + (size_t)s * a.ld);
        const double v = __ldg(vp + (size_t)s * a.ld);
        double xc;
        if (a.P > 1 && c < a.lo_end)
            xc = __ldcg(xv.lo + c);
        else if (a.P > 1 && c >= a.hi_begin)
            xc = __ldcg(xv.hi + c);
        else
            xc = xb[c];
        acc = fma(a.slot_coef[s] * v, xc, acc);
    }
    double dg = 0.0;
    for (int d = 0; d < a.ndiag; ++d) dg = fma(a.diag_coef[d], __ldg(a.diag + (size_t)d * a.ld + i), dg);
    jd = dg;
    return fma(dg, xv.x[i], acc);
}

// sh.sinkS[r] = c_r * sum_k sink_val[k] x[sink_row[k]], sh.sinkA[r] = same with |x|, in every CTA; ends with a barrier
template <bool MULTI>
__device__ __forceinline__ void sink_rows(const StepArgs& a, const double* x, Shared& sh, SyncState& sy) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (!MULTI) {
        for (int r = wid; r < a.R; r += FW) {
            double s0 = 0.0, s1 = 0.0;
            for (int64_t k = a.sink_ptr[r] + lane; k < a.sink_ptr[r + 1]; k += 32) {
                const double v = __ldg(a.sink_val + k), xv = x[__ldg(a.sink_row + k)];
                s0 = fma(v, xv, s0);
                s1 = fma(v, fabs(xv), s1);
            }
            s0 = warp_sum(s0);
            s1 = warp_sum(s1);
            if (lane == 0) {
                sh.sinkS[r] = a.sink_coef[r] * s0;
                sh.sinkA[r] = a.sink_coef[r] * s1;
            }
        }
        __syncthreads();
    } else {
        for (int r = blockIdx.x; r < a.R; r += a.G) {
            double s0 = 0.0, s1 = 0.0;
            for (int64_t k = a.sink_ptr[r] + threadIdx.x; k < a.sink_ptr[r + 1]; k += FBT) {
                const double v = __ldg(a.sink_val + k), xv = x[__ldg(a.sink_row + k)];
                s0 = fma(v, xv, s0);
                s1 = fma(v, fabs(xv), s1);
            }
            s0 = warp_sum(s0);
            s1 = warp_sum(s1);
            __syncthreads();   // sh.w may still be read by the previous round
            if (lane == 0) {
                sh.w[wid][0] = s0;
                sh.w[wid][1] = s1;
            }
            __syncthreads();
            if (threadIdx.x < 2) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < FW; ++w) t += sh.w[w][threadIdx.x];
                a.sinkbuf[threadIdx.x * MAXR + r] = a.sink_coef[r] * t;
            }
        }
        cg::this_grid().sync();
        if ((int)threadIdx.x < a.R) {
            sh.sinkS[threadIdx.x] = __ldcg(a.sinkbuf + threadIdx.x);
            sh.sinkA[threadIdx.x] = __ldcg(a.sinkbuf + MAXR + threadIdx.x);
        }
        __syncthreads();
        if (a.P > 1) {   // sink rows are sums over ALL states: combine the ranks' parts (every rank keeps the totals)
            if ((int)threadIdx.x < a.R) {
                sh.xg[threadIdx.x] = sh.sinkS[threadIdx.x];
                sh.xg[a.R + threadIdx.x] = sh.sinkA[threadIdx.x];
            }
            __syncthreads();
            xg_allreduce(a, sy, sh.xg, 2 * a.R);
            if ((int)threadIdx.x < a.R) {
                sh.sinkS[threadIdx.x] = sh.xg[threadIdx.x];
                sh.sinkA[threadIdx.x] = sh.xg[a.R + threadIdx.x];
            }
            __syncthreads();
        }
    }
}

// Krylov iteration k: w = (z - c A z) ps, inner products with V_0..V_k (KB = k+1 rounded up to a multiple of 4; the
// surplus products are garbage-free duplicates of V_0 and ignored), <w,w> in slot KB.
template <int KB, bool MULTI>
__device__ __forceinline__ void krylov_apply(const StepArgs& a, const Vecs& v, const StepDyn& dyn, int k, SyncState& sy, Shared& sh) {
    double acc[KB + 1];
#pragma unroll
    for (int s = 0; s <= KB; ++s) acc[s] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * FBT + threadIdx.x; i < a.n; i += (int64_t)a.G * FBT) {
        double jd;
        const double Az = row_apply(a, XView{v.z, a.z_lo, a.z_hi}, i, jd);
        const double w = (v.z[i] - dyn.c * Az) * v.ps[i];
        v.w[i] = w;
#pragma unroll
        for (int j = 0; j < KB; ++j) {
            const double vj = v.V[(size_t)(j <= k ? j : 0) * v.vs + i];
            acc[j] = fma(w, vj, acc[j]);
        }
        acc[KB] = fma(w, w, acc[KB]);
    }
    reduce_all<KB + 1, MULTI>(acc, KB + 1, a, sy, sh);
    // results: sh.res[0..k] = h_j, sh.res[KB] = <w,w>.  Thread 0 runs the Hessenberg/Givens recurrence: a serial chain, so
    // the rotation coefficients are prefetched, and sqrt/one reciprocal replace hypot and the divisions
    if (threadIdx.x == 0) {
        const double ww = sh.res[KB];
        double hsq = 0.0;
        for (int j = 0; j <= k; ++j) {
            const double hj = sh.res[j];
            sh.hcol[j] = hj;
            hsq = fma(hj, hj, hsq);
        }
        double hk1sq = ww - hsq;   // |w - sum h_j v_j|^2 by Pythagoras
        if (!(hk1sq > 0.0)) hk1sq = 0.0;
        const double hk1 = sqrt(hk1sq);
        double hj = sh.res[0];     // running H[j][k] under the previous rotations
        for (int j = 0; j < k; ++j) {
            const double c = sh.cs[j], sn = sh.sn[j], hn = sh.res[j + 1];
            sh.H[j][k] = fma(c, hj, sn * hn);
            hj = fma(-sn, hj, c * hn);
        }
        const double den = sqrt(fma(hj, hj, hk1sq));
        if (!(den > 0.0) || !(ww == ww) || !(den < 1e300)) {
            sh.ok = 0;
            sh.stop = 1;
        } else {
            const double rden = 1.0 / den;
            const double c = hj * rden, sn = hk1 * rden;
            sh.cs[k] = c;
            sh.sn[k] = sn;
            sh.H[k][k] = den;
            const double gk = sh.g[k];
            sh.g[k + 1] = -sn * gk;
            sh.g[k] = c * gk;
            sh.resid = fabs(sn * gk);
            const bool happy = hk1sq <= 1e-28 * fmax(ww, 1e-300);
            sh.stop = (sh.resid / a.sqrtn <= a.lin_tol || happy || k + 1 == GM) ? 1 : 0;
            sh.inv_hk1 = hk1 > 0.0 ? 1.0 / hk1 : 0.0;
        }
    }
    __syncthreads();
}

struct StepOut {   // identical in every thread of the grid
    int lin_ok, accepted, rhs_evals;
    double error_norm, ord_sm, ord_sp;
};

// One BDF step attempt described by `dyn` (see the file header); `publish`: write the sequence number of the result record
template <bool MULTI>
__device__ __forceinline__ StepOut step_body(const StepArgs& a, const Vecs& v, const StepDyn& dyn, Shared& sh, SyncState& sy,
                                             const bool publish) {
    // publish: single-step launch -- the result record goes to pinned host memory and its sequence number is published.
    // In multi-step mode the in-kernel controller consumes StepOut and writes the record only for the step it leaves
    // to the host.
    StepOut out;
    out.lin_ok = out.accepted = out.rhs_evals = 0;
    out.error_norm = 1e300;
    out.ord_sm = out.ord_sp = 0.0;
    const int64_t gtid = (int64_t)blockIdx.x * FBT + threadIdx.x, gstride = (int64_t)a.G * FBT;
    const int order = dyn.order;
    double* const D = a.D;
    const int64_t st = a.stride;
    int rhs_evals = 0, kiters = 0, restarts = 0;

    // ---- P0: pending rescaling of the differences, predictor y_pred = sum D_j, psi = sum gamma_j D_j / alpha_k
    for (int64_t i = gtid; i < a.N; i += gstride) {
        double in[MAXO + 1];
#pragma unroll
        for (int j = 0; j <= MAXO; ++j) in[j] = j <= order ? D[(size_t)j * st + i] : 0.0;
        if (dyn.have_change) {
            double out[MAXO + 1];
#pragma unroll
            for (int r = 0; r <= MAXO; ++r) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= MAXO; ++j) s = fma(dyn.P[j][r], in[j], s);
                out[r] = s;
            }
#pragma unroll
            for (int r = 0; r <= MAXO; ++r)
                if (r <= order) {
                    D[(size_t)r * st + i] = out[r];
                    in[r] = out[r];
                }
        }
        double y = in[0], p = 0.0;
#pragma unroll
        for (int j = 1; j <= MAXO; ++j)
            if (j <= order) {
                y += in[j];
                p = fma(a.gamma[j], in[j], p);
            }
        v.ypred[i] = y;
        v.psi[i] = p * dyn.inv_alpha;
    }
    if (publish && blockIdx.x == 0 && (int)threadIdx.x < a.R) a.res->sink_old0[threadIdx.x] = D[a.n + threadIdx.x];
    grid_barrier<MULTI>();
    if (MULTI && a.P > 1) xg_barrier(a, sy);   // the neighbours' y_pred is complete before P1 gathers its halo

    // ---- P1: A y_pred, Jacobi preconditioner and error weights, right-hand side w0, beta^2 = |w0|^2
    double v1[1] = {0.0};
    for (int64_t i = gtid; i < a.n; i += gstride) {
        double jd;
        const double Ay = row_apply(a, XView{v.ypred, a.ypred_lo, a.ypred_hi}, i, jd);
        const double sc = a.atol + a.rtol * fabs(v.ypred[i]);
        const double p = 1.0 / ((1.0 - dyn.c * jd) * sc);
        const double w = (dyn.c * Ay - v.psi[i]) * p;
        v.scale[i] = sc;
        v.ps[i] = p;
        v.w[i] = w;
        v1[0] = fma(w, w, v1[0]);
    }
    rhs_evals++;
    reduce_all<1, MULTI>(v1, 1, a, sy, sh);
    double beta = sqrt(fmax(sh.res[0], 0.0));
    const double beta0 = beta;
    bool lin_ok = (beta == beta);
    double resid = beta;
    double ms0 = 0.0, ms1 = 0.0;

    if (!(beta / a.sqrtn > a.lin_tol * 1e-3)) {
        // right-hand side already negligible: d = 0
        double v2[2] = {0.0, 0.0};
        for (int64_t i = gtid; i < a.n; i += gstride) {
            v.d[i] = 0.0;
            const double yn = v.ypred[i];
            v.ynew[i] = yn;
            v2[0] += v.psi[i];
            v2[1] += fabs(yn);
        }
        reduce_all<2, MULTI>(v2, 2, a, sy, sh);
        ms0 = sh.res[0];
        ms1 = sh.res[1];
    } else {
        bool have_d = false;
        while (true) {
            // ---- P2: v_0 = w / beta, z = v_0 * scale
            __syncthreads();
            if (threadIdx.x == 0) {
                sh.g[0] = beta;
                sh.ok = 1;
                sh.stop = 0;
                sh.resid = beta;
            }
            {
                const double inv = 1.0 / beta;
                for (int64_t i = gtid; i < a.n; i += gstride) {
                    const double x = v.w[i] * inv;
                    v.V[i] = x;
                    v.z[i] = x * v.scale[i];
                }
            }
            grid_barrier<MULTI>();
            if (MULTI && a.P > 1) xg_barrier(a, sy);   // z complete on the neighbours (they finished READING the old z
                                                       // before the all-reduce that precedes this phase)
            int k = 0;
            for (; k < GM; ++k) {
                // ---- P3: fused matvec + Gram-Schmidt inner products + Givens update
                const int kb = ((k + 1 + 3) / 4) * 4;
                switch (kb) {
                    case 4: krylov_apply<4, MULTI>(a, v, dyn, k, sy, sh); break;
                    case 8: krylov_apply<8, MULTI>(a, v, dyn, k, sy, sh); break;
                    case 12: krylov_apply<12, MULTI>(a, v, dyn, k, sy, sh); break;
                    case 16: krylov_apply<16, MULTI>(a, v, dyn, k, sy, sh); break;
                    case 20: krylov_apply<20, MULTI>(a, v, dyn, k, sy, sh); break;
                    default: krylov_apply<24, MULTI>(a, v, dyn, k, sy, sh); break;
                }
                rhs_evals++;
                kiters++;
                if (sh.stop) {
                    ++k;
                    break;
                }
                // ---- P4: v_{k+1} = (w - sum h_j v_j) / h_{k+1,k}, z = v_{k+1} * scale
                {
                    const double inv = sh.inv_hk1;
                    for (int64_t i = gtid; i < a.n; i += gstride) {
                        double x = v.w[i];
                        for (int j = 0; j <= k; ++j) x = fma(-sh.hcol[j], v.V[(size_t)j * v.vs + i], x);
                        x *= inv;
                        v.V[(size_t)(k + 1) * v.vs + i] = x;
                        v.z[i] = x * v.scale[i];
                    }
                }
                grid_barrier<MULTI>();
                if (MULTI && a.P > 1) xg_barrier(a, sy);
            }
            if (!sh.ok) {
                lin_ok = false;
                break;
            }
            resid = sh.resid;
            // ---- back substitution H y = g (thread 0), then P5: d = scale * sum y_j V_j, ynew = ypred + d
            if (threadIdx.x == 0) {
                for (int i = k - 1; i >= 0; --i) {
                    double acc = sh.g[i];
                    for (int j = i + 1; j < k; ++j) acc -= sh.H[i][j] * sh.y[j];
                    sh.y[i] = acc / sh.H[i][i];
                }
            }
            __syncthreads();
            double v2[2] = {0.0, 0.0};
            for (int64_t i = gtid; i < a.n; i += gstride) {
                double x = 0.0;
                for (int j = 0; j < k; ++j) x = fma(sh.y[j], v.V[(size_t)j * v.vs + i], x);
                x *= v.scale[i];
                if (have_d) x += v.d[i];
                v.d[i] = x;
                const double yn = v.ypred[i] + x;
                v.ynew[i] = yn;
                v2[0] += x + v.psi[i];
                v2[1] += fabs(yn);
            }
            have_d = true;
            reduce_all<2, MULTI>(v2, 2, a, sy, sh);
            ms0 = sh.res[0];
            ms1 = sh.res[1];
            if (resid / a.sqrtn <= a.lin_tol) break;
            if (++restarts > 3) {
                lin_ok = false;
                break;
            }
            // restart: residual of the current d, r = (c A ynew - psi - d) ps
            double v3[1] = {0.0};
            for (int64_t i = gtid; i < a.n; i += gstride) {
                double jd;
                const double Ay = row_apply(a, XView{v.ynew, a.ynew_lo, a.ynew_hi}, i, jd);
                const double w = (dyn.c * Ay - v.psi[i] - v.d[i]) * v.ps[i];
                v.w[i] = w;
                v3[0] = fma(w, w, v3[0]);
            }
            rhs_evals++;
            reduce_all<1, MULTI>(v3, 1, a, sy, sh);
            beta = sqrt(fmax(sh.res[0], 0.0));
            if (!(beta > 0.0)) break;
        }
    }

    StepResult* res = a.res;
    if (!lin_ok) {
        if (publish && blockIdx.x == 0 && threadIdx.x == 0) {
            res->lin_ok = (a.P > 1 && *((volatile unsigned int*)&a.flags[a.me]->error)) ? -1 : 0;
            res->accepted = 0;
            res->kiters = kiters;
            res->rhs_evals = rhs_evals;
            res->restarts = restarts;
            res->beta0 = beta0;
            res->resid = resid;
            res->error_norm = 1e300;
            res->error_sumsq = 1e300;
            __threadfence_system();
            res->seq = a.seq;
        }
        __syncthreads();
        out.lin_ok = 0;
        out.rhs_evals = rhs_evals;
        return out;
    }

    // ---- P6: explicit sink rows (they never feed back) and the total-mass projection of the inexact solve.
    // 1^T A = 0, so an exact step has sum_all(d + psi) = 0; the defect is removed by the relative rescaling
    // d_i -= defect |ynew_i| / sum|ynew| (see bdf.cu).  The sink rows are linear in ynew: S(ynew + delta) follows from
    // S(ynew) and S(|ynew|) without a second pass.
    sink_rows<MULTI>(a, v.ynew, sh, sy);
    double ms2 = 0.0;
    for (int r = 0; r < a.R; ++r) ms2 += dyn.c * sh.sinkS[r];
    const double defect = ms0 + ms2;
    const bool fix = (fabs(defect) <= a.massfix_limit * ms1) && (ms1 > 0.0);
    const double fixfac = fix ? -defect / ms1 : 0.0;
    // ---- P7: apply the projection to the state rows, local error test
    double v4[1] = {0.0};
    for (int64_t i = gtid; i < a.n; i += gstride) {
        double dd = v.d[i], yn = v.ynew[i];
        if (fix) {
            const double delta = fixfac * fabs(yn);
            dd += delta;
            yn += delta;
            v.d[i] = dd;
            v.ynew[i] = yn;
        }
        const double q = dd / (a.atol + a.rtol * fabs(yn));
        v4[0] = fma(q, q, v4[0]);
    }
    if ((int)threadIdx.x < a.R) {   // every CTA keeps the sink increments (needed for the update of rows n..n+R)
        const int r = threadIdx.x;
        const double S = sh.sinkS[r] + fixfac * sh.sinkA[r];
        sh.sink_d[r] = dyn.c * S - v.psi[a.n + r];
    }
    reduce_all<1, MULTI>(v4, 1, a, sy, sh);
    double sumsq = sh.res[0];
    for (int r = 0; r < a.R; ++r) {
        const double ds = sh.sink_d[r], yn = v.ypred[a.n + r] + ds;
        const double q = ds / (a.atol + a.rtol * fabs(yn));
        sumsq += q * q;
    }
    const double error_norm = dyn.err_const * sqrt(sumsq / a.Nglob);
    const bool accept = (error_norm <= 1.0);

    double ord_sm = 0.0, ord_sp = 0.0;
    if (accept) {
        // ---- P8: D_{k+2} = d - D_{k+1}; D_{k+1} = d; D_j += D_{j+1} (j = k..0); order-selection norms of the new D
        double v5[2] = {0.0, 0.0};
        for (int64_t i = gtid; i < a.N; i += gstride) {
            const double dd = i < a.n ? v.d[i] : sh.sink_d[i - a.n];
            const double dp = dd - D[(size_t)(order + 1) * st + i];
            D[(size_t)(order + 2) * st + i] = dp;
            D[(size_t)(order + 1) * st + i] = dd;
            double run = dd, dm = 0.0;
            for (int j = order; j >= 0; --j) {
                run += D[(size_t)j * st + i];
                D[(size_t)j * st + i] = run;
                if (j == order) dm = run;
            }
            if (i < a.n) {
                const double inv = 1.0 / (a.atol + a.rtol * fabs(run));
                const double qm = order > 1 ? dm * inv : 0.0, qp = order < MAXO ? dp * inv : 0.0;
                v5[0] = fma(qm, qm, v5[0]);
                v5[1] = fma(qp, qp, v5[1]);
            }
        }
        reduce_all<2, MULTI>(v5, 2, a, sy, sh);
        ord_sm = sh.res[0];
        ord_sp = sh.res[1];
        // new sink tails for the host's event function (rows n..n+R were updated before the barrier above)
        if (publish && blockIdx.x == 0 && (int)threadIdx.x < a.R)
            for (int j = 0; j <= order + 2; ++j) res->nd[j][threadIdx.x] = __ldcg(D + (size_t)j * st + a.n + threadIdx.x);
    }
    __syncthreads();
    if (publish && blockIdx.x == 0 && threadIdx.x == 0) {
        res->lin_ok = (a.P > 1 && *((volatile unsigned int*)&a.flags[a.me]->error)) ? -1 : 1;   // -1: a peer was lost
        res->accepted = accept ? 1 : 0;
        res->kiters = kiters;
        res->rhs_evals = rhs_evals;
        res->restarts = restarts;
        res->beta0 = beta0;
        res->resid = resid;
        res->error_sumsq = sumsq;
        res->error_norm = (error_norm == error_norm) ? error_norm : 1e300;
        res->ord_sm = ord_sm;
        res->ord_sp = ord_sp;
        res->ms0 = ms0;
        res->ms1 = ms1;
        res->ms2 = ms2;
        __threadfence_system();   // cumulative: also orders the tails written by the other threads of this CTA
        res->seq = a.seq;
    }
    __syncthreads();
    out.lin_ok = 1;
    out.accepted = accept ? 1 : 0;
    out.error_norm = (error_norm == error_norm) ? error_norm : 1e300;
    out.ord_sm = ord_sm;
    out.ord_sp = ord_sp;
    out.rhs_evals = rhs_evals;
    return out;
}

// sink entries of the difference array as the controller sees them: nd(j, r) = D_j[n + r]
struct NdView {
    const double* D;
    int64_t st, n;
    __device__ __forceinline__ double operator()(int j, int r) const { return __ldcg(D + (size_t)j * st + n + r); }
};

// SMEMV (single CTA, N <= ~800): the per-step work vectors and the Krylov basis live in shared memory, only the
// difference array D stays in global memory.
template <bool MULTI, bool SMEMV>
__global__ void __launch_bounds__(FBT) k_bdf_step(const __grid_constant__ StepArgs a) {
    __shared__ Shared sh;
    extern __shared__ __align__(16) double dyn_smem[];
    Vecs v;
    if (SMEMV) {
        const int64_t vs = (a.N + 7) / 8 * 8;
        double* q = dyn_smem;
        v.ypred = q; q += vs;
        v.ynew = q; q += vs;
        v.z = q; q += vs;
        v.psi = q; q += vs;
        v.scale = q; q += vs;
        v.ps = q; q += vs;
        v.w = q; q += vs;
        v.d = q; q += vs;
        v.V = q;
        v.vs = vs;
    } else {
        v.ypred = a.ypred; v.ynew = a.ynew; v.z = a.z; v.psi = a.psi; v.scale = a.scale; v.ps = a.ps; v.w = a.w; v.d = a.d;
        v.V = a.V;
        v.vs = a.stride;
    }
    SyncState sy;
    sy.parity = 0;
    sy.xe = (a.P > 1) ? a.flags[a.me]->fz_epoch : 0u;
    if (MULTI || a.ctl == nullptr) {   // one step attempt per launch, the host keeps the controller
        step_body<MULTI>(a, v, a.dyn, sh, sy, true);
        // every CTA read fz_epoch before its first grid barrier: safe to store the advanced value now
        if (MULTI && a.P > 1 && blockIdx.x == 0 && threadIdx.x == 0) a.flags[a.me]->fz_epoch = sy.xe;
        return;
    }
    // ---- multi-step mode (single CTA): the controller of bdf_ctl.h runs here, thread 0 holds its state in shared memory
    __shared__ BdfCtl ctl;
    __shared__ StepDyn dyn;
    __shared__ int s_save;
    __shared__ double ctl_ws[BDF_SCRATCH];
    __shared__ double s_nd[MAXO + 3][MAXR];   // sink entries of the differences, staged for the controller
    static_assert(sizeof(BdfCtl) % 8 == 0, "BdfCtl is copied as doubles");
    for (int i = threadIdx.x; i < (int)(sizeof(BdfCtl) / 8); i += FBT)
        reinterpret_cast<double*>(&ctl)[i] = reinterpret_cast<const volatile double*>(a.ctl)[i];
    __syncthreads();
    const int R = a.R;
    const NdView ndg{a.D, a.stride, a.n};
    auto nd = [&](int j, int r) { return s_nd[j][r]; };
    for (int attempt = 0;; ++attempt) {
        if (threadIdx.x == 0) {
            int stt = BDF_RUN;
            if (ctl.steps + ctl.rejected >= ctl.max_steps)
                stt = BDF_MAXSTEPS;
            else if (ctl.h_abs < ctl.hmin)
                stt = BDF_UNDERFLOW;
            else if (attempt >= a.max_attempts)
                stt = BDF_YIELD;
            else {
                if (ctl.check_event && R > 0 && !ctl.have_g) {   // g(t) at the start of the segment's first step
                    double s0 = 0.0;
                    for (int r = 0; r < R; ++r) s0 += ndg(0, r);
                    ctl.g_prev = s0 - ctl.event_slope * ctl.t;
                    ctl.have_g = 1;
                }
                bdf_begin_step(ctl, a.kc, dyn, ctl_ws);
            }
            ctl.status = stt;
            s_save = 0;
        }
        __syncthreads();
        if (ctl.status != BDF_RUN) break;
        const StepOut o = step_body<false>(a, v, dyn, sh, sy, false);
        if (threadIdx.x < 32) {   // warp 0: reaction of the controller
            const int lane = threadIdx.x;
            if (!o.lin_ok) {
                if (lane == 0) {
                    ctl.rhs_evals += o.rhs_evals;
                    bdf_after_linfail(ctl, ctl_ws);
                }
            } else if (!o.accepted) {
                if (lane == 0) {
                    ctl.rhs_evals += o.rhs_evals;
                    bdf_after_reject(ctl, o.error_norm, ctl_ws);
                }
            } else {
                const int order = ctl.order;
                const double t = ctl.t, t_new = ctl.t_new, h = ctl.h;
                // one L2 round trip for all sink tails of the new differences (rows n..n+R of D_0..D_{order+2})
                for (int q = lane; q < (order + 3) * R; q += 32) s_nd[q / R][q % R] = ndg(q / R, q % R);
                __syncwarp();
                bool need_host = t_new >= ctl.t1;          // the last step of the segment is finished by the host
                double g_last = ctl.g_prev;
                if (!need_host && ctl.check_event && R > 0) {
                    // sign test of the sink event on 16 samples of the dense output (fspsolve.jl:145-156), one per lane
                    const int q = lane < 16 ? lane + 1 : 16;
                    const double tb = t + (t_new - t) * q / 16;
                    const double gb = bdf_sink_sum_at(nd, R, order, t_new, h, tb) - ctl.event_slope * tb;
                    double ga = __shfl_up_sync(0xffffffffu, gb, 1);
                    if (lane == 0) ga = ctl.g_prev;
                    const bool cross = lane < 16 && ga <= 0.0 && gb > 0.0;
                    if (__ballot_sync(0xffffffffu, cross)) need_host = true;
                    g_last = __shfl_sync(0xffffffffu, gb, 15);
                }
                __syncwarp();
                if (need_host && lane < R)   // the record the host's bookkeeping of this step reads
                    for (int j = 0; j <= order + 2; ++j) a.res->nd[j][lane] = s_nd[j][lane];
                if (lane == 0) {
                    ctl.rhs_evals += o.rhs_evals;
                    ctl.error_norm = o.error_norm;
                    if (need_host) {
                        ctl.status = BDF_HOST_STEP;
                        a.res->lin_ok = 1;
                        a.res->accepted = 1;
                        a.res->error_norm = o.error_norm;
                        a.res->ord_sm = o.ord_sm;
                        a.res->ord_sp = o.ord_sp;
                    } else {
                        ctl.steps++;
                        ctl.n_equal_steps++;
                        if (ctl.check_event && R > 0) ctl.g_prev = g_last;
                        ctl.t = t_new;
                        ctl.h_last = ctl.h_abs;
                        if (ctl.save_every_step) {
                            ctl.ring_t[ctl.ring_count] = t_new;
                            s_save = 1;
                        }
                        if (ctl.n_equal_steps >= order + 1) {
                            double sm = o.ord_sm, sp = o.ord_sp;
                            for (int r = 0; r < R; ++r) {
                                const double inv = 1.0 / (a.atol + a.rtol * fabs(nd(0, r)));
                                if (order > 1) sm += (nd(order, r) * inv) * (nd(order, r) * inv);
                                if (order < MAXO) sp += (nd(order + 2, r) * inv) * (nd(order + 2, r) * inv);
                            }
                            bdf_select_order(ctl, a.kc, o.error_norm, sm, sp, ctl_ws);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (s_save) {   // every-step output: the new solution D_0 goes into the ring, the host drains it after the launch
            double* dst = a.ring + (size_t)ctl.ring_count * a.stride;
            for (int64_t i = threadIdx.x; i < a.N; i += FBT) dst[i] = a.D[i];
            __syncthreads();
            if (threadIdx.x == 0 && ++ctl.ring_count == BDF_RING && ctl.status == BDF_RUN) ctl.status = BDF_YIELD;
            __syncthreads();
        }
        // racecheck (round 2): thread 0 rewrites ctl.status at the top of the next attempt -- every thread must have
        // read it before that, or a slow warp could leave the loop on the NEXT attempt's status
        const bool keep_going = (ctl.status == BDF_RUN);
        __syncthreads();
        if (!keep_going) break;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)(sizeof(BdfCtl) / 8); i += FBT)
        reinterpret_cast<volatile double*>(a.ctl)[i] = reinterpret_cast<const double*>(&ctl)[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        a.res->seq = a.seq;
    }
}

// output slices with the device->host copies queued behind the step kernels and the host callbacks deferred to the
// next synchronisation that happens anyway (no extra stream sync per saved step)
struct LazySaver {
    ncme_ctx* ctx = nullptr;
    ncme_save_fn fn = nullptr;
    void* user = nullptr;
    ncme_solve_stats* st = nullptr;
    double* pinned = nullptr;
    size_t len = 0;
    int nslots = 0;
    std::vector<std::pair<double, int>> pending;
    // sharded: a slice handed to the callback is the GLOBAL vector [all state rows | R sinks] on every rank (as
    // SliceSaver of solve.cu delivers it): all-gather of the ranks' rows; the sink entries are already replicated
    ncme_comm* comm = nullptr;
    int64_t n_local = 0, n_global = 0;
    int R = 0;
    double* full = nullptr;
    std::vector<int64_t> counts, displs;
    int init_sharded(ncme_comm* cm, int64_t nloc, int64_t nglob, int nr) {
        comm = cm;
        n_local = nloc;
        n_global = nglob;
        R = nr;
        if (!fn) return NCME_OK;
        NCME_TRY(cache_reserve(&ctx->solve_full, &ctx->solve_full_bytes, (size_t)(nglob + nr) * sizeof(double), false));
        full = ctx->solve_full;
        counts.assign((size_t)cm->nranks, 0);
        displs.assign((size_t)cm->nranks, 0);
        for (int r = 0; r < cm->nranks; ++r) {   // the row cuts of matrix_build (multiples of 64)
            auto cut = [&](int q) -> int64_t {
                if (q <= 0) return 0;
                if (q >= cm->nranks) return nglob;
                return std::min<int64_t>(nglob, round_up<int64_t>((int64_t)((__int128)nglob * q / cm->nranks), 64));
            };
            displs[(size_t)r] = cut(r);
            counts[(size_t)r] = cut(r + 1) - cut(r);
        }
        NCME_REQUIRE(counts[(size_t)cm->rank] == nloc, "fused BDF step (sharded): unexpected row partition");
        return NCME_OK;
    }
    int init(ncme_ctx* c, size_t n, ncme_save_fn f, void* u, ncme_solve_stats* stats) {
        ctx = c;
        fn = f;
        user = u;
        st = stats;
        len = n;
        if (!fn) return NCME_OK;
        nslots = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)16 << 20) / (len * sizeof(double))));
        NCME_TRY(cache_reserve(&ctx->solve_pinned, &ctx->solve_pinned_bytes, (size_t)nslots * len * sizeof(double), true));
        pinned = ctx->solve_pinned;
        return NCME_OK;
    }
    void delivered() {   // call right after a stream synchronisation
        for (auto& p : pending) {
            fn(p.first, pinned + (size_t)p.second * len, user);
            st->nsaved++;
        }
        pending.clear();
    }
    int flush() {
        if (pending.empty()) return NCME_OK;
        NCME_CUDA(cudaStreamSynchronize(ctx->stream));
        delivered();
        return NCME_OK;
    }
    int save(double t, const double* v_dev) {
        if (!fn) return NCME_OK;
        if ((int)pending.size() == nslots) NCME_TRY(flush());
        const int slot = (int)pending.size();
        const double* src = v_dev;
        if (comm) {
            NCME_TRY(comm_allgatherv(comm, v_dev, full, counts.data(), displs.data(), ctx->stream));
            NCME_CUDA(cudaMemcpyAsync(full + n_global, v_dev + n_local, (size_t)R * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            src = full;
        }
        NCME_CUDA(cudaMemcpyAsync(pinned + (size_t)slot * len, src, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        pending.emplace_back(t, slot);
        if (comm) NCME_TRY(flush());   // `full` is reused by the next slice
        return NCME_OK;
    }
};

}  // namespace

// Sharded matrices: the ranks' cooperative grids exchange halos, barriers and all-reduces through peer memory
// (CUDA IPC); needs the peer-memory transport, a halo that lives on the two neighbouring ranks, and room in the
// exchange slots.  The answer must be the same on every rank: it only depends on replicated facts.
static bool fused_sharded_ok(const ncme_matrix* A) {
    const ncme_comm* c = A->comm;
    static const bool off = getenv("NCME_BDF_NO_SHARDED_FUSED") != nullptr;
    return !off && c && c->nranks > 1 && c->nranks <= NCME_RED_RANKS && c->p2p_ok && c->my_flags && A->p2p_eligible &&
           2 * A->nr <= NCME_FZ_VALS && comm_hostreduce_available(c);
}
bool bdf_fused_eligible(const ncme_matrix* A) {
    if (A->n_global < 1 || A->nr > MAXR) return false;
    if (A->comm == nullptr) return A->hl == 0 && A->hh == 0 && A->n >= 1;
    return fused_sharded_ok(A);
}

int solve_bdf_fused(ncme_matrix* A, ncme_coef_fn coef_fn, ncme_save_fn save_fn, void* user, double t0, double t1,
                    double* u, const ncme_solve_opts* o, ncme_solve_stats* st) {
    ncme_ctx* ctx = A->ctx;
    cudaStream_t s = ctx->stream;
    NCME_REQUIRE(bdf_fused_eligible(A), "fused BDF step kernel: needs a single-GPU matrix or the peer-memory transport");
    ncme_comm* comm = A->comm;            // non-null: row-sharded, one cooperative grid per rank
    const bool sharded = comm != nullptr;
    const int64_t n = A->n, N = A->N;     // local rows, local vector length (rows + R sink entries, sinks replicated)
    const int R = A->nr;
    const int64_t Nglob = A->n_global + R;
    const int64_t launches0 = ctx->launches;

    // ---- grid: one CTA up to 2 rows per thread, a cooperative grid above (bounded by co-residency)
    static int occ_cached = 0;
    if (!occ_cached) {
        int occ = 0;
        NCME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bdf_step<true, false>, FBT, 0));
        occ_cached = std::max(1, occ);
        NCME_CUDA(cudaFuncSetAttribute(k_bdf_step<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX_BYTES));
    }
    const int maxG = std::min(occ_cached, 2) * ctx->sm_count;
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(maxG, (n + (int64_t)FBT * 2 - 1) / ((int64_t)FBT * 2)));
    // tiny systems: work vectors + Krylov basis in shared memory (8 + GM + 1 vectors of round_up(N, 8) doubles)
    const size_t smem_need = (size_t)(8 + GM + 1) * (size_t)((N + 7) / 8 * 8) * sizeof(double);
    const size_t smem_bytes = (!sharded && G == 1 && smem_need <= (size_t)SMEM_MAX_BYTES && !getenv("NCME_BDF_NO_SMEM")) ? smem_need : 0;
    // multi-step launches: time-invariant matrix (no host callback per step), one CTA, no saveat list
    bool need_coef = false;
    for (int r = 0; r < R; ++r) need_coef |= (A->kind[r] != NCME_TIME_INVARIANT);
    const bool multi = !sharded && G == 1 && !need_coef && o->nsave == 0 && !getenv("NCME_BDF_SINGLE_STEP");

    // ---- workspace
    // sharded: every vector carries the halo margins a matvec input needs and the workspace is exposed to the
    // neighbours (CUDA IPC, collective registration), exactly as in bdf.cu
    const size_t hlp = sharded ? round_up<size_t>((size_t)A->hl, 32) : 0, hhp = sharded ? round_up<size_t>((size_t)A->hh, 32) : 0;
    const size_t stride = hlp + round_up<size_t>((size_t)N, 32) + hhp;
    const int NV = (MAXO + 3) + 8 + (GM + 1) + (multi ? BDF_RING : 0);
    const int Gmax_all = 2 * ctx->sm_count;   // the scratch size must not depend on this rank's row count
    const size_t extra = (size_t)2 * (sharded ? Gmax_all : G) * NSLOT + 2 * MAXR + 64;
    double* base = nullptr;
    if (sharded) {
        const int peers[2] = {A->plo, A->phi};
        NCME_TRY(comm_workspace(comm, (stride * NV + extra) * sizeof(double), (int64_t)hlp, (int64_t)stride, NV, peers, 2, &base));
    } else {
        NCME_TRY(cache_reserve(&ctx->solve_ws, &ctx->solve_ws_bytes, (stride * NV + extra) * sizeof(double), false));
        base = ctx->solve_ws;
    }
    NCME_CUDA(cudaMemsetAsync(base, 0, (stride * ((MAXO + 3) + 8 + (GM + 1)) + extra) * sizeof(double), s));
    int slot = 0;
    auto vec = [&]() { return base + stride * (slot++) + hlp; };
    double* D[MAXO + 3];
    for (int j = 0; j < MAXO + 3; ++j) D[j] = vec();
    double* ypred = vec();
    double* ynew = vec();
    double* z = vec();
    double* psi = vec();
    double* scale = vec();
    double* ps = vec();
    double* w = vec();
    double* d = vec();
    double* V = base + stride * slot + hlp;
    slot += GM + 1;
    double* ring = base + stride * slot;       // multi-step mode only (never sharded: hlp == 0)
    double* partials = base + stride * NV;
    double* sinkbuf = partials + (size_t)2 * G * NSLOT;
    // pinned scalars of the context: [StepResult | BdfCtl]
    constexpr size_t CTL_OFF = (sizeof(StepResult) + 63) / 64 * 64;
    static_assert(CTL_OFF + sizeof(BdfCtl) <= 1024 * sizeof(double), "result record + controller must fit the pinned scalars");
    StepResult* res_host = reinterpret_cast<StepResult*>(ctx->red_result_host);
    BdfCtl& c = *reinterpret_cast<BdfCtl*>(reinterpret_cast<char*>(ctx->red_result_host) + CTL_OFF);
    StepResult* res_dev = nullptr;
    NCME_CUDA(cudaHostGetDevicePointer((void**)&res_dev, (void*)res_host, 0));
    BdfCtl* ctl_dev = reinterpret_cast<BdfCtl*>(reinterpret_cast<char*>(res_dev) + CTL_OFF);
    unsigned int seq = 0;
    res_host->seq = 0;

    LazySaver saver;
    NCME_TRY(saver.init(ctx, (size_t)(sharded ? Nglob : N), save_fn, user, st));
    if (sharded) NCME_TRY(saver.init_sharded(comm, n, A->n_global, R));
    double* ring_pinned = nullptr;
    if (multi && save_fn) {   // the lazy saver's pinned block is at least 4 MB: the ring's host image lives behind its slots
        const size_t need = ((size_t)saver.nslots + BDF_RING) * (size_t)N * sizeof(double);
        NCME_TRY(cache_reserve(&ctx->solve_pinned, &ctx->solve_pinned_bytes, need, true));
        saver.pinned = ctx->solve_pinned;
        ring_pinned = ctx->solve_pinned + (size_t)saver.nslots * (size_t)N;
    }
    const double rtol = o->rtol > 0 ? o->rtol : 1e-4, atol = o->atol > 0 ? o->atol : 1e-6;
    const double tspan = t1 - t0;

    double ctl_ws[BDF_SCRATCH];
    BdfConst kc;
    bdf_constants(kc);
    kc.atol = atol;
    kc.rtol = rtol;
    kc.Nglob = (double)Nglob;

    double coef[NCME_MAX_REACTIONS];
    for (int r = 0; r < NCME_MAX_REACTIONS; ++r) coef[r] = 1.0;

    auto finish = [&](const double* src) -> int {
        if (src != u) NCME_CUDA(cudaMemcpyAsync(u, src, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s));
        NCME_CUDA(cudaStreamSynchronize(s));
        if (sharded) {   // sink entries are replicated (all-reduced inside the kernel): nothing to combine here
            unsigned int perr = 0;
            NCME_CUDA(cudaMemcpy(&perr, &comm->my_flags->error, sizeof(perr), cudaMemcpyDeviceToHost));
            if (perr) {
                set_error("fused BDF step (sharded): a peer rank did not arrive at an exchange within 2 s");
                return NCME_ERR_COMM;
            }
        }
        saver.delivered();
        st->steps = c.steps;
        st->rejected = c.rejected;
        st->rhs_evals = c.rhs_evals;
        st->h_last = c.h_last;
        st->launches = ctx->launches - launches0;
        return NCME_OK;
    };

    memset(&c, 0, sizeof(c));
    c.t = t0;
    c.t1 = t1;
    c.tspan = tspan;
    c.hmin = std::max(1e-14 * fabs(t0), 1e-20 * fabs(t1 - t0));   // see bdf.cu
    c.order = 1;
    c.max_steps = o->max_steps > 0 ? o->max_steps : 100000000;
    c.check_event = (o->check_event && R > 0) ? 1 : 0;
    c.save_every_step = (o->save_every_step && save_fn) ? 1 : 0;
    c.event_slope = o->event_slope;

    NCME_CUDA(cudaMemcpyAsync(D[0], u, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s));
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

    // ---- initial step size (Hairer's rule on the WRMS norms of u and f(t0, u)), D_1 = h f(t0, u)
    if (coef_fn) coef_fn(t0, coef, user);
    if (abort_requested()) return abort_status();
    if (sharded) {
        // The ranks' exchange counters must agree before the first cooperative launch (they do unless an earlier
        // segment failed half way): everybody continues from the largest one.
        unsigned int mine = 0;
        NCME_CUDA(cudaMemsetAsync(&comm->my_flags->error, 0, sizeof(unsigned int), s));   // (only this rank ever sets it)
        NCME_CUDA(cudaMemcpyAsync(&mine, &comm->my_flags->fz_epoch, sizeof(mine), cudaMemcpyDeviceToHost, s));
        NCME_CUDA(cudaStreamSynchronize(s));
        double all[NCME_RED_RANKS] = {0.0};
        all[comm->rank] = (double)mine;
        NCME_TRY(comm_hostreduce_sum(comm, all, (size_t)comm->nranks));
        double mx = 0.0, mn = 1e300;
        for (int q = 0; q < comm->nranks; ++q) {
            mx = std::max(mx, all[q]);
            mn = std::min(mn, all[q]);
        }
        if (mx != mn) {
            const unsigned int next = (unsigned int)mx + 1024u;
            NCME_CUDA(cudaMemcpyAsync(&comm->my_flags->fz_epoch, &next, sizeof(next), cudaMemcpyHostToDevice, s));
            NCME_CUDA(cudaStreamSynchronize(s));
        }
    }
    NCME_TRY(matvec_dist(A, coef, D[0], ynew, 0.0, sharded ? 1 : 0));   // sharded: all-reduced (replicated) sink rows
    c.rhs_evals++;
    c.h_abs = o->h_init;
    if (!(c.h_abs > 0)) {
        double d0 = 0, d1 = 0;
        const int64_t nn = sharded ? n : N;   // sharded: state rows only (the sinks are replicated), summed over the ranks
        NCME_TRY(ncme_vec_wrms(ctx, nn > 0 ? nn : 1, D[0], D[0], D[0], atol, rtol, &d0));
        NCME_TRY(ncme_vec_wrms(ctx, nn > 0 ? nn : 1, ynew, D[0], D[0], atol, rtol, &d1));
        saver.delivered();
        if (sharded) {
            double ss[2] = {d0 * d0 * (double)nn, d1 * d1 * (double)nn};
            NCME_TRY(comm_hostreduce_sum(comm, ss, 2));
            d0 = sqrt(ss[0] / (double)Nglob);
            d1 = sqrt(ss[1] / (double)Nglob);
        }
        c.h_abs = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        c.h_abs = std::min(c.h_abs, tspan);
    }
    {
        const double cs[1] = {c.h_abs};
        const double* xs[1] = {ynew};
        NCME_TRY(ncme_vec_lincomb(ctx, N, 1, cs, xs, D[1]));
    }
    const double lin_tol = 5e-3;   // WRMS residual of the linear solve = CVODE's 0.05 x Newton tolerance 0.1

    StepArgs sa{};
    sa.col = A->col.p;
    sa.val = A->val.p;
    sa.diag = A->diag.p;
    sa.n = n;
    sa.ld = A->ld;
    sa.N = N;
    sa.nslots = A->nslots;
    sa.ndiag = A->ndiag;
    sa.R = R;
    sa.G = G;
    sa.sink_row = A->sink_row.p;
    sa.sink_val = A->sink_val.p;
    for (int r = 0; r <= R; ++r) sa.sink_ptr[r] = A->sink_ptr[r];
    sa.D = D[0];
    sa.V = V;
    sa.ypred = ypred;
    sa.ynew = ynew;
    sa.z = z;
    sa.psi = psi;
    sa.scale = scale;
    sa.ps = ps;
    sa.w = w;
    sa.d = d;
    sa.stride = (int64_t)stride;
    for (int k = 0; k <= MAXO; ++k) sa.gamma[k] = kc.gamma[k];
    sa.atol = atol;
    sa.rtol = rtol;
    sa.lin_tol = lin_tol;
    sa.sqrtn = sqrt((double)std::max<int64_t>(1, A->n_global));
    sa.massfix_limit = 10.0 * rtol;
    sa.Nglob = (double)Nglob;
    sa.P = 1;
    if (sharded) {
        sa.P = comm->nranks;
        sa.me = comm->rank;
        sa.self_off = (uint32_t)A->hl;
        sa.lo_end = (uint32_t)A->hl;
        sa.hi_begin = (uint32_t)(A->hl + A->n + A->nr);
        for (int q = 0; q < comm->nranks; ++q) sa.flags[q] = (q == comm->rank) ? comm->my_flags : comm->peer_flags[q];
        // peer views of the matvec inputs: view[c] = entry at padded position c (see matvec_dist_p2p)
        auto views = [&](const double* x, const double** lo, const double** hi) -> int {
            const double* xlo = A->plo >= 0 ? comm_peer_vector(comm, x, A->plo) : nullptr;
            const double* xhi = A->phi >= 0 ? comm_peer_vector(comm, x, A->phi) : nullptr;
            NCME_REQUIRE((A->plo < 0 || xlo) && (A->phi < 0 || xhi), "fused BDF step (sharded): workspace is not registered on a neighbour");
            *lo = xlo ? xlo + (A->ext_lo - A->plo_row_lo) : x - A->hl;
            *hi = xhi ? xhi + (A->row_hi - A->phi_row_lo) - (int64_t)sa.hi_begin : x - A->hl;
            return NCME_OK;
        };
        NCME_TRY(views(ypred, &sa.ypred_lo, &sa.ypred_hi));
        NCME_TRY(views(z, &sa.z_lo, &sa.z_hi));
        NCME_TRY(views(ynew, &sa.ynew_lo, &sa.ynew_hi));
    }
    sa.partials = partials;
    sa.sinkbuf = sinkbuf;
    sa.res = res_dev;
    sa.ctl = multi ? ctl_dev : nullptr;
    sa.kc = kc;
    sa.ring = ring;
    sa.max_attempts = BDF_RING;
    {   // coefficients of a time-invariant matrix never change
        MatvecArgs ma;
        matvec_fill_args(A, coef, &ma);
        for (int q = 0; q < A->nslots; ++q) sa.slot_coef[q] = ma.slot_coef[q];
        for (int q = 0; q < A->ndiag; ++q) sa.diag_coef[q] = ma.diag_coef[q];
        for (int q = 0; q < R; ++q) sa.sink_coef[q] = ma.sink_coef[q];
    }

    auto launch_and_wait = [&]() -> int {
        sa.seq = ++seq;
        if (smem_bytes) {
            k_bdf_step<false, true><<<1, FBT, smem_bytes, s>>>(sa);
        } else if (G == 1 && !sharded) {
            k_bdf_step<false, false><<<1, FBT, 0, s>>>(sa);
        } else {
            void* kargs[1] = {(void*)&sa};
            NCME_CUDA(cudaLaunchCooperativeKernel((const void*)k_bdf_step<true, false>, dim3(G), dim3(FBT), kargs, 0, s));
        }
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        // the kernel writes its result record (and, in multi-step mode, the controller state) into pinned host memory
        // and publishes the sequence number last: poll it (cheaper than a copy + stream synchronisation); the stream is
        // queried now and then to catch faults
        for (unsigned spin = 1; res_host->seq != seq; ++spin) {
            if ((spin & 0xFFFF) == 0) {
                const cudaError_t q = cudaStreamQuery(s);
                if (q == cudaSuccess) {
                    if (res_host->seq == seq) break;
                    set_error("fused BDF step kernel finished without publishing its result");
                    return NCME_ERR_SOLVER;
                }
                if (q != cudaErrorNotReady) NCME_CUDA(q);
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        saver.delivered();   // copies queued before the kernel have completed (stream order)
        return NCME_OK;
    };

    while (c.t < t1) {
        if (abort_requested()) return abort_status();   // a save callback failed
        if (!multi) {
            if (c.steps + c.rejected >= c.max_steps) {
                set_error("integrator: maximum number of steps (%lld) reached at t = %g", (long long)c.max_steps, c.t);
                return NCME_ERR_SOLVER;
            }
            if (c.h_abs < c.hmin) {
                set_error("integrator (BDF): step size underflow at t = %g", c.t);
                return NCME_ERR_SOLVER;
            }
            // ---- one launch: the whole step attempt at t_new
            bdf_begin_step(c, kc, sa.dyn, ctl_ws);
            if (coef_fn) {
                coef_fn(c.t_new, coef, user);
                if (abort_requested()) return abort_status();
                MatvecArgs ma;
                matvec_fill_args(A, coef, &ma);
                for (int q = 0; q < A->nslots; ++q) sa.slot_coef[q] = ma.slot_coef[q];
                for (int q = 0; q < A->ndiag; ++q) sa.diag_coef[q] = ma.diag_coef[q];
                for (int q = 0; q < R; ++q) sa.sink_coef[q] = ma.sink_coef[q];
            }
            NCME_TRY(launch_and_wait());
            c.rhs_evals += res_host->rhs_evals;
            if (res_host->lin_ok < 0) {
                set_error("fused BDF step (sharded): a peer rank did not arrive at an exchange within 2 s");
                return NCME_ERR_COMM;
            }
            if (!res_host->lin_ok) {
                bdf_after_linfail(c, ctl_ws);
                continue;
            }
            c.error_norm = res_host->error_norm;
            if (!res_host->accepted) {
                bdf_after_reject(c, c.error_norm, ctl_ws);
                continue;
            }
        } else {
            // ---- one launch: up to BDF_RING step attempts with the controller on the device
            c.status = BDF_RUN;
            c.ring_count = 0;
            NCME_TRY(launch_and_wait());
            if (c.ring_count > 0 && save_fn) {   // every-step output slices produced inside the launch
                NCME_CUDA(cudaMemcpy2DAsync(ring_pinned, (size_t)N * sizeof(double), ring, stride * sizeof(double),
                                            (size_t)N * sizeof(double), (size_t)c.ring_count, cudaMemcpyDeviceToHost, s));
                NCME_CUDA(cudaStreamSynchronize(s));
                for (int k = 0; k < c.ring_count; ++k) {
                    save_fn(c.ring_t[k], ring_pinned + (size_t)k * (size_t)N, user);
                    st->nsaved++;
                }
            }
            if (c.status == BDF_MAXSTEPS) {
                set_error("integrator: maximum number of steps (%lld) reached at t = %g", (long long)c.max_steps, c.t);
                return NCME_ERR_SOLVER;
            }
            if (c.status == BDF_UNDERFLOW) {
                set_error("integrator (BDF): step size underflow at t = %g", c.t);
                return NCME_ERR_SOLVER;
            }
            if (c.status != BDF_HOST_STEP) continue;   // BDF_YIELD: ring full / attempts used up
        }
        // ---- accepted step (the kernel has already updated the differences): bookkeeping, event, output
        const StepResult& rs = *res_host;
        const int order = c.order;
        const double t = c.t, t_new = c.t_new, h = c.h;
        c.steps++;
        c.n_equal_steps++;
        const double(*nd)[MAXR] = rs.nd;
        auto ndv = [&](int j, int r) { return nd[j][r]; };
        auto sink_sum_at = [&](double tt) { return bdf_sink_sum_at(ndv, R, order, t_new, h, tt); };
        auto dense_to = [&](double tt, double* out) -> int {
            double cs[MAXO + 1];
            const double* xs[MAXO + 1];
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
        if (c.check_event) {
            if (!c.have_g) {
                double s0 = 0.0;
                for (int r = 0; r < R; ++r) s0 += rs.sink_old0[r];
                c.g_prev = s0 - c.event_slope * t;
                c.have_g = 1;
            }
            const int NS_ = 16;
            double ga = c.g_prev, ta = t;
            for (int q = 1; q <= NS_; ++q) {
                const double tb = t + (t_new - t) * q / NS_;
                const double gb = sink_sum_at(tb) - c.event_slope * tb;
                if (ga <= 0.0 && gb > 0.0) {
                    double lo = ta, hi = tb;
                    for (int it = 0; it < 60; ++it) {
                        const double mid = 0.5 * (lo + hi);
                        if (sink_sum_at(mid) - c.event_slope * mid > 0.0)
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
            if (!event) c.g_prev = ga;
        }
        while (isave < o->nsave && o->save_t[isave] <= t_hi + 1e-14 * fabs(t_hi)) {
            const double ts = o->save_t[isave];
            if (ts >= t_new && !event) {
                NCME_TRY(saver.save(ts, D[0]));
            } else {
                NCME_TRY(saver.flush());   // z is about to be overwritten: deliver what still points at it
                NCME_TRY(dense_to(std::min(ts, t_new), z));
                NCME_TRY(saver.save(ts, z));
            }
            ++isave;
        }
        if (event) {
            NCME_TRY(saver.flush());
            NCME_TRY(dense_to(t_hi, z));
            st->t_final = t_hi;
            st->event_hit = 1;
            c.h_last = c.h_abs;
            return finish(z);
        }
        c.t = t_new;
        c.h_last = c.h_abs;
        if (o->save_every_step) NCME_TRY(saver.save(c.t, D[0]));
        if (c.t >= t1) break;
        if (c.n_equal_steps < order + 1) continue;
        // ---- order / step-size selection (norms of the new differences came back with the step result)
        double sm = rs.ord_sm, sp = rs.ord_sp;
        for (int r = 0; r < R; ++r) {
            const double inv = 1.0 / (atol + rtol * fabs(nd[0][r]));
            if (order > 1) sm += (nd[order][r] * inv) * (nd[order][r] * inv);
            if (order < MAXO) sp += (nd[order + 2][r] * inv) * (nd[order + 2][r] * inv);
        }
        bdf_select_order(c, kc, c.error_norm, sm, sp, ctl_ws);
    }
    st->t_final = c.t;
    return finish(D[0]);
}

}  // namespace ncme
