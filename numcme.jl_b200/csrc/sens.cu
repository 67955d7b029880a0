// K2: ForwardSensFspMatrixSparse -- fused block matvec over [p; s_1; ...; s_P].
// Reference: src/forwardsensfspmatrix/forwardsensfspmatrixsparse/sensfspmatrixsparse.jl
//   constructor :31-95, matvec! :97-142.
//
// The reference runs (P+1) full matvec! passes (re-reading A every time) plus one CSC pass per
// (reaction, parameter) entry.  Here one kernel reads A's col/val/diag ONCE per row into registers,
// applies it to all P+1 blocks, and streams each dA entry exactly once:
//   Y_0  = A(t) p
//   Y_ip = A(t) s_ip + sum_{e=(r,ip)} [ c_r(t) dA_e p + (d c_r/d theta_ip)(t) A_r p ]
// dA_e has the sparsity of reaction r's slot, so it shares col[slot(r)] with A.
#include "matrix.cuh"

// (reaction, parameter) entries of the sparsity pattern a sensitivity matrix may hold (Hog1p: 14); bounded by the
// 4 KB kernel parameter block of K2, which carries c_r(t) and dc_r/dtheta(t) of every entry by value
#define NCME_SENS_MAX_ENTRIES 96

struct ncme_sensmatrix {
    ncme_matrix* A = nullptr;
    int64_t A_n = 0;   // states of A at build time (stride of dG)
    int npar = 0;
    int nent = 0;
    int ent_reaction[NCME_SENS_MAX_ENTRIES];   // internal order: sorted by parameter
    int ent_param[NCME_SENS_MAX_ENTRIES];
    int ent_user[NCME_SENS_MAX_ENTRIES];       // internal -> user entry index
    int user_ent[NCME_SENS_MAX_ENTRIES];       // user -> internal
    int ent_ptr[NCME_SENS_MAX_ENTRIES + 2];    // entries of parameter ip: [ent_ptr[ip], ent_ptr[ip+1])
    ncme::DevArray<double> dval;    // [nent][ld]  d(state factor)/d theta at the predecessor
    ncme::DevArray<double> ddiag;   // [nent][ld]  minus d(state factor)/d theta at the state itself
    ncme::DevArray<double> dsink;   // per entry: values along the sink list of its reaction
    // d(state factor)/d theta of ALL states per entry (internal entry order, stride A->n), kept for the incremental
    // constructor of the sensitivity matrix that follows an adapt! (only the appended states are evaluated on the host)
    ncme::DevArray<double> dG;
    int64_t dsink_ptr[NCME_SENS_MAX_ENTRIES + 1];
    ncme::DevArray<double> partial;   // [ntasks][P+1]
    ncme::DevArray<double> dpartial;  // [ntasks][nent]
    unsigned int* counter = nullptr;
    int* meta_dev = nullptr;  // ent_reaction | ent_param | ent_slot | ent_diag | ent_ptr  (device copy)
    int tune_rows = 0;        // K2 rows per thread (0 = auto)
};

namespace ncme {

constexpr int SMAX_ENT = NCME_SENS_MAX_ENTRIES;
constexpr int SV_THREADS = 256;

struct SensArgs {
    MatvecArgs m;          // x = X block 0, y = Y block 0
    int npar, nent;
    int64_t N;             // block length
    const double* dval;
    const double* ddiag;
    const double* dsink;
    const int* ent_reaction;  // device meta (static per matrix)
    const int* ent_param;
    const int* ent_slot;
    const int* ent_diag;
    const int* ent_ptr;
    const int64_t* dsink_ptr;  // device
    double* partial;
    double* dpartial;
    unsigned int* counter;
    // time-dependent scalars travel BY VALUE with the launch (no copy, no synchronisation per right-hand side)
    double ent_c[SMAX_ENT];    // c_r(t) per entry (1 for time-invariant / joint reactions)
    double ent_dc[SMAX_ENT];   // d c_r / d theta (t) per entry (separable reactions only)
};
static_assert(sizeof(SensArgs) <= 4096, "kernel parameter block must stay below 4 KB");

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// block-wide sum, result valid in thread 0
__device__ __forceinline__ double bsum(double v, double* sh) {
    v = wsum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < SV_THREADS / 32; ++w) t += sh[w];
    return t;
}

__device__ void sens_sink_task(const SensArgs& a) {
    __shared__ double sh[SV_THREADS / 32];
    __shared__ bool is_last;
    const MatvecArgs& m = a.m;
    const int4 t = m.tasks[blockIdx.x];
    const int r = t.x;
    for (int b = 0; b <= a.npar; ++b) {
        const double* xb = m.x + (int64_t)b * a.N;
        double s = 0.0;
        for (int k = t.y + (int)threadIdx.x; k < t.z; k += SV_THREADS) s += m.sink_val[k] * xb[m.sink_row[k]];
        s = bsum(s, sh);
        if (threadIdx.x == 0) a.partial[(int64_t)blockIdx.x * (a.npar + 1) + b] = s;
    }
    for (int e = 0; e < a.nent; ++e) {
        if (a.ent_reaction[e] != r) continue;  // uniform
        const double* ds = a.dsink + a.dsink_ptr[e];
        const int rbeg = m.tasks[m.task_ptr[r]].y;  // first sink entry of reaction r
        double s = 0.0;
        for (int k = t.y + (int)threadIdx.x; k < t.z; k += SV_THREADS) s += ds[k - rbeg] * m.x[m.sink_row[k]];
        s = bsum(s, sh);
        if (threadIdx.x == 0) a.dpartial[(int64_t)blockIdx.x * a.nent + e] = s;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(a.counter, 1u) == (unsigned)m.ntasks - 1u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const volatile double* part = a.partial;
    const volatile double* dpart = a.dpartial;
    for (int q = threadIdx.x; q < m.nr * (a.npar + 1); q += SV_THREADS) {
        const int rr = q / (a.npar + 1), b = q % (a.npar + 1);
        double tot = 0.0;
        for (int k = m.task_ptr[rr]; k < m.task_ptr[rr + 1]; ++k) tot += part[(int64_t)k * (a.npar + 1) + b];
        m.y[(int64_t)b * a.N + m.n + rr] = m.sink_coef[rr] * tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int e = 0; e < a.nent; ++e) {
            const int rr = a.ent_reaction[e], ip = a.ent_param[e];
            double dt = 0.0, s0 = 0.0;
            for (int k = m.task_ptr[rr]; k < m.task_ptr[rr + 1]; ++k) {
                dt += dpart[(int64_t)k * a.nent + e];
                s0 += part[(int64_t)k * (a.npar + 1)];
            }
            m.y[(int64_t)(ip + 1) * a.N + m.n + rr] += a.ent_c[e] * dt + a.ent_dc[e] * s0;
        }
        *a.counter = 0u;
    }
}

// streaming (evict-first) vector loads of ROWS consecutive entries: the once-read matrix must not evict the block
// vectors, whose gathers live in L1/L2
template <int ROWS>
__device__ __forceinline__ void lds(const double* __restrict__ p, double (&v)[ROWS]);
template <>
__device__ __forceinline__ void lds<1>(const double* __restrict__ p, double (&v)[1]) {
    v[0] = __ldcs(p);
}
template <>
__device__ __forceinline__ void lds<2>(const double* __restrict__ p, double (&v)[2]) {
    const double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    v[0] = t.x;
    v[1] = t.y;
}
template <int ROWS>
__device__ __forceinline__ void ldsc(const uint32_t* __restrict__ p, uint32_t (&v)[ROWS]);
template <>
__device__ __forceinline__ void ldsc<1>(const uint32_t* __restrict__ p, uint32_t (&v)[1]) {
    v[0] = __ldcs(p);
}
template <>
__device__ __forceinline__ void ldsc<2>(const uint32_t* __restrict__ p, uint32_t (&v)[2]) {
    const uint2 t = __ldcs(reinterpret_cast<const uint2*>(p));
    v[0] = t.x;
    v[1] = t.y;
}

// K2.  One thread owns ROWS consecutive rows.  The row of A (column indices, values, combined diagonal) is loaded
// once into registers with vector loads and applied to all P + 1 blocks; per parameter the derivative entries
// (dval, ddiag: 16 B per row and entry) are streamed once.  All loads of one block iteration are independent and are
// issued before the first use; the gathered p values a derivative entry needs are re-gathered (an L1 hit: the same
// addresses were read for block 0) instead of being held in S registers per row.
template <int S, int ROWS>
__global__ void __launch_bounds__(SV_THREADS) k_sens_matvec(const __grid_constant__ SensArgs a) {
    const MatvecArgs& m = a.m;
    if ((int)blockIdx.x < m.ntasks) {
        sens_sink_task(a);
        return;
    }
    const int64_t i0 = ((int64_t)(blockIdx.x - m.ntasks) * SV_THREADS + threadIdx.x) * ROWS;
    if (i0 >= m.n) return;
    // ---- the row of A: first-level loads, all independent
    uint32_t c[S][ROWS];
    double v[S][ROWS];
#pragma unroll
    for (int s = 0; s < S; ++s) ldsc<ROWS>(m.col + (int64_t)s * m.ld + i0, c[s]);
#pragma unroll
    for (int s = 0; s < S; ++s) lds<ROWS>(m.val + (int64_t)s * m.ld + i0, v[s]);
    double pi[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) pi[j] = (i0 + j < m.n) ? __ldg(m.x + i0 + j) : 0.0;
    double d[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) d[j] = 0.0;
    for (int k = 0; k < m.ndiag; ++k) {
        double t[ROWS];
        lds<ROWS>(m.diag + (int64_t)k * m.ld + i0, t);
#pragma unroll
        for (int j = 0; j < ROWS; ++j) d[j] = fma(m.diag_coef[k], t[j], d[j]);
    }
    // padding rows (i0 + j >= n, only inside the last thread) carry col = self, val = 0: their gathers stay in range
    // ---- block 0: y_0 = A p
    {
        double acc[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) acc[j] = d[j] * pi[j];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double cs = m.slot_coef[s];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) acc[j] = fma(cs * v[s][j], __ldg(m.x + c[s][j]), acc[j]);
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (i0 + j < m.n) m.y[i0 + j] = acc[j];
    }
    // ---- blocks 1..P: y_ip = A s_ip + sum_e [ c_e (dA_e p) + dc_e (A_r p) ]
    for (int ip = 0; ip < a.npar; ++ip) {
        const double* xb = m.x + (int64_t)(ip + 1) * a.N;
        double xs[ROWS], g[S][ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) xs[j] = (i0 + j < m.n) ? __ldg(xb + i0 + j) : 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int j = 0; j < ROWS; ++j) g[s][j] = __ldg(xb + c[s][j]);
        double acc[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) acc[j] = 0.0;
        const int e0 = a.ent_ptr[ip], e1 = a.ent_ptr[ip + 1];
        for (int e = e0; e < e1; ++e) {
            const int se = a.ent_slot[e];
            if (se < 0) continue;   // zero-stoichiometry reaction: contributes nothing
            double dv[ROWS], dd[ROWS], gp[ROWS], vr[ROWS];
            lds<ROWS>(a.dval + (int64_t)e * m.ld + i0, dv);
            lds<ROWS>(a.ddiag + (int64_t)e * m.ld + i0, dd);
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                uint32_t cc = c[0][j];
                double vv = v[0][j];
#pragma unroll
                for (int q = 1; q < S; ++q)
                    if (q == se) {
                        cc = c[q][j];
                        vv = v[q][j];
                    }
                gp[j] = __ldg(m.x + cc);   // p at the predecessor through the entry's reaction (L1 hit)
                vr[j] = vv;
            }
            const double ce = a.ent_c[e], dc = a.ent_dc[e];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) acc[j] = fma(ce, fma(dv[j], gp[j], dd[j] * pi[j]), acc[j]);
            if (dc != 0.0) {
                // (d c_r / d theta) A_r p: the reaction's own slot (raw values, without c_r) and diagonal array
                double gr[ROWS];
                lds<ROWS>(m.diag + (int64_t)a.ent_diag[e] * m.ld + i0, gr);
#pragma unroll
                for (int j = 0; j < ROWS; ++j) acc[j] = fma(dc, fma(vr[j], gp[j], gr[j] * pi[j]), acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j) acc[j] = fma(d[j], xs[j], acc[j]);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double cs = m.slot_coef[s];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) acc[j] = fma(cs * v[s][j], g[s][j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (i0 + j < m.n) m.y[(int64_t)(ip + 1) * a.N + i0 + j] = acc[j];
    }
}

__global__ void k_sens_entry_fill(const uint32_t* __restrict__ col, const double* __restrict__ dG, int64_t n, int64_t ld,
                                  double* __restrict__ dval, double* __restrict__ ddiag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    double a = 0.0, b = 0.0;
    if (i < n) {
        const uint32_t c = col[i];
        a = (c == (uint32_t)i) ? 0.0 : dG[c];
        b = -dG[i];
    }
    dval[i] = a;
    ddiag[i] = b;
}

__global__ void k_sens_sink_fill(const uint32_t* __restrict__ sink_row, int64_t begin, int64_t end,
                                 const double* __restrict__ dG, double* __restrict__ out) {
    int64_t k = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < end) out[k - begin] = dG[sink_row[k]];
}

static inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// dval / ddiag / dsink of entry e from its derivative factors dG_row (DEVICE, length A->n)
static int sens_fill_entry(ncme_sensmatrix* SA, int e, const double* dG_row) {
    ncme_matrix* A = SA->A;
    ncme_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    const int r = SA->ent_reaction[e];
    const int s = A->reaction_slot[r];
    if (s >= 0) {
        k_sens_entry_fill<<<nblk(A->ld), 256, 0, st>>>(A->col.p + (size_t)s * A->ld, dG_row, A->n, A->ld,
                                                       SA->dval.p + (size_t)e * A->ld, SA->ddiag.p + (size_t)e * A->ld);
        ctx->launches++;
        const int64_t b = A->sink_ptr[r], en = A->sink_ptr[r + 1];
        if (en > b) {
            k_sens_sink_fill<<<nblk(en - b), 256, 0, st>>>(A->sink_row.p, b, en, dG_row, SA->dsink.p + SA->dsink_ptr[e]);
            ctx->launches++;
        }
    } else {
        NCME_CUDA(cudaMemsetAsync(SA->dval.p + (size_t)e * A->ld, 0, (size_t)A->ld * 8, st));
        NCME_CUDA(cudaMemsetAsync(SA->ddiag.p + (size_t)e * A->ld, 0, (size_t)A->ld * 8, st));
    }
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int sens_describe(ncme_sensmatrix* SA, ncme_matrix** A, int* npar, int* nent) {
    NCME_REQUIRE(SA, "null sensitivity matrix");
    *A = SA->A;
    *npar = SA->npar;
    *nent = SA->nent;
    return NCME_OK;
}

}  // namespace ncme

using namespace ncme;

extern "C" {

int ncme_sensmatrix_destroy(ncme_sensmatrix* SA) {
    if (!SA) return NCME_OK;
    if (SA->A && SA->A->ctx) cudaStreamSynchronize(SA->A->ctx->stream);
    SA->dval.release();
    SA->ddiag.release();
    SA->dsink.release();
    SA->dG.release();
    SA->partial.release();
    SA->dpartial.release();
    if (SA->counter) cudaFree(SA->counter);
    if (SA->meta_dev) cudaFree(SA->meta_dev);
    delete SA;
    return NCME_OK;
}

// prev != nullptr: incremental build after an adapt! -- dpropvals then holds the derivative factors of the
// nnew = A->n - A->carry_nkept appended states only (entry-major nentries x nnew); the rows of the surviving states are
// carried over on the device through the origin map A's own incremental constructor used.
static int sens_build(ncme_matrix* A, const ncme_sensmatrix* prev, int npar, int nentries, const int32_t* ent_reaction,
                      const int32_t* ent_param, const double* dpropvals, ncme_sensmatrix** out) {
    NCME_REQUIRE(A && out && npar >= 0 && nentries >= 0, "bad arguments");
    NCME_REQUIRE(!A->comm, "the sensitivity matrix is single-GPU only");
    NCME_REQUIRE(nentries <= SMAX_ENT, "too many (reaction, parameter) entries (max %d)", SMAX_ENT);
    NCME_REQUIRE(npar <= SMAX_ENT, "too many parameters (max %d)", SMAX_ENT);
    NCME_REQUIRE(nentries == 0 || (ent_reaction && ent_param && dpropvals), "null entry arrays");
    ncme_sensmatrix* SA = new ncme_sensmatrix();
    SA->A = A;
    SA->A_n = A->n;
    SA->npar = npar;
    SA->nent = nentries;
    ncme_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    int rc = NCME_OK;
    do {
        // internal order: by parameter (stable)
        int k = 0;
        for (int ip = 0; ip < npar; ++ip) {
            SA->ent_ptr[ip] = k;
            for (int e = 0; e < nentries; ++e)
                if (ent_param[e] == ip + 1) {
                    if (ent_reaction[e] < 1 || ent_reaction[e] > A->nr) {
                        set_error("entry %d: reaction out of range", e);
                        rc = NCME_ERR_ARG;
                    }
                    SA->ent_reaction[k] = ent_reaction[e] - 1;
                    SA->ent_param[k] = ip;
                    SA->ent_user[k] = e;
                    SA->user_ent[e] = k;
                    ++k;
                }
        }
        SA->ent_ptr[npar] = k;
        if (rc != NCME_OK) break;
        if (k != nentries) {
            set_error("entry parameter index out of range 1..%d", npar);
            rc = NCME_ERR_ARG;
            break;
        }
        SA->dsink_ptr[0] = 0;
        for (int e = 0; e < nentries; ++e) {
            const int r = SA->ent_reaction[e];
            SA->dsink_ptr[e + 1] = SA->dsink_ptr[e] + (A->sink_ptr[r + 1] - A->sink_ptr[r]);
        }
        const size_t ne = (size_t)(nentries > 0 ? nentries : 1);
        if ((rc = SA->dval.reserve(ne * A->ld, st, false)) != NCME_OK) break;
        if ((rc = SA->ddiag.reserve(ne * A->ld, st, false)) != NCME_OK) break;
        if ((rc = SA->dsink.reserve((size_t)SA->dsink_ptr[nentries] + 1, st, false)) != NCME_OK) break;
        if ((rc = SA->partial.reserve((size_t)A->ntasks * (npar + 1), st, false)) != NCME_OK) break;
        if ((rc = SA->dpartial.reserve((size_t)A->ntasks * ne, st, false)) != NCME_OK) break;
        const int64_t n = A->n;
        if ((rc = SA->dG.reserve(ne * (size_t)(n > 0 ? n : 1), st, false)) != NCME_OK) break;
        if (prev) {
            bool same = prev->nent == nentries && prev->npar == npar;
            for (int e = 0; same && e < nentries; ++e)
                same = prev->ent_reaction[e] == SA->ent_reaction[e] && prev->ent_param[e] == SA->ent_param[e] &&
                       prev->ent_user[e] == SA->ent_user[e];
            if (!same) {
                set_error("incremental sensitivity build: the (reaction, parameter) pattern changed");
                rc = NCME_ERR_ARG;
                break;
            }
            const int64_t nkept = A->carry_nkept, nnew = n - nkept;
            DevArray<double> tail;
            if ((rc = tail.reserve(ne * (size_t)(nnew > 0 ? nnew : 1), st, false)) != NCME_OK) break;
            bool ok = true;
            for (int e = 0; e < nentries && nnew > 0; ++e) {
                const int r = SA->ent_reaction[e];
                if (A->kind[r] == NCME_JOINT_TV)
                    ok &= cudaMemsetAsync(tail.p + (size_t)e * nnew, 0, (size_t)nnew * 8, st) == cudaSuccess;
                else
                    ok &= cudaMemcpyAsync(tail.p + (size_t)e * nnew, dpropvals + (size_t)SA->ent_user[e] * nnew, (size_t)nnew * 8,
                                          cudaMemcpyHostToDevice, st) == cudaSuccess;
            }
            if (ok && nentries > 0)
                rc = carry_rows(ctx, prev->dG.p, prev->A_n, A->carry_origin.p, nkept, tail.p, nnew, nentries, SA->dG.p, n);
            ok &= cudaStreamSynchronize(st) == cudaSuccess;   // `tail` is released right below
            tail.release();
            if (!ok) {
                set_error("incremental sensitivity build: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = NCME_ERR_CUDA;
            }
            if (rc != NCME_OK) break;
        } else {
            bool ok = true;
            for (int e = 0; e < nentries && n > 0; ++e) {
                const int r = SA->ent_reaction[e];
                if (A->kind[r] == NCME_JOINT_TV)
                    ok &= cudaMemsetAsync(SA->dG.p + (size_t)e * n, 0, (size_t)n * 8, st) == cudaSuccess;
                else
                    ok &= cudaMemcpyAsync(SA->dG.p + (size_t)e * n, dpropvals + (size_t)SA->ent_user[e] * n, (size_t)n * 8,
                                          cudaMemcpyHostToDevice, st) == cudaSuccess;
            }
            if (!ok) {
                set_error("sensitivity build: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = NCME_ERR_CUDA;
                break;
            }
        }
        for (int e = 0; e < nentries && rc == NCME_OK; ++e) rc = sens_fill_entry(SA, e, SA->dG.p + (size_t)e * n);
        if (rc != NCME_OK) break;
        // device meta: ent_reaction | ent_param | ent_slot | ent_diag | ent_ptr(npar+1) ; then dsink_ptr (int64) ; then 2*nent doubles
        std::vector<int> meta((size_t)4 * SMAX_ENT + SMAX_ENT + 2, 0);
        for (int e = 0; e < nentries; ++e) {
            meta[(size_t)e] = SA->ent_reaction[e];
            meta[(size_t)SMAX_ENT + e] = SA->ent_param[e];
            meta[(size_t)2 * SMAX_ENT + e] = A->reaction_slot[SA->ent_reaction[e]];
            meta[(size_t)3 * SMAX_ENT + e] = A->reaction_diag[SA->ent_reaction[e]];
        }
        for (int ip = 0; ip <= npar; ++ip) meta[(size_t)4 * SMAX_ENT + ip] = SA->ent_ptr[ip];
        const size_t meta_bytes = meta.size() * sizeof(int);
        const size_t total = meta_bytes + (SMAX_ENT + 1) * sizeof(int64_t) + 2 * SMAX_ENT * sizeof(double) + 64;
        if (cudaMalloc(&SA->meta_dev, total) != cudaSuccess || cudaMalloc(&SA->counter, sizeof(unsigned)) != cudaSuccess) {
            set_error("cudaMalloc failed");
            rc = NCME_ERR_NOMEM;
            break;
        }
        cudaMemsetAsync(SA->counter, 0, sizeof(unsigned), st);
        cudaMemcpyAsync(SA->meta_dev, meta.data(), meta_bytes, cudaMemcpyHostToDevice, st);
        char* base = (char*)SA->meta_dev + round_up<size_t>(meta_bytes, 16);
        cudaMemcpyAsync(base, SA->dsink_ptr, (SMAX_ENT + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("sens meta upload failed");
            rc = NCME_ERR_CUDA;
        }
    } while (0);
    if (rc != NCME_OK) {
        ncme_sensmatrix_destroy(SA);
        return rc;
    }
    *out = SA;
    return NCME_OK;
}

int ncme_sensmatrix_create(ncme_matrix* A, int npar, int nentries, const int32_t* ent_reaction, const int32_t* ent_param,
                           const double* dpropvals, ncme_sensmatrix** out) {
    NCME_RANGE("ncme_sensmatrix_create");
    return sens_build(A, nullptr, npar, nentries, ent_reaction, ent_param, dpropvals, out);
}

int ncme_sensmatrix_create_incremental(ncme_matrix* A, ncme_sensmatrix* prev, int npar, int nentries,
                                       const int32_t* ent_reaction, const int32_t* ent_param, const double* dpropvals_new,
                                       ncme_sensmatrix** out) {
    NCME_RANGE("ncme_sensmatrix_create_incremental");
    NCME_REQUIRE(A && prev && out, "null argument");
    NCME_REQUIRE(A->carry_nkept >= 0 && A->carry_prev == prev->A && prev->dG.p,
                 "incremental sensitivity build: `mat` was not built incrementally from the matrix of `prev`");
    NCME_REQUIRE(dpropvals_new || nentries == 0 || A->n == A->carry_nkept, "dpropvals_new is null");
    return sens_build(A, prev, npar, nentries, ent_reaction, ent_param, dpropvals_new, out);
}

// Reference-structure statistics of the derivative terms ("Algorithmic bytes, sensitivity matvec"):
// the reference holds, per parameter, ONE summed CSC of its time-invariant reactions' derivatives
// (sensfspmatrixsparse.jl:44-58) and one CSC per separable / joint (reaction, parameter) entry (:60-93).
//   B_sens = bytes(A) + sum_dterms (8 nnz + 4 nnz_offdiag) + 16 N (P + 1)
int ncme_sensmatrix_stats(ncme_sensmatrix* SA, int* ndterms, int64_t* nnz_per_dterm, int64_t* algorithmic_bytes,
                          int64_t* device_bytes) {
    NCME_REQUIRE(SA, "null sensitivity matrix");
    ncme_matrix* A = SA->A;
    const int64_t n = A->n;
    int nt = 0;
    int64_t bytes = A->algorithmic_bytes - 16 * A->N;
    auto add = [&](int64_t nnz) {
        if (nnz_per_dterm) nnz_per_dterm[nt] = nnz;
        ++nt;
        bytes += 8 * nnz + 4 * (nnz - n);
    };
    for (int ip = 0; ip < SA->npar; ++ip) {
        bool any = false, slot_seen[NCME_MAX_REACTIONS] = {false};
        int64_t nnz = n;
        for (int e = SA->ent_ptr[ip]; e < SA->ent_ptr[ip + 1]; ++e) {
            const int r = SA->ent_reaction[e];
            if (A->kind[r] != NCME_TIME_INVARIANT) continue;
            any = true;
            const int s = A->reaction_slot[r];
            if (s >= 0 && !slot_seen[s]) {   // duplicates (same stoichiometry) are summed by sparse()
                slot_seen[s] = true;
                nnz += A->npred_r[r];
            }
            nnz += A->sink_ptr[r + 1] - A->sink_ptr[r];
        }
        // the reference builds this matrix for every parameter, empty (no stored entry) when no reaction depends on it
        if (any) add(nnz);
    }
    for (int pass = NCME_SEPARABLE_TV; pass <= NCME_JOINT_TV; ++pass)
        for (int e = 0; e < SA->nent; ++e) {
            const int r = SA->ent_reaction[e];
            if (A->kind[r] == pass) add(n + A->npred_r[r] + (A->sink_ptr[r + 1] - A->sink_ptr[r]));
        }
    bytes += 16 * A->N * (int64_t)(SA->npar + 1);
    if (ndterms) *ndterms = nt;
    if (algorithmic_bytes) *algorithmic_bytes = bytes;
    if (device_bytes)   // what the fused kernel actually streams: A once, two fp64 arrays per entry, the block vectors
        *device_bytes = (int64_t)A->ld * (12 * A->nslots + 8 * A->ndiag) + 16 * (int64_t)SA->nent * A->ld +
                        12 * A->nsink + 8 * SA->dsink_ptr[SA->nent] + 16 * A->N * (int64_t)(SA->npar + 1);
    return NCME_OK;
}

int ncme_sensmatrix_set_joint_values(ncme_sensmatrix* SA, int entry, const double* vals) {
    NCME_REQUIRE(SA && vals && entry >= 0 && entry < SA->nent, "bad arguments");
    const int e = SA->user_ent[entry];
    NCME_REQUIRE(SA->A->kind[SA->ent_reaction[e]] == NCME_JOINT_TV, "entry %d does not belong to a joint reaction", entry);
    ncme_matrix* A = SA->A;
    if (A->n == 0) return NCME_OK;
    cudaStream_t st = A->ctx->stream;
    double* row = SA->dG.p + (size_t)e * A->n;
    NCME_CUDA(cudaMemcpyAsync(row, vals, (size_t)A->n * 8, cudaMemcpyHostToDevice, st));
    NCME_TRY(sens_fill_entry(SA, e, row));
    NCME_CUDA(cudaStreamSynchronize(st));   // `vals` is caller memory
    return NCME_OK;
}

int ncme_sens_matvec(ncme_sensmatrix* SA, const double* coef, const double* dcoef, const double* X, double* Y) {
    NCME_RANGE("ncme_sens_matvec");
    NCME_REQUIRE(SA && X && Y && X != Y, "bad arguments");
    ncme_matrix* A = SA->A;
    ncme_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    bool need_coef = false;
    for (int r = 0; r < A->nr; ++r) need_coef |= (A->kind[r] == NCME_SEPARABLE_TV);
    NCME_REQUIRE((coef && (dcoef || SA->nent == 0)) || !need_coef, "coef/dcoef required for separable reactions");
    SensArgs a;
    matvec_fill_args(A, coef, &a.m);
    a.m.x = X;
    a.m.xd = X;
    a.m.y = Y;
    a.m.beta = 0.0;
    a.npar = SA->npar;
    a.nent = SA->nent;
    a.N = A->N;
    a.dval = SA->dval.p;
    a.ddiag = SA->ddiag.p;
    a.dsink = SA->dsink.p;
    const size_t meta_bytes = ((size_t)4 * SMAX_ENT + SMAX_ENT + 2) * sizeof(int);
    a.ent_reaction = SA->meta_dev;
    a.ent_param = SA->meta_dev + SMAX_ENT;
    a.ent_slot = SA->meta_dev + 2 * SMAX_ENT;
    a.ent_diag = SA->meta_dev + 3 * SMAX_ENT;
    a.ent_ptr = SA->meta_dev + 4 * SMAX_ENT;
    char* base = (char*)SA->meta_dev + round_up<size_t>(meta_bytes, 16);
    a.dsink_ptr = (const int64_t*)base;
    for (int e = 0; e < SA->nent; ++e) {
        const int r = SA->ent_reaction[e];
        const bool sep = A->kind[r] == NCME_SEPARABLE_TV;
        a.ent_c[e] = sep ? coef[r] : 1.0;
        a.ent_dc[e] = sep ? dcoef[SA->ent_user[e]] : 0.0;
    }
    a.partial = SA->partial.p;
    a.dpartial = SA->dpartial.p;
    a.counter = SA->counter;
    int rows = SA->tune_rows ? SA->tune_rows : (A->nslots <= 8 ? 2 : 1);
    const int64_t rpb = (int64_t)SV_THREADS * rows;
    const unsigned grid = (unsigned)(A->ntasks + (A->n + rpb - 1) / rpb);
    if (grid == 0) return NCME_OK;
#define NCME_SCASE(SS)                                                   \
    case SS:                                                             \
        if (rows == 2)                                                   \
            k_sens_matvec<SS, 2><<<grid, SV_THREADS, 0, st>>>(a);        \
        else                                                             \
            k_sens_matvec<SS, 1><<<grid, SV_THREADS, 0, st>>>(a);        \
        break;
    switch (A->nslots) {
        NCME_SCASE(1) NCME_SCASE(2) NCME_SCASE(3) NCME_SCASE(4) NCME_SCASE(5) NCME_SCASE(6) NCME_SCASE(7) NCME_SCASE(8)
        NCME_SCASE(9) NCME_SCASE(10) NCME_SCASE(11) NCME_SCASE(12) NCME_SCASE(13) NCME_SCASE(14) NCME_SCASE(15) NCME_SCASE(16)
        NCME_SCASE(17) NCME_SCASE(18) NCME_SCASE(19) NCME_SCASE(20) NCME_SCASE(21) NCME_SCASE(22) NCME_SCASE(23) NCME_SCASE(24)
        NCME_SCASE(25) NCME_SCASE(26) NCME_SCASE(27) NCME_SCASE(28) NCME_SCASE(29) NCME_SCASE(30) NCME_SCASE(31) NCME_SCASE(32)
        default:
            set_error("sens matvec: unsupported slot count %d", A->nslots);
            return NCME_ERR_ARG;
    }
#undef NCME_SCASE
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

// experiments: rows per thread of K2 (0 = auto: 2 up to 8 slots, else 1)
int ncme_sensmatrix_set_tuning(ncme_sensmatrix* SA, int rows_per_thread) {
    NCME_REQUIRE(SA && (rows_per_thread == 0 || rows_per_thread == 1 || rows_per_thread == 2), "rows per thread: 0, 1 or 2");
    SA->tune_rows = rows_per_thread;
    return NCME_OK;
}

}  // extern "C"
