// FspMatrixSparse on the device: assembly (K6) and the fused multi-term fp64 SpMV (K1).
// Reference: src/fspmatrix/sparse/fspsparsematrix.jl  (constructor :47-108, COO generator :120-152,
// joint refresh :154-166, matvec!/matvecadd! :196-247).
#include "matrix.cuh"

#include <algorithm>
#include <stdlib.h>

#include "comm.cuh"

namespace ncme {

constexpr int SINK_CHUNK = 2048;
constexpr int MV_THREADS = 256;

// ------------------------------------------------------------------------------ load helpers ---
// Matrix streams are read exactly once per matvec: streaming (evict-first) loads keep L1/L2 for x.
template <int ROWS>
__device__ __forceinline__ void ld_stream(const double* __restrict__ p, double (&v)[ROWS]);
template <>
__device__ __forceinline__ void ld_stream<1>(const double* __restrict__ p, double (&v)[1]) {
    v[0] = __ldcs(p);
}
template <>
__device__ __forceinline__ void ld_stream<2>(const double* __restrict__ p, double (&v)[2]) {
    double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    v[0] = t.x;
    v[1] = t.y;
}
template <>
__device__ __forceinline__ void ld_stream<4>(const double* __restrict__ p, double (&v)[4]) {
    // 256-bit global load (sm_100+)
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
template <int ROWS>
__device__ __forceinline__ void ld_stream(const uint32_t* __restrict__ p, uint32_t (&v)[ROWS]);
template <>
__device__ __forceinline__ void ld_stream<1>(const uint32_t* __restrict__ p, uint32_t (&v)[1]) {
    v[0] = __ldcs(p);
}
template <>
__device__ __forceinline__ void ld_stream<2>(const uint32_t* __restrict__ p, uint32_t (&v)[2]) {
    uint2 t = __ldcs(reinterpret_cast<const uint2*>(p));
    v[0] = t.x;
    v[1] = t.y;
}
template <>
__device__ __forceinline__ void ld_stream<4>(const uint32_t* __restrict__ p, uint32_t (&v)[4]) {
    uint4 t = __ldcs(reinterpret_cast<const uint4*>(p));
    v[0] = t.x;
    v[1] = t.y;
    v[2] = t.z;
    v[3] = t.w;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ------------------------------------------------------------------------------ sink rows ------
// y[n+r] = c_r * sum over the sink entries of reaction r.  One CTA per (reaction, <=SINK_CHUNK
// entries) task writes a partial; the last CTA to finish adds the partials of each reaction in task
// order, so the result is deterministic (no floating-point atomics).
__device__ __forceinline__ void sink_task(const MatvecArgs& a, const int task) {
    __shared__ double wsum[MV_THREADS / 32];
    __shared__ bool is_last;
    const int4 t = a.tasks[task];
    double s = 0.0;
    for (int k = t.y + (int)threadIdx.x; k < t.z; k += MV_THREADS) s += __ldcs(a.sink_val + k) * __ldg(a.xd + __ldcs(a.sink_row + k));
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < MV_THREADS / 32; ++w) tot += wsum[w];
        a.sink_partial[task] = tot;
        __threadfence();
        const unsigned prev = atomicAdd(a.sink_counter, 1u);
        is_last = (prev == (unsigned)a.ntasks - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if ((int)threadIdx.x < a.nr) {
        const int r = threadIdx.x;
        double tot = 0.0;
        const volatile double* part = a.sink_partial;
        for (int k = a.task_ptr[r]; k < a.task_ptr[r + 1]; ++k) tot += part[k];
        double out = a.sink_coef[r] * tot;
        if (a.beta != 0.0) out += a.beta * a.y[a.n + r];
        a.y[a.n + r] = out;
    }
    if (threadIdx.x == 0) *a.sink_counter = 0u;
}

// ------------------------------------------------------------------------------ K1 -------------
// One thread owns ROWS consecutive rows.  Per slot it streams ROWS column indices and ROWS values
// (vector loads), gathers x through the read-only path (neighbouring rows have neighbouring
// predecessors, so the gathers of a warp coalesce into a few L1 lines), and finally applies the
// time-dependent coefficients that arrive by value in the launch parameters.
template <int ROWS>
__device__ __forceinline__ void ld_stream(const uint8_t* __restrict__ p, uint32_t (&v)[ROWS]);
template <>
__device__ __forceinline__ void ld_stream<1>(const uint8_t* __restrict__ p, uint32_t (&v)[1]) {
    v[0] = __ldcs(p);
}
template <>
__device__ __forceinline__ void ld_stream<2>(const uint8_t* __restrict__ p, uint32_t (&v)[2]) {
    uchar2 t = __ldcs(reinterpret_cast<const uchar2*>(p));
    v[0] = t.x;
    v[1] = t.y;
}
template <>
__device__ __forceinline__ void ld_stream<4>(const uint8_t* __restrict__ p, uint32_t (&v)[4]) {
    uchar4 t = __ldcs(reinterpret_cast<const uchar4*>(p));
    v[0] = t.x;
    v[1] = t.y;
    v[2] = t.z;
    v[3] = t.w;
}

// Compressed column indices.  C8: one byte per entry + a per-(64-row chunk, slot) descriptor {base, mode}:
// mode 1: col = base + d;  2: col = base + d + k;  3: col = base + d + (63 - k)  (k = row within the chunk; the
// reference's LIFO level order makes predecessor indices run against the row index);  d == 255: no predecessor
// (col = the row itself, val = 0);  mode 0: the chunk's range does not fit => 32-bit indices (rare, loaded late).
// The byte and descriptor loads are issued unconditionally and up front, so they add no level to the dependent
// load chain (descriptor -> index -> gather would otherwise be three serialized DRAM latencies).
template <int ROWS>
__device__ __forceinline__ void decode_cols(const MatvecArgs& a, int s, int64_t i0, const int2 desc, const uint32_t (&d)[ROWS],
                                            uint32_t (&c)[ROWS]) {
    if (desc.y == 0) {
        ld_stream<ROWS>(a.col + (int64_t)s * a.ld + i0, c);
        return;
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int k = (int)((i0 + j) & 63);
        const int shift = desc.y == 2 ? k : (desc.y == 3 ? 63 - k : 0);
        // d == 255: no predecessor -> the row itself (padding rows past n are clamped into the buffer)
        c[j] = (d[j] == 255u) ? (a.self_off + (uint32_t)min(i0 + j, a.n - 1)) : (uint32_t)(desc.x + (int)d[j] + shift);
    }
}

// rows [i0, i0 + ROWS) of y (clipped to row_end): the whole per-thread work of K1
template <int S, int ROWS, bool C8, bool P2P>
__device__ __forceinline__ void mv_rows(const MatvecArgs& a, const int64_t i0, const int64_t row_end) {
    if (i0 >= row_end) return;

    // ---- first-level loads: all independent, all issued before the first use (one exposed DRAM latency)
    double xi[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) xi[j] = (i0 + j < row_end) ? __ldg(a.xd + i0 + j) : 0.0;
    uint32_t c[S][ROWS];
    double v[S][ROWS];
    int2 desc[C8 ? S : 1];
    if (C8) {
        const int2* dp = a.cdesc + (i0 >> 6) * S;   // descriptors of one chunk are contiguous: [chunk][slot]
#pragma unroll
        for (int s = 0; s < S; ++s) desc[s] = __ldg(dp + s);
#pragma unroll
        for (int s = 0; s < S; ++s) ld_stream<ROWS>(a.col8 + (int64_t)s * a.ld + i0, c[s]);   // raw bytes for now
    } else {
#pragma unroll
        for (int s = 0; s < S; ++s) ld_stream<ROWS>(a.col + (int64_t)s * a.ld + i0, c[s]);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) ld_stream<ROWS>(a.val + (int64_t)s * a.ld + i0, v[s]);
    constexpr int MAXD = 3;   // diagonal arrays loaded up front (time-invariant sum + two time-varying reactions)
    double dg[MAXD][ROWS];
#pragma unroll
    for (int k = 0; k < MAXD; ++k) {
        if (k < a.ndiag) {
            ld_stream<ROWS>(a.diag + (int64_t)k * a.ld + i0, dg[k]);
        } else {
#pragma unroll
            for (int j = 0; j < ROWS; ++j) dg[k][j] = 0.0;
        }
    }
    // ---- second level: the gathers of x
    if (C8) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            uint32_t d8[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) d8[j] = c[s][j];
            decode_cols<ROWS>(a, s, i0, desc[s], d8, c[s]);
        }
    }
    double g[S][ROWS];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            if (P2P) {   // boundary rows of a sharded matrix: halo entries are read from the neighbour's HBM (NVLink)
                const uint32_t cc = c[s][j];
                const double* src = cc < a.lo_end ? a.x_lo : (cc >= a.hi_begin ? a.x_hi : a.x);
                g[s][j] = src[cc];
            } else {
                g[s][j] = __ldg(a.x + c[s][j]);
            }
        }
    // ---- arithmetic
    double d[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) d[j] = 0.0;
#pragma unroll
    for (int k = 0; k < MAXD; ++k)
#pragma unroll
        for (int j = 0; j < ROWS; ++j) d[j] = fma(k < a.ndiag ? a.diag_coef[k] : 0.0, dg[k][j], d[j]);
    for (int k = MAXD; k < a.ndiag; ++k) {   // rare: more than two time-varying reactions
        double t[ROWS];
        ld_stream<ROWS>(a.diag + (int64_t)k * a.ld + i0, t);
#pragma unroll
        for (int j = 0; j < ROWS; ++j) d[j] = fma(a.diag_coef[k], t[j], d[j]);
    }
    double acc[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) acc[j] = d[j] * xi[j];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const double cs = a.slot_coef[s];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) acc[j] = fma(cs * v[s][j], g[s][j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        if (i0 + j < row_end) {
            double out = acc[j];
            if (a.beta != 0.0) out += a.beta * a.y[i0 + j];
            a.y[i0 + j] = out;
        }
    }
}


template <int S, int ROWS, bool C8, bool P2P = false>
__global__ void __launch_bounds__(MV_THREADS) k_fsp_matvec(const __grid_constant__ MatvecArgs a) {
    if (a.nsig && blockIdx.x == 0 && threadIdx.x == 0) {
        // everything enqueued before this kernel has completed: the input vector is final -> tell the readers
        __threadfence_system();
        for (int k = 0; k < a.nsig; ++k) *((volatile unsigned int*)a.sig_flag[k]) = a.epoch;
    }
    const int nt = a.do_sinks ? a.ntasks : 0;
    if ((int)blockIdx.x < nt) {
        sink_task(a, (int)blockIdx.x);
        return;
    }
    int64_t i0 = a.row_begin + ((int64_t)(blockIdx.x - nt) * MV_THREADS + threadIdx.x) * ROWS;
    int64_t row_end = a.row_end;
    if (P2P) {
        // boundary rows of a sharded matrix: wait (bounded) until the neighbours have published this matvec's input
        if (a.nwait) {
            if (threadIdx.x == 0) {
                const unsigned long long t0 = global_ns();
                for (int k = 0; k < a.nwait; ++k) {
                    const volatile unsigned int* f = a.wait_flag[k];
                    while ((int)(*f - a.epoch) < 0) {
                        if (global_ns() - t0 > 2000000000ull) {
                            atomicExch(a.err_flag, 1u);
                            break;
                        }
                    }
                }
                __threadfence_system();
            }
            __syncthreads();
        }
        const int64_t nb1 = (a.row_end - a.row_begin + MV_THREADS * ROWS - 1) / (MV_THREADS * ROWS);
        if ((int64_t)(blockIdx.x - nt) >= nb1) {   // second range
            i0 = a.row_begin2 + ((int64_t)(blockIdx.x - nt - nb1) * MV_THREADS + threadIdx.x) * ROWS;
            row_end = a.row_end2;
        }
    }
    mv_rows<S, ROWS, C8, P2P>(a, i0, row_end);
}

// ------------------------------------------------------------------------------ K1, sharded, one launch --
// The whole matvec of a row shard in ONE launch on ONE stream (K8): CTA 0 publishes ready(e) ("my x is final"),
// the halo-free rows run exactly like the single-GPU kernel (ROWS rows per thread), the boundary rows (one row per
// thread) wait -- bounded -- for ready(e) of the owners of their halo and gather it straight from the neighbours'
// HBM over NVLink.  The optional "done" handshake (the caller overwrites x right after this matvec) is folded in:
// the last boundary CTA to finish publishes done(e) to the owners and waits for done(e) of its readers.
// Replaces two launches on two streams + two event record/wait pairs + a third launch for the handshake.
__device__ __forceinline__ void flag_wait(const volatile unsigned int* f, unsigned int epoch, unsigned int* err,
                                          unsigned long long t0) {
    while ((int)(*f - epoch) < 0) {
        if (global_ns() - t0 > 2000000000ull) {
            atomicExch(err, 1u);
            break;
        }
    }
}

template <int S, int ROWS, int MINB>
__global__ void __launch_bounds__(MV_THREADS, MINB) k_fsp_matvec_sharded(const __grid_constant__ MatvecArgs a) {
    if (a.nsig && blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence_system();
        for (int k = 0; k < a.nsig; ++k) *((volatile unsigned int*)a.sig_flag[k]) = a.epoch;
    }
    const int nt = a.ntasks;
    int b = (int)blockIdx.x;
    bool boundary;
    if (a.bd_first >= 2) {
        // interleaved: among the first nb_bd * bd_first blocks every bd_first-th one is a boundary CTA, so the
        // latency-bound NVLink gathers share their SM with bandwidth-bound halo-free rows instead of filling a wave
        const int stride = a.bd_first, span = a.nb_bd * stride;
        if (b < span) {
            boundary = (b % stride) == 0;
            b = boundary ? b / stride : b - b / stride - 1;
        } else {
            boundary = false;
            b -= a.nb_bd;
        }
    } else if (a.bd_first) {
        boundary = b < a.nb_bd;
        if (!boundary) b -= a.nb_bd;
    } else {
        boundary = b >= nt + a.nb_int;
        if (boundary) b -= nt + a.nb_int;
    }
    if (!boundary) {   // b indexes [sink tasks | halo-free row tiles]
        if (b < nt) {
            sink_task(a, b);
            return;
        }
        mv_rows<S, ROWS, false, false>(a, a.int_begin + ((int64_t)(b - nt) * MV_THREADS + threadIdx.x) * ROWS, a.int_end);
        return;
    }
    if (a.nwait) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = global_ns();
            for (int k = 0; k < a.nwait; ++k) flag_wait(a.wait_flag[k], a.epoch, a.err_flag, t0);
            __threadfence_system();
        }
        __syncthreads();
    }
    {
        const int nb1 = (int)((a.row_end - a.row_begin + MV_THREADS - 1) / MV_THREADS);
        const bool second = b >= nb1;
        const int64_t i0 = (second ? a.row_begin2 + (int64_t)(b - nb1) * MV_THREADS : a.row_begin + (int64_t)b * MV_THREADS) + threadIdx.x;
        mv_rows<S, 1, false, true>(a, i0, second ? a.row_end2 : a.row_end);
    }
    if (a.ndone_sig || a.ndone_wait) {
        __syncthreads();                      // every gather of this CTA has been consumed
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(a.bd_counter, 1u) == (unsigned)a.nb_bd - 1u) {
                *a.bd_counter = 0u;
                __threadfence_system();
                for (int k = 0; k < a.ndone_sig; ++k) *((volatile unsigned int*)a.done_sig[k]) = a.epoch;
                const unsigned long long t0 = global_ns();
                for (int k = 0; k < a.ndone_wait; ++k) flag_wait(a.done_wait[k], a.epoch, a.err_flag, t0);
                __threadfence_system();
            }
        }
    }
}

template <int ROWS, int MINB>
static int launch_sharded(ncme_matrix* A, MatvecArgs& a) {
    a.nb_int = (int)((a.int_end - a.int_begin + (int64_t)MV_THREADS * ROWS - 1) / ((int64_t)MV_THREADS * ROWS));
    a.nb_bd = (int)((a.row_end - a.row_begin + MV_THREADS - 1) / MV_THREADS + (a.row_end2 - a.row_begin2 + MV_THREADS - 1) / MV_THREADS);
    const unsigned grid = (unsigned)(a.ntasks + a.nb_int + a.nb_bd);
    if (a.bd_first >= 2) {   // the interleaved span nb_bd * stride must fit the grid
        a.bd_first = (int)std::min<unsigned>((unsigned)a.bd_first, a.nb_bd > 0 ? grid / (unsigned)a.nb_bd : 1u);
        if (a.bd_first < 2) a.bd_first = 1;
    }
    cudaStream_t st = A->ctx->stream;
    switch (a.nslots) {
#define NCME_CASE(SS)                                                   \
    case SS:                                                            \
        k_fsp_matvec_sharded<SS, ROWS, MINB><<<grid, MV_THREADS, 0, st>>>(a); \
        break;
        NCME_CASE(1) NCME_CASE(2) NCME_CASE(3) NCME_CASE(4) NCME_CASE(5) NCME_CASE(6) NCME_CASE(7) NCME_CASE(8)
        NCME_CASE(9) NCME_CASE(10) NCME_CASE(11) NCME_CASE(12) NCME_CASE(13) NCME_CASE(14) NCME_CASE(15) NCME_CASE(16)
#undef NCME_CASE
        default:
            return -1;
    }
    A->ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

// ------------------------------------------------------------------------------ K1, pipelined --
// Same arithmetic as k_fsp_matvec, different data movement: the once-read matrix streams (values, diagonals,
// byte-compressed column indices, chunk descriptors) are staged through shared memory with a 4-stage cp.async
// pipeline, so the bytes in flight per SM (3 stages x ~18-36 KB per CTA, 2-3 CTAs) no longer depend on registers or
// occupancy -- which is what kept the byte-compressed variant of the register-staged kernel latency-bound.  Only the
// gathers of x and x_i/y go through the normal load/store path.  Persistent CTAs walk the row tiles round-robin.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

template <int S, int D, int ROWS, int PIPE_STAGES>
struct PipeLayout {
    static constexpr int TILE = MV_THREADS * ROWS;
    static constexpr int VAL_B = S * TILE * 8;
    static constexpr int DIAG_B = D * TILE * 8;
    static constexpr int COL_B = S * TILE;
    static constexpr int DESC_B = ((TILE / 64) * S * 8 + 15) / 16 * 16;
    static constexpr int STAGE_B = VAL_B + DIAG_B + COL_B + DESC_B;
    static constexpr int SMEM_B = STAGE_B * PIPE_STAGES;
};

template <int S, int D, int ROWS, int PIPE_STAGES>
__global__ void __launch_bounds__(MV_THREADS) k_fsp_matvec_pipe(const __grid_constant__ MatvecArgs a) {
    using L = PipeLayout<S, D, ROWS, PIPE_STAGES>;
    constexpr int TILE = L::TILE;
    extern __shared__ __align__(16) unsigned char smem[];
    const int nt = a.do_sinks ? a.ntasks : 0;
    if ((int)blockIdx.x < nt) {
        sink_task(a, (int)blockIdx.x);
        return;
    }
    const int64_t G = (int64_t)gridDim.x - nt;
    const int64_t b = (int64_t)blockIdx.x - nt;
    const int64_t tile_lo = a.row_begin / TILE;
    const int64_t tile_hi = (a.row_end + TILE - 1) / TILE;   // exclusive
    const int tid = threadIdx.x;

    auto issue = [&](int64_t tl) {
        const int64_t T = tile_lo + b + tl * G;
        if (T < tile_hi) {
            const int64_t r0 = T * TILE;
            unsigned char* st = smem + (size_t)(tl % PIPE_STAGES) * L::STAGE_B;
            // values: S rows of TILE doubles
            for (int q = tid; q < L::VAL_B / 16; q += MV_THREADS) {
                const int sl = q / (TILE / 2), o = q % (TILE / 2);
                cp_async16(st + (size_t)q * 16, a.val + (int64_t)sl * a.ld + r0 + o * 2);
            }
            for (int q = tid; q < L::DIAG_B / 16; q += MV_THREADS) {
                const int dl = q / (TILE / 2), o = q % (TILE / 2);
                cp_async16(st + L::VAL_B + (size_t)q * 16, a.diag + (int64_t)dl * a.ld + r0 + o * 2);
            }
            for (int q = tid; q < L::COL_B / 16; q += MV_THREADS) {
                const int sl = q / (TILE / 16), o = q % (TILE / 16);
                cp_async16(st + L::VAL_B + L::DIAG_B + (size_t)q * 16, a.col8 + (int64_t)sl * a.ld + r0 + o * 16);
            }
            for (int q = tid; q < (TILE / 64) * S * 8 / 16; q += MV_THREADS)
                cp_async16(st + L::VAL_B + L::DIAG_B + L::COL_B + (size_t)q * 16,
                           reinterpret_cast<const unsigned char*>(a.cdesc + (r0 >> 6) * S) + (size_t)q * 16);
        }
        cp_async_commit();   // (possibly empty) group: keeps the group count aligned with the tile counter
    };

#pragma unroll
    for (int p = 0; p < PIPE_STAGES - 1; ++p) issue(p);
    for (int64_t tl = 0;; ++tl) {
        const int64_t T = tile_lo + b + tl * G;
        if (T >= tile_hi) break;
        issue(tl + PIPE_STAGES - 1);
        cp_async_wait<PIPE_STAGES - 1>();   // tile tl has landed
        __syncthreads();
        const unsigned char* st = smem + (size_t)(tl % PIPE_STAGES) * L::STAGE_B;
        const double* sval = reinterpret_cast<const double*>(st);
        const double* sdiag = reinterpret_cast<const double*>(st + L::VAL_B);
        const unsigned char* scol = st + L::VAL_B + L::DIAG_B;
        const int2* sdesc = reinterpret_cast<const int2*>(st + L::VAL_B + L::DIAG_B + L::COL_B);
        const int l0 = tid * ROWS;                 // row inside the tile
        const int64_t i0 = T * TILE + l0;
        double xi[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) xi[j] = (i0 + j >= a.row_begin && i0 + j < a.row_end) ? __ldg(a.xd + i0 + j) : 0.0;
        // Column indices.  The rare chunks that stayed on 32-bit indices are handled by a separate (warp-uniform for
        // ROWS <= 2) path so that the common path contains no global load before the gathers: the compiler can then
        // issue all S*ROWS gathers back to back instead of serialising them behind predicated loads.
        uint32_t c[S][ROWS];
        bool wide = false;
#pragma unroll
        for (int sl = 0; sl < S; ++sl)
#pragma unroll
            for (int j = 0; j < ROWS; j += 64) wide |= (sdesc[((l0 + j) >> 6) * S + sl].y == 0);
        if (!wide) {
#pragma unroll
            for (int sl = 0; sl < S; ++sl) {
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    const int2 desc = sdesc[((l0 + j) >> 6) * S + sl];
                    const uint32_t d8 = scol[sl * TILE + l0 + j];
                    const int k = (l0 + j) & 63;
                    const int shift = desc.y == 2 ? k : (desc.y == 3 ? 63 - k : 0);
                    c[sl][j] = (d8 == 255u) ? (a.self_off + (uint32_t)min(i0 + j, a.n - 1)) : (uint32_t)(desc.x + (int)d8 + shift);
                }
            }
        } else {
#pragma unroll
            for (int sl = 0; sl < S; ++sl) ld_stream<ROWS>(a.col + (int64_t)sl * a.ld + i0, c[sl]);
        }
        double g[S][ROWS];
#pragma unroll
        for (int sl = 0; sl < S; ++sl)
#pragma unroll
            for (int j = 0; j < ROWS; ++j) g[sl][j] = __ldg(a.x + c[sl][j]);
        double acc[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            double dsum = 0.0;
#pragma unroll
            for (int dl = 0; dl < D; ++dl) dsum = fma(a.diag_coef[dl], sdiag[dl * TILE + l0 + j], dsum);
            acc[j] = dsum * xi[j];
        }
#pragma unroll
        for (int sl = 0; sl < S; ++sl) {
            const double cs = a.slot_coef[sl];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) acc[j] = fma(cs * sval[sl * TILE + l0 + j], g[sl][j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            if (i0 + j >= a.row_begin && i0 + j < a.row_end) {
                double out = acc[j];
                if (a.beta != 0.0) out += a.beta * a.y[i0 + j];
                a.y[i0 + j] = out;
            }
        }
        __syncthreads();   // the stage may be refilled by the next iteration's issue
    }
    cp_async_wait<0>();
}

template <int S, int D, int ROWS, int PIPE_STAGES>
static int launch_pipe_sdr(ncme_matrix* A, const MatvecArgs& a) {
    using L = PipeLayout<S, D, ROWS, PIPE_STAGES>;
    static bool configured = false;
    if (!configured) {
        NCME_CUDA(cudaFuncSetAttribute(k_fsp_matvec_pipe<S, D, ROWS, PIPE_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM_B));
        configured = true;
    }
    const int per_sm = std::max(1, std::min(4, (int)(220 * 1024 / L::SMEM_B)));
    const int64_t ntiles = (a.row_end + L::TILE - 1) / L::TILE - a.row_begin / L::TILE;
    const int64_t G = std::max<int64_t>(1, std::min<int64_t>((int64_t)A->ctx->sm_count * per_sm, ntiles));
    const unsigned grid = (unsigned)((a.do_sinks ? a.ntasks : 0) + G);
    k_fsp_matvec_pipe<S, D, ROWS, PIPE_STAGES><<<grid, MV_THREADS, L::SMEM_B, A->ctx->stream>>>(a);
    A->ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

// returns 1 if no pipelined instantiation covers (slots, diagonals)
template <int ROWS, int PIPE_STAGES>
static int launch_pipe(ncme_matrix* A, const MatvecArgs& a) {
#define NCME_PIPE(SS, DD) \
    if (a.nslots == SS && a.ndiag == DD) return launch_pipe_sdr<SS, DD, ROWS, PIPE_STAGES>(A, a);
    // experimental kernel: instantiated for the benchmark shapes only (4 / 6 slots, 1 / 2 diagonal arrays)
    NCME_PIPE(4, 1) NCME_PIPE(4, 2) NCME_PIPE(6, 1) NCME_PIPE(6, 2)
#undef NCME_PIPE
    return 1;
}

// Generic slot count (> 16 slots): same data flow without the register tile.
__global__ void __launch_bounds__(MV_THREADS) k_fsp_matvec_generic(const __grid_constant__ MatvecArgs a) {
    const int nt = a.do_sinks ? a.ntasks : 0;
    if ((int)blockIdx.x < nt) {
        sink_task(a, (int)blockIdx.x);
        return;
    }
    const int64_t i = a.row_begin + (int64_t)(blockIdx.x - nt) * MV_THREADS + threadIdx.x;
    if (i >= a.row_end) return;
    double d = 0.0;
    for (int k = 0; k < a.ndiag; ++k) d = fma(a.diag_coef[k], __ldcs(a.diag + (int64_t)k * a.ld + i), d);
    double acc = d * __ldg(a.xd + i);
#pragma unroll 4
    for (int s = 0; s < a.nslots; ++s) {
        const uint32_t c = __ldcs(a.col + (int64_t)s * a.ld + i);
        const double v = __ldcs(a.val + (int64_t)s * a.ld + i);
        acc = fma(a.slot_coef[s] * v, __ldg(a.x + c), acc);
    }
    if (a.beta != 0.0) acc += a.beta * a.y[i];
    a.y[i] = acc;
}

template <int ROWS, bool C8, bool P2P = false>
static int launch_rows(ncme_matrix* A, const MatvecArgs& a) {
    const int64_t rows_per_block = (int64_t)MV_THREADS * ROWS;
    const int64_t nrows = a.row_end - a.row_begin;
    const int64_t nrows2 = P2P ? (a.row_end2 - a.row_begin2) : 0;
    const unsigned grid = (unsigned)((a.do_sinks ? a.ntasks : 0) + (nrows + rows_per_block - 1) / rows_per_block +
                                     (nrows2 + rows_per_block - 1) / rows_per_block);
    if (grid == 0) return 0;
    cudaStream_t st = A->ctx->stream;
    switch (a.nslots) {
#define NCME_CASE(SS)                                              \
    case SS:                                                       \
        k_fsp_matvec<SS, ROWS, C8, P2P><<<grid, MV_THREADS, 0, st>>>(a); \
        break;
        NCME_CASE(1) NCME_CASE(2) NCME_CASE(3) NCME_CASE(4) NCME_CASE(5) NCME_CASE(6) NCME_CASE(7) NCME_CASE(8)
        NCME_CASE(9) NCME_CASE(10) NCME_CASE(11) NCME_CASE(12) NCME_CASE(13) NCME_CASE(14) NCME_CASE(15) NCME_CASE(16)
#undef NCME_CASE
        default:
            return -1;
    }
    return 0;
}

// boundary rows with peer-memory gathers
static int matvec_launch_p2p(ncme_matrix* A, const MatvecArgs& a) {
    NCME_REQUIRE(a.nslots >= 1 && a.nslots <= 16, "peer-memory halo supports up to 16 slots");
    launch_rows<1, false, true>(A, a);
    A->ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int matvec_launch(ncme_matrix* A, const MatvecArgs& a) {
    ncme_ctx* ctx = A->ctx;
    static const bool force_sharded = getenv("NCME_FORCE_SHARDED_KERNEL") != nullptr;   // experiments: code-quality check
    if (force_sharded && a.do_sinks && a.nslots <= 8 && a.row_begin == 0 && a.row_end == A->n && !A->comm) {
        MatvecArgs f = a;
        f.int_begin = 0;
        f.int_end = A->n;
        f.row_begin = f.row_end = f.row_begin2 = f.row_end2 = 0;
        f.bd_first = 1;
        f.ndone_sig = f.ndone_wait = 0;
        f.bd_counter = A->sink_counter + 1;
        return launch_sharded<2, 4>(A, f);
    }
    int rows = A->tune_rows;
    if (rows == 0) rows = (a.nslots <= 8) ? 2 : 1;
    // vector loads of x[i0..] / y are not used (scalar), but the matrix streams need i0 % ROWS == 0 only.
    int rc = -1;
    const bool c8 = A->use_c8 && A->col8.p != nullptr;
    if (A->use_pipe && A->col8.p != nullptr && a.row_end - a.row_begin >= A->pipe_min_rows) {
        int prc = 1;
        switch (A->use_pipe) {   // rows per thread + 10 * stages
            case 41: prc = launch_pipe<1, 4>(A, a); break;
            case 42: prc = launch_pipe<2, 4>(A, a); break;
            case 22: prc = launch_pipe<2, 2>(A, a); break;
            case 21: prc = launch_pipe<1, 2>(A, a); break;
            default: break;
        }
        if (prc <= 0) return prc;   // launched (0) or failed (< 0); 1 = no instantiation, fall through
    }
    if (a.nslots >= 1 && a.nslots <= 16) {
        if (rows == 4 && a.nslots <= 8)
            rc = c8 ? launch_rows<4, true>(A, a) : launch_rows<4, false>(A, a);
        else if (rows >= 2)
            rc = c8 ? launch_rows<2, true>(A, a) : launch_rows<2, false>(A, a);
        else
            rc = c8 ? launch_rows<1, true>(A, a) : launch_rows<1, false>(A, a);
    }
    if (rc != 0) {
        const unsigned grid = (unsigned)((a.do_sinks ? a.ntasks : 0) + (a.row_end - a.row_begin + MV_THREADS - 1) / MV_THREADS);
        if (grid) k_fsp_matvec_generic<<<grid, MV_THREADS, 0, ctx->stream>>>(a);
    }
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int matvec_fill_args(const ncme_matrix* A, const double* coef, MatvecArgs* a) {
    a->col = A->col.p;
    a->col8 = A->col8.p;
    a->cdesc = A->cdesc.p;
    a->nchunks = A->nchunks;
    a->self_off = (uint32_t)A->hl;
    a->val = A->val.p;
    a->diag = A->diag.p;
    a->n = A->n;
    a->ld = A->ld;
    a->nslots = A->nslots;
    a->ndiag = A->ndiag;
    a->nr = A->nr;
    for (int s = 0; s < A->nslots; ++s) {
        const int src = A->slot_coef_src[s];
        a->slot_coef[s] = (src >= 0 && A->kind[src] == NCME_SEPARABLE_TV) ? coef[src] : 1.0;
    }
    for (int d = 0; d < A->ndiag; ++d) {
        const int src = A->diag_coef_src[d];
        a->diag_coef[d] = (src >= 0 && A->kind[src] == NCME_SEPARABLE_TV) ? coef[src] : 1.0;
    }
    for (int r = 0; r < A->nr; ++r) a->sink_coef[r] = (A->kind[r] == NCME_SEPARABLE_TV) ? coef[r] : 1.0;
    a->sink_row = A->sink_row.p;
    a->sink_val = A->sink_val.p;
    a->tasks = A->tasks.p;
    a->ntasks = A->ntasks;
    for (int r = 0; r <= A->nr; ++r) a->task_ptr[r] = A->task_ptr[r];
    a->sink_partial = A->sink_partial.p;
    a->sink_counter = A->sink_counter;
    a->row_begin = 0;
    a->row_end = A->n;
    a->row_begin2 = a->row_end2 = 0;
    a->nwait = 0;
    a->nsig = 0;
    a->epoch = 0;
    a->err_flag = nullptr;
    a->do_sinks = 1;
    return NCME_OK;
}

// Halo of x between row shards: one grouped ncclSend/ncclRecv round with the ranks whose rows this rank's
// predecessor window touches (rank +-1 for level-ordered state spaces).
int halo_exchange(ncme_matrix* A, const double* x_local, cudaStream_t st) {
    if (!A->comm || A->comm->nranks == 1 || (A->halo_send.empty() && A->halo_recv.empty())) return NCME_OK;
    const NcclApi* api = nccl_api();
    double* x = const_cast<double*>(x_local);
    NCME_NCCL(api->GroupStart());
    for (const auto& sg : A->halo_send) {
        ncclResult_t r = api->Send(x + sg.offset, (size_t)sg.count, ncclDouble, sg.peer, A->comm->nccl, st);
        if (r != ncclSuccess) {
            api->GroupEnd();
            set_error("ncclSend failed: %s", api->GetErrorString(r));
            return NCME_ERR_COMM;
        }
        A->comm->bytes_sent += 8 * sg.count;
    }
    for (const auto& sg : A->halo_recv) {
        ncclResult_t r = api->Recv(x + sg.offset, (size_t)sg.count, ncclDouble, sg.peer, A->comm->nccl, st);
        if (r != ncclSuccess) {
            api->GroupEnd();
            set_error("ncclRecv failed: %s", api->GetErrorString(r));
            return NCME_ERR_COMM;
        }
    }
    NCME_NCCL(api->GroupEnd());
    return NCME_OK;
}

// ---- cross-GPU flag synchronisation for the peer-memory halo (no NCCL in the loop) -----------------------------
struct SyncArgs {
    int nsig, nwait;
    unsigned int* sig[8];          // peer memory: slots to store the epoch into
    const unsigned int* wait[8];   // local memory: slots that must reach the epoch
    unsigned int epoch;
    unsigned int* err;
};

// one thread: publish, then (optionally) wait.  Waits are bounded (2 s): a lost peer raises an error flag instead of
// hanging the GPU.
__global__ void k_p2p_sync(const __grid_constant__ SyncArgs a) {
    if (threadIdx.x != 0) return;
    if (a.nsig) {
        __threadfence_system();
        for (int k = 0; k < a.nsig; ++k) *((volatile unsigned int*)a.sig[k]) = a.epoch;
    }
    if (a.nwait) {
        const unsigned long long t0 = global_ns();
        for (int k = 0; k < a.nwait; ++k) {
            const volatile unsigned int* f = a.wait[k];
            while ((int)(*f - a.epoch) < 0) {
                if (global_ns() - t0 > 2000000000ull) {
                    atomicExch(a.err, 1u);
                    break;
                }
            }
        }
        __threadfence_system();
    }
}

static int p2p_sync(ncme_matrix* A, const SyncArgs& sa) {
    if (sa.nsig == 0 && sa.nwait == 0) return NCME_OK;
    k_p2p_sync<<<1, 32, 0, A->ctx->stream>>>(sa);
    A->ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

// Halo of x read straight from the neighbours' HBM by the boundary rows (CUDA IPC + NVLink).  Flags replace the
// collective: ready(e) is published before the interior rows start and awaited inside the boundary-rows kernel;
// done(e) protects the input against being overwritten while a neighbour still reads it -- skipped when the caller
// alternates input buffers, because observing ready(e+1) from a neighbour already implies it finished matvec e.
static int matvec_dist_p2p(ncme_matrix* A, MatvecArgs a, const double* xlo, const double* xhi, int flags) {
    ncme_comm* c = A->comm;
    const unsigned int e = ++c->epoch;
    const int me = c->rank;
    SyncArgs ready_sig{}, done{};
    ready_sig.epoch = done.epoch = e;
    ready_sig.err = done.err = &c->my_flags->error;
    a.nwait = 0;
    a.epoch = e;
    a.err_flag = &c->my_flags->error;
    auto add_wait = [&](int q) {
        for (int k = 0; k < a.nwait; ++k)
            if (a.wait_flag[k] == &c->my_flags->ready[q]) return;
        if (a.nwait < 4) a.wait_flag[a.nwait++] = &c->my_flags->ready[q];
    };
    for (int q : A->readers) {                               // they read my x: tell them it is complete
        ready_sig.sig[ready_sig.nsig++] = &c->peer_flags[q]->ready[me];
        done.wait[done.nwait++] = &c->my_flags->done[q];
        add_wait(q);                                         // (their ready(e) also certifies they are past e-1)
    }
    const bool sig_in_kernel = ready_sig.nsig <= 4;          // published by CTA 0 of the first matvec kernel
    if (sig_in_kernel) {
        a.nsig = ready_sig.nsig;
        for (int k = 0; k < ready_sig.nsig; ++k) a.sig_flag[k] = ready_sig.sig[k];
    }
    for (int q : {A->plo, A->phi}) {
        if (q < 0) continue;
        add_wait(q);
        done.sig[done.nsig++] = &c->peer_flags[q]->done[me];
    }
    // padded position c -> peer address: low halo c in [0, hl): global ext_lo + c; high halo c >= hl + n + R
    a.lo_end = (uint32_t)A->hl;
    a.hi_begin = (uint32_t)(A->hl + A->n + A->nr);
    a.x_lo = xlo ? xlo + (A->ext_lo - A->plo_row_lo) : a.x;
    a.x_hi = xhi ? xhi + (A->row_hi - A->phi_row_lo) - (int64_t)a.hi_begin : a.x;
    const bool interior = A->b1 > A->b0;
    // One launch (k_fsp_matvec_sharded) or two (halo-free rows on the compute stream + boundary rows on the priority
    // stream).  Measured on B200 (profiles/README.md, round 2): the two-launch path is 3-9 % faster at 2 and 4 GPUs
    // (84 vs 87-90 us and 44 vs 48 us per matvec), so it stays the default; NCME_P2P_ONE_LAUNCH=1 selects the other.
    static const bool two_launch = getenv("NCME_P2P_ONE_LAUNCH") == nullptr;
    static const int bd_mode = getenv("NCME_P2P_BD_MODE") ? atoi(getenv("NCME_P2P_BD_MODE")) : 4;   // 0 last, 1 first, >= 2 interleave stride
    const bool handshake = !(flags & 2);
    if (!two_launch && sig_in_kernel && a.nslots <= 16 && done.nsig <= 4 && done.nwait <= 4 && a.nwait > 0) {
        MatvecArgs f = a;
        f.int_begin = interior ? A->b0 : 0;
        f.int_end = interior ? A->b1 : 0;
        f.row_begin = 0;
        f.row_end = interior ? A->b0 : A->n;
        f.row_begin2 = interior ? A->b1 : 0;
        f.row_end2 = interior ? A->n : 0;
        f.bd_first = bd_mode;
        f.ndone_sig = handshake ? done.nsig : 0;
        f.ndone_wait = handshake ? done.nwait : 0;
        for (int k = 0; k < f.ndone_sig; ++k) f.done_sig[k] = done.sig[k];
        for (int k = 0; k < f.ndone_wait; ++k) f.done_wait[k] = done.wait[k];
        f.bd_counter = A->sink_counter + 1;
        if (f.row_end - f.row_begin + f.row_end2 - f.row_begin2 > 0) {
            static const bool minb3 = getenv("NCME_SHARDED_MINB3") != nullptr;   // experiments: 3 CTAs/SM, no spills
            if (minb3)
                NCME_TRY(a.nslots <= 8 ? (launch_sharded<2, 3>(A, f)) : (launch_sharded<1, 3>(A, f)));
            else
                NCME_TRY(a.nslots <= 8 ? (launch_sharded<2, 4>(A, f)) : (launch_sharded<1, 4>(A, f)));
            c->p2p_matvecs++;
            if (flags & 1) NCME_TRY(comm_allreduce_sum(c, a.y + A->n, (size_t)A->nr, A->ctx->stream));
            return NCME_OK;
        }
    }
    if (!sig_in_kernel) NCME_TRY(p2p_sync(A, ready_sig));
    if (interior) {
        // boundary rows (peer loads over NVLink, waiting for the neighbours inside the kernel) run on the
        // high-priority stream CONCURRENTLY with the interior rows; they write disjoint rows of y
        cudaStream_t st = A->ctx->stream;
        NCME_CUDA(cudaEventRecord(c->ev_ready, st));
        NCME_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
        MatvecArgs bd = a;                                   // both boundary ranges in one launch
        bd.row_begin = 0;
        bd.row_end = A->b0;
        bd.row_begin2 = A->b1;
        bd.row_end2 = A->n;
        bd.do_sinks = 0;   // (this kernel's CTA 0 publishes "ready" before it starts waiting: no rank can starve another)
        A->ctx->stream = c->comm_stream;
        int rc = matvec_launch_p2p(A, bd);
        A->ctx->stream = st;
        NCME_TRY(rc);
        NCME_CUDA(cudaEventRecord(c->ev_done, c->comm_stream));
        MatvecArgs in = a;
        in.row_begin = A->b0;
        in.row_end = A->b1;
        in.do_sinks = 1;
        in.nsig = 0;
        NCME_TRY(matvec_launch(A, in));
        NCME_CUDA(cudaStreamWaitEvent(st, c->ev_done, 0));
    } else {
        NCME_TRY(matvec_launch_p2p(A, a));
    }
    if (!(flags & 2)) NCME_TRY(p2p_sync(A, done));
    c->p2p_matvecs++;
    if (flags & 1) NCME_TRY(comm_allreduce_sum(c, a.y + A->n, (size_t)A->nr, A->ctx->stream));
    return NCME_OK;
}

int matvec_dist(ncme_matrix* A, const double* coef, const double* x_local, double* y_local, double beta, int reduce_sinks) {
    MatvecArgs a;
    matvec_fill_args(A, coef, &a);
    a.xd = x_local;
    a.x = x_local - A->hl;
    a.y = y_local;
    a.beta = beta;
    ncme_comm* c = A->comm;
    if (!c || c->nranks == 1) return matvec_launch(A, a);
    cudaStream_t st = A->ctx->stream;
    if (c->p2p_ok && A->p2p_eligible && beta == 0.0) {
        // registered on every rank or on none (registration is collective), so all ranks take the same branch
        const double* xlo = A->plo >= 0 ? comm_peer_vector(c, x_local, A->plo) : nullptr;
        const double* xhi = A->phi >= 0 ? comm_peer_vector(c, x_local, A->phi) : nullptr;
        bool registered = false;
        for (const auto& r : c->regs)
            registered |= ((const char*)x_local >= (const char*)r.base && (const char*)x_local < (const char*)r.base + r.bytes);
        if (registered && (A->plo < 0 || xlo) && (A->phi < 0 || xhi)) return matvec_dist_p2p(A, a, xlo, xhi, reduce_sinks);
    }
    c->nccl_matvecs++;
    static const bool no_overlap_env = getenv("NCME_NO_OVERLAP") != nullptr;   // experiments only
    const bool overlap = A->b1 > A->b0 && !no_overlap_env;
    if (!overlap) {
        NCME_TRY(halo_exchange(A, x_local, st));
        NCME_TRY(matvec_launch(A, a));
    } else {
        // halo on the communication stream while the rows that touch no halo entry are computed
        NCME_CUDA(cudaEventRecord(c->ev_ready, st));
        NCME_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
        NCME_TRY(halo_exchange(A, x_local, c->comm_stream));
        NCME_CUDA(cudaEventRecord(c->ev_done, c->comm_stream));
        MatvecArgs in = a;
        in.row_begin = A->b0;
        in.row_end = A->b1;
        in.do_sinks = 1;     // sink rows read local x only
        NCME_TRY(matvec_launch(A, in));
        NCME_CUDA(cudaStreamWaitEvent(st, c->ev_done, 0));
        if (A->b0 > 0) {
            MatvecArgs lo = a;
            lo.row_begin = 0;
            lo.row_end = A->b0;
            lo.do_sinks = 0;
            NCME_TRY(matvec_launch(A, lo));
        }
        if (A->b1 < A->n) {
            MatvecArgs hi = a;
            hi.row_begin = A->b1;
            hi.row_end = A->n;
            hi.do_sinks = 0;
            NCME_TRY(matvec_launch(A, hi));
        }
    }
    if (reduce_sinks & 1) NCME_TRY(comm_allreduce_sum(c, y_local + A->n, (size_t)A->nr, st));
    return NCME_OK;
}

// diag(A(t))_i = sum_d cd_d(t) diag[d][i]   (Jacobi preconditioner of the BDF/GMRES integrator)
struct DiagArgs {
    int64_t n, ld;
    int ndiag;
    const double* diag;
    double coef[NCME_MAX_REACTIONS + 1];
    double* out;
};
__global__ void k_matrix_diag(const __grid_constant__ DiagArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    double d = 0.0;
    for (int k = 0; k < a.ndiag; ++k) d = fma(a.coef[k], a.diag[(int64_t)k * a.ld + i], d);
    a.out[i] = d;
}

int matrix_diag(ncme_matrix* A, const double* coef, double* out) {
    if (A->n == 0) return NCME_OK;
    MatvecArgs m;
    matvec_fill_args(A, coef, &m);
    DiagArgs a;
    a.n = A->n;
    a.ld = A->ld;
    a.ndiag = A->ndiag;
    a.diag = A->diag.p;
    for (int k = 0; k < A->ndiag; ++k) a.coef[k] = m.diag_coef[k];
    a.out = out;
    k_matrix_diag<<<(unsigned)((A->n + 255) / 256), 256, 0, A->ctx->stream>>>(a);
    A->ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int matvec_sinks_only(ncme_matrix* A, const double* coef, const double* x_local, double* y_local) {
    MatvecArgs a;
    matvec_fill_args(A, coef, &a);
    a.xd = x_local;
    a.x = x_local - A->hl;
    a.y = y_local;
    a.beta = 0.0;
    a.row_begin = a.row_end = 0;
    a.do_sinks = 1;
    return matvec_launch(A, a);
}

// ------------------------------------------------------------------------------ K6 assembly ----
struct SlotReactions {
    int count;
    int r[NCME_MAX_REACTIONS];
};

// col/val of one slot from the predecessor table and the uploaded state factors G[r][i].
struct ShardGeom {
    int64_t n_global, row_lo, row_hi, ext_lo, nloc;
    int nr;
};

// position of global state j inside the halo-padded x buffer [halo_lo | local | sinks | halo_hi]
__device__ __forceinline__ uint32_t padded_pos(const ShardGeom& g, int64_t j) {
    return (uint32_t)(j - g.ext_lo + (j >= g.row_hi ? g.nr : 0));
}

__global__ void k_assemble_slot(const uint32_t* __restrict__ pred_r0 /*global rows*/, const double* __restrict__ G,
                                ShardGeom g, SlotReactions sr, uint32_t* __restrict__ col, double* __restrict__ val,
                                int64_t ld) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    uint32_t c = padded_pos(g, g.row_lo);
    double v = 0.0;
    if (i < g.nloc) {
        const int64_t gi = g.row_lo + i;
        const uint32_t p = pred_r0[gi];
        c = padded_pos(g, p == NONE32 ? gi : (int64_t)p);
        if (p != NONE32)
            for (int k = 0; k < sr.count; ++k) v += G[(int64_t)sr.r[k] * g.n_global + p];
    }
    col[i] = c;
    val[i] = v;
}

__global__ void k_assemble_diag(const double* __restrict__ G, ShardGeom g, SlotReactions sr, double* __restrict__ diag,
                                int64_t ld) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    double v = 0.0;
    if (i < g.nloc)
        for (int k = 0; k < sr.count; ++k) v -= G[(int64_t)sr.r[k] * g.n_global + g.row_lo + i];
    diag[i] = v;
}

// predecessor window of the local rows (min / max global predecessor index) and the interior row range
__global__ void k_pred_window(const uint32_t* __restrict__ pred_r /*global rows*/, ShardGeom g, unsigned int* mnmx /*[4]*/) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nloc) return;
    const uint32_t p = pred_r[g.row_lo + i];
    if (p == NONE32) return;
    if ((int64_t)p < g.row_lo) {
        atomicMin(&mnmx[0], p);
        atomicMax(&mnmx[2], (unsigned int)(i + 1));   // rows [0, b0) touch the low halo
    }
    if ((int64_t)p >= g.row_hi) {
        atomicMax(&mnmx[1], p);
        atomicMin(&mnmx[3], (unsigned int)i);         // rows [b1, nloc) touch the high halo
    }
}

// One warp per (slot, 64-row chunk): pick the addressing mode whose residual range fits one byte.
__global__ void k_compress_cols(const uint32_t* __restrict__ col, int64_t ld, int64_t n, int64_t nchunks, int nslots,
                                uint32_t self_off, uint8_t* __restrict__ col8, int2* __restrict__ cdesc,
                                unsigned long long* wide_count) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nchunks * nslots) return;
    const int s = (int)(w / nchunks);
    const int64_t chunk = w % nchunks;
    const int64_t r0 = chunk * 64;
    long long mn[3] = {(1ll << 40), (1ll << 40), (1ll << 40)}, mx[3] = {-(1ll << 40), -(1ll << 40), -(1ll << 40)};
    long long cv[2];
    bool has[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        const int64_t i = r0 + k;
        const uint32_t c = col[(int64_t)s * ld + i];
        has[h] = (i < n) && c != self_off + (uint32_t)i;
        cv[h] = (long long)c;
        if (has[h]) {
            const long long r[3] = {cv[h], cv[h] - k, cv[h] - (63 - k)};
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                mn[m] = min(mn[m], r[m]);
                mx[m] = max(mx[m], r[m]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[m] = min(mn[m], __shfl_xor_sync(0xffffffffu, mn[m], d));
            mx[m] = max(mx[m], __shfl_xor_sync(0xffffffffu, mx[m], d));
        }
    int mode = 0;
    long long base = 0;
    if (mx[0] < mn[0]) {          // no predecessor in the whole chunk
        mode = 1;
        base = 0;
    } else {
        for (int m = 2; m >= 0; --m)   // prefer the plain mode when several fit
            if (mx[m] - mn[m] <= 254 && mn[m] > -(1ll << 31) && mx[m] < (1ll << 31)) {
                mode = m + 1;
                base = mn[m];
            }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        uint8_t d = 255;
        if (mode != 0 && has[h]) {
            const long long shift = mode == 2 ? k : (mode == 3 ? 63 - k : 0);
            d = (uint8_t)(cv[h] - shift - base);
        }
        col8[(int64_t)s * ld + r0 + k] = d;
    }
    if (lane == 0) {
        cdesc[chunk * nslots + s] = make_int2((int)base, mode);
        if (mode == 0) atomicAdd(wide_count, 1ull);
    }
}

// per row chunk: the highest padded x position its gathers touch (host-buffer pipeline)
// row chunks of the host-buffer pipeline: chunk c = rows [row[c], row[c+1])
struct PipeChunks {
    int nc;
    int64_t row[17];
};
__global__ void k_chunk_reach(const uint32_t* __restrict__ col, int64_t ld, int64_t n, int nslots, PipeChunks pc,
                              unsigned int* __restrict__ reach /*[16]*/) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int m = 0;
    for (int s = 0; s < nslots; ++s) m = max(m, col[(int64_t)s * ld + i]);
    int c = 0;
    while (c < pc.nc - 1 && i >= pc.row[c + 1]) ++c;
    atomicMax(&reach[c], m);
}

// counts[r] = rows with a predecessor through reaction r; counts[NCME_MAX_REACTIONS + r] = rows whose sink flag r is set
__global__ void __launch_bounds__(256) k_count_structure(const uint32_t* __restrict__ pred, int64_t ld, int64_t row_lo,
                                                         const smask_t* __restrict__ sinkmask, int64_t n, int nr,
                                                         smask_t validmask, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int sh[2 * NCME_MAX_REACTIONS];
    if (threadIdx.x < 2 * NCME_MAX_REACTIONS) sh[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n;
    const smask_t m = in ? (sinkmask[row_lo + i] & validmask) : 0;
    for (int r = 0; r < nr; ++r) {
        const bool hp = in && pred[(int64_t)r * ld + row_lo + i] != NONE32;
        const unsigned bp = __ballot_sync(0xffffffffu, hp), bs = __ballot_sync(0xffffffffu, (m >> r) & 1u);
        if ((threadIdx.x & 31) == 0) {
            if (bp) atomicAdd(&sh[r], (unsigned)__popc(bp));
            if (bs) atomicAdd(&sh[NCME_MAX_REACTIONS + r], (unsigned)__popc(bs));
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * NCME_MAX_REACTIONS && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// flags[q * n + i] = sink flag of reaction r0 + q at local row i
__global__ void k_sink_flags_group(const smask_t* __restrict__ mask, int64_t n, int r0, int g, smask_t validmask,
                                   uint32_t* __restrict__ flags) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)g * n) return;
    const int q = (int)(t / n);
    const int64_t i = t - (int64_t)q * n;
    flags[t] = (uint32_t)(((mask[i] & validmask) >> (r0 + q)) & 1u);
}

__global__ void k_fill_sinks_group(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, int64_t n, int g,
                                   int r0, const double* __restrict__ G, int64_t ng, int64_t row_lo, int64_t base,
                                   uint32_t* __restrict__ sink_row, double* __restrict__ sink_val) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)g * n || !flags[t]) return;
    const int q = (int)(t / n);
    const int64_t i = t - (int64_t)q * n;
    const int64_t k = base + pos[t];
    sink_row[k] = (uint32_t)i;
    sink_val[k] = G[(int64_t)(r0 + q) * ng + row_lo + i];
}




// _update_sparsematrix! (:154-166) for one joint reaction.  vals holds f(t, x, theta) over the states this rank's
// rows can reach: the window [ext_lo, ext_hi) = [low halo (hl) | local rows (n) | high halo]; on a single GPU that is
// the whole state list (hl = 0, no halo).  Column indices are positions in the padded matvec input
// [low halo | local rows | nr sinks | high halo], so positions behind the local rows skip the nr sink entries.
__global__ void k_set_joint(const double* __restrict__ vals, int64_t n, uint32_t hl, uint32_t nr,
                            const uint32_t* __restrict__ col, double* __restrict__ val, double* __restrict__ diag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = col[i], self = hl + (uint32_t)i;
    const uint32_t w = (c < hl + (uint32_t)n) ? c : c - nr;
    val[i] = (c == self) ? 0.0 : vals[w];
    diag[i] = -vals[self];
}

__global__ void k_set_joint_sinks(const double* __restrict__ vals, uint32_t hl, const uint32_t* __restrict__ sink_row,
                                  int64_t begin, int64_t end, double* __restrict__ sink_val) {
    int64_t k = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < end) sink_val[k] = vals[hl + sink_row[k]];
}

static inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

static bool same_stoich(const ncme_space* sp, int r1, int r2) {
    for (int s = 0; s < sp->ns; ++s)
        if (sp->stoich[(size_t)r1 * sp->ns + s] != sp->stoich[(size_t)r2 * sp->ns + s]) return false;
    return true;
}
static bool zero_stoich(const ncme_space* sp, int r) {
    for (int s = 0; s < sp->ns; ++s)
        if (sp->stoich[(size_t)r * sp->ns + s] != 0) return false;
    return true;
}

// Byte-compressed column indices for the experimental matvec variants (ncme_matrix_set_tuning(+16/+32/+64),
// ncme_matrix_set_pipe): built on first use, the default 32-bit kernel never needs them.
static int matrix_ensure_compressed(ncme_matrix* A) {
    if (A->col8.p || A->nslots == 0) return NCME_OK;
    ncme_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    const int nslots = A->nslots;
    const int64_t n = A->n;
    A->nchunks = A->ld / 64;
    NCME_TRY(A->col8.reserve((size_t)A->ld * nslots, st, false));
    NCME_TRY(A->cdesc.reserve((size_t)A->nchunks * nslots, st, false));
    unsigned long long* d_wide = nullptr;
    NCME_CUDA(cudaMalloc(&d_wide, sizeof(unsigned long long)));
    NCME_CUDA(cudaMemsetAsync(d_wide, 0, sizeof(unsigned long long), st));
    const int64_t nwarps = A->nchunks * nslots;
    k_compress_cols<<<nblk(nwarps * 32), 256, 0, st>>>(A->col.p, A->ld, n, A->nchunks, nslots, (uint32_t)A->hl,
                                                       A->col8.p, A->cdesc.p, d_wide);
    ctx->launches++;
    unsigned long long hw = 0;
    NCME_CUDA(cudaMemcpyAsync(&hw, d_wide, sizeof(hw), cudaMemcpyDeviceToHost, st));
    NCME_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_wide);
    A->wide_chunks = (int64_t)hw;
    return NCME_OK;
}

// G_new[r][i] = G_prev[r][origin[i]] for the surviving states (a prefix), the uploaded factors for the new tail
__global__ void k_carry_factors(const double* __restrict__ Gprev, int64_t ngprev, const uint32_t* __restrict__ origin,
                                int64_t nkept, const double* __restrict__ Gnew_tail, int64_t nnew, int nr,
                                double* __restrict__ G, int64_t ng) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ng * nr) return;
    const int r = (int)(t / ng);
    const int64_t i = t - (int64_t)r * ng;
    G[t] = i < nkept ? Gprev[(int64_t)r * ngprev + origin[i]] : Gnew_tail[(int64_t)r * nnew + (i - nkept)];
}

int carry_rows(ncme_ctx* ctx, const double* prev, int64_t ngprev, const uint32_t* origin, int64_t nkept, const double* tail,
               int64_t nnew, int nrows, double* out, int64_t ng) {
    if (ng <= 0 || nrows <= 0) return NCME_OK;
    k_carry_factors<<<nblk(ng * nrows), 256, 0, ctx->stream>>>(prev, ngprev, origin, nkept, tail, nnew, nrows, out, ng);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

// prev != nullptr: incremental build -- `propvals` then holds the factors of the nnew = n - nkept states appended since
// prev was built (reaction-major nnew x nr), everything else is carried over on the device.
// win_lo >= 0: windowed build of a row shard -- `propvals` holds the factors of the states [win_lo, win_hi) only
// (reaction-major, stride win_hi - win_lo); the window must cover this rank's rows and their predecessor window.
static int matrix_build(ncme_space* sp, const int32_t* kind, const double* propvals, ncme_matrix* A, ncme_comm* comm,
                        const ncme_matrix* prev = nullptr, int64_t nkept = 0, int64_t win_lo = -1, int64_t win_hi = -1) {
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t st = ctx->stream;
    const int nr = sp->nr;
    const int64_t ng = sp->n;             // global number of states
    // contiguous row blocks, boundaries on multiples of 64 rows (vector-load alignment of the row ranges)
    const int P = comm ? comm->nranks : 1, me = comm ? comm->rank : 0;
    auto cut = [&](int r) -> int64_t {
        if (r <= 0) return 0;
        if (r >= P) return ng;
        return std::min<int64_t>(ng, round_up<int64_t>((int64_t)((__int128)ng * r / P), 64));
    };
    const int64_t row_lo = cut(me), row_hi = cut(me + 1);
    const int64_t n = row_hi - row_lo;    // local rows
    A->ctx = ctx;
    A->comm = (comm && comm->nranks > 1) ? comm : nullptr;
    A->ns = sp->ns;
    A->nr = nr;
    A->n = n;
    A->N = n + nr;
    A->n_global = ng;
    A->row_lo = row_lo;
    A->row_hi = row_hi;
    A->ld = round_up<int64_t>(n > 0 ? n : 1, 512);   // multiple of the largest row tile of the pipelined kernel
    for (int r = 0; r < nr; ++r) {
        NCME_REQUIRE(kind[r] >= 0 && kind[r] <= 2, "bad reaction kind");
        A->kind[r] = kind[r];
    }
    // ---- slot / diagonal plan
    SlotReactions slots[NCME_MAX_REACTIONS];
    SlotReactions diags[NCME_MAX_REACTIONS + 1];
    int nslots = 0, ndiag = 0;
    bool any_ti = false;
    for (int r = 0; r < nr; ++r) any_ti |= (kind[r] == NCME_TIME_INVARIANT);
    if (any_ti) {
        diags[0].count = 0;
        A->diag_coef_src[0] = -1;
        ndiag = 1;
    }
    for (int r = 0; r < nr; ++r) {
        A->reaction_slot[r] = -1;
        A->reaction_diag[r] = -1;
        if (zero_stoich(sp, r)) continue;  // x -> x: the +a and -a entries cancel, contributes nothing
        if (kind[r] == NCME_TIME_INVARIANT) {
            diags[0].r[diags[0].count++] = r;
            A->reaction_diag[r] = 0;
            int found = -1;
            for (int s = 0; s < nslots; ++s)
                if (A->slot_coef_src[s] == -1 && same_stoich(sp, slots[s].r[0], r)) found = s;
            if (found >= 0) {
                slots[found].r[slots[found].count++] = r;
                A->reaction_slot[r] = found;
                continue;
            }
            A->slot_coef_src[nslots] = -1;
        } else {
            A->slot_coef_src[nslots] = r;
            diags[ndiag].count = 1;
            diags[ndiag].r[0] = r;
            A->diag_coef_src[ndiag] = r;
            A->reaction_diag[r] = ndiag;
            ndiag++;
        }
        slots[nslots].count = 1;
        slots[nslots].r[0] = r;
        A->slot_first_reaction[nslots] = r;
        A->reaction_slot[r] = nslots;
        nslots++;
    }
    A->nslots = nslots;
    A->ndiag = ndiag;

    // ---- upload the state factors of ALL states (the values at predecessors outside the shard are needed;
    //      joint reactions start at zero, :87,129)
    DevArray<double>& G = A->G;
    NCME_TRY(G.reserve((size_t)(ng > 0 ? ng : 1) * nr, st, false));
    if (prev && ng > 0) {
        const int64_t nnew = ng - nkept;
        DevArray<double> tail;
        NCME_TRY(tail.reserve((size_t)(nnew > 0 ? nnew : 1) * nr, st, false));
        for (int r = 0; r < nr && nnew > 0; ++r) {
            if (kind[r] == NCME_JOINT_TV || !propvals)
                NCME_CUDA(cudaMemsetAsync(tail.p + (size_t)r * nnew, 0, (size_t)nnew * 8, st));
            else
                NCME_CUDA(cudaMemcpyAsync(tail.p + (size_t)r * nnew, propvals + (size_t)r * nnew, (size_t)nnew * 8, cudaMemcpyHostToDevice, st));
        }
        NCME_TRY(carry_rows(ctx, prev->G.p, prev->n_global, sp->origin.p, nkept, tail.p, nnew, nr, G.p, ng));
        tail.release();
        NCME_TRY(A->carry_origin.reserve((size_t)(nkept > 0 ? nkept : 1), st, false));
        if (nkept > 0) NCME_CUDA(cudaMemcpyAsync(A->carry_origin.p, sp->origin.p, (size_t)nkept * 4, cudaMemcpyDeviceToDevice, st));
        A->carry_nkept = nkept;
        A->carry_prev = prev;
    } else if (win_lo >= 0) {
        NCME_REQUIRE(win_lo <= row_lo && row_hi <= win_hi && win_hi <= ng, "state-factor window does not cover this rank's rows");
        const int64_t nw = win_hi - win_lo;
        A->g_window = true;     // G is only valid inside the window: no incremental rebuild from this matrix
        for (int r = 0; r < nr && nw > 0; ++r) {
            if (kind[r] == NCME_JOINT_TV || !propvals)
                NCME_CUDA(cudaMemsetAsync(G.p + (size_t)r * ng + win_lo, 0, (size_t)nw * 8, st));
            else
                NCME_CUDA(cudaMemcpyAsync(G.p + (size_t)r * ng + win_lo, propvals + (size_t)r * nw, (size_t)nw * 8, cudaMemcpyHostToDevice, st));
        }
    } else {
        for (int r = 0; r < nr && ng > 0; ++r) {
            if (kind[r] == NCME_JOINT_TV || !propvals)
                NCME_CUDA(cudaMemsetAsync(G.p + (size_t)r * ng, 0, (size_t)ng * 8, st));
            else
                NCME_CUDA(cudaMemcpyAsync(G.p + (size_t)r * ng, propvals + (size_t)r * ng, (size_t)ng * 8, cudaMemcpyHostToDevice, st));
        }
    }
    // ---- predecessor window of the local rows -> halo extents and the halo-free interior row range
    ShardGeom geom{ng, row_lo, row_hi, row_lo, n, nr};
    A->ext_lo = row_lo;
    A->ext_hi = row_hi;
    A->b0 = 0;
    A->b1 = n;
    if (A->comm && n > 0) {
        unsigned int h_mm[4] = {0xFFFFFFFFu, 0u, 0u, (unsigned int)n};
        unsigned int* d_mm = nullptr;
        NCME_CUDA(cudaMalloc(&d_mm, sizeof(h_mm)));
        NCME_CUDA(cudaMemcpyAsync(d_mm, h_mm, sizeof(h_mm), cudaMemcpyHostToDevice, st));
        for (int s = 0; s < nslots; ++s) {
            k_pred_window<<<nblk(n), 256, 0, st>>>(sp->pred.p + (size_t)slots[s].r[0] * sp->ld, geom, d_mm);
            ctx->launches++;
        }
        NCME_CUDA(cudaMemcpyAsync(h_mm, d_mm, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
        NCME_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_mm);
        if (h_mm[0] != 0xFFFFFFFFu) A->ext_lo = std::min<int64_t>(row_lo, (int64_t)h_mm[0]);
        if (h_mm[1] != 0u) A->ext_hi = std::max<int64_t>(row_hi, (int64_t)h_mm[1] + 1);
        A->b0 = round_up<int64_t>((int64_t)h_mm[2], 64);
        A->b1 = (int64_t)h_mm[3] / 64 * 64;
        if (A->b0 >= A->b1) A->b0 = A->b1 = 0;   // no halo-free interior: everything waits for the halo
    }
    A->hl = row_lo - A->ext_lo;
    A->hh = A->ext_hi - row_hi;
    NCME_REQUIRE(win_lo < 0 || (win_lo <= A->ext_lo && A->ext_hi <= win_hi),
                 "state-factor window does not cover the predecessor window of this rank's rows (see ncme_matrix_shard_window)");
    geom.ext_lo = A->ext_lo;
    NCME_REQUIRE(A->ext_hi - A->ext_lo + nr < 0xFFFFFFF0ll, "padded window exceeds the 32-bit index range");
    NCME_TRY(A->col.reserve((size_t)A->ld * (nslots > 0 ? nslots : 1), st, false));
    NCME_TRY(A->val.reserve((size_t)A->ld * (nslots > 0 ? nslots : 1), st, false));
    NCME_TRY(A->diag.reserve((size_t)A->ld * (ndiag > 0 ? ndiag : 1), st, false));
    for (int s = 0; s < nslots; ++s) {
        k_assemble_slot<<<nblk(A->ld), 256, 0, st>>>(sp->pred.p + (size_t)slots[s].r[0] * sp->ld, G.p, geom, slots[s],
                                                     A->col.p + (size_t)s * A->ld, A->val.p + (size_t)s * A->ld, A->ld);
        ctx->launches++;
    }
    for (int d = 0; d < ndiag; ++d) {
        k_assemble_diag<<<nblk(A->ld), 256, 0, st>>>(G.p, geom, diags[d], A->diag.p + (size_t)d * A->ld, A->ld);
        ctx->launches++;
    }
    NCME_CUDA(cudaGetLastError());
    // ---- row chunks of the host-buffer pipeline (single GPU): how far ahead each chunk's gathers reach
    A->pipe_chunks = 0;
    if (!A->comm && n >= 16 * 4096 && nslots > 0) {
        static const int NC = [] {   // experiments: NCME_HOST_PIPE_CHUNKS = 2..16 (default 16)
            const char* e = getenv("NCME_HOST_PIPE_CHUNKS");
            const int v = e ? atoi(e) : 16;
            return v < 2 ? 2 : (v > 16 ? 16 : v);
        }();
        // Chunk sizes ramp up and down (1 : 2 : 4 : 8 ... 8 : 4 : 2 : 1): the first download can start after ~3 % of the
        // upload instead of 2/NC of it, and what is left after the last upload is ~3 % of the download.  Both directions
        // of the link then overlap for almost the whole call (measured floor of this link for 80 MB each way at once:
        // 1.80 ms, tools/pcie_ceiling.py).  NCME_HOST_PIPE_UNIFORM=1: equal chunks (round 1).
        static const bool uniform = getenv("NCME_HOST_PIPE_UNIFORM") != nullptr;
        PipeChunks pc;
        {
            static const int ramp[4] = {1, 2, 4, 8};
            double w[16], wsum = 0.0;
            for (int c = 0; c < NC; ++c) {
                const int d = std::min(c, NC - 1 - c);
                w[c] = uniform ? 1.0 : (double)ramp[std::min(d, 3)];
                wsum += w[c];
            }
            int nc = 0;
            int64_t r0 = 0;
            double acc = 0.0;
            pc.row[0] = 0;
            for (int c = 0; c < NC && r0 < n; ++c) {
                acc += w[c];
                int64_t r1 = (c == NC - 1) ? n : std::min<int64_t>(n, round_up<int64_t>((int64_t)((double)n * acc / wsum), 512));
                if (r1 <= r0) continue;
                pc.row[++nc] = r1;
                r0 = r1;
            }
            pc.row[nc] = n;
            pc.nc = nc;
        }
        unsigned int* d_reach = nullptr;
        NCME_CUDA(cudaMalloc(&d_reach, 16 * sizeof(unsigned int)));
        NCME_CUDA(cudaMemsetAsync(d_reach, 0, 16 * sizeof(unsigned int), st));
        k_chunk_reach<<<nblk(n), 256, 0, st>>>(A->col.p, A->ld, n, nslots, pc, d_reach);
        ctx->launches++;
        unsigned int h_reach[16];
        NCME_CUDA(cudaMemcpyAsync(h_reach, d_reach, sizeof(h_reach), cudaMemcpyDeviceToHost, st));
        NCME_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_reach);
        for (int c = 0; c < pc.nc; ++c) {
            A->pipe_row[c] = pc.row[c];
            A->pipe_need_hi[c] = (int64_t)h_reach[c] + 1;
        }
        A->pipe_row[pc.nc] = n;
        A->pipe_chunks = pc.nc;
    }
    // byte-compressed column indices (experimental kernel variants) are built on demand: matrix_ensure_compressed
    A->nchunks = A->ld / 64;
    A->wide_chunks = 0;
    // ---- halo plan: who owns what I need, who needs what I own
    if (A->comm) {
        double mine[4] = {(double)row_lo, (double)row_hi, (double)A->ext_lo, (double)A->ext_hi};
        double* all = A->comm->scratch;
        NCME_CUDA(cudaMemcpyAsync(all + 4 * me, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
        NCME_NCCL(nccl_api()->AllGather(all + 4 * me, all, 4, ncclDouble, A->comm->nccl, st));
        std::vector<double> h((size_t)4 * P);
        NCME_CUDA(cudaMemcpyAsync(h.data(), all, sizeof(double) * 4 * P, cudaMemcpyDeviceToHost, st));
        NCME_CUDA(cudaStreamSynchronize(st));
        auto isect = [](int64_t a0, int64_t a1, int64_t b0, int64_t b1, int64_t* s0, int64_t* s1) {
            *s0 = std::max(a0, b0);
            *s1 = std::min(a1, b1);
            return *s1 > *s0;
        };
        A->halo_send.clear();
        A->halo_recv.clear();
        for (int q = 0; q < P; ++q) {
            if (q == me) continue;
            const int64_t qlo = (int64_t)h[4 * q], qhi = (int64_t)h[4 * q + 1], qelo = (int64_t)h[4 * q + 2],
                          qehi = (int64_t)h[4 * q + 3];
            int64_t s0, s1;
            // what q needs from my rows: its low halo [qelo, qlo) and its high halo [qhi, qehi)
            if (isect(qelo, qlo, row_lo, row_hi, &s0, &s1)) A->halo_send.push_back({q, s0 - row_lo, s1 - s0});
            if (isect(qhi, qehi, row_lo, row_hi, &s0, &s1)) A->halo_send.push_back({q, s0 - row_lo, s1 - s0});
            // what I need from q's rows
            if (isect(A->ext_lo, row_lo, qlo, qhi, &s0, &s1)) A->halo_recv.push_back({q, s0 - row_lo, s1 - s0});
            if (isect(row_hi, A->ext_hi, qlo, qhi, &s0, &s1)) A->halo_recv.push_back({q, s0 - row_lo + nr, s1 - s0});
        }
        // peer-memory eligibility, evaluated for EVERY rank from the gathered table so that all ranks agree
        bool all_ok = true;
        A->readers.clear();
        A->plo = A->phi = -1;
        for (int r = 0; r < P; ++r) {
            const int64_t rlo = (int64_t)h[4 * r], rhi = (int64_t)h[4 * r + 1], relo = (int64_t)h[4 * r + 2],
                          rehi = (int64_t)h[4 * r + 3];
            int owner_lo = -1, owner_hi = -1;
            for (int q = 0; q < P; ++q) {
                if (q == r) continue;
                const int64_t qlo = (int64_t)h[4 * q], qhi = (int64_t)h[4 * q + 1];
                if (relo < rlo && qlo <= relo && rlo <= qhi) owner_lo = q;
                if (rehi > rhi && qlo <= rhi && rehi <= qhi) owner_hi = q;
            }
            if ((relo < rlo && owner_lo < 0) || (rehi > rhi && owner_hi < 0)) all_ok = false;
            if (r == me) {
                A->plo = owner_lo;
                A->phi = owner_hi;
                if (owner_lo >= 0) A->plo_row_lo = (int64_t)h[4 * owner_lo];
                if (owner_hi >= 0) A->phi_row_lo = (int64_t)h[4 * owner_hi];
            }
            if (owner_lo == me || owner_hi == me) A->readers.push_back(r);
        }
        if (A->readers.size() > 8) all_ok = false;
        A->p2p_eligible = all_ok;
    }

    // ---- structural counts (one kernel, one device->host fetch) and sink lists (rows ascending inside each reaction)
    int64_t npred[NCME_MAX_REACTIONS] = {0};
    std::vector<uint64_t> nsink_r((size_t)nr, 0);
    smask_t validmask = 0;
    for (int r = 0; r < nr; ++r)
        if (!zero_stoich(sp, r)) validmask |= SMASK1(r);
    A->sink_ptr[0] = 0;
    if (n > 0) {
        unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ctx->red_result_dev);
        unsigned long long* h_cnt = reinterpret_cast<unsigned long long*>(ctx->red_result_host);
        NCME_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 2 * NCME_MAX_REACTIONS, st));
        k_count_structure<<<nblk(n), 256, 0, st>>>(sp->pred.p, sp->ld, row_lo, sp->sinkmask.p, n, nr, validmask, d_cnt);
        ctx->launches++;
        NCME_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(unsigned long long) * 2 * NCME_MAX_REACTIONS, cudaMemcpyDeviceToHost, st));
        NCME_CUDA(cudaStreamSynchronize(st));
        for (int r = 0; r < nr; ++r) {
            npred[r] = (int64_t)h_cnt[r];
            nsink_r[(size_t)r] = h_cnt[NCME_MAX_REACTIONS + r];
        }
    }
    for (int r = 0; r < nr; ++r) A->sink_ptr[r + 1] = A->sink_ptr[r] + (int64_t)nsink_r[(size_t)r];
    A->nsink = A->sink_ptr[nr];
    NCME_TRY(A->sink_row.reserve((size_t)(A->nsink > 0 ? A->nsink : 1), st, false));
    NCME_TRY(A->sink_val.reserve((size_t)(A->nsink > 0 ? A->nsink : 1), st, false));
    DevArray<uint32_t> flags, pos, scratch;
    int rc = NCME_OK;
    if (A->nsink > 0) {
        // reactions are processed in groups whose reaction-major flag array (g x n) is scanned at once: the scan
        // position of an entry is then its offset inside the group's part of the sink list
        const int64_t group = std::max<int64_t>(1, std::min<int64_t>(nr, ((int64_t)1 << 24) / n));
        NCME_TRY(flags.reserve((size_t)(group * n), st, false));
        NCME_TRY(pos.reserve((size_t)(group * n), st, false));
        NCME_TRY(scratch.reserve(scan_scratch_elems(group * n), st, false));
        for (int r0 = 0; r0 < nr && rc == NCME_OK; r0 += (int)group) {
            const int g = (int)std::min<int64_t>(group, nr - r0);
            if (A->sink_ptr[r0 + g] == A->sink_ptr[r0]) continue;
            k_sink_flags_group<<<nblk((int64_t)g * n), 256, 0, st>>>(sp->sinkmask.p + row_lo, n, r0, g, validmask, flags.p);
            ctx->launches++;
            rc = exclusive_scan_u32(ctx, flags.p, pos.p, (int64_t)g * n, scratch.p, scratch.cap, nullptr);
            k_fill_sinks_group<<<nblk((int64_t)g * n), 256, 0, st>>>(flags.p, pos.p, n, g, r0, G.p, ng, row_lo, A->sink_ptr[r0],
                                                                    A->sink_row.p, A->sink_val.p);
            ctx->launches++;
        }
    }
    if (rc == NCME_OK && cudaGetLastError() != cudaSuccess) {
        set_error("matrix assembly failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = NCME_ERR_CUDA;
    }
    flags.release();
    pos.release();
    scratch.release();
    NCME_TRY(rc);

    // ---- sink tasks
    std::vector<int4> tasks;
    for (int r = 0; r < nr; ++r) {
        A->task_ptr[r] = (int)tasks.size();
        int64_t b = A->sink_ptr[r], e = A->sink_ptr[r + 1];
        do {
            int64_t e2 = (e - b > SINK_CHUNK) ? b + SINK_CHUNK : e;
            tasks.push_back(make_int4(r, (int)b, (int)e2, 0));
            b = e2;
        } while (b < e);
    }
    A->task_ptr[nr] = (int)tasks.size();
    A->ntasks = (int)tasks.size();
    NCME_REQUIRE(A->nsink < 0x7FFFFFFF, "too many sink entries");
    NCME_TRY(A->tasks.reserve(tasks.size(), st, false));
    NCME_TRY(A->sink_partial.reserve(tasks.size(), st, false));
    NCME_CUDA(cudaMemcpyAsync(A->tasks.p, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    NCME_TRY(A->sink_counter_mem.reserve(2, st, false));   // [0] sink tasks, [1] boundary CTAs of the sharded launch
    A->sink_counter = A->sink_counter_mem.p;
    NCME_CUDA(cudaMemsetAsync(A->sink_counter, 0, 2 * sizeof(unsigned), st));
    NCME_CUDA(cudaStreamSynchronize(st));

    // ---- reference-structure statistics
    int nt = 0;
    if (any_ti) {
        int64_t nnz = n;
        for (int s = 0; s < nslots; ++s)
            if (A->slot_coef_src[s] == -1) nnz += npred[slots[s].r[0]];
        for (int r = 0; r < nr; ++r)
            if (kind[r] == NCME_TIME_INVARIANT) nnz += (int64_t)nsink_r[(size_t)r];
        A->nnz_term[nt++] = nnz;
    }
    for (int pass = NCME_SEPARABLE_TV; pass <= NCME_JOINT_TV; ++pass)
        for (int r = 0; r < nr; ++r)
            if (kind[r] == pass) A->nnz_term[nt++] = n + npred[r] + (int64_t)nsink_r[(size_t)r];
    A->nterms = nt;
    for (int r = 0; r < nr; ++r) A->npred_r[r] = npred[r];
    A->algorithmic_bytes = 16 * A->N;
    for (int k = 0; k < nt; ++k) A->algorithmic_bytes += 8 * A->nnz_term[k] + 4 * (A->nnz_term[k] - n);
    // from now on the space tracks where its states were at this build (incremental constructor of the next matrix)
    NCME_TRY(space_mark(sp));
    A->space_mark = sp->mark_id;
    return NCME_OK;
}

}  // namespace ncme

using namespace ncme;

extern "C" {

int ncme_matrix_create(ncme_space* space, const int32_t* kind, const double* propvals, ncme_matrix** out) {
    return ncme_matrix_create_sharded(space, nullptr, kind, propvals, out);
}

int ncme_matrix_create_sharded(ncme_space* space, ncme_comm* comm, const int32_t* kind, const double* propvals,
                               ncme_matrix** out) {
    NCME_RANGE("ncme_matrix_create_sharded");
    NCME_REQUIRE(space && kind && out, "null argument");
    NCME_REQUIRE(propvals || space->n == 0, "propvals is null");
    ncme_matrix* A = new ncme_matrix();
    int st = matrix_build(space, kind, propvals, A, comm);
    if (st != NCME_OK) {
        ncme_matrix_destroy(A);
        return st;
    }
    *out = A;
    return NCME_OK;
}

// Rows and predecessor window of this rank's shard, BEFORE any propensity is evaluated: the host then evaluates the
// state factors of the states [ext_lo, ext_hi) only and builds with ncme_matrix_create_window (sharded build: host
// evaluation and upload shrink with the number of ranks).  out = {row_lo, row_hi, ext_lo, ext_hi}.
int ncme_matrix_shard_window(ncme_space* sp, ncme_comm* comm, int64_t out[4]) {
    NCME_RANGE("ncme_matrix_shard_window");
    NCME_REQUIRE(sp && out, "null argument");
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t ng = sp->n;
    const int P = comm ? comm->nranks : 1, me = comm ? comm->rank : 0;
    auto cut = [&](int r) -> int64_t {
        if (r <= 0) return 0;
        if (r >= P) return ng;
        return std::min<int64_t>(ng, round_up<int64_t>((int64_t)((__int128)ng * r / P), 64));
    };
    const int64_t row_lo = cut(me), row_hi = cut(me + 1), n = row_hi - row_lo;
    out[0] = out[2] = row_lo;
    out[1] = out[3] = row_hi;
    if (P == 1 || n <= 0) return NCME_OK;
    ShardGeom geom{ng, row_lo, row_hi, row_lo, n, sp->nr};
    unsigned int h_mm[4] = {0xFFFFFFFFu, 0u, 0u, (unsigned int)n};
    unsigned int* d_mm = nullptr;
    NCME_CUDA(cudaMalloc(&d_mm, sizeof(h_mm)));
    NCME_CUDA(cudaMemcpyAsync(d_mm, h_mm, sizeof(h_mm), cudaMemcpyHostToDevice, st));
    for (int r = 0; r < sp->nr; ++r) {
        if (zero_stoich(sp, r)) continue;
        k_pred_window<<<nblk(n), 256, 0, st>>>(sp->pred.p + (size_t)r * sp->ld, geom, d_mm);
        ctx->launches++;
    }
    NCME_CUDA(cudaMemcpyAsync(h_mm, d_mm, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    NCME_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_mm);
    if (h_mm[0] != 0xFFFFFFFFu) out[2] = std::min<int64_t>(row_lo, (int64_t)h_mm[0]);
    if (h_mm[1] != 0u) out[3] = std::max<int64_t>(row_hi, (int64_t)h_mm[1] + 1);
    return NCME_OK;
}

int ncme_matrix_create_window(ncme_space* space, ncme_comm* comm, const int32_t* kind, const double* propvals_window,
                              int64_t win_lo, int64_t win_hi, ncme_matrix** out) {
    NCME_RANGE("ncme_matrix_create_window");
    NCME_REQUIRE(space && kind && out, "null argument");
    NCME_REQUIRE(win_lo >= 0 && win_lo <= win_hi && win_hi <= space->n, "bad window");
    NCME_REQUIRE(propvals_window || win_hi == win_lo, "propvals is null");
    ncme_matrix* A = new ncme_matrix();
    int st = matrix_build(space, kind, propvals_window, A, comm, nullptr, 0, win_lo, win_hi);
    if (st != NCME_OK) {
        ncme_matrix_destroy(A);
        return st;
    }
    *out = A;
    return NCME_OK;
}

int ncme_space_new_count(ncme_space* space, int64_t* n_kept, int64_t* n_new) {
    NCME_REQUIRE(space && n_kept && n_new, "null argument");
    if (space->mark_n < 0) {   // never marked: everything is new
        *n_kept = 0;
        *n_new = space->n;
        return NCME_OK;
    }
    NCME_TRY(space_count_kept(space, n_kept));
    *n_new = space->n - *n_kept;
    return NCME_OK;
}

int ncme_matrix_create_incremental(ncme_space* space, ncme_comm* comm, ncme_matrix* prev, const int32_t* kind,
                                   const double* propvals_new, ncme_matrix** out) {
    NCME_RANGE("ncme_matrix_create_incremental");
    NCME_REQUIRE(space && prev && kind && out, "null argument");
    NCME_REQUIRE(prev->space_mark == space->mark_id && space->mark_n >= 0 && prev->G.p,
                 "incremental build: `prev` is not the matrix this space was last assembled into");
    NCME_REQUIRE(!prev->g_window, "incremental build: `prev` was built from a state-factor window (sharded build)");
    NCME_REQUIRE(prev->nr == space->nr, "incremental build: reaction count changed");
    for (int r = 0; r < space->nr; ++r) NCME_REQUIRE(prev->kind[r] == kind[r], "incremental build: reaction kinds changed");
    int64_t nkept = 0;
    NCME_TRY(space_count_kept(space, &nkept));
    NCME_REQUIRE(propvals_new || space->n == nkept, "propvals_new is null");
    ncme_matrix* A = new ncme_matrix();
    int st = matrix_build(space, kind, propvals_new, A, comm, prev, nkept);
    if (st != NCME_OK) {
        ncme_matrix_destroy(A);
        return st;
    }
    *out = A;
    return NCME_OK;
}

int ncme_matrix_register_buffer(ncme_matrix* A, void* base_dev, size_t bytes, int64_t local0) {
    NCME_REQUIRE(A && base_dev, "null argument");
    if (!A->comm) return NCME_OK;
    const int peers[2] = {A->plo, A->phi};
    return comm_register(A->comm, base_dev, bytes, local0, 0, 1, peers, 2);
}

int ncme_matrix_unregister_buffer(ncme_matrix* A, void* base_dev) {
    NCME_REQUIRE(A && base_dev, "null argument");
    if (!A->comm) return NCME_OK;
    return comm_unregister(A->comm, base_dev);
}

int ncme_matrix_set_pipe(ncme_matrix* A, int rows, int stages) {   // experiments: shared-memory pipelined kernel variant
    NCME_REQUIRE(A, "null matrix");
    A->use_pipe = rows > 0 ? rows + 10 * stages : 0;
    A->pipe_min_rows = 0;
    if (A->use_pipe) NCME_TRY(matrix_ensure_compressed(A));
    return NCME_OK;
}

int ncme_matrix_shard_info(ncme_matrix* A, int64_t info[8]) {
    NCME_REQUIRE(A && info, "null argument");
    info[0] = A->row_lo;
    info[1] = A->row_hi;
    info[2] = A->hl;
    info[3] = A->hh;
    info[4] = A->n_global;
    info[5] = A->b0;
    info[6] = A->b1;
    info[7] = A->comm ? A->comm->nranks : 1;
    return NCME_OK;
}

int ncme_matrix_destroy(ncme_matrix* A) {
    if (!A) return NCME_OK;
    if (A->ctx) cudaStreamSynchronize(A->ctx->stream);
    A->col.release();
    A->col8.release();
    A->cdesc.release();
    A->val.release();
    A->diag.release();
    A->G.release();
    A->carry_origin.release();
    A->sink_row.release();
    A->sink_val.release();
    A->tasks.release();
    A->sink_partial.release();
    A->sink_counter_mem.release();
    delete A;
    return NCME_OK;
}

int ncme_matrix_size(ncme_matrix* A, int64_t* rows, int64_t* cols) {
    NCME_REQUIRE(A, "null matrix");
    if (rows) *rows = A->N;
    if (cols) *cols = A->N;
    return NCME_OK;
}

int ncme_matrix_set_tuning(ncme_matrix* A, int rows_per_thread) {
    // rows_per_thread: 0 (auto), 1, 2, 4; +16: byte-compressed column indices in the register-staged kernel;
    // +32 / +64: shared-memory pipelined kernel with 1 / 2 rows per thread; +128: never use the pipelined kernel
    const int rows = rows_per_thread & 15;
    NCME_REQUIRE(A && (rows == 0 || rows == 1 || rows == 2 || rows == 4) && (rows_per_thread & ~255) == 0,
                 "rows_per_thread must be 0, 1, 2 or 4 (+16 / +32 / +64 / +128 select kernel variants)");
    A->tune_rows = rows;
    A->use_c8 = (rows_per_thread & 16) ? 1 : 0;
    if (rows_per_thread & 32) A->use_pipe = 41;
    if (rows_per_thread & 64) A->use_pipe = 42;
    if (rows_per_thread & 128) A->use_pipe = 0;
    if (rows_per_thread & (32 | 64)) A->pipe_min_rows = 0;
    if (A->use_c8 || A->use_pipe) NCME_TRY(matrix_ensure_compressed(A));
    return NCME_OK;
}

int ncme_matrix_set_joint_values(ncme_matrix* A, int reaction, const double* vals) {
    NCME_RANGE("ncme_matrix_set_joint_values");
    NCME_REQUIRE(A && vals, "null argument");
    NCME_REQUIRE(reaction >= 1 && reaction <= A->nr && A->kind[reaction - 1] == NCME_JOINT_TV,
                 "reaction %d is not a joint time-varying reaction", reaction);
    const int r = reaction - 1;
    const int s = A->reaction_slot[r], d = A->reaction_diag[r];
    if (s < 0 || A->n == 0) return NCME_OK;
    ncme_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    // sharded matrices take the values over this rank's window [ext_lo, ext_hi) (ncme_matrix_shard_info), see k_set_joint
    const int64_t nw = A->hl + A->n + A->hh;
    DevArray<double> tmp;
    NCME_TRY(tmp.reserve((size_t)nw, st, false));
    NCME_CUDA(cudaMemcpyAsync(tmp.p, vals, (size_t)nw * 8, cudaMemcpyHostToDevice, st));
    k_set_joint<<<nblk(A->n), 256, 0, st>>>(tmp.p, A->n, (uint32_t)A->hl, (uint32_t)A->nr, A->col.p + (size_t)s * A->ld,
                                            A->val.p + (size_t)s * A->ld, A->diag.p + (size_t)d * A->ld);
    ctx->launches++;
    const int64_t b = A->sink_ptr[r], e = A->sink_ptr[r + 1];
    if (e > b) {
        k_set_joint_sinks<<<nblk(e - b), 256, 0, st>>>(tmp.p, (uint32_t)A->hl, A->sink_row.p, b, e, A->sink_val.p);
        ctx->launches++;
    }
    NCME_CUDA(cudaGetLastError());
    NCME_CUDA(cudaStreamSynchronize(st));
    tmp.release();
    return NCME_OK;
}

int ncme_matvec(ncme_matrix* A, const double* coef, const double* x_dev, double* y_dev, double beta) {
    NCME_RANGE("ncme_matvec");
    NCME_REQUIRE(A && x_dev && y_dev, "null argument");
    NCME_REQUIRE(x_dev != y_dev, "matvec!: input and output must not alias");
    bool need_coef = false;
    for (int r = 0; r < A->nr; ++r) need_coef |= (A->kind[r] == NCME_SEPARABLE_TV);
    NCME_REQUIRE(coef || !need_coef, "coef is null but the matrix has separable time-varying reactions");
    NCME_REQUIRE(!(A->comm && beta != 0.0), "matvecadd! is not supported on row-sharded matrices");
    return matvec_dist(A, coef, x_dev, y_dev, beta, 1);
}

int ncme_matvec_local(ncme_matrix* A, const double* coef, const double* x_dev, double* y_dev) {
    NCME_REQUIRE(A && x_dev && y_dev && x_dev != y_dev, "bad arguments");
    bool need_coef = false;
    for (int r = 0; r < A->nr; ++r) need_coef |= (A->kind[r] == NCME_SEPARABLE_TV);
    NCME_REQUIRE(coef || !need_coef, "coef is null but the matrix has separable time-varying reactions");
    return matvec_dist(A, coef, x_dev, y_dev, 0.0, 2);
}

int ncme_matvec_host(ncme_matrix* A, const double* coef, const double* x_host, double* y_host, double beta) {
    NCME_RANGE("ncme_matvec_host");
    NCME_REQUIRE(A && x_host && y_host, "null argument");
    NCME_REQUIRE(!A->comm, "the host-buffer matvec is single-GPU only");
    ncme_ctx* ctx = A->ctx;
    const size_t bytes = (size_t)A->N * sizeof(double);
    if (ctx->stage_dev_bytes < bytes) {
        if (ctx->stage_dev_x) cudaFree(ctx->stage_dev_x);
        if (ctx->stage_dev_y) cudaFree(ctx->stage_dev_y);
        ctx->stage_dev_x = ctx->stage_dev_y = nullptr;
        ctx->stage_dev_bytes = 0;
        NCME_CUDA(cudaMalloc(&ctx->stage_dev_x, bytes));
        NCME_CUDA(cudaMalloc(&ctx->stage_dev_y, bytes));
        ctx->stage_dev_bytes = bytes;
    }
    cudaStream_t st = ctx->stream;
    if (beta != 0.0 || A->pipe_chunks < 2) {
        NCME_CUDA(cudaMemcpyAsync(ctx->stage_dev_x, x_host, bytes, cudaMemcpyHostToDevice, st));
        if (beta != 0.0) NCME_CUDA(cudaMemcpyAsync(ctx->stage_dev_y, y_host, bytes, cudaMemcpyHostToDevice, st));
        NCME_TRY(ncme_matvec(A, coef, ctx->stage_dev_x, ctx->stage_dev_y, beta));
        NCME_CUDA(cudaMemcpyAsync(y_host, ctx->stage_dev_y, bytes, cudaMemcpyDeviceToHost, st));
        NCME_CUDA(cudaStreamSynchronize(st));
        return NCME_OK;
    }
    // NCME_HOST_PIPE_TRACE=1 (diagnostics): events carry timestamps and the 5th call prints, per chunk, when its upload,
    // kernel and download finished relative to the start of the call
    static const bool pipe_trace = getenv("NCME_HOST_PIPE_TRACE") != nullptr;
    static cudaEvent_t trace_d2h[16];
    static int trace_calls = 0;
    // ---- pipelined: H2D of x (chunk c+1), rows of chunk c, D2H of y (chunk c-1) overlap on three streams.
    // Chunk c may start once x is on the device up to the furthest entry its gathers reach (pipe_need_hi).
    if (!ctx->h2d_stream) {
        NCME_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
        NCME_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        const unsigned evflags = pipe_trace ? cudaEventDefault : cudaEventDisableTiming;
        for (int k = 0; k < 16; ++k) {
            NCME_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[k], evflags));
            NCME_CUDA(cudaEventCreateWithFlags(&ctx->ev_comp[k], evflags));
            if (pipe_trace) NCME_CUDA(cudaEventCreate(&trace_d2h[k]));
        }
        NCME_CUDA(cudaEventCreateWithFlags(&ctx->ev_start, evflags));
    }
    const int nc = A->pipe_chunks;
    double* xd = ctx->stage_dev_x;
    double* yd = ctx->stage_dev_y;
    // Experiment (NCME_HOST_ZEROCOPY=1, off by default): with a page-locked (mapped) output buffer the row kernels can
    // store y straight into host memory over the link, which removes the device->host copy stage and its dependency
    // lag.  Measured on B200 / PCIe Gen5 (M-3D, 80.3 MB each way): 2.60 ms per matvec against 2.13 ms with the copy
    // engines -- SM stores to host memory do not reach the DMA rate -- so the copy stage stays (profiles/README.md).
    static const bool zc_enabled = [] {
        const char* e = getenv("NCME_HOST_ZEROCOPY");
        return e && e[0] == '1';
    }();
    double* y_map = nullptr;
    if (zc_enabled) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, y_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            y_map = static_cast<double*>(at.devicePointer);
        else
            cudaGetLastError();   // pageable memory: not an error
    }
    NCME_CUDA(cudaEventRecord(ctx->ev_start, st));                       // order after earlier work on the context
    NCME_CUDA(cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_start, 0));
    NCME_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_start, 0));
    // Uploads are ISSUED in dependency order, interleaved with the kernels and downloads below (issuing all 16 uploads
    // first kept the host busy for ~130 us before the first kernel / download could even be queued).
    int issued = 0;
    auto issue_uploads = [&](int upto) -> int {   // chunks [issued, upto]
        for (; issued <= upto && issued < nc; ++issued) {
            const int64_t r0 = A->pipe_row[issued];
            const int64_t r1 = (issued == nc - 1) ? A->N : A->pipe_row[issued + 1];  // the last chunk carries the sink entries
            NCME_CUDA(cudaMemcpyAsync(xd + r0, x_host + r0, (size_t)(r1 - r0) * 8, cudaMemcpyHostToDevice, ctx->h2d_stream));
            NCME_CUDA(cudaEventRecord(ctx->ev_h2d[issued], ctx->h2d_stream));
        }
        return NCME_OK;
    };
    MatvecArgs a;
    matvec_fill_args(A, coef, &a);
    a.xd = xd;
    a.x = xd;
    a.y = y_map ? y_map : yd;
    a.beta = 0.0;
    for (int c = 0; c < nc; ++c) {
        int need = c;
        while (need < nc - 1 && A->pipe_row[need + 1] < A->pipe_need_hi[c]) ++need;
        NCME_TRY(issue_uploads(std::min(nc - 1, need + 2)));             // keep the upload queue two chunks ahead
        NCME_CUDA(cudaStreamWaitEvent(st, ctx->ev_h2d[need], 0));
        MatvecArgs ac = a;
        ac.row_begin = A->pipe_row[c];
        ac.row_end = A->pipe_row[c + 1];
        ac.do_sinks = (c == nc - 1) ? 1 : 0;                             // sink rows read x everywhere: last
        static const bool trace_nokernel = getenv("NCME_HOST_PIPE_NOKERNEL") != nullptr;   // diagnostics: copies only
        if (!(pipe_trace && trace_nokernel)) NCME_TRY(matvec_launch(A, ac));
        if (y_map) continue;                                             // y already went to the host buffer
        NCME_CUDA(cudaEventRecord(ctx->ev_comp[c], st));
        NCME_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_comp[c], 0));
        const int64_t r0 = A->pipe_row[c];
        const int64_t r1 = (c == nc - 1) ? A->N : A->pipe_row[c + 1];
        NCME_CUDA(cudaMemcpyAsync(y_host + r0, yd + r0, (size_t)(r1 - r0) * 8, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        if (pipe_trace) NCME_CUDA(cudaEventRecord(trace_d2h[c], ctx->d2h_stream));
    }
    NCME_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
    NCME_CUDA(cudaStreamSynchronize(st));
    if (pipe_trace && !y_map && ++trace_calls == 5) {
        fprintf(stderr, "[ncme host pipe] chunk rows MB | upload done | kernel done | download done (ms after the call's start)\n");
        for (int c = 0; c < nc; ++c) {
            float a_ms = 0, b_ms = 0, c_ms = 0;
            cudaEventElapsedTime(&a_ms, ctx->ev_start, ctx->ev_h2d[c]);
            cudaEventElapsedTime(&b_ms, ctx->ev_start, ctx->ev_comp[c]);
            cudaEventElapsedTime(&c_ms, ctx->ev_start, trace_d2h[c]);
            const int64_t rows = A->pipe_row[c + 1] - A->pipe_row[c];
            fprintf(stderr, "[ncme host pipe] %2d %9lld %6.2f | %7.3f | %7.3f | %7.3f\n", c, (long long)rows, rows * 8e-6, a_ms, b_ms, c_ms);
        }
    }
    return NCME_OK;
}

int ncme_matrix_compression_info(ncme_matrix* A, int64_t info[4]) {
    NCME_REQUIRE(A && info, "null argument");
    NCME_TRY(matrix_ensure_compressed(A));
    info[0] = A->nchunks * A->nslots;   // (64-row chunk, slot) pairs
    info[1] = A->wide_chunks;           // of which kept on 32-bit column indices
    info[2] = A->use_c8;
    info[3] = A->nslots;
    return NCME_OK;
}

int ncme_matrix_stats(ncme_matrix* A, int* nterms, int64_t* nnz_per_term, int64_t* algorithmic_bytes, int64_t* device_bytes) {
    NCME_REQUIRE(A, "null matrix");
    if (nterms) *nterms = A->nterms;
    if (nnz_per_term)
        for (int k = 0; k < A->nterms; ++k) nnz_per_term[k] = A->nnz_term[k];
    if (algorithmic_bytes) *algorithmic_bytes = A->algorithmic_bytes;
    if (device_bytes)   // bytes one matvec actually streams with the compressed indices
        *device_bytes = (int64_t)A->ld * (9 * A->nslots + 8 * A->ndiag) + A->wide_chunks * 256 + 8 * A->nchunks * A->nslots +
                        12 * A->nsink + 16 * A->N;
    return NCME_OK;
}

}  // extern "C"
