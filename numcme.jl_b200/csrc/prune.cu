// K5 (prune rule): the dropstates branch of adapt!  -- src/transientcme/sparse/rstepadapters.jl:40-46, :93-99
//   pids = sortperm(p); dropcount = sum((sum(p) .- cumsum(p[pids])) .>= threshold); deleteat!(space, sort(pids[1:dropcount]))
// On the device: bitonic sort of (p_i, i) pairs (ties broken by index == Julia's stable sortperm),
// exclusive scan of the sorted values, count of the positions that satisfy the inequality, keep-flag
// scatter, then the shared deletion path of space.cu.
#include <math.h>

#include "space.cuh"
#include "vec.cuh"

namespace ncme {

constexpr int BS_THREADS = 512;
constexpr int BS_TILE = 2 * BS_THREADS;  // elements sorted in shared memory by one CTA

__device__ __forceinline__ bool pair_less(double ka, uint32_t ia, double kb, uint32_t ib) {
    return ka < kb || (ka == kb && ia < ib);
}

__device__ __forceinline__ void cmp_swap(double& ka, uint32_t& ia, double& kb, uint32_t& ib, bool ascending) {
    const bool sw = ascending ? pair_less(kb, ib, ka, ia) : pair_less(ka, ia, kb, ib);
    if (sw) {
        double tk = ka;
        ka = kb;
        kb = tk;
        uint32_t ti = ia;
        ia = ib;
        ib = ti;
    }
}

__global__ void k_sort_init(const double* __restrict__ p, int64_t n, int64_t n2, double* __restrict__ key,
                            uint32_t* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    key[i] = i < n ? p[i] : INFINITY;
    idx[i] = i < n ? (uint32_t)i : NONE32;
}

// All bitonic stages with k <= BS_TILE, entirely in shared memory.
__global__ void __launch_bounds__(BS_THREADS) k_bitonic_local_sort(double* __restrict__ key, uint32_t* __restrict__ idx) {
    __shared__ double sk[BS_TILE];
    __shared__ uint32_t si[BS_TILE];
    const int64_t base = (int64_t)blockIdx.x * BS_TILE;
    for (int q = threadIdx.x; q < BS_TILE; q += BS_THREADS) {
        sk[q] = key[base + q];
        si[q] = idx[base + q];
    }
    __syncthreads();
    for (int k = 2; k <= BS_TILE; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int t = threadIdx.x;
            const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const int hi = lo | j;
            const bool asc = (((base + lo) & k) == 0);
            cmp_swap(sk[lo], si[lo], sk[hi], si[hi], asc);
            __syncthreads();
        }
    }
    for (int q = threadIdx.x; q < BS_TILE; q += BS_THREADS) {
        key[base + q] = sk[q];
        idx[base + q] = si[q];
    }
}

// One global compare-exchange pass (j >= BS_TILE).
__global__ void k_bitonic_global(double* __restrict__ key, uint32_t* __restrict__ idx, int64_t n2, int64_t k, int64_t j) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n2 / 2) return;
    const int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const int64_t hi = lo | j;
    const bool asc = ((lo & k) == 0);
    double ka = key[lo], kb = key[hi];
    uint32_t ia = idx[lo], ib = idx[hi];
    const double ka0 = ka;
    const uint32_t ia0 = ia;
    cmp_swap(ka, ia, kb, ib, asc);
    if (ka != ka0 || ia != ia0) {
        key[lo] = ka;
        idx[lo] = ia;
        key[hi] = kb;
        idx[hi] = ib;
    }
}

// The passes j = BS_TILE/2 .. 1 of merge stage k, in shared memory.
__global__ void __launch_bounds__(BS_THREADS) k_bitonic_local_merge(double* __restrict__ key, uint32_t* __restrict__ idx,
                                                                    int64_t k) {
    __shared__ double sk[BS_TILE];
    __shared__ uint32_t si[BS_TILE];
    const int64_t base = (int64_t)blockIdx.x * BS_TILE;
    for (int q = threadIdx.x; q < BS_TILE; q += BS_THREADS) {
        sk[q] = key[base + q];
        si[q] = idx[base + q];
    }
    __syncthreads();
    const bool asc = ((base & k) == 0);
    for (int j = BS_TILE >> 1; j > 0; j >>= 1) {
        const int t = threadIdx.x;
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        cmp_swap(sk[lo], si[lo], sk[hi], si[hi], asc);
        __syncthreads();
    }
    for (int q = threadIdx.x; q < BS_TILE; q += BS_THREADS) {
        key[base + q] = sk[q];
        idx[base + q] = si[q];
    }
}

// count of k in [0,n) with total - (excl[k] + key[k]) >= thr  (or > thr when strict)
__global__ void k_count_tail(const double* __restrict__ key, const double* __restrict__ excl, int64_t n, double total,
                             double thr, int strict, unsigned long long* count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < n) {
        const double tail = total - (excl[i] + key[i]);
        ok = strict ? (tail > thr) : (tail >= thr);
    }
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}

__global__ void k_flag_dropped(const uint32_t* __restrict__ sorted_idx, int64_t dropcount, uint32_t* __restrict__ keep) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < dropcount) keep[sorted_idx[k]] = 0u;
}

__global__ void k_fill_keep(uint32_t* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1u;
}

static inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace ncme

using namespace ncme;

extern "C" int ncme_space_prune_by_mass(ncme_space* sp, const double* p_dev, double threshold, int strict,
                                        int64_t* dropcount) {
    NCME_RANGE("ncme_space_prune_by_mass");
    NCME_REQUIRE(sp && p_dev && dropcount, "null argument");
    *dropcount = 0;
    const int64_t n = sp->n;
    if (n == 0) return NCME_OK;
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t s = ctx->stream;
    int64_t n2 = BS_TILE;
    while (n2 < n) n2 <<= 1;
    DevArray<double> key, excl, scratch;
    DevArray<uint32_t> idx;
    unsigned long long* cnt = nullptr;
    int rc = NCME_OK;
    do {
        if ((rc = key.reserve((size_t)n2, s, false)) != NCME_OK) break;
        if ((rc = idx.reserve((size_t)n2, s, false)) != NCME_OK) break;
        if ((rc = excl.reserve((size_t)n, s, false)) != NCME_OK) break;
        if ((rc = scratch.reserve(scan_scratch_elems(n), s, false)) != NCME_OK) break;
        cnt = reinterpret_cast<unsigned long long*>(ctx->red_result_dev + 900);   // context scratch scalar
        cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), s);
        k_sort_init<<<nblk(n2), 256, 0, s>>>(p_dev, n, n2, key.p, idx.p);
        k_bitonic_local_sort<<<(unsigned)(n2 / BS_TILE), BS_THREADS, 0, s>>>(key.p, idx.p);
        ctx->launches += 2;
        for (int64_t k = 2 * BS_TILE; k <= n2; k <<= 1) {
            for (int64_t j = k >> 1; j >= BS_TILE; j >>= 1) {
                k_bitonic_global<<<nblk(n2 / 2), 256, 0, s>>>(key.p, idx.p, n2, k, j);
                ctx->launches++;
            }
            k_bitonic_local_merge<<<(unsigned)(n2 / BS_TILE), BS_THREADS, 0, s>>>(key.p, idx.p, k);
            ctx->launches++;
        }
        double total = 0.0;
        if ((rc = vec_sum(ctx, n, p_dev, &total)) != NCME_OK) break;
        if ((rc = exclusive_scan_f64(ctx, key.p, excl.p, n, scratch.p, scratch.cap)) != NCME_OK) break;
        k_count_tail<<<nblk(n), 256, 0, s>>>(key.p, excl.p, n, total, threshold, strict, cnt);
        ctx->launches++;
        unsigned long long hc = 0;
        if (cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            set_error("prune: %s", cudaGetErrorString(cudaGetLastError()));
            rc = NCME_ERR_CUDA;
            break;
        }
        *dropcount = (int64_t)hc;
        if ((rc = sp->flags.reserve((size_t)n, s, false)) != NCME_OK) break;
        k_fill_keep<<<nblk(n), 256, 0, s>>>(sp->flags.p, n);
        ctx->launches++;
        if (hc > 0) {
            k_flag_dropped<<<nblk((int64_t)hc), 256, 0, s>>>(idx.p, (int64_t)hc, sp->flags.p);
            ctx->launches++;
        }
        if (cudaStreamSynchronize(s) != cudaSuccess) {
            set_error("prune: %s", cudaGetErrorString(cudaGetLastError()));
            rc = NCME_ERR_CUDA;
            break;
        }
        rc = space_delete_flagged(sp);
    } while (0);
    key.release();
    idx.release();
    excl.release();
    scratch.release();
    return rc;
}
