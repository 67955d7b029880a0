// Exclusive prefix sum (uint32) used by the stream-compaction steps of expand!/deleteat!/assembly.
// Three-phase scan: per-block reduce -> recursive scan of block sums -> per-block scan + offset.
#include "common.cuh"

namespace ncme {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *block_total.
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T* block_total) {
    __shared__ T warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T incl = warp_incl_scan<T>(v);
    T excl = __shfl_up_sync(0xffffffffu, incl, 1);  // exclusive by shifting (no subtraction: exact for doubles)
    if (lane == 0) excl = T(0);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        T w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : T(0);
        T wi = warp_incl_scan<T>(w);
        T we = __shfl_up_sync(0xffffffffu, wi, 1);
        if (lane == 0) we = T(0);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = we;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) *block_total = wi;
    }
    __syncthreads();
    return excl + warp_sums[wid];
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const T* __restrict__ in, int64_t n,
                                                                T* __restrict__ block_sums) {
    __shared__ T total;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    block_excl_scan<T>(s, &total);
    __syncthreads();
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const T* in, int64_t n, const T* block_offsets, T* out) {
    __shared__ T total;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    T run = block_excl_scan<T>(s, &total) + (block_offsets ? block_offsets[blockIdx.x] : T(0));
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

size_t scan_scratch_elems(int64_t n) {
    size_t total = 0;
    int64_t m = n;
    while (m > SCAN_TILE) {
        m = (m + SCAN_TILE - 1) / SCAN_TILE;
        total += round_up<size_t>((size_t)m, 64);
    }
    return total + 64;
}

template <typename T>
static int scan_rec(ncme_ctx* ctx, const T* in, T* out, int64_t n, T* scratch) {
    if (n <= 0) return NCME_OK;
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        k_scan_down<T><<<1, SCAN_THREADS, 0, ctx->stream>>>(in, n, nullptr, out);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        return NCME_OK;
    }
    k_scan_reduce<T><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, n, scratch);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    NCME_TRY(scan_rec<T>(ctx, scratch, scratch, nb, scratch + round_up<size_t>((size_t)nb, 64)));
    k_scan_down<T><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, n, scratch, out);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int exclusive_scan_u32(ncme_ctx* ctx, const uint32_t* in_dev, uint32_t* out_dev, int64_t n, uint32_t* scratch_dev,
                       size_t scratch_elems, uint64_t* total_host) {
    if (n <= 0) {
        if (total_host) *total_host = 0;
        return NCME_OK;
    }
    NCME_REQUIRE(scratch_elems >= scan_scratch_elems(n), "scan scratch too small");
    uint32_t last_in = 0, last_out = 0;
    if (total_host)  // in may alias out: fetch the last input before scanning
        NCME_CUDA(cudaMemcpyAsync(&last_in, in_dev + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    NCME_TRY(scan_rec<uint32_t>(ctx, in_dev, out_dev, n, scratch_dev));
    if (total_host) {
        NCME_CUDA(cudaMemcpyAsync(&last_out, out_dev + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        NCME_CUDA(cudaStreamSynchronize(ctx->stream));
        *total_host = (uint64_t)last_in + (uint64_t)last_out;
    }
    return NCME_OK;
}

// exclusive prefix sum of doubles (prune: cumsum of the sorted probabilities); scratch_elems as above
int exclusive_scan_f64(ncme_ctx* ctx, const double* in_dev, double* out_dev, int64_t n, double* scratch_dev,
                       size_t scratch_elems) {
    if (n <= 0) return NCME_OK;
    NCME_REQUIRE(scratch_elems >= scan_scratch_elems(n), "scan scratch too small");
    return scan_rec<double>(ctx, in_dev, out_dev, n, scratch_dev);
}

}  // namespace ncme
