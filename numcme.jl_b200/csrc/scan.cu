// Exclusive prefix sum (uint32) used by the stream-compaction steps of expand!/deleteat!/assembly.
// Three-phase scan: per-block reduce -> recursive scan of block sums -> per-block scan + offset.
#include "common.cuh"

namespace ncme {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *block_total.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* block_total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) *block_total = wi;
    }
    __syncthreads();
    return incl - v + warp_sums[wid];
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, int64_t n,
                                                                uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t total;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    block_excl_scan(s, &total);
    __syncthreads();
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const uint32_t* __restrict__ in, int64_t n,
                                                              const uint32_t* __restrict__ block_offsets,
                                                              uint32_t* __restrict__ out) {
    __shared__ uint32_t total;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t run = block_excl_scan(s, &total) + (block_offsets ? block_offsets[blockIdx.x] : 0);
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

static int scan_rec(ncme_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch) {
    if (n <= 0) return NCME_OK;
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        k_scan_down<<<1, SCAN_THREADS, 0, ctx->stream>>>(in, n, nullptr, out);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        return NCME_OK;
    }
    k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, n, scratch);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    NCME_TRY(scan_rec(ctx, scratch, scratch, nb, scratch + round_up<size_t>((size_t)nb, 64)));
    k_scan_down<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, n, scratch, out);
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
    NCME_TRY(scan_rec(ctx, in_dev, out_dev, n, scratch_dev));
    if (total_host) {
        NCME_CUDA(cudaMemcpyAsync(&last_out, out_dev + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        NCME_CUDA(cudaStreamSynchronize(ctx->stream));
        *total_host = (uint64_t)last_in + (uint64_t)last_out;
    }
    return NCME_OK;
}

}  // namespace ncme
