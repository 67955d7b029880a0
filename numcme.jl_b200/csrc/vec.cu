// K7: the device-resident vector operations an ODE integrator performs on the FSP vector
// (reference: the integrator internals driven from src/transientcme/sparse/fspsolve.jl:158-161).
// All are single-pass, HBM-bound, vectorised; reductions are deterministic (fixed tree + ordered
// last-block combine, no floating-point atomics).
#include "vec.cuh"

namespace ncme {

constexpr int VT = 256;
constexpr int RED_MAX_BLOCKS = 4096;

static inline unsigned ew_grid(int64_t n, int per_thread) {
    int64_t b = (n + (int64_t)VT * per_thread - 1) / ((int64_t)VT * per_thread);
    return (unsigned)(b < 1 ? 1 : b);
}

__global__ void __launch_bounds__(VT) k_fill(int64_t n, double a, double* __restrict__ x) {
    int64_t i = ((int64_t)blockIdx.x * VT + threadIdx.x) * 2;
    if (i + 1 < n && ((uintptr_t)x & 15) == 0) {
        *reinterpret_cast<double2*>(x + i) = make_double2(a, a);
    } else {
        if (i < n) x[i] = a;
        if (i + 1 < n) x[i + 1] = a;
    }
}

__global__ void __launch_bounds__(VT) k_scale(int64_t n, double a, double* __restrict__ x) {
    int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x;
    if (i < n) x[i] *= a;
}

__global__ void __launch_bounds__(VT) k_axpy(int64_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x;
    if (i < n) y[i] = fma(a, x[i], y[i]);
}

struct LincombArgs {
    int k;
    double c[8];
    const double* x[8];
};

template <int K>
__global__ void __launch_bounds__(VT) k_lincomb(int64_t n, const __grid_constant__ LincombArgs a, double* out) {
    int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) s = fma(a.c[j], a.x[j][i], s);
    out[i] = s;
}

// ---- reductions ------------------------------------------------------------------------------
enum RedOp { RED_SUM = 0, RED_DOT = 1, RED_WRMS = 2, RED_NONFINITE = 3 };

struct RedArgs {
    int64_t n;
    const double* x;
    const double* y;
    const double* z;
    double atol, rtol;
    double* partials;
    unsigned int* counter;
    double* result;
};

template <int OP>
__device__ __forceinline__ double red_term(const RedArgs& a, int64_t i) {
    if (OP == RED_SUM) return a.x[i];
    if (OP == RED_DOT) return a.x[i] * a.y[i];
    if (OP == RED_WRMS) {
        const double w = a.atol + a.rtol * fmax(fabs(a.y[i]), fabs(a.z[i]));
        const double q = a.x[i] / w;
        return q * q;
    }
    return isfinite(a.x[i]) ? 0.0 : 1.0;
}

template <int OP>
__global__ void __launch_bounds__(VT) k_reduce(const __grid_constant__ RedArgs a) {
    __shared__ double wsum[VT / 32];
    __shared__ bool is_last;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * VT) s += red_term<OP>(a, i);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < VT / 32; ++w) t += wsum[w];
        a.partials[blockIdx.x] = t;
        __threadfence();
        is_last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ordered combine by one warp: lane l sums partials l, l+32, ... then a fixed shuffle tree
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
}

template <int OP>
static int reduce_launch(ncme_ctx* ctx, RedArgs a, double* out_host) {
    a.partials = ctx->red_partials;
    a.counter = ctx->red_counter;
    a.result = ctx->red_result_dev;
    if (a.n <= 0) {
        *out_host = 0.0;
        return NCME_OK;
    }
    int64_t b = (a.n + (int64_t)VT * 8 - 1) / ((int64_t)VT * 8);
    if (b > RED_MAX_BLOCKS) b = RED_MAX_BLOCKS;
    if (b < 1) b = 1;
    k_reduce<OP><<<(unsigned)b, VT, 0, ctx->stream>>>(a);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    NCME_CUDA(cudaMemcpyAsync(ctx->red_result_host, ctx->red_result_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_host = ctx->red_result_host[0];
    return NCME_OK;
}

int vec_sum(ncme_ctx* ctx, int64_t n, const double* x, double* out) {
    RedArgs a{};
    a.n = n;
    a.x = x;
    return reduce_launch<RED_SUM>(ctx, a, out);
}

}  // namespace ncme

using namespace ncme;

extern "C" {

int ncme_vec_fill(ncme_ctx* ctx, int64_t n, double a, double* x) {
    NCME_REQUIRE(ctx && (n == 0 || x) && n >= 0, "bad arguments");
    if (n == 0) return NCME_OK;
    k_fill<<<ew_grid(n, 2), VT, 0, ctx->stream>>>(n, a, x);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int ncme_vec_copy(ncme_ctx* ctx, int64_t n, const double* x, double* y) {
    NCME_REQUIRE(ctx && n >= 0 && (n == 0 || (x && y)), "bad arguments");
    if (n) NCME_CUDA(cudaMemcpyAsync(y, x, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return NCME_OK;
}

int ncme_vec_scale(ncme_ctx* ctx, int64_t n, double a, double* x) {
    NCME_REQUIRE(ctx && n >= 0 && (n == 0 || x), "bad arguments");
    if (n == 0) return NCME_OK;
    k_scale<<<ew_grid(n, 1), VT, 0, ctx->stream>>>(n, a, x);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

// out_i = x_i / (atol + rtol max(|u0_i|, |u1_i|)) : OrdinaryDiffEq's calculate_residuals!(out, x, u0, u1, atol, rtol, ...)
__global__ void k_residuals(int64_t n, const double* __restrict__ x, const double* __restrict__ u0,
                            const double* __restrict__ u1, double atol, double rtol, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = x[i] / (atol + rtol * fmax(fabs(u0[i]), fabs(u1[i])));
}
// x_i += a  (the constant term of a broadcast expression)
__global__ void k_shift(int64_t n, double a, double* __restrict__ x) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] += a;
}

int ncme_vec_residuals(ncme_ctx* ctx, int64_t n, const double* x, const double* u0, const double* u1, double atol,
                       double rtol, double* out) {
    NCME_REQUIRE(ctx && n >= 0 && (n == 0 || (x && u0 && u1 && out)), "bad arguments");
    if (n == 0) return NCME_OK;
    k_residuals<<<ew_grid(n, 1), VT, 0, ctx->stream>>>(n, x, u0, u1, atol, rtol, out);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int ncme_vec_shift(ncme_ctx* ctx, int64_t n, double a, double* x) {
    NCME_REQUIRE(ctx && n >= 0 && (n == 0 || x), "bad arguments");
    if (n == 0) return NCME_OK;
    k_shift<<<ew_grid(n, 1), VT, 0, ctx->stream>>>(n, a, x);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int ncme_vec_axpy(ncme_ctx* ctx, int64_t n, double a, const double* x, double* y) {
    NCME_REQUIRE(ctx && n >= 0 && (n == 0 || (x && y)), "bad arguments");
    if (n == 0) return NCME_OK;
    k_axpy<<<ew_grid(n, 1), VT, 0, ctx->stream>>>(n, a, x, y);
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int ncme_vec_lincomb(ncme_ctx* ctx, int64_t n, int k, const double* coefs, const double* const* xs, double* out) {
    NCME_REQUIRE(ctx && n >= 0 && k >= 1 && k <= 8 && coefs && xs && (n == 0 || out), "bad arguments (k must be 1..8)");
    if (n == 0) return NCME_OK;
    LincombArgs a;
    a.k = k;
    for (int j = 0; j < 8; ++j) {
        a.c[j] = j < k ? coefs[j] : 0.0;
        a.x[j] = j < k ? xs[j] : nullptr;
    }
    const unsigned g = ew_grid(n, 1);
    switch (k) {
#define NCME_LC(K)                                          \
    case K:                                                 \
        k_lincomb<K><<<g, VT, 0, ctx->stream>>>(n, a, out); \
        break;
        NCME_LC(1) NCME_LC(2) NCME_LC(3) NCME_LC(4) NCME_LC(5) NCME_LC(6) NCME_LC(7) NCME_LC(8)
#undef NCME_LC
    }
    ctx->launches++;
    NCME_CUDA(cudaGetLastError());
    return NCME_OK;
}

int ncme_vec_sum(ncme_ctx* ctx, int64_t n, const double* x, double* out) {
    NCME_REQUIRE(ctx && out && n >= 0 && (n == 0 || x), "bad arguments");
    return vec_sum(ctx, n, x, out);
}

int ncme_vec_dot(ncme_ctx* ctx, int64_t n, const double* x, const double* y, double* out) {
    NCME_REQUIRE(ctx && out && n >= 0 && (n == 0 || (x && y)), "bad arguments");
    RedArgs a{};
    a.n = n;
    a.x = x;
    a.y = y;
    return reduce_launch<RED_DOT>(ctx, a, out);
}

int ncme_vec_wrms(ncme_ctx* ctx, int64_t n, const double* x, const double* u0, const double* u1, double atol, double rtol,
                  double* out) {
    NCME_REQUIRE(ctx && out && n > 0 && x && u0 && u1, "bad arguments");
    RedArgs a{};
    a.n = n;
    a.x = x;
    a.y = u0;
    a.z = u1;
    a.atol = atol;
    a.rtol = rtol;
    double s = 0.0;
    NCME_TRY(reduce_launch<RED_WRMS>(ctx, a, &s));
    *out = sqrt(s / (double)n);
    return NCME_OK;
}

int ncme_vec_any_nonfinite(ncme_ctx* ctx, int64_t n, const double* x, int* out) {
    NCME_REQUIRE(ctx && out && n >= 0 && (n == 0 || x), "bad arguments");
    RedArgs a{};
    a.n = n;
    a.x = x;
    double s = 0.0;
    NCME_TRY(reduce_launch<RED_NONFINITE>(ctx, a, &s));
    *out = s > 0.0 ? 1 : 0;
    return NCME_OK;
}

}  // extern "C"
