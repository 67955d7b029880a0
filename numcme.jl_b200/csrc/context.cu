// Context, error reporting, raw device memory helpers of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace ncme {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local int g_abort = 0;
bool abort_requested() { return g_abort != 0; }
void clear_abort() { g_abort = 0; }
int abort_status() {
    set_error("solve aborted by a host callback (ncme_request_abort)");
    return NCME_ERR_ABORTED;
}

}  // namespace ncme

using namespace ncme;

extern "C" {

int ncme_version(void) { return NCME_VERSION; }

void ncme_request_abort(void) { g_abort = 1; }

const char* ncme_last_error(void) { return g_err; }

int ncme_ctx_create(int device, ncme_ctx** out) {
    NCME_REQUIRE(out, "null out pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s); libncme has no CPU fallback", cudaGetErrorString(e));
        return NCME_ERR_CUDA;
    }
    NCME_REQUIRE(device >= 0 && device < ndev, "device %d out of range (0..%d)", device, ndev - 1);
    NCME_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NCME_CUDA(cudaGetDeviceProperties(&prop, device));
    {   // keep freed blocks of the stream-ordered pool (DevArray) cached instead of returning them to the driver
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaGetLastError();
    }
    ncme_ctx* ctx = new ncme_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    ctx->total_mem = prop.totalGlobalMem;
    ctx->cc = prop.major * 10 + prop.minor;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->red_partials, sizeof(double) * 4 * 4096) != cudaSuccess ||
        cudaMalloc(&ctx->red_counter, sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&ctx->red_result_dev, sizeof(double) * 1024) != cudaSuccess ||
        cudaMallocHost(&ctx->red_result_host, sizeof(double) * 1024) != cudaSuccess ||
        cudaMemset(ctx->red_counter, 0, sizeof(unsigned int)) != cudaSuccess) {
        set_error("context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        ncme_ctx_destroy(ctx);
        return NCME_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return NCME_OK;
}

int ncme_ctx_destroy(ncme_ctx* ctx) {
    if (!ctx) return NCME_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->red_partials) cudaFree(ctx->red_partials);
    if (ctx->red_counter) cudaFree(ctx->red_counter);
    if (ctx->red_result_dev) cudaFree(ctx->red_result_dev);
    if (ctx->red_result_host) cudaFreeHost(ctx->red_result_host);
    if (ctx->stage_dev_x) cudaFree(ctx->stage_dev_x);
    if (ctx->stage_dev_y) cudaFree(ctx->stage_dev_y);
    if (ctx->solve_ws) cudaFree(ctx->solve_ws);
    if (ctx->solve_full) cudaFree(ctx->solve_full);
    if (ctx->gm_partials) cudaFree(ctx->gm_partials);
    if (ctx->solve_pinned) cudaFreeHost(ctx->solve_pinned);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    for (int k = 0; k < 16; ++k) {
        if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]);
        if (ctx->ev_comp[k]) cudaEventDestroy(ctx->ev_comp[k]);
    }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return NCME_OK;
}

int ncme_ctx_set_stream(ncme_ctx* ctx, void* cuda_stream) {
    NCME_REQUIRE(ctx, "null context");
    NCME_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return NCME_OK;
}

int ncme_ctx_sync(ncme_ctx* ctx) {
    NCME_REQUIRE(ctx, "null context");
    NCME_CUDA(cudaStreamSynchronize(ctx->stream));
    return NCME_OK;
}

int ncme_ctx_device_info(ncme_ctx* ctx, int64_t info[4]) {
    NCME_REQUIRE(ctx && info, "null argument");
    info[0] = ctx->sm_count;
    info[1] = (int64_t)ctx->l2_bytes;
    info[2] = (int64_t)ctx->total_mem;
    info[3] = ctx->cc;
    return NCME_OK;
}

int ncme_ctx_launch_count(ncme_ctx* ctx, int64_t* count) {
    NCME_REQUIRE(ctx && count, "null argument");
    *count = ctx->launches;
    return NCME_OK;
}

int ncme_dmalloc(ncme_ctx* ctx, size_t bytes, void** dptr) {
    NCME_REQUIRE(ctx && dptr, "null argument");
    NCME_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 8);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return NCME_ERR_NOMEM;
    }
    return NCME_OK;
}

int ncme_dfree(ncme_ctx* ctx, void* dptr) {
    NCME_REQUIRE(ctx, "null context");
    if (dptr) {
        NCME_CUDA(cudaStreamSynchronize(ctx->stream));
        NCME_CUDA(cudaFree(dptr));
    }
    return NCME_OK;
}

int ncme_h2d(ncme_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    NCME_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_host)), "null argument");
    NCME_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(ctx->stream));
    return NCME_OK;
}

int ncme_d2h(ncme_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    NCME_REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), "null argument");
    NCME_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(ctx->stream));
    return NCME_OK;
}

int ncme_host_alloc(size_t bytes, void** hptr) {
    NCME_REQUIRE(hptr, "null argument");
    cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 8);
    if (e != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return NCME_ERR_NOMEM;
    }
    return NCME_OK;
}

int ncme_host_free(void* hptr) {
    if (hptr) NCME_CUDA(cudaFreeHost(hptr));
    return NCME_OK;
}

}  // extern "C"
