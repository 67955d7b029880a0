// Internal declarations shared by the libncme translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "ncme.h"

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: a no-op (one predictable branch) unless a profiler injects itself

#include <chrono>
namespace ncme {
// NVTX range around an ABI entry point (named ranges for nsys / ncu --nvtx): NCME_RANGE("ncme_matvec");
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define NCME_RANGE(name) ncme::NvtxRange _ncme_nvtx_range(name)

inline double wall_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void set_error(const char* fmt, ...);
bool abort_requested();   // a host callback called ncme_request_abort() (context.cu)
void clear_abort();
int abort_status();        // sets the message, returns NCME_ERR_ABORTED

#define NCME_CUDA(expr)                                                                           \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ncme::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return NCME_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

#define NCME_TRY(expr)              \
    do {                            \
        int _s = (expr);            \
        if (_s != NCME_OK) return _s; \
    } while (0)

#define NCME_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            ncme::set_error(__VA_ARGS__); \
            return NCME_ERR_ARG;         \
        }                                \
    } while (0)

constexpr uint32_t NONE32 = 0xFFFFFFFFu;
constexpr uint64_t EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
// one bit per reaction (sink_connectivity of a state, sets of reactions): NCME_MAX_REACTIONS <= 64
typedef unsigned long long smask_t;
#define SMASK1(r) ((ncme::smask_t)1 << (r))
static_assert(NCME_MAX_REACTIONS <= 64, "sink masks hold one bit per reaction");

template <typename T>
static inline T round_up(T a, T b) {
    return (a + b - 1) / b * b;
}

}  // namespace ncme

struct ncme_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    size_t l2_bytes = 0;
    size_t total_mem = 0;
    int cc = 100;
    int64_t launches = 0;
    // small device scratch for reductions (partials + counters), and a pinned host mirror
    double* red_partials = nullptr;   // [RED_MAX_BLOCKS * 4]
    unsigned int* red_counter = nullptr;
    double* red_result_dev = nullptr; // [1024]
    double* red_result_host = nullptr;  // pinned [1024]
    // pinned staging for the host-buffer matvec path
    double* stage_host = nullptr;
    size_t stage_host_bytes = 0;
    double* stage_dev_x = nullptr;
    double* stage_dev_y = nullptr;
    size_t stage_dev_bytes = 0;
    // host-buffer matvec pipeline: copy streams + per-chunk events
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_h2d[16] = {nullptr}, ev_comp[16] = {nullptr}, ev_start = nullptr;
    // grow-only caches of the native integrator (re-used by every segment of a solve)
    double* solve_ws = nullptr;
    size_t solve_ws_bytes = 0;
    double* solve_pinned = nullptr;
    size_t solve_pinned_bytes = 0;
    double* solve_full = nullptr;
    size_t solve_full_bytes = 0;
    double* gm_partials = nullptr;      // per-CTA partial sums of the fused Krylov inner products
    size_t gm_partials_bytes = 0;
};

namespace ncme {

// Device array with capacity growth (contents preserved).  Storage comes from the device's stream-ordered memory
// pool (cudaMallocAsync; release threshold raised by ncme_ctx_create so freed blocks are re-used without returning to
// the driver): the per-adapt rebuilds of spaces and matrices (fspsolve.jl:175-176) make dozens of small allocations,
// and cudaMalloc/cudaFree (implicit device synchronisation) dominated the small example configurations.
template <typename T>
struct DevArray {
    T* p = nullptr;
    size_t cap = 0;
    cudaStream_t last_stream = nullptr;   // stream the storage was last (re)allocated on; frees are ordered on it
    int reserve(size_t n, cudaStream_t s, bool keep = true) {
        if (n <= cap) return NCME_OK;
        size_t ncap = cap ? cap : 256;
        while (ncap < n) ncap = ncap + ncap / 2 + 256;
        T* q = nullptr;
        cudaError_t e = cudaMallocAsync((void**)&q, ncap * sizeof(T), s);
        if (e != cudaSuccess) {
            set_error("cudaMallocAsync(%zu bytes) failed: %s", ncap * sizeof(T), cudaGetErrorString(e));
            return NCME_ERR_NOMEM;
        }
        if (p && keep && cap) {
            e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) {
                set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
                return NCME_ERR_CUDA;
            }
        }
        if (p) {
            if (last_stream != s) cudaStreamSynchronize(last_stream);   // work queued on the old stream may still use p
            cudaFreeAsync(p, s);
        }
        p = q;
        cap = ncap;
        last_stream = s;
        return NCME_OK;
    }
    void release() {
        if (p) cudaFreeAsync(p, last_stream);
        p = nullptr;
        cap = 0;
    }
};

// ---- scan / compaction primitives (scan.cu)
// exclusive prefix sum of n uint32 flags into out (uint32), total returned on host (synchronises).
int exclusive_scan_u32(ncme_ctx* ctx, const uint32_t* in_dev, uint32_t* out_dev, int64_t n, uint32_t* scratch_dev,
                       size_t scratch_elems, uint64_t* total_host);
size_t scan_scratch_elems(int64_t n);
int exclusive_scan_f64(ncme_ctx* ctx, const double* in_dev, double* out_dev, int64_t n, double* scratch_dev,
                       size_t scratch_elems);

}  // namespace ncme
