// NCCL communicator of libncme (K8): creation from a unique id distributed by the host language
// (torch.distributed / MPI / Julia Distributed), scalar all-reduce, all-gather-v.
#include "comm.cuh"
#include "hostreduce.h"
static_assert(ncme::HR_MAX_RANKS >= NCME_MAX_RANKS && ncme::HR_MAX_VALUES >= NCME_HOSTREDUCE_MAX, "hostreduce.h limits");

#include <dlfcn.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>

namespace ncme {

static NcclApi g_api;
static bool g_api_ok = false;

const NcclApi* nccl_api() {
    if (g_api_ok) return &g_api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return nullptr;
    }
#define NCME_SYM(field, name)                                       \
    *(void**)(&g_api.field) = dlsym(h, name);                       \
    if (!g_api.field) {                                             \
        set_error("libnccl: missing symbol %s", name);              \
        return nullptr;                                             \
    }
    NCME_SYM(GetUniqueId, "ncclGetUniqueId")
    NCME_SYM(CommInitRank, "ncclCommInitRank")
    NCME_SYM(CommDestroy, "ncclCommDestroy")
    NCME_SYM(GetErrorString, "ncclGetErrorString")
    NCME_SYM(AllReduce, "ncclAllReduce")
    NCME_SYM(AllGather, "ncclAllGather")
    NCME_SYM(Broadcast, "ncclBroadcast")
    NCME_SYM(Send, "ncclSend")
    NCME_SYM(Recv, "ncclRecv")
    NCME_SYM(GroupStart, "ncclGroupStart")
    NCME_SYM(GroupEnd, "ncclGroupEnd")
#undef NCME_SYM
    g_api_ok = true;
    return &g_api;
}

int comm_allreduce_sum(ncme_comm* c, double* buf, size_t count, cudaStream_t st) {
    if (!c || c->nranks == 1 || count == 0) return NCME_OK;
    NCME_NCCL(nccl_api()->AllReduce(buf, buf, count, ncclDouble, ncclSum, c->nccl, st));
    return NCME_OK;
}

int comm_allgatherv(ncme_comm* c, const double* send, double* recv, const int64_t* counts, const int64_t* displs,
                    cudaStream_t st) {
    if (!c || c->nranks == 1) {
        if (send != recv + displs[0])
            NCME_CUDA(cudaMemcpyAsync(recv + displs[0], send, (size_t)counts[0] * 8, cudaMemcpyDeviceToDevice, st));
        return NCME_OK;
    }
    const NcclApi* api = nccl_api();
    NCME_NCCL(api->GroupStart());
    for (int r = 0; r < c->nranks; ++r) {
        // every rank broadcasts its slice: grouped broadcasts == all-gather-v
        const void* src = (r == c->rank) ? (const void*)send : (const void*)(recv + displs[r]);
        ncclResult_t rc = api->Broadcast(src, recv + displs[r], (size_t)counts[r], ncclDouble, r, c->nccl, st);
        if (rc != ncclSuccess) {
            api->GroupEnd();
            set_error("ncclBroadcast failed: %s", api->GetErrorString(rc));
            return NCME_ERR_COMM;
        }
    }
    NCME_NCCL(api->GroupEnd());
    return NCME_OK;
}

// byte all-gather through the device scratch (handles, layout descriptors)
static int allgather_bytes(ncme_comm* c, const void* mine, size_t nbytes, std::vector<char>* all) {
    all->resize(nbytes * c->nranks);
    if (c->nranks == 1) {
        memcpy(all->data(), mine, nbytes);
        return NCME_OK;
    }
    NCME_REQUIRE(nbytes * c->nranks <= 4096 * sizeof(double), "allgather_bytes: message too large");
    cudaStream_t st = c->ctx->stream;
    char* dev = (char*)c->scratch;
    NCME_CUDA(cudaMemcpyAsync(dev + nbytes * c->rank, mine, nbytes, cudaMemcpyHostToDevice, st));
    NCME_NCCL(nccl_api()->AllGather(dev + nbytes * c->rank, dev, nbytes, ncclChar, c->nccl, st));
    NCME_CUDA(cudaMemcpyAsync(all->data(), dev, nbytes * c->nranks, cudaMemcpyDeviceToHost, st));
    NCME_CUDA(cudaStreamSynchronize(st));
    return NCME_OK;
}

struct RegMsg {
    cudaIpcMemHandle_t handle;
    int64_t local0, stride, nvec;
    int64_t valid;
};

int comm_register(ncme_comm* c, void* base, size_t bytes, int64_t local0, int64_t stride, int64_t nvec, const int* peers,
                  int npeers) {
    if (!c || c->nranks == 1 || !c->p2p_ok) return NCME_OK;
    for (auto& r : c->regs)
        if (r.base == base) return NCME_OK;
    RegMsg mine;
    memset(&mine, 0, sizeof(mine));
    cudaError_t e = cudaIpcGetMemHandle(&mine.handle, base);
    mine.valid = (e == cudaSuccess) ? 1 : 0;
    if (e != cudaSuccess) cudaGetLastError();
    mine.local0 = local0;
    mine.stride = stride;
    mine.nvec = nvec;
    std::vector<char> all;
    NCME_TRY(allgather_bytes(c, &mine, sizeof(mine), &all));
    const RegMsg* msgs = (const RegMsg*)all.data();
    bool ok = true;
    for (int q = 0; q < c->nranks; ++q) ok &= msgs[q].valid != 0;
    if (!ok) return NCME_OK;   // some rank could not export: nobody registers, matvecs stay on NCCL
    RegBuf rb;
    rb.base = base;
    rb.bytes = bytes;
    rb.local0 = local0;
    rb.stride = stride;
    rb.nvec = nvec;
    for (int k = 0; k < npeers; ++k) {
        const int q = peers[k];
        if (q < 0 || q >= c->nranks || q == c->rank || rb.peer_base[q]) continue;
        void* pb = nullptr;
        e = cudaIpcOpenMemHandle(&pb, msgs[q].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e));
            cudaGetLastError();
            return NCME_ERR_COMM;
        }
        rb.peer_base[q] = pb;
        rb.peer_local0[q] = msgs[q].local0;
        rb.peer_stride[q] = msgs[q].stride;
    }
    c->regs.push_back(rb);
    return NCME_OK;
}

int comm_unregister(ncme_comm* c, void* base) {
    if (!c || c->nranks == 1 || !c->p2p_ok) return NCME_OK;
    for (size_t k = 0; k < c->regs.size(); ++k) {
        if (c->regs[k].base != base) continue;
        NCME_CUDA(cudaStreamSynchronize(c->ctx->stream));
        for (int q = 0; q < c->nranks; ++q)
            if (c->regs[k].peer_base[q]) cudaIpcCloseMemHandle(c->regs[k].peer_base[q]);
        c->regs.erase(c->regs.begin() + k);
        // nobody may free before every rank has unmapped: a tiny all-reduce is the barrier
        NCME_TRY(comm_allreduce_sum(c, c->scratch, 1, c->ctx->stream));
        NCME_CUDA(cudaStreamSynchronize(c->ctx->stream));
        return NCME_OK;
    }
    return NCME_OK;
}

int comm_workspace(ncme_comm* c, size_t bytes, int64_t local0, int64_t stride, int64_t nvec, const int* peers, int npeers,
                   double** out) {
    cudaStream_t st = c->ctx->stream;
    // layout or size changed anywhere => everybody re-allocates (keeps the registration symmetric)
    const bool same = c->ws_base && c->ws_bytes >= bytes && c->ws_local0 == local0 && c->ws_stride == stride &&
                      c->ws_nvec == nvec;
    double v = same ? 0.0 : 1.0;
    NCME_CUDA(cudaMemcpyAsync(c->scratch, &v, sizeof(double), cudaMemcpyHostToDevice, st));
    NCME_TRY(comm_allreduce_sum(c, c->scratch, 1, st));
    NCME_CUDA(cudaMemcpyAsync(&v, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
    NCME_CUDA(cudaStreamSynchronize(st));
    if (v != 0.0) {
        if (c->ws_base) {
            NCME_TRY(comm_unregister(c, c->ws_base));
            cudaFree(c->ws_base);
            c->ws_base = nullptr;
            c->ws_bytes = 0;
        }
        const size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&c->ws_base, want);
        if (e != cudaSuccess) {
            set_error("integrator workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            return NCME_ERR_NOMEM;
        }
        c->ws_bytes = want;
        c->ws_local0 = local0;
        c->ws_stride = stride;
        c->ws_nvec = nvec;
        NCME_CUDA(cudaMemsetAsync(c->ws_base, 0, want, st));
        NCME_TRY(comm_register(c, c->ws_base, want, local0, stride, nvec, peers, npeers));
    }
    *out = c->ws_base;
    return NCME_OK;
}

const double* comm_peer_vector(const ncme_comm* c, const double* x_local, int q) {
    for (const auto& r : c->regs) {
        const char* b = (const char*)r.base;
        const char* x = (const char*)x_local;
        if (x < b || x >= b + r.bytes || !r.peer_base[q]) continue;
        const int64_t off = (int64_t)((x - b) / 8) - r.local0;
        if (off < 0) continue;
        int64_t v = 0;
        if (r.stride > 0) {
            if (off % r.stride != 0) continue;
            v = off / r.stride;
        } else if (off != 0) {
            continue;
        }
        if (v >= r.nvec) continue;
        return (const double*)r.peer_base[q] + r.peer_local0[q] + v * r.peer_stride[q];
    }
    return nullptr;
}

// ---- host-side scalar all-reduce over POSIX shared memory (protocol: hostreduce.h) ------------------------------
static int comm_setup_hostreduce(ncme_comm* c) {
    c->hr = nullptr;
    if (c->nranks == 1 || c->nranks > NCME_MAX_RANKS || getenv("NCME_NO_HOSTREDUCE")) return NCME_OK;
    struct {
        char name[64];
        char host[64];
    } mine;
    memset(&mine, 0, sizeof(mine));
    gethostname(mine.host, sizeof(mine.host) - 1);
    if (c->rank == 0) snprintf(mine.name, sizeof(mine.name), "/ncme_hr_%d_%p", (int)getpid(), (void*)c);
    std::vector<char> all;
    NCME_TRY(allgather_bytes(c, &mine, sizeof(mine), &all));
    const decltype(mine)* g = (const decltype(mine)*)all.data();
    bool same_host = true;
    for (int q = 0; q < c->nranks; ++q) same_host &= strncmp(g[q].host, g[0].host, sizeof(mine.host)) == 0;
    HostReduce* hr = nullptr;
    bool created = false;
    if (same_host && c->rank == 0) {
        hr = hr_map(g[0].name, true);
        created = hr != nullptr;
    }
    // barrier 1: the segment exists (or not) before the others open it
    double v = (c->rank == 0 && same_host && !created) ? 1.0 : 0.0;
    NCME_CUDA(cudaMemcpyAsync(c->scratch, &v, sizeof(double), cudaMemcpyHostToDevice, c->ctx->stream));
    NCME_TRY(comm_allreduce_sum(c, c->scratch, 1, c->ctx->stream));
    NCME_CUDA(cudaMemcpyAsync(&v, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(c->ctx->stream));
    const bool ok = same_host && v == 0.0;
    if (ok && c->rank != 0) hr = hr_map(g[0].name, false);
    // barrier 2: everybody mapped it (or the feature is off everywhere); then the name can go
    v = (ok && hr) ? 0.0 : 1.0;
    NCME_CUDA(cudaMemcpyAsync(c->scratch, &v, sizeof(double), cudaMemcpyHostToDevice, c->ctx->stream));
    NCME_TRY(comm_allreduce_sum(c, c->scratch, 1, c->ctx->stream));
    NCME_CUDA(cudaMemcpyAsync(&v, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(c->ctx->stream));
    if (created) shm_unlink(g[0].name);
    if (v != 0.0) {
        hr_unmap(hr);
        hr = nullptr;
    }
    c->hr = hr;
    c->hr_epoch = 0;
    return NCME_OK;
}

bool comm_hostreduce_available(const ncme_comm* c) { return c && c->nranks > 1 && c->hr != nullptr; }

int comm_hostreduce_sum(ncme_comm* c, double* vals, size_t count) {
    if (!c || c->nranks == 1 || count == 0) return NCME_OK;
    NCME_REQUIRE(c->hr && count <= (size_t)NCME_HOSTREDUCE_MAX, "host reduce unavailable or too many values");
    const unsigned long long e = ++c->hr_epoch;
    const int late = hr_sum(c->hr, c->rank, c->nranks, e, vals, count, 1ull << 34);   // ~ a minute of spinning
    if (late) {
        set_error("host all-reduce: rank %d never arrived (epoch %llu)", late - 1, e);
        return NCME_ERR_COMM;
    }
    c->hr_reduces++;
    return NCME_OK;
}

bool comm_dev_allreduce(ncme_comm* c, DevAllreduce* ar) {
    ar->nranks = 1;
    ar->me = 0;
    ar->epoch = 0;
    for (int q = 0; q < NCME_RED_RANKS; ++q) ar->flags[q] = nullptr;
    if (!c || c->nranks == 1) return true;
    static const bool off = getenv("NCME_NO_DEV_ALLREDUCE") != nullptr;
    if (off || !c->p2p_ok || c->nranks > NCME_RED_RANKS || !c->my_flags) return false;
    ar->nranks = c->nranks;
    ar->me = c->rank;
    ar->epoch = ++c->red_epoch;
    for (int q = 0; q < c->nranks; ++q) ar->flags[q] = (q == c->rank) ? c->my_flags : c->peer_flags[q];
    return true;
}

// Flags: every rank exports its PeerFlags block and maps everybody else's.
static int comm_setup_p2p(ncme_comm* c) {
    c->p2p_ok = false;
    if (c->nranks == 1 || c->nranks > NCME_MAX_RANKS || getenv("NCME_HALO_NCCL")) return NCME_OK;
    NCME_CUDA(cudaMalloc(&c->my_flags, sizeof(PeerFlags)));
    NCME_CUDA(cudaMemset(c->my_flags, 0, sizeof(PeerFlags)));
    struct {
        cudaIpcMemHandle_t h;
        int64_t valid;
    } mine;
    memset(&mine, 0, sizeof(mine));
    cudaError_t e = cudaIpcGetMemHandle(&mine.h, c->my_flags);
    mine.valid = e == cudaSuccess;
    if (e != cudaSuccess) cudaGetLastError();
    std::vector<char> all;
    NCME_TRY(allgather_bytes(c, &mine, sizeof(mine), &all));
    bool ok = true;
    for (int q = 0; q < c->nranks; ++q) ok &= ((const decltype(mine)*)all.data())[q].valid != 0;
    int opened = 1;
    if (ok) {
        for (int q = 0; q < c->nranks && opened; ++q) {
            if (q == c->rank) continue;
            void* pb = nullptr;
            e = cudaIpcOpenMemHandle(&pb, ((const decltype(mine)*)all.data())[q].h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                opened = 0;
            } else {
                c->peer_flags[q] = (PeerFlags*)pb;
            }
        }
    } else {
        opened = 0;
    }
    // everybody must agree
    double v = opened ? 0.0 : 1.0;
    NCME_CUDA(cudaMemcpyAsync(c->scratch, &v, sizeof(double), cudaMemcpyHostToDevice, c->ctx->stream));
    NCME_TRY(comm_allreduce_sum(c, c->scratch, 1, c->ctx->stream));
    NCME_CUDA(cudaMemcpyAsync(&v, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(c->ctx->stream));
    c->p2p_ok = (v == 0.0);
    return NCME_OK;
}

}  // namespace ncme

using namespace ncme;

extern "C" {

int ncme_comm_unique_id(char* out128) {
    NCME_REQUIRE(out128, "null argument");
    const NcclApi* api = nccl_api();
    if (!api) return NCME_ERR_COMM;
    ncclUniqueId id;
    NCME_NCCL(api->GetUniqueId(&id));
    memcpy(out128, &id, NCCL_UNIQUE_ID_BYTES);
    return NCME_OK;
}

int ncme_comm_destroy(ncme_comm* c) {
    if (!c) return NCME_OK;
    if (c->ctx) cudaStreamSynchronize(c->ctx->stream);
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    for (auto& r : c->regs)
        for (int q = 0; q < c->nranks; ++q)
            if (r.peer_base[q]) cudaIpcCloseMemHandle(r.peer_base[q]);
    c->regs.clear();
    for (int q = 0; q < NCME_MAX_RANKS; ++q)
        if (c->peer_flags[q]) cudaIpcCloseMemHandle(c->peer_flags[q]);
    hr_unmap(c->hr);
    if (c->my_flags) cudaFree(c->my_flags);
    if (c->ws_base) cudaFree(c->ws_base);
    cudaGetLastError();
    if (c->nccl && nccl_api()) nccl_api()->CommDestroy(c->nccl);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    if (c->scratch) cudaFree(c->scratch);
    delete c;
    return NCME_OK;
}

int ncme_comm_create(ncme_ctx* ctx, int rank, int nranks, const char* uid128, ncme_comm** out) {
    NCME_REQUIRE(ctx && out && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
    ncme_comm* c = new ncme_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->nranks = nranks;
    NCME_CUDA(cudaSetDevice(ctx->device));
    // highest priority: the halo transfer must get SM slots ahead of the (much larger) interior-rows kernel
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc(&c->scratch, 4096 * sizeof(double)) != cudaSuccess) {
        set_error("comm allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        ncme_comm_destroy(c);
        return NCME_ERR_CUDA;
    }
    if (nranks > 1) {
        NCME_REQUIRE(uid128, "a unique id is required for nranks > 1");
        const NcclApi* api = nccl_api();
        if (!api) {
            ncme_comm_destroy(c);
            return NCME_ERR_COMM;
        }
        ncclUniqueId id;
        memcpy(&id, uid128, NCCL_UNIQUE_ID_BYTES);
        ncclResult_t r = api->CommInitRank(&c->nccl, nranks, id, rank);
        if (r != ncclSuccess) {
            set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
            ncme_comm_destroy(c);
            return NCME_ERR_COMM;
        }
        int st = comm_setup_p2p(c);
        if (st == NCME_OK) st = comm_setup_hostreduce(c);
        if (st != NCME_OK) {
            ncme_comm_destroy(c);
            return st;
        }
    }
    *out = c;
    return NCME_OK;
}

int ncme_comm_info(ncme_comm* c, int64_t info[4]) {
    NCME_REQUIRE(c && info, "null argument");
    info[0] = c->p2p_ok ? 1 : 0;
    info[1] = c->p2p_matvecs;
    info[2] = c->nccl_matvecs;
    info[3] = c->bytes_sent;
    return NCME_OK;
}

int ncme_comm_rank(ncme_comm* c, int* rank, int* nranks) {
    NCME_REQUIRE(c, "null communicator");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return NCME_OK;
}

int ncme_comm_allreduce_sum(ncme_comm* c, double* buf_dev, int64_t count) {
    NCME_REQUIRE(c && (count == 0 || buf_dev) && count >= 0, "bad arguments");
    return comm_allreduce_sum(c, buf_dev, (size_t)count, c->ctx->stream);
}

int ncme_comm_allgatherv(ncme_comm* c, const double* send_dev, double* recv_dev, const int64_t* counts,
                         const int64_t* displs) {
    NCME_REQUIRE(c && send_dev && recv_dev && counts && displs, "null argument");
    return comm_allgatherv(c, send_dev, recv_dev, counts, displs, c->ctx->stream);
}

}  // extern "C"
