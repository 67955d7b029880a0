// NCCL communicator of libncme (K8): creation from a unique id distributed by the host language
// (torch.distributed / MPI / Julia Distributed), scalar all-reduce, all-gather-v.
#include "comm.cuh"

#include <dlfcn.h>

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
    }
    *out = c;
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
