// K8 plumbing: NCCL communicator owned by libncme (one process per GPU).  NCCL is resolved with
// dlopen("libnccl.so.2") at communicator creation, so single-GPU use never needs it and a process
// that already loaded torch's bundled NCCL shares that copy.
#pragma once
#include <nccl.h>

#include "common.cuh"

struct ncme_comm {
    ncme_ctx* ctx = nullptr;
    int rank = 0, nranks = 1;
    ncclComm_t nccl = nullptr;
    cudaStream_t comm_stream = nullptr;   // halo traffic (overlaps interior rows on ctx->stream)
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    double* scratch = nullptr;            // device, small
    int64_t bytes_sent = 0;
};

namespace ncme {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
};
const NcclApi* nccl_api();  // nullptr (+ error message) if libnccl cannot be loaded

#define NCME_NCCL(expr)                                                                              \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != ncclSuccess) {                                                                     \
            ncme::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncme::nccl_api()->GetErrorString(_r)); \
            return NCME_ERR_COMM;                                                                    \
        }                                                                                            \
    } while (0)

// in-place sum over ranks of `count` doubles on the device (stream-ordered on `st`)
int comm_allreduce_sum(ncme_comm* c, double* buf_dev, size_t count, cudaStream_t st);
// gather variable-size slices: recv_dev[displs[r] .. +counts[r]) <- rank r's send_dev[0..counts[r])
int comm_allgatherv(ncme_comm* c, const double* send_dev, double* recv_dev, const int64_t* counts, const int64_t* displs,
                    cudaStream_t st);

}  // namespace ncme
