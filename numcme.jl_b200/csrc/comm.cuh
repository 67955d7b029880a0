// K8 plumbing: NCCL communicator owned by libncme (one process per GPU).  NCCL is resolved with
// dlopen("libnccl.so.2") at communicator creation, so single-GPU use never needs it and a process
// that already loaded torch's bundled NCCL shares that copy.
#pragma once
#include <nccl.h>

#include "common.cuh"

constexpr int NCME_MAX_RANKS = 64;
constexpr int NCME_HOSTREDUCE_MAX = 320;   // doubles per host-side all-reduce (1 + 9 R sink tails for R <= 35; wider sets fall back to the device all-reduce)

constexpr int NCME_RED_RANKS = 16;         // device-side all-reduce of small scalar sets: ranks, buffers, values
constexpr int NCME_RED_BUFS = 4;
constexpr int NCME_RED_VALS = 32;
constexpr int NCME_FZ_VALS = 64;           // widest in-kernel all-reduce of the fused BDF step (2 R sink sums; the fused kernel handles R <= 32)

// Flags each rank exposes to its peers through CUDA IPC (peer GPUs store into them over NVLink).
struct PeerFlags {
    unsigned int ready[NCME_MAX_RANKS];   // ready[q] = e: rank q's input vector of matvec #e is complete
    unsigned int done[NCME_MAX_RANKS];    // done[q]  = e: rank q has finished reading my vector in matvec #e
    unsigned int error;                   // a wait timed out
    // in-kernel all-reduce (Krylov inner products): rank q stores its partial sums of reduction #e into
    // red_slot[e % BUFS][q][..] of EVERY rank and then publishes red_flag[e % BUFS][q] = e there
    unsigned int red_flag[NCME_RED_BUFS][NCME_RED_RANKS];
    double red_slot[NCME_RED_BUFS][NCME_RED_RANKS][NCME_RED_VALS];
    // the same exchange driven from INSIDE the persistent fused BDF step kernel (bdf_fused.cu, sharded mode): barriers
    // and all-reduces between the ranks' cooperative grids; fz_epoch is this rank's running sequence number (local use)
    unsigned int fz_flag[NCME_RED_BUFS][NCME_RED_RANKS];
    double fz_slot[NCME_RED_BUFS][NCME_RED_RANKS][NCME_FZ_VALS];
    unsigned int fz_epoch;
};

// By-value kernel argument of the in-kernel all-reduce.  nranks == 1: no exchange.
struct DevAllreduce {
    int nranks, me;
    unsigned int epoch;                   // sequence number of this reduction (same on every rank)
    PeerFlags* flags[NCME_RED_RANKS];     // flags[q] = rank q's PeerFlags as mapped into this process (flags[me] = my own)
};

// A device allocation whose vectors the neighbouring ranks may read directly (halo pull over NVLink).
// Vector v of rank q starts (its local rows) at base_q + (local0_q + v * stride_q) doubles.
struct RegBuf {
    void* base = nullptr;
    size_t bytes = 0;
    int64_t local0 = 0, stride = 0, nvec = 1;
    void* peer_base[NCME_MAX_RANKS] = {nullptr};
    int64_t peer_local0[NCME_MAX_RANKS] = {0};
    int64_t peer_stride[NCME_MAX_RANKS] = {0};
};

namespace ncme {
struct HostReduce;
}

struct ncme_comm {
    ncme_ctx* ctx = nullptr;
    int rank = 0, nranks = 1;
    ncclComm_t nccl = nullptr;
    cudaStream_t comm_stream = nullptr;   // halo traffic (overlaps interior rows on ctx->stream)
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    double* scratch = nullptr;            // device, small
    int64_t bytes_sent = 0;
    // peer-memory transport (CUDA IPC): flags + registered buffers
    bool p2p_ok = false;
    PeerFlags* my_flags = nullptr;
    PeerFlags* peer_flags[NCME_MAX_RANKS] = {nullptr};
    unsigned int epoch = 0;
    unsigned int red_epoch = 0;           // sequence number of the in-kernel all-reduces
    std::vector<RegBuf> regs;
    int64_t p2p_matvecs = 0, nccl_matvecs = 0;
    // host-side all-reduce of step-control scalars through POSIX shared memory (all ranks live on one node)
    ncme::HostReduce* hr = nullptr;
    unsigned long long hr_epoch = 0;
    int64_t hr_reduces = 0;
    // grow-only integrator workspace, registered for peer access (re-allocation is a collective decision)
    double* ws_base = nullptr;
    size_t ws_bytes = 0;
    int64_t ws_local0 = -1, ws_stride = -1, ws_nvec = -1;
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

// Collective: expose [base, base+bytes) to the ranks in `peers` (and map theirs).  local0/stride/nvec describe where
// the vectors live inside the allocation (see RegBuf).  Idempotent for an already registered base.
int comm_register(ncme_comm* c, void* base, size_t bytes, int64_t local0, int64_t stride, int64_t nvec, const int* peers,
                  int npeers);
// Collective: unmap the peers' views of `base` on every rank (call before freeing it).
int comm_unregister(ncme_comm* c, void* base);
// Collective: make sure the registered integrator workspace holds `bytes` with the given vector layout.
int comm_workspace(ncme_comm* c, size_t bytes, int64_t local0, int64_t stride, int64_t nvec, const int* peers, int npeers,
                   double** out);
// pointer to the local rows of the same vector on rank q, or nullptr if x_local is not inside a registered buffer
const double* comm_peer_vector(const ncme_comm* c, const double* x_local, int q);

// Sum over ranks of `count` (<= NCME_HOSTREDUCE_MAX) HOST doubles, in place, summed in rank order on every rank
// (bitwise identical results everywhere).  The integrators' step-control scalars (error norms, Krylov inner products,
// sink tails) are needed on the HOST; an NCCL all-reduce costs a kernel launch + ~25 us of device latency before the
// device->host copy can even start, the exchange through a shared-memory segment costs ~2 us.
// Returns false when the shared segment is unavailable (the caller then uses comm_allreduce_sum on the device).
bool comm_hostreduce_available(const ncme_comm* c);
int comm_hostreduce_sum(ncme_comm* c, double* host_vals, size_t count);

// Descriptor of the next in-kernel all-reduce (advances the sequence number).  Returns false when the peer-memory
// transport is unavailable or there are more than NCME_RED_RANKS ranks: the caller then reduces with NCCL.
bool comm_dev_allreduce(ncme_comm* c, DevAllreduce* ar);

// in-place sum over ranks of `count` doubles on the device (stream-ordered on `st`)
int comm_allreduce_sum(ncme_comm* c, double* buf_dev, size_t count, cudaStream_t st);
// gather variable-size slices: recv_dev[displs[r] .. +counts[r]) <- rank r's send_dev[0..counts[r])
int comm_allgatherv(ncme_comm* c, const double* send_dev, double* recv_dev, const int64_t* counts, const int64_t* displs,
                    cudaStream_t st);

}  // namespace ncme
