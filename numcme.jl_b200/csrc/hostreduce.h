// Host-side all-reduce of a few doubles between the ranks of one node through a POSIX shared-memory segment.
// Used by the sharded integrators for their step-control scalars (error norms, Krylov inner products, sink tails),
// which are needed on the HOST: an NCCL all-reduce costs a kernel launch plus ~25 us of device latency before the
// device->host copy can even start; this exchange costs ~2 us.  Plain C++ (no CUDA): comm.cu uses it, and
// tests/test_hostreduce.py runs it between real processes on the CPU.
//
// Protocol for reduction #e (every rank counts its calls, so e agrees everywhere): rank r writes its values into
// slot[e & 1][r] and then publishes seq = e (release); it then waits, for q = 0..P-1 in order, until slot[e & 1][q].seq
// == e (acquire) and adds that rank's values -- rank order, so every rank gets bitwise the same sums.  Two buffers
// suffice: a rank can only start reduction e + 1 after it has seen everybody's seq == e, i.e. after everybody has
// finished reading the buffer of reduction e - 1 that e + 1 is going to overwrite.
#pragma once
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>

namespace ncme {

constexpr int HR_MAX_RANKS = 64;
constexpr int HR_MAX_VALUES = 320;

struct HostReduce {
    struct alignas(128) Slot {
        std::atomic<unsigned long long> seq;
        double vals[HR_MAX_VALUES];
    };
    Slot slot[2][HR_MAX_RANKS];
};

// create (rank 0) or open (others) the segment `name`; nullptr on failure.  A fresh segment is zero-filled: seq = 0.
inline HostReduce* hr_map(const char* name, bool create) {
    int fd = create ? shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600) : shm_open(name, O_RDWR, 0600);
    if (fd < 0) return nullptr;
    if (create && ftruncate(fd, sizeof(HostReduce)) != 0) {
        close(fd);
        shm_unlink(name);
        return nullptr;
    }
    void* m = mmap(nullptr, sizeof(HostReduce), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    return m == MAP_FAILED ? nullptr : static_cast<HostReduce*>(m);
}

inline void hr_unmap(HostReduce* hr) {
    if (hr) munmap(hr, sizeof(HostReduce));
}

// in-place sum over ranks of vals[0..count); returns 0, or the 1-based rank that never arrived within spin_limit polls
inline int hr_sum(HostReduce* hr, int rank, int nranks, unsigned long long epoch, double* vals, size_t count,
                  unsigned long long spin_limit) {
    HostReduce::Slot* buf = hr->slot[epoch & 1];
    HostReduce::Slot& me = buf[rank];
    memcpy(me.vals, vals, count * sizeof(double));
    me.seq.store(epoch, std::memory_order_release);
    double acc[HR_MAX_VALUES];
    for (size_t k = 0; k < count; ++k) acc[k] = 0.0;
    for (int q = 0; q < nranks; ++q) {
        unsigned long long spins = 0;
        while (buf[q].seq.load(std::memory_order_acquire) != epoch) {
            if (++spins > spin_limit) return q + 1;
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        for (size_t k = 0; k < count; ++k) acc[k] += buf[q].vals[k];   // rank order: identical bits on every rank
    }
    memcpy(vals, acc, count * sizeof(double));
    return 0;
}

}  // namespace ncme
