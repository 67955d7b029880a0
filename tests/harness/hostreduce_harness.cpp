// CPU harness for numcme.jl_b200/csrc/hostreduce.h (the shared-memory all-reduce of the sharded integrators' step
// scalars): each process opens the segment and runs `rounds` reductions of `count` values; see tests/test_hostreduce.py.
#include "hostreduce.h"

#include <stdio.h>
#include <stdlib.h>

using namespace ncme;

extern "C" {

void* h_open(const char* name, int create) { return hr_map(name, create != 0); }
void h_close(void* hr) { hr_unmap(static_cast<HostReduce*>(hr)); }
void h_unlink(const char* name) { shm_unlink(name); }

// rounds of reductions; vals(round, k) = (rank + 1) * 1e-3 * (k + 1) + round * 0.5 + jitter(rank, round, k).  The sums are
// written to out[round * count + k].  Returns 0, or 1-based rank that timed out.
int h_run(void* hrp, int rank, int nranks, int rounds, int count, double* out) {
    HostReduce* hr = static_cast<HostReduce*>(hrp);
    double vals[HR_MAX_VALUES];
    for (int e = 1; e <= rounds; ++e) {
        for (int k = 0; k < count; ++k) {
            unsigned long long z = 0x9E3779B97F4A7C15ull * (unsigned long long)(rank * 1000003 + e * 7919 + k + 1);
            z ^= z >> 31;
            vals[k] = (rank + 1) * 1e-3 * (k + 1) + e * 0.5 + (double)(z % 1000003) * 1e-9;
        }
        const int late = hr_sum(hr, rank, nranks, (unsigned long long)e, vals, (size_t)count, 1ull << 33);
        if (late) return late;
        for (int k = 0; k < count; ++k) out[(size_t)(e - 1) * count + k] = vals[k];
    }
    return 0;
}
}
