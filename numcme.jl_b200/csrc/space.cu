// StateSpaceSparse on the device: GPU hash table (K3), frontier expansion with stream compaction (K4),
// deletion/reindexing (K5).  Reference: src/statespace/sparse/sparsestatespace.jl
//   constructor :103-124, expand! :153-194, _addstates! :208-267, deleteat! :276-331.
//
// Ordering contract: new states are appended in exactly the reference's insertion order.  The
// reference walks the candidate list sequentially and gives the next index to the first occurrence
// of every unseen non-negative state (:220-226).  Here every candidate c carries its position in
// that list as a rank; duplicates race with atomicMin(rank) on the hash-table value, the winners
// (rank == stored value) are stream-compacted in rank order, which reproduces the sequential result.
#include "space.cuh"

#include <algorithm>

namespace ncme {

// ------------------------------------------------------------------------------------ kernels ---
__global__ void k_fill_u32(uint32_t* p, int64_t n, uint32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// insert key -> value i for i in [first, n)
__global__ void k_table_insert_range(HashView h, const uint64_t* __restrict__ keys, int64_t first, int64_t n) {
    int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    uint64_t slot = hash64(key) & h.capmask;
    while (true) {
        uint64_t prev = atomicCAS((unsigned long long*)&h.keys[slot], (unsigned long long)EMPTY_KEY,
                                  (unsigned long long)key);
        if (prev == EMPTY_KEY || prev == key) {
            h.vals[slot] = (uint32_t)i;
            return;
        }
        slot = (slot + 1) & h.capmask;
    }
}

// key of x + sign * s_r.  Returns 0 if representable, 1 if a component would be negative (not a
// state at all), 2 if non-negative but too wide for the packed key.
__device__ __forceinline__ int shifted_key(const KeyLayout& L, const StoichDev& S, uint64_t key, int r, int sign,
                                           uint64_t* out) {
    uint64_t nk = 0;
    bool neg = false, wide = false;
    for (int s = 0; s < L.ns; ++s) {
        long long v = (long long)((key >> L.shift[s]) & L.mask[s]) + (long long)sign * S.s[r][s];
        neg |= v < 0;
        wide |= v > (long long)L.mask[s];
        nk |= ((uint64_t)(v < 0 ? 0 : v)) << L.shift[s];
    }
    *out = nk;
    return neg ? 1 : (wide ? 2 : 0);
}

// Candidate generation (expand! :176-186): candidate rank c <-> (frontier entry popped LIFO, reaction k).
// frontier == nullptr means the frontier is the index range [fbase, fbase + F).
__global__ void k_gen_candidates(KeyLayout L, StoichDev S, const uint64_t* __restrict__ keys,
                                 const uint32_t* __restrict__ frontier, int64_t fbase, int64_t F, int nreact,
                                 const ReactList reacts /*0-based*/, uint64_t* __restrict__ cand_key, int* err_flag) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= F * nreact) return;
    const int64_t f = F - 1 - c / nreact;
    const int r = reacts.r[c % nreact];
    const int64_t idx = frontier ? (int64_t)frontier[f] : fbase + f;
    uint64_t nk;
    const int rc = shifted_key(L, S, keys[idx], r, +1, &nk);
    if (rc == 2) atomicExch(err_flag, 1);
    cand_key[c] = rc == 0 ? nk : EMPTY_KEY;
}

// _addstates! first loop (:219-226), parallel: probe; existing states drop out; unseen keys are
// inserted with value = min over duplicates of (n_old + rank).
__global__ void k_insert_candidates(HashView h, const uint64_t* __restrict__ cand_key, int64_t ncand, uint32_t n_old,
                                    uint32_t* __restrict__ cand_slot) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    const uint64_t key = cand_key[c];
    if (key == EMPTY_KEY) {
        cand_slot[c] = NONE32;
        return;
    }
    uint64_t slot = hash64(key) & h.capmask;
    while (true) {
        uint64_t prev = atomicCAS((unsigned long long*)&h.keys[slot], (unsigned long long)EMPTY_KEY,
                                  (unsigned long long)key);
        if (prev == EMPTY_KEY || prev == key) {
            // committed states have values < n_old and are never modified by atomicMin with >= n_old
            uint32_t before = atomicMin(&h.vals[slot], n_old + (uint32_t)c);
            cand_slot[c] = (before < n_old) ? NONE32 : (uint32_t)slot;
            return;
        }
        slot = (slot + 1) & h.capmask;
    }
}

__global__ void k_mark_winners(HashView h, const uint32_t* __restrict__ cand_slot, int64_t ncand, uint32_t n_old,
                               uint32_t* __restrict__ flags) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    const uint32_t slot = cand_slot[c];
    flags[c] = (slot != NONE32 && h.vals[slot] == n_old + (uint32_t)c) ? 1u : 0u;
}

__global__ void k_commit_new(HashView h, const uint64_t* __restrict__ cand_key, const uint32_t* __restrict__ cand_slot,
                             const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, int64_t ncand,
                             uint32_t n_old, uint64_t* __restrict__ keys, uint32_t* __restrict__ pred, int64_t ld, int nr,
                             smask_t* __restrict__ sinkmask) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand || !flags[c]) return;
    const uint32_t i = n_old + pos[c];
    keys[i] = cand_key[c];
    h.vals[cand_slot[c]] = i;
    sinkmask[i] = 0;
    for (int r = 0; r < nr; ++r) pred[(int64_t)r * ld + i] = NONE32;
}

// _addstates! connectivity loops (:241-266), one thread per (new state, reaction).
__global__ void k_connect_new(HashView h, KeyLayout L, StoichDev S, const uint64_t* __restrict__ keys, int64_t n_old,
                              int64_t n_new, uint32_t* __restrict__ pred, int64_t ld, smask_t* __restrict__ sinkmask) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nr = S.nr;
    if (t >= (n_new - n_old) * nr) return;
    const int64_t i = n_old + t / nr;
    const int r = (int)(t % nr);
    const uint64_t key = keys[i];
    uint64_t k2;
    if (shifted_key(L, S, key, r, -1, &k2) == 0) {  // predecessor x_i - s_r
        uint32_t j = hash_lookup(h, k2);
        if (j != NONE32) {
            pred[(int64_t)r * ld + i] = j;
            atomicAnd(&sinkmask[j], ~SMASK1(r));
        }
    }
    const int rc = shifted_key(L, S, key, r, +1, &k2);  // successor x_i + s_r
    if (rc == 0) {
        uint32_t j = hash_lookup(h, k2);
        if (j == NONE32)
            atomicOr(&sinkmask[i], SMASK1(r));
        else
            pred[(int64_t)r * ld + j] = (uint32_t)i;
    } else if (rc == 2) {
        // successor is non-negative but does not fit the key: it cannot be in the space => it is a sink
        atomicOr(&sinkmask[i], SMASK1(r));
    }
}

__global__ void k_frontier_flags(const smask_t* __restrict__ sinkmask, int64_t n, smask_t reactmask,
                                 uint32_t* __restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (sinkmask[i] & reactmask) ? 1u : 0u;
}

__global__ void k_compact_indices(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, int64_t n,
                                  uint32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[pos[i]] = (uint32_t)i;
}

// deleteat! :283-317 -- compact keys / sink masks / predecessor table through the old->new index map.
__global__ void k_clear_flags_at(const uint32_t* __restrict__ ids0 /*0-based*/, int64_t nids, uint32_t* __restrict__ keep) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nids) keep[ids0[k]] = 0u;
}

__global__ void k_compact_space(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, int64_t n,
                                const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pred, int64_t ld,
                                const smask_t* __restrict__ sinkmask, int nr, uint64_t* __restrict__ keys2,
                                uint32_t* __restrict__ pred2, int64_t ld2, smask_t* __restrict__ sinkmask2,
                                const uint32_t* __restrict__ origin, uint32_t* __restrict__ origin2) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const uint32_t j = pos[i];
    keys2[j] = keys[i];
    sinkmask2[j] = sinkmask[i];
    if (origin) origin2[j] = origin[i];
    for (int r = 0; r < nr; ++r) {
        uint32_t p = pred[(int64_t)r * ld + i];
        pred2[(int64_t)r * ld2 + j] = (p != NONE32 && keep[p]) ? pos[p] : NONE32;
    }
}

// deleteat! :318-329 -- re-derive sink flags of the surviving states.
__global__ void k_rederive_sinks(HashView h, KeyLayout L, StoichDev S, const uint64_t* __restrict__ keys, int64_t n,
                                 smask_t* __restrict__ sinkmask) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nr = S.nr;
    if (t >= n * nr) return;
    const int64_t i = t / nr;
    const int r = (int)(t % nr);
    if (sinkmask[i] & SMASK1(r)) return;
    uint64_t k2;
    const int rc = shifted_key(L, S, keys[i], r, +1, &k2);
    if ((rc == 0 && hash_lookup(h, k2) == NONE32) || rc == 2) atomicOr(&sinkmask[i], SMASK1(r));
}

__global__ void k_lookup(HashView h, const uint64_t* __restrict__ q, int64_t m, uint32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t k = q[i];
    uint32_t v = (k == EMPTY_KEY) ? NONE32 : hash_lookup(h, k);
    out[i] = (v == NONE32) ? 0u : v + 1u;
}

// exact per-species maxima of the packed keys (key re-layout decisions)
__global__ void k_species_max(KeyLayout L, const uint64_t* __restrict__ keys, int64_t n, unsigned long long* __restrict__ mx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t key = i < n ? keys[i] : 0ull;     // no early return: the whole warp takes part in the shuffles
    for (int s = 0; s < L.ns; ++s) {
        unsigned long long v = (key >> L.shift[s]) & L.mask[s];
        for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
        if ((threadIdx.x & 31) == 0) atomicMax(&mx[s], v);
    }
}

__global__ void k_repack_keys(KeyLayout from, KeyLayout to, uint64_t* __restrict__ keys, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    uint64_t nk = 0;
    for (int s = 0; s < from.ns; ++s) nk |= ((key >> from.shift[s]) & from.mask[s]) << to.shift[s];
    keys[i] = nk;
}

// --------------------------------------------------------------------------------- host side ---
static inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }


// ---------------------------------------------------------------- small spaces: all levels in ONE launch ----
// expand! on a space of a few thousand states is pure launch/synchronisation latency in the per-level pipeline above
// (8 launches + a host round trip per level; the reference's example configurations expand 5-20 levels per adapt
// event).  One CTA runs the SAME algorithm -- candidate ranks, atomicMin on the table value, ordered compaction of the
// winners, connectivity of the new states -- for as many levels as the pre-reserved capacities allow, with
// __syncthreads() between the phases; the host continues with the general path if a level does not fit.
struct ExpandSmallArgs {
    HashView h;
    KeyLayout L;
    StoichDev S;
    ReactList reacts;
    smask_t reactmask;
    uint64_t* keys;
    uint32_t* pred;
    int64_t ld;
    smask_t* sinkmask;
    uint64_t* cand_key;
    uint32_t* cand_slot;
    uint32_t* frontier;
    int64_t n0, cand_cap;
    uint64_t tcap;
    int levels;
    int* err_flag;
    long long* out;   // [0] states after, [1] levels done, [2] states added by the last level done, [3] stopped for capacity
};
constexpr int XS_THREADS = 1024;

__device__ __forceinline__ uint32_t block_flag_scan(bool flag, uint32_t* warp_tot, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    const uint32_t excl = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[wid] = __popc(b);
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < XS_THREADS / 32; ++w) {
        const uint32_t v = warp_tot[w];
        if (w < wid) off += v;
        tot += v;
    }
    __syncthreads();
    *total = tot;
    return off + excl;
}

__global__ void __launch_bounds__(XS_THREADS) k_expand_small(const __grid_constant__ ExpandSmallArgs a) {
    __shared__ uint32_t warp_tot[XS_THREADS / 32];
    const int tid = threadIdx.x;
    const int nreact = a.reacts.n, nr = a.S.nr;
    int64_t n = a.n0;
    // explorables = states with a sink flag on one of the expansion reactions (:165-173), ascending
    uint32_t running = 0;
    for (int64_t base = 0; base < n; base += XS_THREADS) {
        const int64_t i = base + tid;
        const bool flag = i < n && (a.sinkmask[i] & a.reactmask);
        uint32_t total;
        const uint32_t pos = block_flag_scan(flag, warp_tot, &total);
        if (flag) a.frontier[running + pos] = (uint32_t)i;
        running += total;
    }
    __syncthreads();
    int64_t F = running, fbase = 0;
    bool use_list = true;
    int done = 0, stopped = 0;
    int64_t last_added = 0;
    for (int level = 0; level < a.levels && F > 0; ++level) {
        const int64_t ncand = F * nreact;
        if (ncand > a.cand_cap || n + ncand > a.ld || 2ull * (uint64_t)(n + ncand) > a.tcap) {
            stopped = 1;
            break;
        }
        const uint32_t n_old = (uint32_t)n;
        // candidates (LIFO over the frontier, reactions ascending :176-186) and the first loop of _addstates! (:219-226)
        for (int64_t c = tid; c < ncand; c += XS_THREADS) {
            const int64_t f = F - 1 - c / nreact;
            const int r = a.reacts.r[c % nreact];
            const int64_t idx = use_list ? (int64_t)a.frontier[f] : fbase + f;
            uint64_t nk;
            const int rc = shifted_key(a.L, a.S, a.keys[idx], r, +1, &nk);
            if (rc == 2) atomicExch(a.err_flag, 1);
            const uint64_t key = rc == 0 ? nk : EMPTY_KEY;
            a.cand_key[c] = key;
            uint32_t cs = NONE32;
            if (key != EMPTY_KEY) {
                uint64_t slot = hash64(key) & a.h.capmask;
                while (true) {
                    const uint64_t prev = atomicCAS((unsigned long long*)&a.h.keys[slot], (unsigned long long)EMPTY_KEY,
                                                    (unsigned long long)key);
                    if (prev == EMPTY_KEY || prev == key) {
                        const uint32_t before = atomicMin(&a.h.vals[slot], n_old + (uint32_t)c);
                        cs = (before < n_old) ? NONE32 : (uint32_t)slot;
                        break;
                    }
                    slot = (slot + 1) & a.h.capmask;
                }
            }
            a.cand_slot[c] = cs;
        }
        __syncthreads();
        // winners in rank order -> new indices; commit
        running = 0;
        for (int64_t base = 0; base < ncand; base += XS_THREADS) {
            const int64_t c = base + tid;
            uint32_t slot = NONE32;
            bool flag = false;
            if (c < ncand) {
                slot = a.cand_slot[c];
                flag = slot != NONE32 && a.h.vals[slot] == n_old + (uint32_t)c;
            }
            uint32_t total;
            const uint32_t pos = block_flag_scan(flag, warp_tot, &total);
            if (flag) {
                const uint32_t i = n_old + running + pos;
                a.keys[i] = a.cand_key[c];
                a.h.vals[slot] = i;
                a.sinkmask[i] = 0;
                for (int r = 0; r < nr; ++r) a.pred[(int64_t)r * a.ld + i] = NONE32;
            }
            running += total;
        }
        __syncthreads();
        const int64_t m = running, n_new = n + m;
        // connectivity of the new states (:241-266)
        for (int64_t t = tid; t < m * nr; t += XS_THREADS) {
            const int64_t i = n + t / nr;
            const int r = (int)(t % nr);
            const uint64_t key = a.keys[i];
            uint64_t k2;
            if (shifted_key(a.L, a.S, key, r, -1, &k2) == 0) {
                const uint32_t j = hash_lookup(a.h, k2);
                if (j != NONE32) {
                    a.pred[(int64_t)r * a.ld + i] = j;
                    atomicAnd(&a.sinkmask[j], ~SMASK1(r));
                }
            }
            const int rc = shifted_key(a.L, a.S, key, r, +1, &k2);
            if (rc == 0) {
                const uint32_t j = hash_lookup(a.h, k2);
                if (j == NONE32)
                    atomicOr(&a.sinkmask[i], SMASK1(r));
                else
                    a.pred[(int64_t)r * a.ld + j] = (uint32_t)i;
            } else if (rc == 2) {
                atomicOr(&a.sinkmask[i], SMASK1(r));
            }
        }
        __syncthreads();
        use_list = false;
        fbase = n;
        F = m;
        last_added = m;
        n = n_new;
        ++done;
    }
    if (tid == 0) {
        a.out[0] = n;
        a.out[1] = done;
        a.out[2] = last_added;
        a.out[3] = stopped;
    }
}

// ------------------------------------------------------------ marginals: sum(p, dims) on the device ----------
// Reference: Base.sum(p::FspVectorSparse, dims) (src/fspvector/fspvector.jl:66-99): reduced states in order of first
// occurrence, values accumulated in state order.  Device version: the reduced key of a state is its packed key with
// the summed-out fields masked away; a temporary open-addressing table maps every reduced key to the smallest state
// index carrying it (atomicMin), those first occurrences are stream-compacted in index order (= the reference's
// order of the reduced states), then every state adds its probability to its bucket.
__global__ void k_marg_insert(HashView h, const uint64_t* __restrict__ keys, int64_t n, uint64_t keepmask,
                              uint32_t* __restrict__ slot_of) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i] & keepmask;
    uint64_t slot = hash64(key) & h.capmask;
    while (true) {
        const uint64_t prev = atomicCAS((unsigned long long*)&h.keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
        if (prev == EMPTY_KEY || prev == key) {
            atomicMin(&h.vals[slot], (uint32_t)i);
            slot_of[i] = (uint32_t)slot;
            return;
        }
        slot = (slot + 1) & h.capmask;
    }
}
__global__ void k_marg_flags(HashView h, const uint32_t* __restrict__ slot_of, int64_t n, uint32_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = h.vals[slot_of[i]] == (uint32_t)i ? 1u : 0u;
}
__global__ void k_marg_accumulate(HashView h, const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ flags,
                                  const uint32_t* __restrict__ pos, const uint64_t* __restrict__ keys, uint64_t keepmask,
                                  const double* __restrict__ p, int64_t n, uint64_t* __restrict__ rkeys,
                                  double* __restrict__ rvals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t first = h.vals[slot_of[i]];
    const uint32_t j = pos[first];
    if (flags[i]) rkeys[j] = keys[i] & keepmask;
    atomicAdd(&rvals[j], p[i]);
}

#define LAUNCH(ctx, kern, n, ...)                                          \
    do {                                                                   \
        if ((n) > 0) {                                                     \
            kern<<<nblk(n), 256, 0, (ctx)->stream>>>(__VA_ARGS__);          \
            (ctx)->launches++;                                             \
            NCME_CUDA(cudaGetLastError());                                 \
        }                                                                  \
    } while (0)

int space_pack_host(const ncme_space* sp, const int64_t* state, uint64_t* key_out) {
    uint64_t k = 0;
    for (int s = 0; s < sp->ns; ++s) {
        if (state[s] < 0) return 1;
        if ((uint64_t)state[s] > sp->layout.mask[s]) {
            set_error("state component %lld of species %d does not fit the %d-species packed key", (long long)state[s],
                      s + 1, sp->ns);
            return NCME_ERR_KEYWIDTH;
        }
        k |= (uint64_t)state[s] << sp->layout.shift[s];
    }
    *key_out = k;
    return 0;
}

int space_reserve_rows(ncme_space* sp, int64_t nrows) {
    if (nrows <= sp->ld) return NCME_OK;
    cudaStream_t st = sp->ctx->stream;
    int64_t nld = sp->ld ? sp->ld : 1024;
    while (nld < nrows) nld = nld + nld / 2;
    nld = round_up<int64_t>(nld, 64);
    NCME_TRY(sp->keys.reserve((size_t)nld, st));
    NCME_TRY(sp->sinkmask.reserve((size_t)nld, st));
    if (sp->mark_n >= 0) NCME_TRY(sp->origin.reserve((size_t)nld, st));
    // the slot-major table changes stride: re-lay it out
    DevArray<uint32_t> np;
    NCME_TRY(np.reserve((size_t)nld * sp->nr, st, false));
    for (int r = 0; r < sp->nr && sp->n > 0; ++r)
        NCME_CUDA(cudaMemcpyAsync(np.p + (size_t)r * nld, sp->pred.p + (size_t)r * sp->ld, (size_t)sp->n * sizeof(uint32_t),
                                  cudaMemcpyDeviceToDevice, st));
    NCME_CUDA(cudaStreamSynchronize(st));
    sp->pred.release();
    sp->pred = np;
    sp->ld = nld;
    return NCME_OK;
}

int space_rebuild_table(ncme_space* sp, uint64_t min_slots) {
    ncme_ctx* ctx = sp->ctx;
    uint64_t cap = 1024;
    while (cap < min_slots) cap <<= 1;
    if (cap < sp->tcap) cap = sp->tcap;
    NCME_TRY(sp->tkeys.reserve(cap, ctx->stream, false));
    NCME_TRY(sp->tvals.reserve(cap, ctx->stream, false));
    sp->tcap = cap;
    NCME_CUDA(cudaMemsetAsync(sp->tkeys.p, 0xFF, cap * sizeof(uint64_t), ctx->stream));
    NCME_CUDA(cudaMemsetAsync(sp->tvals.p, 0xFF, cap * sizeof(uint32_t), ctx->stream));
    LAUNCH(ctx, k_table_insert_range, sp->n, sp->hview(), sp->keys.p, (int64_t)0, sp->n);
    return NCME_OK;
}

static int ensure_scratch(ncme_space* sp, int64_t m) {
    cudaStream_t st = sp->ctx->stream;
    NCME_TRY(sp->cand_key.reserve((size_t)m, st, true));
    NCME_TRY(sp->cand_slot.reserve((size_t)m, st, false));
    NCME_TRY(sp->flags.reserve((size_t)m, st, false));
    NCME_TRY(sp->pos.reserve((size_t)m, st, false));
    NCME_TRY(sp->scan_scratch.reserve(scan_scratch_elems(m), st, false));
    return NCME_OK;
}

int space_addstates(ncme_space* sp, int64_t ncand, int64_t* added) {
    ncme_ctx* ctx = sp->ctx;
    *added = 0;
    sp->last_delete_nold = -1;  // the flags/pos scratch is about to be reused
    if (ncand <= 0) return NCME_OK;
    NCME_REQUIRE((uint64_t)sp->n + (uint64_t)ncand < 0xFFFFFFF0ull, "state space exceeds the 32-bit index range");
    NCME_TRY(ensure_scratch(sp, ncand));
    if (2 * (uint64_t)(sp->n + ncand) > sp->tcap) NCME_TRY(space_rebuild_table(sp, 4 * (uint64_t)(sp->n + ncand)));
    const uint32_t n_old = (uint32_t)sp->n;
    LAUNCH(ctx, k_insert_candidates, ncand, sp->hview(), sp->cand_key.p, ncand, n_old, sp->cand_slot.p);
    LAUNCH(ctx, k_mark_winners, ncand, sp->hview(), sp->cand_slot.p, ncand, n_old, sp->flags.p);
    uint64_t m = 0;
    NCME_TRY(exclusive_scan_u32(ctx, sp->flags.p, sp->pos.p, ncand, sp->scan_scratch.p, sp->scan_scratch.cap, &m));
    if (m == 0) return NCME_OK;
    NCME_TRY(space_reserve_rows(sp, sp->n + (int64_t)m));
    LAUNCH(ctx, k_commit_new, ncand, sp->hview(), sp->cand_key.p, sp->cand_slot.p, sp->flags.p, sp->pos.p, ncand, n_old,
           sp->keys.p, sp->pred.p, sp->ld, sp->nr, sp->sinkmask.p);
    const int64_t n_new = sp->n + (int64_t)m;
    LAUNCH(ctx, k_connect_new, (int64_t)m * sp->nr, sp->hview(), sp->layout, sp->sdev, sp->keys.p, sp->n, n_new,
           sp->pred.p, sp->ld, sp->sinkmask.p);
    sp->n = n_new;
    sp->version++;
    *added = (int64_t)m;
    return NCME_OK;
}

// deleteat! with the keep flags already in sp->flags[0..n): compaction through the old->new map
// (:283-317), hash-table rebuild, sink re-derivation (:318-329).  sp->flags / sp->pos stay valid
// afterwards (last_delete_nold) so that vectors can be compacted with the same map.
int space_delete_flagged(ncme_space* sp) {
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t s = ctx->stream;
    const int64_t n = sp->n;
    NCME_TRY(sp->pos.reserve((size_t)n, s, false));
    NCME_TRY(sp->scan_scratch.reserve(scan_scratch_elems(n), s, false));
    uint64_t m = 0;
    NCME_TRY(exclusive_scan_u32(ctx, sp->flags.p, sp->pos.p, n, sp->scan_scratch.p, sp->scan_scratch.cap, &m));
    sp->last_delete_nold = n;
    sp->last_delete_nnew = (int64_t)m;
    if ((int64_t)m == n) return NCME_OK;
    DevArray<uint64_t> k2;
    DevArray<uint32_t> p2, o2;
    DevArray<smask_t> m2;
    // keep head-room for the expansion that follows every prune (adapt!, rstepadapters.jl:44-49) instead of shrinking
    // to fit and re-growing (re-layout of the slot-major predecessor table) a moment later
    const int64_t ld2 = round_up<int64_t>(std::max<int64_t>(1024, std::min<int64_t>(sp->ld, 2 * (int64_t)m + 8192)), 64);
    NCME_TRY(k2.reserve((size_t)ld2, s, false));
    NCME_TRY(m2.reserve((size_t)ld2, s, false));
    NCME_TRY(p2.reserve((size_t)ld2 * sp->nr, s, false));
    const bool tracked = sp->mark_n >= 0;
    if (tracked) NCME_TRY(o2.reserve((size_t)ld2, s, false));
    LAUNCH(ctx, k_compact_space, n, sp->flags.p, sp->pos.p, n, sp->keys.p, sp->pred.p, sp->ld, sp->sinkmask.p, sp->nr, k2.p,
           p2.p, ld2, m2.p, tracked ? sp->origin.p : nullptr, o2.p);
    NCME_CUDA(cudaStreamSynchronize(s));
    sp->keys.release();
    sp->pred.release();
    sp->sinkmask.release();
    sp->keys = k2;
    sp->pred = p2;
    sp->sinkmask = m2;
    if (tracked) {
        sp->origin.release();
        sp->origin = o2;
    }
    sp->ld = ld2;
    sp->n = (int64_t)m;
    sp->version++;
    if (sp->n == 0) {
        NCME_CUDA(cudaMemsetAsync(sp->tkeys.p, 0xFF, sp->tcap * sizeof(uint64_t), s));
        NCME_CUDA(cudaMemsetAsync(sp->tvals.p, 0xFF, sp->tcap * sizeof(uint32_t), s));
        return NCME_OK;
    }
    NCME_TRY(space_rebuild_table(sp, 4 * (uint64_t)(sp->n + 256)));
    LAUNCH(ctx, k_rederive_sinks, sp->n * sp->nr, sp->hview(), sp->layout, sp->sdev, sp->keys.p, sp->n, sp->sinkmask.p);
    NCME_CUDA(cudaStreamSynchronize(s));
    return NCME_OK;
}

__global__ void k_compact_vector(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, int64_t n,
                                 const double* __restrict__ in, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && keep[i]) out[pos[i]] = in[i];
}

static int bitlen(uint64_t v) {
    int b = 0;
    while (v) {
        ++b;
        v >>= 1;
    }
    return b;
}

// The 63 key bits start as an equal split; when a species is about to outgrow its field the budget is re-divided
// from the observed maxima (all keys re-packed, table rebuilt).  NCME_ERR_KEYWIDTH only if 63 bits cannot hold it.
int space_ensure_key_room(ncme_space* sp, const int64_t* inc) {
    const int ns = sp->ns;
    bool tight = false;
    for (int s = 0; s < ns; ++s) tight |= (uint64_t)(sp->ub[s] + inc[s]) > sp->layout.mask[s];
    if (!tight) return NCME_OK;
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t st = ctx->stream;
    // the host bound is pessimistic (+inc per level): get the exact maxima first
    if (sp->n > 0) {
        unsigned long long* d_mx = nullptr;
        NCME_CUDA(cudaMalloc(&d_mx, sizeof(unsigned long long) * NCME_MAX_SPECIES));
        NCME_CUDA(cudaMemsetAsync(d_mx, 0, sizeof(unsigned long long) * NCME_MAX_SPECIES, st));
        k_species_max<<<(unsigned)((sp->n + 255) / 256), 256, 0, st>>>(sp->layout, sp->keys.p, sp->n, d_mx);
        ctx->launches++;
        unsigned long long h_mx[NCME_MAX_SPECIES];
        NCME_CUDA(cudaMemcpyAsync(h_mx, d_mx, sizeof(h_mx), cudaMemcpyDeviceToHost, st));
        NCME_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_mx);
        for (int s = 0; s < ns; ++s) sp->ub[s] = (int64_t)h_mx[s];
    }
    tight = false;
    for (int s = 0; s < ns; ++s) tight |= (uint64_t)(sp->ub[s] + inc[s]) > sp->layout.mask[s];
    if (!tight) return NCME_OK;
    int need[NCME_MAX_SPECIES], total = 0;
    for (int s = 0; s < ns; ++s) {
        need[s] = std::max(1, bitlen((uint64_t)(sp->ub[s] + inc[s])));
        total += need[s];
    }
    if (total > 63) {
        set_error("the reachable states no longer fit a 64-bit packed key (%d species need %d bits)", ns, total);
        return NCME_ERR_KEYWIDTH;
    }
    // hand the spare bits to the species that can still grow, round robin (one doubling each per round)
    int spare = 63 - total;
    while (spare > 0) {
        bool any = false;
        for (int s = 0; s < ns && spare > 0; ++s)
            if (inc[s] > 0 && need[s] < 62) {
                need[s]++;
                spare--;
                any = true;
            }
        if (!any) break;
    }
    KeyLayout to{};
    to.ns = ns;
    int sh = 0;
    for (int s = 0; s < ns; ++s) {
        to.shift[s] = sh;
        to.mask[s] = (1ull << need[s]) - 1;
        sh += need[s];
    }
    if (sp->n > 0) {
        k_repack_keys<<<(unsigned)((sp->n + 255) / 256), 256, 0, st>>>(sp->layout, to, sp->keys.p, sp->n);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
    }
    sp->layout = to;
    sp->relayouts++;
    return space_rebuild_table(sp, 4 * (uint64_t)(sp->n + 256));
}

static int space_alloc(ncme_ctx* ctx, int ns, int nr, const int64_t* stoich, ncme_space** out) {
    NCME_REQUIRE(ctx && out && stoich, "null argument");
    NCME_REQUIRE(ns >= 1 && ns <= NCME_MAX_SPECIES, "species count must be in 1..%d", NCME_MAX_SPECIES);
    NCME_REQUIRE(nr >= 1 && nr <= NCME_MAX_REACTIONS, "reaction count must be in 1..%d", NCME_MAX_REACTIONS);
    ncme_space* sp = new ncme_space();
    sp->ctx = ctx;
    sp->ns = ns;
    sp->nr = nr;
    sp->stoich.assign(stoich, stoich + (size_t)ns * nr);
    const int bits = 63 / ns;  // top bit stays 0 so that EMPTY_KEY is never a valid key
    sp->layout.ns = ns;
    for (int s = 0; s < ns; ++s) {
        sp->layout.shift[s] = s * bits;
        sp->layout.mask[s] = (bits >= 63) ? 0x7FFFFFFFFFFFFFFFull : ((1ull << bits) - 1);
    }
    sp->sdev.nr = nr;
    sp->sdev.ns = ns;
    for (int r = 0; r < nr; ++r)
        for (int s = 0; s < ns; ++s) {
            int64_t v = stoich[(size_t)r * ns + s];
            if (v < -32768 || v > 32767) {
                delete sp;
                set_error("stoichiometry entry out of int16 range");
                return NCME_ERR_ARG;
            }
            sp->sdev.s[r][s] = (int16_t)v;
        }
    cudaError_t e = cudaMalloc(&sp->err_flag, sizeof(int));
    if (e != cudaSuccess) {
        delete sp;
        set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
        return NCME_ERR_CUDA;
    }
    cudaMemsetAsync(sp->err_flag, 0, sizeof(int), ctx->stream);
    *out = sp;
    return NCME_OK;
}

__global__ void k_iota_u32(uint32_t* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void k_count_has_origin(const uint32_t* __restrict__ origin, int64_t n, unsigned long long* __restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned b = __ballot_sync(0xffffffffu, i < n && origin[i] != NONE32);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}

int space_mark(ncme_space* sp) {
    cudaStream_t s = sp->ctx->stream;
    NCME_TRY(sp->origin.reserve((size_t)(sp->ld > 0 ? sp->ld : 1), s, false));
    if (sp->n > 0) {
        k_iota_u32<<<nblk(sp->n), 256, 0, s>>>(sp->origin.p, sp->n);
        sp->ctx->launches++;
        NCME_CUDA(cudaGetLastError());
    }
    sp->mark_n = sp->n;
    sp->mark_id++;
    return NCME_OK;
}

int space_count_kept(ncme_space* sp, int64_t* n_kept) {
    *n_kept = 0;
    if (sp->mark_n < 0 || sp->n == 0) return NCME_OK;
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t s = ctx->stream;
    unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->red_result_dev + 910);
    unsigned long long* h = reinterpret_cast<unsigned long long*>(ctx->red_result_host + 910);
    NCME_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), s));
    k_count_has_origin<<<nblk(sp->n), 256, 0, s>>>(sp->origin.p, sp->n, d);
    ctx->launches++;
    NCME_CUDA(cudaMemcpyAsync(h, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    NCME_CUDA(cudaStreamSynchronize(s));
    *n_kept = (int64_t)*h;
    return NCME_OK;
}

static int check_overflow(ncme_space* sp) {
    int flag = 0;
    NCME_CUDA(cudaMemcpyAsync(&flag, sp->err_flag, sizeof(int), cudaMemcpyDeviceToHost, sp->ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(sp->ctx->stream));
    if (flag) {
        cudaMemsetAsync(sp->err_flag, 0, sizeof(int), sp->ctx->stream);
        set_error("a reachable state does not fit the 64-bit packed key (%d species, %d bits each)", sp->ns, 63 / sp->ns);
        return NCME_ERR_KEYWIDTH;
    }
    return NCME_OK;
}

}  // namespace ncme

using namespace ncme;

extern "C" {

int ncme_space_create(ncme_ctx* ctx, int ns, int nr, const int64_t* stoich, int64_t n0, const int64_t* states0,
                      ncme_space** out) {
    NCME_RANGE("ncme_space_create");
    NCME_REQUIRE(n0 >= 0 && (n0 == 0 || states0), "bad initial state list");
    ncme_space* sp = nullptr;
    NCME_TRY(space_alloc(ctx, ns, nr, stoich, &sp));
    int st = NCME_OK;
    do {
        if ((st = space_reserve_rows(sp, n0 > 1024 ? n0 : 1024)) != NCME_OK) break;
        if ((st = space_rebuild_table(sp, 4 * (uint64_t)(n0 + 256))) != NCME_OK) break;
        if (n0 == 0) break;
        {   // widen the key fields first if an initial state needs it
            int64_t mx[NCME_MAX_SPECIES] = {0}, zero[NCME_MAX_SPECIES] = {0};
            for (int64_t i = 0; i < n0; ++i)
                for (int s2 = 0; s2 < ns; ++s2) mx[s2] = std::max(mx[s2], states0[(size_t)i * ns + s2]);
            for (int s2 = 0; s2 < ns; ++s2) sp->ub[s2] = mx[s2];
            if ((st = space_ensure_key_room(sp, zero)) != NCME_OK) break;
        }
        std::vector<uint64_t> ck((size_t)n0);
        for (int64_t i = 0; i < n0; ++i) {
            int rc = space_pack_host(sp, states0 + (size_t)i * ns, &ck[(size_t)i]);
            if (rc < 0) {
                st = rc;
                break;
            }
            if (rc == 1) ck[(size_t)i] = EMPTY_KEY;  // negative component: silently dropped (:221)
        }
        if (st != NCME_OK) break;
        if ((st = sp->cand_key.reserve((size_t)n0, ctx->stream, false)) != NCME_OK) break;
        if (cudaMemcpyAsync(sp->cand_key.p, ck.data(), (size_t)n0 * sizeof(uint64_t), cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            set_error("upload of initial states failed");
            st = NCME_ERR_CUDA;
            break;
        }
        int64_t added = 0;
        st = space_addstates(sp, n0, &added);
    } while (0);
    if (st != NCME_OK) {
        ncme_space_destroy(sp);
        return st;
    }
    *out = sp;
    return NCME_OK;
}

int ncme_space_from_host(ncme_ctx* ctx, int ns, int nr, const int64_t* stoich, int64_t n, const int64_t* states,
                         const uint32_t* state_connectivity, const uint32_t* sink_connectivity, ncme_space** out) {
    NCME_RANGE("ncme_space_from_host");
    NCME_REQUIRE(n >= 0 && (n == 0 || (states && state_connectivity && sink_connectivity)), "bad arguments");
    ncme_space* sp = nullptr;
    NCME_TRY(space_alloc(ctx, ns, nr, stoich, &sp));
    int st = NCME_OK;
    do {
        if ((st = space_reserve_rows(sp, n > 1024 ? n : 1024)) != NCME_OK) break;
        {
            int64_t mx[NCME_MAX_SPECIES] = {0}, zero[NCME_MAX_SPECIES] = {0};
            for (int64_t i = 0; i < n; ++i)
                for (int s2 = 0; s2 < ns; ++s2) mx[s2] = std::max(mx[s2], states[(size_t)i * ns + s2]);
            for (int s2 = 0; s2 < ns; ++s2) sp->ub[s2] = mx[s2];
            if ((st = space_ensure_key_room(sp, zero)) != NCME_OK) break;
        }
        std::vector<uint64_t> hk((size_t)n);
        std::vector<smask_t> hm((size_t)n, 0);
        std::vector<uint32_t> hp((size_t)n * nr);
        for (int64_t i = 0; i < n && st == NCME_OK; ++i) {
            int rc = space_pack_host(sp, states + (size_t)i * ns, &hk[(size_t)i]);
            if (rc != 0) {
                if (rc == 1) set_error("negative state component in imported state space");
                st = rc < 0 ? rc : NCME_ERR_ARG;
            }
            for (int r = 0; r < nr; ++r) {
                uint32_t c = state_connectivity[(size_t)i * nr + r];
                if (c > (uint32_t)n) {
                    set_error("state_connectivity entry out of range");
                    st = NCME_ERR_ARG;
                }
                hp[(size_t)r * n + i] = c ? c - 1 : NONE32;
                if (sink_connectivity[(size_t)i * nr + r]) hm[(size_t)i] |= SMASK1(r);
            }
        }
        if (st != NCME_OK) break;
        cudaStream_t s = ctx->stream;
        bool ok = true;
        if (n > 0) {
            ok &= cudaMemcpyAsync(sp->keys.p, hk.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s) == cudaSuccess;
            ok &= cudaMemcpyAsync(sp->sinkmask.p, hm.data(), (size_t)n * sizeof(smask_t), cudaMemcpyHostToDevice, s) == cudaSuccess;
            for (int r = 0; r < nr; ++r)
                ok &= cudaMemcpyAsync(sp->pred.p + (size_t)r * sp->ld, hp.data() + (size_t)r * n, (size_t)n * 4,
                                      cudaMemcpyHostToDevice, s) == cudaSuccess;
            ok &= cudaStreamSynchronize(s) == cudaSuccess;
        }
        if (!ok) {
            set_error("upload of imported state space failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = NCME_ERR_CUDA;
            break;
        }
        sp->n = n;
        st = space_rebuild_table(sp, 4 * (uint64_t)(n + 256));
    } while (0);
    if (st != NCME_OK) {
        ncme_space_destroy(sp);
        return st;
    }
    *out = sp;
    return NCME_OK;
}

int ncme_space_destroy(ncme_space* sp) {
    if (!sp) return NCME_OK;
    if (sp->ctx) cudaStreamSynchronize(sp->ctx->stream);
    sp->keys.release();
    sp->pred.release();
    sp->sinkmask.release();
    sp->tkeys.release();
    sp->tvals.release();
    sp->cand_key.release();
    sp->cand_slot.release();
    sp->flags.release();
    sp->pos.release();
    sp->scan_scratch.release();
    sp->frontier.release();
    sp->origin.release();
    if (sp->err_flag) cudaFree(sp->err_flag);
    delete sp;
    return NCME_OK;
}

// small spaces: run as many levels as fit the (generously pre-reserved) capacities in one single-CTA launch.
// Returns the levels done and the size of the last level's batch of new states (the next frontier).
static int expand_small(ncme_space* sp, int levels, const ReactList& reacts, smask_t reactmask, const int64_t* inc,
                        int* levels_done, int64_t* last_added, bool* frontier_empty) {
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t s = ctx->stream;
    *levels_done = 0;
    *last_added = 0;
    *frontier_empty = false;
    constexpr int64_t SMALL_N = 1 << 16, CAND_CAP = 1 << 16;
    while (*levels_done < levels && sp->n <= SMALL_N) {
        // how many levels can the key fields take without a re-layout?
        int64_t safe = levels - *levels_done;
        for (int q = 0; q < sp->ns; ++q)
            if (inc[q] > 0) safe = std::min<int64_t>(safe, ((int64_t)sp->layout.mask[q] - sp->ub[q]) / inc[q]);
        if (safe <= 0) {
            NCME_TRY(space_ensure_key_room(sp, inc));
            safe = 1;
        }
        const int64_t want_rows = 2 * sp->n + 8192;
        NCME_TRY(space_reserve_rows(sp, want_rows));
        if (4 * (uint64_t)sp->ld > sp->tcap) NCME_TRY(space_rebuild_table(sp, 4 * (uint64_t)sp->ld));
        NCME_TRY(sp->cand_key.reserve((size_t)CAND_CAP, s, false));
        NCME_TRY(sp->cand_slot.reserve((size_t)CAND_CAP, s, false));
        NCME_TRY(sp->frontier.reserve((size_t)sp->ld, s, false));
        ExpandSmallArgs a{};
        a.h = sp->hview();
        a.L = sp->layout;
        a.S = sp->sdev;
        a.reacts = reacts;
        a.reactmask = reactmask;
        a.keys = sp->keys.p;
        a.pred = sp->pred.p;
        a.ld = sp->ld;
        a.sinkmask = sp->sinkmask.p;
        a.cand_key = sp->cand_key.p;
        a.cand_slot = sp->cand_slot.p;
        a.frontier = sp->frontier.p;
        a.n0 = sp->n;
        a.cand_cap = CAND_CAP;
        a.tcap = sp->tcap;
        a.levels = (int)safe;
        a.err_flag = sp->err_flag;
        long long* out_dev = reinterpret_cast<long long*>(ctx->red_result_dev);
        long long* out_host = reinterpret_cast<long long*>(ctx->red_result_host);
        a.out = out_dev;
        k_expand_small<<<1, XS_THREADS, 0, s>>>(a);
        ctx->launches++;
        NCME_CUDA(cudaGetLastError());
        NCME_CUDA(cudaMemcpyAsync(out_host, out_dev, 4 * sizeof(long long), cudaMemcpyDeviceToHost, s));
        NCME_CUDA(cudaStreamSynchronize(s));
        const int done = (int)out_host[1];
        if (out_host[0] != sp->n) sp->version++;
        sp->n = out_host[0];
        for (int q = 0; q < sp->ns; ++q) sp->ub[q] += inc[q] * done;
        *levels_done += done;
        if (done > 0) *last_added = out_host[2];
        if (done > 0 && out_host[2] == 0) {   // the frontier died out
            *frontier_empty = true;
            return NCME_OK;
        }
        if (done == 0 && !out_host[3]) {      // nothing to explore at all
            *frontier_empty = true;
            return NCME_OK;
        }
        if (out_host[3] && done == 0) break;  // a single level exceeds the small-path capacities: general path
    }
    return NCME_OK;
}

int ncme_space_expand(ncme_space* sp, int expansionlevel, int nonly, const int32_t* onlyreactions) {
    NCME_RANGE("ncme_space_expand");
    NCME_REQUIRE(sp, "null space");
    if (expansionlevel <= 0 || sp->n == 0) return NCME_OK;
    ncme_ctx* ctx = sp->ctx;
    sp->last_delete_nold = -1;
    ReactList reacts{};
    smask_t reactmask = 0;
    if (nonly <= 0) {
        for (int r = 0; r < sp->nr; ++r) reacts.r[reacts.n++] = r;
    } else {
        NCME_REQUIRE(onlyreactions && nonly <= NCME_MAX_REACTIONS, "bad onlyreactions");
        for (int k = 0; k < nonly; ++k) {
            NCME_REQUIRE(onlyreactions[k] >= 1 && onlyreactions[k] <= sp->nr, "onlyreactions entry out of range");
            reacts.r[reacts.n++] = onlyreactions[k] - 1;
        }
    }
    const int nreact = reacts.n;
    for (int k = 0; k < nreact; ++k) reactmask |= SMASK1(reacts.r[k]);
    int64_t inc[NCME_MAX_SPECIES] = {0};   // largest possible growth of each species in one level
    for (int k = 0; k < nreact; ++k)
        for (int s2 = 0; s2 < sp->ns; ++s2) inc[s2] = std::max(inc[s2], sp->stoich[(size_t)reacts.r[k] * sp->ns + s2]);
    cudaStream_t s = ctx->stream;
    const int64_t n_at_entry = sp->n;
    int st = NCME_OK;
    do {
        int level0 = 0;
        int64_t last_added = 0;
        bool dead = false;
        if ((st = expand_small(sp, expansionlevel, reacts, reactmask, inc, &level0, &last_added, &dead)) != NCME_OK) break;
        if (dead || level0 >= expansionlevel) {
            st = check_overflow(sp);
            break;
        }
        // ---- general path (one kernel pipeline per level)
        const uint32_t* frontier = nullptr;
        int64_t fbase = 0;
        uint64_t F = 0;
        if (level0 > 0) {   // continue from the last batch of new states
            fbase = sp->n - last_added;
            F = (uint64_t)last_added;
        } else {
            // explorables = states with a sink flag on one of the expansion reactions (:165-173)
            if ((st = sp->flags.reserve((size_t)sp->n, s, false)) != NCME_OK) break;
            if ((st = sp->pos.reserve((size_t)sp->n, s, false)) != NCME_OK) break;
            if ((st = sp->scan_scratch.reserve(scan_scratch_elems(sp->n), s, false)) != NCME_OK) break;
            k_frontier_flags<<<nblk(sp->n), 256, 0, s>>>(sp->sinkmask.p, sp->n, reactmask, sp->flags.p);
            ctx->launches++;
            if ((st = exclusive_scan_u32(ctx, sp->flags.p, sp->pos.p, sp->n, sp->scan_scratch.p, sp->scan_scratch.cap, &F)) != NCME_OK)
                break;
            if (F == 0) break;
            if ((st = sp->frontier.reserve((size_t)F, s, false)) != NCME_OK) break;
            k_compact_indices<<<nblk(sp->n), 256, 0, s>>>(sp->flags.p, sp->pos.p, sp->n, sp->frontier.p);
            ctx->launches++;
            frontier = sp->frontier.p;
        }
        for (int level = level0; level < expansionlevel && F > 0; ++level) {
            if ((st = space_ensure_key_room(sp, inc)) != NCME_OK) break;
            for (int s2 = 0; s2 < sp->ns; ++s2) sp->ub[s2] += inc[s2];
            const int64_t ncand = (int64_t)F * nreact;
            if ((st = sp->cand_key.reserve((size_t)ncand, s, false)) != NCME_OK) break;
            k_gen_candidates<<<nblk(ncand), 256, 0, s>>>(sp->layout, sp->sdev, sp->keys.p, frontier, fbase, (int64_t)F,
                                                         nreact, reacts, sp->cand_key.p, sp->err_flag);
            ctx->launches++;
            int64_t added = 0;
            const int64_t n_before = sp->n;
            if ((st = space_addstates(sp, ncand, &added)) != NCME_OK) break;
            frontier = nullptr;
            fbase = n_before;
            F = (uint64_t)added;
        }
        if (st != NCME_OK) break;
        if (cudaGetLastError() != cudaSuccess) {
            set_error("kernel launch failed in expand");
            st = NCME_ERR_CUDA;
            break;
        }
        st = check_overflow(sp);
    } while (0);
    if (st == NCME_OK && sp->mark_n >= 0 && sp->n > n_at_entry) {   // states added by this expansion have no origin
        if ((st = sp->origin.reserve((size_t)sp->ld, s)) == NCME_OK) {
            k_fill_u32<<<nblk(sp->n - n_at_entry), 256, 0, s>>>(sp->origin.p + n_at_entry, sp->n - n_at_entry, NONE32);
            ctx->launches++;
        }
    }
    cudaStreamSynchronize(ctx->stream);
    return st;
}

int ncme_space_delete(ncme_space* sp, int64_t nids, const int64_t* ids) {
    NCME_RANGE("ncme_space_delete");
    NCME_REQUIRE(sp && (nids == 0 || ids), "bad arguments");
    if (nids <= 0 || sp->n == 0) return NCME_OK;
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t s = ctx->stream;
    const int64_t n = sp->n;
    std::vector<uint32_t> ids0((size_t)nids);
    for (int64_t k = 0; k < nids; ++k) {
        NCME_REQUIRE(ids[k] >= 1 && ids[k] <= n, "deleteat!: index %lld out of range 1..%lld", (long long)ids[k], (long long)n);
        ids0[(size_t)k] = (uint32_t)(ids[k] - 1);
    }
    NCME_TRY(sp->flags.reserve((size_t)n, s, false));
    NCME_TRY(sp->cand_slot.reserve((size_t)nids, s, false));
    NCME_CUDA(cudaMemcpyAsync(sp->cand_slot.p, ids0.data(), (size_t)nids * 4, cudaMemcpyHostToDevice, s));
    LAUNCH(ctx, k_fill_u32, n, sp->flags.p, n, 1u);
    LAUNCH(ctx, k_clear_flags_at, nids, sp->cand_slot.p, nids, sp->flags.p);
    NCME_CUDA(cudaStreamSynchronize(s));  // ids0 is a host temporary
    return space_delete_flagged(sp);
}

int ncme_space_compact_vector(ncme_space* sp, const double* in_dev, double* out_dev) {
    NCME_REQUIRE(sp && in_dev && out_dev && in_dev != out_dev, "bad arguments");
    NCME_REQUIRE(sp->last_delete_nold >= 0, "no deletion to apply (the space was mutated since the last delete/prune)");
    LAUNCH(sp->ctx, k_compact_vector, sp->last_delete_nold, sp->flags.p, sp->pos.p, sp->last_delete_nold, in_dev, out_dev);
    return NCME_OK;
}

int ncme_space_state_count(ncme_space* sp, int64_t* n) {
    NCME_REQUIRE(sp && n, "null argument");
    *n = sp->n;
    return NCME_OK;
}

int ncme_space_sink_count(ncme_space* sp, int64_t* r) {
    NCME_REQUIRE(sp && r, "null argument");
    *r = sp->nr;
    return NCME_OK;
}

int ncme_space_download_states(ncme_space* sp, int64_t first, int64_t count, int64_t* out) {
    NCME_REQUIRE(sp && first >= 0 && count >= 0 && first + count <= sp->n && (count == 0 || out), "bad range");
    if (count == 0) return NCME_OK;
    std::vector<uint64_t> hk((size_t)count);
    NCME_CUDA(cudaMemcpyAsync(hk.data(), sp->keys.p + first, (size_t)count * 8, cudaMemcpyDeviceToHost, sp->ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(sp->ctx->stream));
    for (int64_t i = 0; i < count; ++i)
        for (int s = 0; s < sp->ns; ++s)
            out[(size_t)i * sp->ns + s] = (int64_t)((hk[(size_t)i] >> sp->layout.shift[s]) & sp->layout.mask[s]);
    return NCME_OK;
}

int ncme_space_download_state_columns(ncme_space* sp, int64_t first, int64_t count, double* out) {
    NCME_REQUIRE(sp && first >= 0 && count >= 0 && first + count <= sp->n && (count == 0 || out), "bad range");
    if (count == 0) return NCME_OK;
    std::vector<uint64_t> hk((size_t)count);
    NCME_CUDA(cudaMemcpyAsync(hk.data(), sp->keys.p + first, (size_t)count * 8, cudaMemcpyDeviceToHost, sp->ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(sp->ctx->stream));
    for (int s = 0; s < sp->ns; ++s) {
        const int sh = sp->layout.shift[s];
        const uint64_t mk = sp->layout.mask[s];
        double* col = out + (size_t)s * count;
        for (int64_t i = 0; i < count; ++i) col[i] = (double)((hk[(size_t)i] >> sh) & mk);
    }
    return NCME_OK;
}

int ncme_space_download_connectivity(ncme_space* sp, int64_t first, int64_t count, uint32_t* sc_out, uint32_t* kc_out) {
    NCME_REQUIRE(sp && first >= 0 && count >= 0 && first + count <= sp->n, "bad range");
    if (count == 0) return NCME_OK;
    cudaStream_t s = sp->ctx->stream;
    const int nr = sp->nr;
    std::vector<uint32_t> hp((size_t)count * nr);
    std::vector<smask_t> hm((size_t)count);
    for (int r = 0; r < nr; ++r)
        NCME_CUDA(cudaMemcpyAsync(hp.data() + (size_t)r * count, sp->pred.p + (size_t)r * sp->ld + first, (size_t)count * 4,
                                  cudaMemcpyDeviceToHost, s));
    NCME_CUDA(cudaMemcpyAsync(hm.data(), sp->sinkmask.p + first, (size_t)count * sizeof(smask_t), cudaMemcpyDeviceToHost, s));
    NCME_CUDA(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < count; ++i)
        for (int r = 0; r < nr; ++r) {
            uint32_t p = hp[(size_t)r * count + i];
            if (sc_out) sc_out[(size_t)i * nr + r] = (p == NONE32) ? 0u : p + 1u;
            if (kc_out) kc_out[(size_t)i * nr + r] = ((hm[(size_t)i] >> r) & 1u) ? (uint32_t)(r + 1) : 0u;
        }
    return NCME_OK;
}

int ncme_space_lookup(ncme_space* sp, int64_t m, const int64_t* states, uint32_t* idx_out) {
    NCME_REQUIRE(sp && m >= 0 && (m == 0 || (states && idx_out)), "bad arguments");
    if (m == 0) return NCME_OK;
    ncme_ctx* ctx = sp->ctx;
    std::vector<uint64_t> q((size_t)m);
    for (int64_t i = 0; i < m; ++i) {
        uint64_t k = EMPTY_KEY;
        bool fits = true;
        for (int s = 0; s < sp->ns; ++s) {
            int64_t v = states[(size_t)i * sp->ns + s];
            if (v < 0 || (uint64_t)v > sp->layout.mask[s]) fits = false;
        }
        if (fits) space_pack_host(sp, states + (size_t)i * sp->ns, &k);
        q[(size_t)i] = k;
    }
    NCME_TRY(sp->cand_key.reserve((size_t)m, ctx->stream, false));
    NCME_TRY(sp->cand_slot.reserve((size_t)m, ctx->stream, false));
    NCME_CUDA(cudaMemcpyAsync(sp->cand_key.p, q.data(), (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_lookup, m, sp->hview(), sp->cand_key.p, m, sp->cand_slot.p);
    NCME_CUDA(cudaMemcpyAsync(idx_out, sp->cand_slot.p, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    NCME_CUDA(cudaStreamSynchronize(ctx->stream));
    return NCME_OK;
}

int ncme_space_marginal(ncme_space* sp, const double* p_dev, int ndims, const int32_t* dims, int64_t cap, int64_t* nred,
                        int64_t* states_out, double* vals_out) {
    NCME_RANGE("ncme_space_marginal");
    NCME_REQUIRE(sp && p_dev && nred && ndims >= 1 && dims, "null argument");
    uint32_t drop = 0;
    for (int k = 0; k < ndims; ++k) {
        NCME_REQUIRE(dims[k] >= 1 && dims[k] <= sp->ns, "Input dimensions must be between 1 and %d.", sp->ns);
        drop |= 1u << (dims[k] - 1);
    }
    int keep[NCME_MAX_SPECIES], nkeep = 0;
    uint64_t keepmask = 0;
    for (int q = 0; q < sp->ns; ++q)
        if (!(drop & (1u << q))) {
            keep[nkeep++] = q;
            keepmask |= sp->layout.mask[q] << sp->layout.shift[q];
        }
    const int64_t n = sp->n;
    *nred = 0;
    if (n == 0) return NCME_OK;
    ncme_ctx* ctx = sp->ctx;
    cudaStream_t s = ctx->stream;
    sp->last_delete_nold = -1;   // flags/pos scratch is reused
    uint64_t tcap = 1024;
    while (tcap < 2 * (uint64_t)n) tcap <<= 1;
    DevArray<uint64_t> tk, rkeys;
    DevArray<uint32_t> tv, slot_of;
    DevArray<double> rvals;
    NCME_TRY(tk.reserve(tcap, s, false));
    NCME_TRY(tv.reserve(tcap, s, false));
    NCME_TRY(slot_of.reserve((size_t)n, s, false));
    NCME_TRY(sp->flags.reserve((size_t)n, s, false));
    NCME_TRY(sp->pos.reserve((size_t)n, s, false));
    NCME_TRY(sp->scan_scratch.reserve(scan_scratch_elems(n), s, false));
    NCME_CUDA(cudaMemsetAsync(tk.p, 0xFF, tcap * sizeof(uint64_t), s));
    NCME_CUDA(cudaMemsetAsync(tv.p, 0xFF, tcap * sizeof(uint32_t), s));
    const HashView h{tk.p, tv.p, tcap - 1};
    int rc = NCME_OK;
    do {
        k_marg_insert<<<nblk(n), 256, 0, s>>>(h, sp->keys.p, n, keepmask, slot_of.p);
        k_marg_flags<<<nblk(n), 256, 0, s>>>(h, slot_of.p, n, sp->flags.p);
        ctx->launches += 2;
        uint64_t m = 0;
        if ((rc = exclusive_scan_u32(ctx, sp->flags.p, sp->pos.p, n, sp->scan_scratch.p, sp->scan_scratch.cap, &m)) != NCME_OK) break;
        *nred = (int64_t)m;
        if (cap < (int64_t)m || !states_out || !vals_out) break;   // size query (or too small): count only
        if ((rc = rkeys.reserve((size_t)m, s, false)) != NCME_OK) break;
        if ((rc = rvals.reserve((size_t)m, s, false)) != NCME_OK) break;
        cudaMemsetAsync(rvals.p, 0, (size_t)m * sizeof(double), s);
        k_marg_accumulate<<<nblk(n), 256, 0, s>>>(h, slot_of.p, sp->flags.p, sp->pos.p, sp->keys.p, keepmask, p_dev, n, rkeys.p,
                                                  rvals.p);
        ctx->launches++;
        std::vector<uint64_t> hk((size_t)m);
        if (cudaMemcpyAsync(hk.data(), rkeys.p, (size_t)m * 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaMemcpyAsync(vals_out, rvals.p, (size_t)m * 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            set_error("marginal: %s", cudaGetErrorString(cudaGetLastError()));
            rc = NCME_ERR_CUDA;
            break;
        }
        for (uint64_t j = 0; j < m; ++j)
            for (int q = 0; q < nkeep; ++q)
                states_out[(size_t)j * nkeep + q] = (int64_t)((hk[(size_t)j] >> sp->layout.shift[keep[q]]) & sp->layout.mask[keep[q]]);
    } while (0);
    tk.release();
    tv.release();
    slot_of.release();
    rkeys.release();
    rvals.release();
    return rc;
}

}  // extern "C"
