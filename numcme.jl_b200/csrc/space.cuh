// Device-resident StateSpaceSparse (reference: src/statespace/sparse/sparsestatespace.jl:22-40).
#pragma once
#include "common.cuh"

namespace ncme {

// How a state (NS non-negative integers) is packed into one 64-bit hash key.
struct KeyLayout {
    int ns;
    int shift[NCME_MAX_SPECIES];
    uint64_t mask[NCME_MAX_SPECIES];
};

// Net stoichiometry, reaction-major, small integers.
struct StoichDev {
    int nr;
    int ns;
    int16_t s[NCME_MAX_REACTIONS][NCME_MAX_SPECIES];
};

// reactions an expansion goes through (0-based), passed by value to the kernels
struct ReactList {
    int n;
    int r[NCME_MAX_REACTIONS];
};

__host__ __device__ inline uint64_t hash64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

struct HashView {
    uint64_t* keys;
    uint32_t* vals;
    uint64_t capmask;  // capacity - 1 (capacity is a power of two)
};

__device__ inline uint32_t hash_lookup(const HashView& h, uint64_t key) {
    uint64_t slot = hash64(key) & h.capmask;
    while (true) {
        uint64_t kk = h.keys[slot];
        if (kk == key) return h.vals[slot];
        if (kk == EMPTY_KEY) return NONE32;
        slot = (slot + 1) & h.capmask;
    }
}

}  // namespace ncme

struct ncme_space {
    ncme_ctx* ctx = nullptr;
    int ns = 0, nr = 0;
    std::vector<int64_t> stoich;  // reaction-major: stoich[r*ns + s]  (== Julia column-major Matrix)
    ncme::KeyLayout layout{};
    ncme::StoichDev sdev{};
    int64_t n = 0;    // number of states
    int64_t ld = 0;   // row stride of the slot-major pred table (capacity)
    uint64_t version = 0;  // bumped by every mutation

    ncme::DevArray<uint64_t> keys;      // [ld]      packed state of index i (insertion order)
    ncme::DevArray<uint32_t> pred;      // [nr][ld]  pred[r][i] = j with x_i = x_j + s_r, NONE32 if absent
    ncme::DevArray<ncme::smask_t> sinkmask;  // [ld]      bit r set <=> x_i + s_r >= 0 and not in the space
    // incremental matrix assembly (H8): origin[i] = index state i had when the space was last marked (by a matrix
    // build), NONE32 for states added since.  Deletions compact it with the states, expansions append NONE32 -- the
    // surviving old states therefore always form a prefix and the new ones the tail.
    ncme::DevArray<uint32_t> origin;    // [ld]
    uint64_t mark_id = 0;               // bumped by every mark; a matrix remembers the mark it was built at
    int64_t mark_n = -1;                // number of states at the mark (-1: never marked)

    ncme::DevArray<uint64_t> tkeys;     // open-addressing table
    ncme::DevArray<uint32_t> tvals;
    uint64_t tcap = 0;

    // scratch for expansion / deletion
    ncme::DevArray<uint64_t> cand_key;
    ncme::DevArray<uint32_t> cand_slot;
    ncme::DevArray<uint32_t> flags;
    ncme::DevArray<uint32_t> pos;
    ncme::DevArray<uint32_t> scan_scratch;
    ncme::DevArray<uint32_t> frontier;
    int* err_flag = nullptr;  // device int: key-width overflow seen
    int relayouts = 0;
    int64_t ub[NCME_MAX_SPECIES] = {0};  // host upper bound of every species count present in the space (key re-layout)
    int64_t last_delete_nold = -1;  // >= 0: flags/pos hold the keep flags / compaction map of the last deletion
    int64_t last_delete_nnew = 0;

    ncme::HashView hview() const { return ncme::HashView{tkeys.p, tvals.p, tcap - 1}; }
};

namespace ncme {
// Append candidate states (packed keys in cand_key[0..ncand), EMPTY_KEY = invalid) in candidate
// order, first occurrence wins; updates connectivity.  Returns number added through *added.
int space_addstates(ncme_space* sp, int64_t ncand, int64_t* added);
int space_reserve_rows(ncme_space* sp, int64_t nrows);
int space_delete_flagged(ncme_space* sp);  // keep flags in sp->flags[0..n)
int space_rebuild_table(ncme_space* sp, uint64_t min_slots);
// make sure every species can grow by inc[s] without overflowing its key field (re-packs all keys if needed)
int space_ensure_key_room(ncme_space* sp, const int64_t* inc);
int space_pack_host(const ncme_space* sp, const int64_t* state, uint64_t* key_out);  // 0 ok, 1 negative, <0 error
int space_mark(ncme_space* sp);                          // origin = identity; returns through sp->mark_id
int space_count_kept(ncme_space* sp, int64_t* n_kept);   // states that already existed at the last mark (a prefix)
}  // namespace ncme
