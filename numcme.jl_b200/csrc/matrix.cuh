// Device layout of FspMatrixSparse (reference: src/fspmatrix/sparse/fspsparsematrix.jl:9-27).
//
// The reference keeps one CSC matrix per term (summed time-invariant, one per separable reaction,
// one per joint reaction).  Here all terms live in ONE slot-major ELL structure that a single
// kernel streams once:
//
//   y_i = sum_s  c_s(t) * val[s][i] * x[col[s][i]]  +  ( sum_d cd_d(t) * diag[d][i] ) * x_i        i < n
//   y_{n+r} = c_r(t) * sum_{k in sinks(r)} sink_val[k] * x[sink_row[k]]                            r < R
//
// slot s   = one reaction (time-invariant reactions with identical stoichiometry share a slot, which is
//            exactly the duplicate-summing of Julia's sparse(), :74).  col[s][i] is the predecessor of
//            state i through the slot's stoichiometry (state_connectivity, sparsestatespace.jl:35) or i
//            itself with val = 0 where there is none.  val[s][i] = state factor at the predecessor.
// diag d   = d = 0: minus the sum of all time-invariant state factors (if any); then one per
//            time-varying reaction (minus its state factor / joint value).
// Arrays are slot-major with row stride ld (multiple of 64 rows => every slot is 256-byte aligned),
// so a warp reads 32*ROWS consecutive values per slot: fully coalesced, vectorisable.
#pragma once
#include "space.cuh"

struct ncme_comm;

struct ncme_matrix {
    ncme_ctx* ctx = nullptr;
    int ns = 0, nr = 0;
    int64_t n = 0, N = 0, ld = 0;   // n = rows held by this rank, N = n + nr (local vector length)
    int kind[NCME_MAX_REACTIONS] = {0};

    // row sharding (K8).  Single GPU: comm == nullptr, row_lo = 0, row_hi = n_global = n, no halo.
    // x vectors are addressed through xb = x_local - hl: [halo_lo (hl) | local rows (n) | sinks (nr) | halo_hi (hh)]
    ncme_comm* comm = nullptr;
    int64_t n_global = 0, row_lo = 0, row_hi = 0, ext_lo = 0, ext_hi = 0, hl = 0, hh = 0;
    int64_t b0 = 0, b1 = 0;         // rows [b0, b1) touch no halo entry (computed while the halo is in flight)
    struct HaloSeg {
        int peer;
        int64_t offset;  // doubles, relative to x_local (may be negative for the low halo)
        int64_t count;
    };
    std::vector<HaloSeg> halo_send, halo_recv;
    // peer-memory halo (CUDA IPC): possible when each halo side lies inside ONE neighbouring rank's block
    bool p2p_eligible = false;        // same value on every rank
    int plo = -1, phi = -1;           // ranks that own my low / high halo (-1: no halo on that side)
    int64_t plo_row_lo = 0, phi_row_lo = 0;
    std::vector<int> readers;         // ranks whose halo lies in my block (they read my vectors)

    int nslots = 0;
    int slot_coef_src[NCME_MAX_REACTIONS] = {0};   // reaction whose time factor scales the slot, -1 => 1.0
    int slot_first_reaction[NCME_MAX_REACTIONS] = {0};
    int reaction_slot[NCME_MAX_REACTIONS] = {0};   // slot holding reaction r (-1: zero-stoichiometry, contributes nothing)
    int ndiag = 0;
    int diag_coef_src[NCME_MAX_REACTIONS + 1] = {0};
    int reaction_diag[NCME_MAX_REACTIONS] = {0};   // diag array of reaction r

    ncme::DevArray<uint32_t> col;   // [nslots][ld]
    ncme::DevArray<double> val;     // [nslots][ld]
    ncme::DevArray<double> diag;    // [ndiag][ld]
    // state factors G[r][i] of ALL states (reaction-major, stride n_global), kept for the incremental constructor of the
    // matrix that follows an adapt! (H8): only the states added since are evaluated on the host
    ncme::DevArray<double> G;
    // origin map the incremental constructor used (old index of the surviving states, a prefix of length carry_nkept) and
    // the matrix it carried the factors from (identity only, never dereferenced): the sensitivity matrix built on top
    // of this matrix carries its derivative factors over through the same map (ncme_sensmatrix_create_incremental)
    ncme::DevArray<uint32_t> carry_origin;
    int64_t carry_nkept = -1;
    const ncme_matrix* carry_prev = nullptr;
    bool g_window = false;          // G only holds the factors of this rank's rows + halo window (ncme_matrix_create_window)
    uint64_t space_mark = 0;        // mark of the space this matrix was built at
    // compressed column indices (K1 fast path): one byte per (slot,row) relative to a per-(slot, 64-row chunk)
    // descriptor {base:i32, mode:u32}; mode 0 = chunk stays on the 32-bit array (range does not fit)
    ncme::DevArray<uint8_t> col8;     // [nslots][ld]
    ncme::DevArray<int2> cdesc;       // [nslots][ld/64]
    int64_t nchunks = 0;
    int64_t wide_chunks = 0;          // chunk-slots that fell back to 32-bit indices
    int use_c8 = 0;                   // off by default: measured slower than 32-bit indices (profiles/README.md)
    int use_pipe = 0;                 // shared-memory pipelined kernel: 0 off, 1 / 2 = rows per thread
    int64_t pipe_min_rows = 1 << 18;  // below this the register-staged kernel is used

    // sink rows: entries grouped by reaction, rows ascending inside a reaction
    int64_t nsink = 0;
    int64_t sink_ptr[NCME_MAX_REACTIONS + 1] = {0};
    ncme::DevArray<uint32_t> sink_row;
    ncme::DevArray<double> sink_val;
    int ntasks = 0;
    int task_ptr[NCME_MAX_REACTIONS + 1] = {0};
    ncme::DevArray<int4> tasks;          // (reaction, begin, end, -)
    ncme::DevArray<double> sink_partial; // [ntasks]
    unsigned int* sink_counter = nullptr;
    ncme::DevArray<unsigned int> sink_counter_mem;

    // host-buffer pipeline: row chunks and, per chunk, the last row index its gathers reach (+1)
    int pipe_chunks = 0;
    int64_t pipe_row[17] = {0};
    int64_t pipe_need_hi[16] = {0};

    // reference-structure statistics
    int nterms = 0;
    int64_t nnz_term[NCME_MAX_REACTIONS + 1] = {0};
    int64_t algorithmic_bytes = 0;
    int64_t npred_r[NCME_MAX_REACTIONS] = {0};   // rows with a predecessor through reaction r (off-diagonal entries)

    // launch tuning (experiments): rows per thread (0 = auto), threads per block
    int tune_rows = 0;
    int tune_block = 256;
};

namespace ncme {

struct MatvecArgs {
    const uint32_t* col;
    const double* val;
    const double* diag;
    int64_t n, ld;
    int nslots, ndiag, nr;
    double slot_coef[NCME_MAX_REACTIONS];
    double diag_coef[NCME_MAX_REACTIONS + 1];
    double sink_coef[NCME_MAX_REACTIONS];
    const uint32_t* sink_row;
    const double* sink_val;
    const int4* tasks;
    int ntasks;
    int task_ptr[NCME_MAX_REACTIONS + 1];
    double* sink_partial;
    unsigned int* sink_counter;
    const uint8_t* col8;
    const int2* cdesc;
    int64_t nchunks;
    uint32_t self_off;  // padded position of row i is self_off + i
    const double* x;    // xb: base of the halo-padded input (== xd on a single GPU)
    const double* xd;   // x_local: entry of row i is xd[i]
    double* y;
    double beta;
    const double* x_lo;   // P2P launches: x_lo[c] for padded positions c < lo_end (peer memory over NVLink)
    const double* x_hi;   //               x_hi[c] for c >= hi_begin
    uint32_t lo_end, hi_begin;
    int nsig;                          // sharded: CTA 0 first publishes `epoch` into these peer flags ("my x is ready")
    unsigned int* sig_flag[4];
    int nwait;                         // P2P launches: every CTA first waits until these flags reach `epoch`
    const unsigned int* wait_flag[4];
    unsigned int epoch;
    unsigned int* err_flag;
    int64_t row_begin2, row_end2;      // P2P launches: optional second row range (high boundary rows)
    int64_t row_begin, row_end;  // rows handled by this launch
    int do_sinks;                // 1: the sink-task CTAs run in this launch
    // one-launch sharded matvec (k_fsp_matvec_sharded): halo-free rows [int_begin, int_end) + both boundary ranges
    int64_t int_begin, int_end;
    int nb_int, nb_bd;                 // CTAs of the halo-free rows / of the boundary rows
    int bd_first;                      // 1: the boundary CTAs take the lowest block indices (scheduled first)
    int ndone_sig;                     // "done" handshake folded into the launch: the last boundary CTA publishes
    unsigned int* done_sig[4];         // `epoch` into these peer flags ("I no longer read your x") and then waits
    int ndone_wait;                    // until these local flags reach `epoch` ("nobody reads my x any more")
    const unsigned int* done_wait[4];
    unsigned int* bd_counter;
};

int matvec_fill_args(const ncme_matrix* A, const double* coef, MatvecArgs* a);
int matvec_launch(ncme_matrix* A, const MatvecArgs& a);
// y_local = A(t) x_local for a (possibly sharded) matrix.  Sharded: exchanges the halo of x in place (x_local must
// have hl doubles of margin before it and hh after its nr sink entries); the nr sink entries of y hold this rank's
// partial sums unless reduce_sinks != 0.
// reduce_sinks: bit 0 = all-reduce the sink entries; bit 1 = the caller alternates input buffers between consecutive
// matvecs (the integrator does), so the peer-memory path may skip the "done" handshake.
int matvec_dist(ncme_matrix* A, const double* coef, const double* x_local, double* y_local, double beta, int reduce_sinks);
int halo_exchange(ncme_matrix* A, const double* x_local, cudaStream_t st);
int matrix_diag(ncme_matrix* A, const double* coef, double* out);                                 // out[0..n) = diag(A(t))
int matvec_sinks_only(ncme_matrix* A, const double* coef, const double* x_local, double* y_local);  // the nr sink rows only
int sens_describe(ncme_sensmatrix* SA, ncme_matrix** A, int* npar, int* nent);
// out[r][i] = i < nkept ? prev[r][origin[i]] : tail[r][i - nkept]   (nrows arrays of length ng; prev has stride ngprev)
int carry_rows(ncme_ctx* ctx, const double* prev, int64_t ngprev, const uint32_t* origin, int64_t nkept, const double* tail,
               int64_t nnew, int nrows, double* out, int64_t ng);

}  // namespace ncme
