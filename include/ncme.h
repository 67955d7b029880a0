/* ncme.h -- C ABI of libncme: the B200-native FSP right-hand-side path of NumCME.jl.
 *
 * Every entry point is what a Julia `ccall` (or Python ctypes) binding for the corresponding
 * reference function would bind.  Citations are file:line under the reference repository
 * (voduchuy/NumCME.jl v0.1.4).  See INTEGRATION.md for the Julia-side glue.
 *
 * Conventions
 *   - All functions return 0 on success and a negative ncme_status on failure; the message of the
 *     last failure on the calling thread is available from ncme_last_error().  Nothing throws.
 *   - Indices crossing the ABI follow the reference: 1-based, 0 = "none".
 *   - `stoich` is the reference's `Matrix{Int}` in Julia (column-major) layout: entry
 *     (species s, reaction r) at stoich[r*ns + s].
 *   - State lists are state-major: state i occupies states[i*ns .. i*ns+ns-1] (a Julia
 *     Vector{MVector{NS,Int64}} is laid out exactly like this).
 *   - Pointers named *_dev are device pointers (from ncme_dmalloc or any CUDA allocator, e.g. a
 *     torch tensor's data_ptr); all others are host pointers.
 *   - Kernels are launched on the context's stream; calls that return a scalar or fill a host
 *     buffer synchronise that stream, all others are asynchronous.
 *   - A context is not thread-safe (neither is the reference: matvec! mutates A.t_cache,
 *     src/fspmatrix/sparse/fspsparsematrix.jl:206-210).
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     NCME_ERR_CUDA.
 */
#ifndef NCME_H
#define NCME_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCME_VERSION 100
#define NCME_MAX_REACTIONS 64
#define NCME_MAX_SPECIES 16

typedef enum {
    NCME_OK = 0,
    NCME_ERR_ARG = -1,      /* ArgumentError / DimensionMismatch in the reference */
    NCME_ERR_CUDA = -2,     /* CUDA runtime failure (incl. no device) */
    NCME_ERR_NOMEM = -3,
    NCME_ERR_KEYWIDTH = -4, /* a state component does not fit the 64-bit packed key */
    NCME_ERR_COMM = -5,     /* NCCL failure */
    NCME_ERR_SOLVER = -6,   /* integrator failure (step size underflow, GMRES breakdown ...) */
    NCME_ERR_ABORTED = -7   /* a host callback called ncme_request_abort() (it caught an exception) */
} ncme_status;

typedef struct ncme_ctx ncme_ctx;
typedef struct ncme_space ncme_space;
typedef struct ncme_matrix ncme_matrix;
typedef struct ncme_sensmatrix ncme_sensmatrix;

/* reaction kinds: src/cmemodel/propensity.jl:8-153 */
enum { NCME_TIME_INVARIANT = 0, NCME_SEPARABLE_TV = 1, NCME_JOINT_TV = 2 };

int ncme_version(void);
const char* ncme_last_error(void);

/* ---------------------------------------------------------------- context / memory ----------- */
int ncme_ctx_create(int device, ncme_ctx** out);
int ncme_ctx_destroy(ncme_ctx* ctx);
/* Adopt an external cudaStream_t (e.g. torch's current stream); NULL restores the context's own.
 * The legacy default stream has handle 0, so pass NCME_STREAM_LEGACY ((void*)1 == cudaStreamLegacy) for it. */
#define NCME_STREAM_LEGACY ((void*)0x1)
int ncme_ctx_set_stream(ncme_ctx* ctx, void* cuda_stream);
int ncme_ctx_sync(ncme_ctx* ctx);
/* info[0]=SM count, [1]=L2 bytes, [2]=total global memory bytes, [3]=compute capability*10 */
int ncme_ctx_device_info(ncme_ctx* ctx, int64_t info[4]);
/* Number of kernels this library has launched on the context since creation (bench.py's gpu_launches). */
int ncme_ctx_launch_count(ncme_ctx* ctx, int64_t* count);

int ncme_dmalloc(ncme_ctx* ctx, size_t bytes, void** dptr);
int ncme_dfree(ncme_ctx* ctx, void* dptr);
int ncme_h2d(ncme_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int ncme_d2h(ncme_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
/* page-locked host staging buffers for the end-to-end (host-buffer) path */
int ncme_host_alloc(size_t bytes, void** hptr);
int ncme_host_free(void* hptr);

/* ---------------------------------------------------------------- StateSpaceSparse ----------- */
/* StateSpaceSparse(stoich_matrix, initstates)      src/statespace/sparse/sparsestatespace.jl:103-124
 * Duplicate and negative initial states are silently dropped (:221). */
int ncme_space_create(ncme_ctx* ctx, int ns, int nr, const int64_t* stoich, int64_t n0, const int64_t* states0,
                      ncme_space** out);
/* Import a state space built elsewhere (states + both connectivity tables, reference conventions:
 * struct fields of sparsestatespace.jl:22-40).  The hash table is rebuilt on the device. */
int ncme_space_from_host(ncme_ctx* ctx, int ns, int nr, const int64_t* stoich, int64_t n, const int64_t* states,
                         const uint32_t* state_connectivity, const uint32_t* sink_connectivity, ncme_space** out);
int ncme_space_destroy(ncme_space* space);
/* expand!(space, expansionlevel; onlyreactions)    sparsestatespace.jl:153-194 (+ _addstates! :208-267).
 * onlyreactions: 1-based reaction ids, nonly = 0 means all.  New states are appended in the
 * reference's exact insertion order (LIFO frontier, reactions ascending, first occurrence wins). */
int ncme_space_expand(ncme_space* space, int expansionlevel, int nonly, const int32_t* onlyreactions);
/* deleteat!(space, ids)                            sparsestatespace.jl:276-331.  ids 1-based, any order. */
int ncme_space_delete(ncme_space* space, int64_t nids, const int64_t* ids);
/* get_state_count / get_sink_count                 sparsestatespace.jl:55,62 */
int ncme_space_state_count(ncme_space* space, int64_t* n);
int ncme_space_sink_count(ncme_space* space, int64_t* r);
/* get_states (rows [first, first+count), 0-based range)   sparsestatespace.jl:69 */
int ncme_space_download_states(ncme_space* space, int64_t first, int64_t count, int64_t* states_out);
/* The same states species-major as doubles, cols_out[s * count + i] = x_{first+i}[s]: the layout a vectorised host
 * evaluation of the propensities over many states wants (one contiguous column per species, no strided gathers). */
int ncme_space_download_state_columns(ncme_space* space, int64_t first, int64_t count, double* cols_out);
/* get_state_connectivity / get_sink_connectivity   sparsestatespace.jl:78-80.  Row-major count x nr. */
int ncme_space_download_connectivity(ncme_space* space, int64_t first, int64_t count, uint32_t* state_conn_out,
                                     uint32_t* sink_conn_out);
/* get(state2idx, x, 0) for m states                sparsestatespace.jl:33 (Dict lookups :221,247,258) */
int ncme_space_lookup(ncme_space* space, int64_t m, const int64_t* states, uint32_t* idx_out);
/* sum(p, dims): marginal of a device-resident probability vector over the (1-based) species `dims`
 *                                                   src/fspvector/fspvector.jl:66-99 (used by examples/hog1p.jl:104-105)
 * p_dev: device, n doubles (entry i belongs to state i).  The reduced states come out in the order of their first
 * occurrence in the state list, like the reference; states_out is row-major nred x (NS - #dims), vals_out nred doubles
 * (both host).  *nred is always set; nothing is written when cap < *nred or an output pointer is NULL (size query).
 * Values are accumulated with fp64 atomics: equal to the reference's sequential sums up to summation order. */
int ncme_space_marginal(ncme_space* space, const double* p_dev, int ndims, const int32_t* dims, int64_t cap, int64_t* nred,
                        int64_t* states_out, double* vals_out);

/* ---------------------------------------------------------------- FspMatrixSparse ------------ */
/* FspMatrixSparse{Float64}(space, propensities; parameters)   src/fspmatrix/sparse/fspsparsematrix.jl:47-108
 * kind[r]      : NCME_TIME_INVARIANT / NCME_SEPARABLE_TV / NCME_JOINT_TV for reaction r (0-based array)
 * propvals     : host, reaction-major n x nr: propvals[r*n + i] = state factor of reaction r at state i
 *                (f(x,p) for time-invariant, statefactor(x,p) for separable, ignored for joint --
 *                joint terms start with all-zero values like the reference, :87,129).
 * The matrix snapshots the space (the reference deep-copies the states, :97). */
int ncme_matrix_create(ncme_space* space, const int32_t* kind, const double* propvals, ncme_matrix** out);
int ncme_matrix_destroy(ncme_matrix* mat);
/* size(A)                                          fspsparsematrix.jl:174-186 */
int ncme_matrix_size(ncme_matrix* mat, int64_t* rows, int64_t* cols);
/* _update_sparsematrix!(A_joint[r], states, prop, t, theta)   fspsparsematrix.jl:154-166
 * vals[i] = f(t, x_i, theta) evaluated by the host for the 1-based joint reaction `reaction`.
 * Row-sharded matrices take the values over this rank's rows and predecessor window only: states
 * [row_lo - halo_lo, row_hi + halo_hi) of ncme_matrix_shard_info, in that order. */
int ncme_matrix_set_joint_values(ncme_matrix* mat, int reaction, const double* vals);
/* matvec!(out,t,A,v) (beta=0) / matvecadd!(out,t,A,v) (beta=1)   fspsparsematrix.jl:196-247
 * coef[r] (host, nr entries) = tfactor_r(t,theta) for separable reactions; entries of time-invariant
 * and joint reactions are ignored (treated as 1).  x_dev,y_dev: device, length n + nr, must not alias.
 * One fused kernel launch: all terms, diagonal and the nr sink rows. */
int ncme_matvec(ncme_matrix* mat, const double* coef, const double* x_dev, double* y_dev, double beta);
/* Same through host buffers (H2D of x, kernel, D2H of y inside the call; synchronous).
 * This is what `matvec!(out::Vector, t, A, v::Vector)` binds when the vectors live in host memory. */
int ncme_matvec_host(ncme_matrix* mat, const double* coef, const double* x_host, double* y_host, double beta);
/* Structural counts as the reference stores them: nnz[k], k = term index in reference order
 * (summed time-invariant matrix first if any, then separable, then joint); returns nterms.
 * algorithmic_bytes = sum_k (8 nnz_k + 4 (nnz_k - n)) + 16 (n + nr)       (SURVEY.md 8(d)). */
int ncme_matrix_stats(ncme_matrix* mat, int* nterms, int64_t* nnz_per_term, int64_t* algorithmic_bytes,
                      int64_t* device_bytes);

/* Launch tuning for experiments: rows each thread owns in the matvec kernel (0 = auto, 1, 2 or 4);
 * +16 selects the experimental byte-compressed column indices (1 byte per index + per-chunk descriptor; 15 % less
 * DRAM traffic but latency-bound in its current form, so off by default). */
int ncme_matrix_set_tuning(ncme_matrix* mat, int rows_per_thread);

/* Experiments: select the shared-memory (cp.async) pipelined matvec kernel, `rows` rows per thread and `stages`
 * pipeline stages ((1,4), (2,4), (2,2), (1,2) are instantiated, for 4 or 6 slots); rows = 0 switches it off.
 * Bitwise identical results; measured slower than the default register-staged kernel (profiles/README.md). */
int ncme_matrix_set_pipe(ncme_matrix* mat, int rows, int stages);
/* info = {#(64-row chunk, slot) pairs, #pairs that fell back to 32-bit indices, compression enabled, #slots} */
int ncme_matrix_compression_info(ncme_matrix* mat, int64_t info[4]);

/* ---------------------------------------------------------------- ForwardSensFspMatrixSparse -- */
/* ForwardSensFspMatrixSparse{Float64}(model, space)
 *                      src/forwardsensfspmatrix/forwardsensfspmatrixsparse/sensfspmatrixsparse.jl:31-95
 * npar          : number of parameters P
 * nentries      : number of (reaction, parameter) pairs of the gradient sparsity pattern
 * ent_reaction  : 1-based reaction of each entry;  ent_param: 1-based parameter of each entry
 * dpropvals     : host, entry-major nentries x n: d(state factor)/d(theta_param) at each state
 *                 (d f / d theta for time-invariant; d statefactor / d theta for separable; ignored for joint) */
int ncme_sensmatrix_create(ncme_matrix* mat, int npar, int nentries, const int32_t* ent_reaction,
                           const int32_t* ent_param, const double* dpropvals, ncme_sensmatrix** out);
/* The rebuild after an adapt! (src/forwardsenscme/sparse/forwardsenscmesparse.jl:140 rebuilds from scratch inside the loop).  `mat` must have been built with
 * ncme_matrix_create_incremental from the matrix `prev` sits on; dpropvals_new then holds the derivative factors of
 * the states appended since (entry-major nentries x n_new, n_new as ncme_space_new_count reported before `mat` was
 * built); the rows of the surviving states are carried over on the device.  The (reaction, parameter) pattern must be
 * the one of `prev`.  Same result, bit for bit, as ncme_sensmatrix_create on all states. */
int ncme_sensmatrix_create_incremental(ncme_matrix* mat, ncme_sensmatrix* prev, int npar, int nentries,
                                       const int32_t* ent_reaction, const int32_t* ent_param,
                                       const double* dpropvals_new, ncme_sensmatrix** out);
int ncme_sensmatrix_destroy(ncme_sensmatrix* smat);
/* joint-TV entries: host evaluates d f / d theta_param (t, x_i, theta) and uploads (entry is 0-based). */
int ncme_sensmatrix_set_joint_values(ncme_sensmatrix* smat, int entry, const double* vals);
/* matvec!(out, t, SA, vs)                          sensfspmatrixsparse.jl:97-142
 * coef[r]   : as ncme_matvec.   dcoef[e] (host, nentries): d tfactor_r / d theta_param (t,theta) for
 * separable entries, ignored otherwise.  X_dev, Y_dev: device, (P+1)*(n+nr) doubles, block layout
 * [p; s_1; ...; s_P].  One fused launch; A is read once for all P+1 blocks. */
int ncme_sens_matvec(ncme_sensmatrix* smat, const double* coef, const double* dcoef, const double* X_dev,
                     double* Y_dev);
/* Structure of the derivative terms as the reference stores them (one summed CSC of the time-invariant derivatives
 * per parameter, sensfspmatrixsparse.jl:44-58; one CSC per separable / joint entry, :60-93): stored entries per
 * term (up to 2 * nentries values), B_sens of SURVEY.md 8(d), and the bytes the fused kernel streams.  Any output
 * pointer may be NULL. */
/* Experiments: rows per thread of the fused sensitivity kernel (0 = automatic, 1, 2). */
int ncme_sensmatrix_set_tuning(ncme_sensmatrix* smat, int rows_per_thread);
int ncme_sensmatrix_stats(ncme_sensmatrix* smat, int* ndterms, int64_t* nnz_per_dterm, int64_t* algorithmic_bytes,
                          int64_t* device_bytes);

/* ---------------------------------------------------------------- device vector ops (K7) ------ */
/* The operations an ODE integrator performs on the FSP vector (DiffEq/CVODE internals driven from
 * src/transientcme/sparse/fspsolve.jl:158-161), so that `u` can stay device-resident. */
int ncme_vec_fill(ncme_ctx* ctx, int64_t n, double a, double* x_dev);
int ncme_vec_copy(ncme_ctx* ctx, int64_t n, const double* x_dev, double* y_dev);
int ncme_vec_scale(ncme_ctx* ctx, int64_t n, double a, double* x_dev);
int ncme_vec_axpy(ncme_ctx* ctx, int64_t n, double a, const double* x_dev, double* y_dev);
/* out = sum_k coefs[k] * xs[k]   (k <= 8; out may alias any xs[k]) */
int ncme_vec_lincomb(ncme_ctx* ctx, int64_t n, int k, const double* coefs, const double* const* xs_dev,
                     double* out_dev);
int ncme_vec_sum(ncme_ctx* ctx, int64_t n, const double* x_dev, double* out);
int ncme_vec_dot(ncme_ctx* ctx, int64_t n, const double* x_dev, const double* y_dev, double* out);
/* sqrt(mean((x_i / (atol + rtol*max(|u0_i|,|u1_i|)))^2)) -- the integrator's weighted RMS error norm */
int ncme_vec_wrms(ncme_ctx* ctx, int64_t n, const double* x_dev, const double* u0_dev, const double* u1_dev,
                  double atol, double rtol, double* out);
int ncme_vec_any_nonfinite(ncme_ctx* ctx, int64_t n, const double* x_dev, int* out);
/* out_i = x_i / (atol + rtol*max(|u0_i|,|u1_i|)): the elementwise form of OrdinaryDiffEq's calculate_residuals!, what
 * a DifferentialEquations.jl algorithm broadcasts before it takes its error norm (Julia glue: Broadcast surface) */
int ncme_vec_residuals(ncme_ctx* ctx, int64_t n, const double* x_dev, const double* u0_dev, const double* u1_dev,
                       double atol, double rtol, double* out_dev);
/* x_i += a (the constant term of a broadcast linear expression) */
int ncme_vec_shift(ncme_ctx* ctx, int64_t n, double a, double* x_dev);

/* ---------------------------------------------------------------- multi-GPU (K8) --------------- */
/* One process per GPU.  The reference is single-process (no collective call site exists in it); large state
 * spaces are row-sharded here: rank r owns a contiguous block of state rows of A and the matching slices of every
 * FSP vector; a matvec exchanges the halo of x with the neighbouring shards (grouped ncclSend/ncclRecv, overlapped
 * with the rows that touch no halo entry) and the nr sink rows are partial sums that are all-reduced.
 * The state space itself is replicated (every rank runs the same expand!/prune; they are rare next to matvecs). */
typedef struct ncme_comm ncme_comm;
/* rank 0 creates the id and the host language broadcasts the 128 bytes (torch.distributed / MPI / Distributed.jl) */
int ncme_comm_unique_id(char* out128);
int ncme_comm_create(ncme_ctx* ctx, int rank, int nranks, const char* uid128, ncme_comm** out);
int ncme_comm_destroy(ncme_comm* comm);
int ncme_comm_rank(ncme_comm* comm, int* rank, int* nranks);
int ncme_comm_allreduce_sum(ncme_comm* comm, double* buf_dev, int64_t count);
int ncme_comm_allgatherv(ncme_comm* comm, const double* send_dev, double* recv_dev, const int64_t* counts,
                         const int64_t* displs);
/* FspMatrixSparse restricted to this rank's rows.  propvals covers ALL states (n_global x nr, as ncme_matrix_create).
 * Vectors used with a sharded matrix hold [local rows | nr sink entries]; a matvec INPUT must be allocated with
 * info[2] doubles of margin before it and info[3] after it (the halo is received in place).
 * ncme_matvec on a sharded matrix returns globally reduced sink entries on every rank. */
int ncme_matrix_create_sharded(ncme_space* space, ncme_comm* comm, const int32_t* kind, const double* propvals,
                               ncme_matrix** out);
/* Sharded BUILD (SURVEY.md 8(e) "State-space ops"; the reference's constructor evaluates every propensity at every
 * state, src/fspmatrix/sparse/fspsparsematrix.jl:47-108): ncme_matrix_shard_window returns this rank's row block and
 * the predecessor window of its rows, out = {row_lo, row_hi, ext_lo, ext_hi}, before any propensity is evaluated; the
 * host then evaluates the state factors of the states [ext_lo, ext_hi) only and passes them (reaction-major,
 * (win_hi - win_lo) x nr) to ncme_matrix_create_window.  Host evaluation and upload shrink with the number of ranks.
 * The result is the matrix ncme_matrix_create_sharded builds; it cannot seed ncme_matrix_create_incremental. */
int ncme_matrix_shard_window(ncme_space* space, ncme_comm* comm, int64_t out[4]);
int ncme_matrix_create_window(ncme_space* space, ncme_comm* comm, const int32_t* kind, const double* propvals_window,
                              int64_t win_lo, int64_t win_hi, ncme_matrix** out);
/* Incremental constructor for the matrix that follows an adapt! (the reference rebuilds from scratch and re-evaluates
 * every propensity at every state, src/transientcme/sparse/fspsolve.jl:176).  `prev` must be the matrix
 * this space was last assembled into.  States that survived the prunes since then (always a prefix of the state list,
 * *n_kept of them) take their state factors from `prev` on the device; propvals_new holds the factors of the *n_new
 * states appended since (always the tail), reaction-major n_new x nr.  comm may be NULL. */
int ncme_space_new_count(ncme_space* space, int64_t* n_kept, int64_t* n_new);
int ncme_matrix_create_incremental(ncme_space* space, ncme_comm* comm, ncme_matrix* prev, const int32_t* kind,
                                   const double* propvals_new, ncme_matrix** out);
/* Peer-memory halo (CUDA IPC over NVLink).  COLLECTIVE: every rank registers the allocation that holds its matvec
 * input (base_dev from ncme_dmalloc / cudaMalloc, local rows starting local0 doubles into it, i.e. the halo_lo
 * margin).  Matvecs whose input lies in a registered allocation skip NCCL: the boundary rows read the neighbours'
 * HBM directly and cross-GPU flags (ready / done epochs, bounded waits) replace the collective.  Unregister
 * (collective) before freeing.  The native integrator registers its own workspace. */
int ncme_matrix_register_buffer(ncme_matrix* mat, void* base_dev, size_t bytes, int64_t local0);
int ncme_matrix_unregister_buffer(ncme_matrix* mat, void* base_dev);
/* info = {peer-memory transport available, #matvecs through peer memory, #matvecs through NCCL, halo bytes sent by NCCL} */
int ncme_comm_info(ncme_comm* comm, int64_t info[4]);
/* The matvec as the native integrator issues it per stage: the nr sink entries of y_dev are left as this rank's
 * partial sums (their sum over the ranks is the result; one all-reduce per step instead of one per stage), and the
 * caller promises not to overwrite x_dev before its NEXT matvec on this matrix has been issued (the integrator
 * alternates input buffers), which lets the peer-memory halo drop its "done" handshake. */
int ncme_matvec_local(ncme_matrix* mat, const double* coef, const double* x_dev, double* y_dev);
/* info = {row_lo, row_hi, halo_lo, halo_hi, n_global, interior_begin, interior_end, nranks} */
int ncme_matrix_shard_info(ncme_matrix* mat, int64_t info[8]);

/* ---------------------------------------------------------------- prune (adapt!) --------------- */
/* The dropstates branch of adapt!           src/transientcme/sparse/rstepadapters.jl:40-46 (RStepAdapter, strict=0:
 * `>=`) and :93-99 (SelectiveRStepAdapter, strict=1: `>`):
 *   pids = sortperm(p); dropcount = #{k : sum(p) - cumsum(p[pids])_k >= threshold}; deleteat!(space, sort(pids[1:dropcount]))
 * computed on the device (key-index bitonic sort, scan, count) followed by the deletion of those states.
 * p_dev: device, n doubles.  The compaction map of this deletion stays valid until the next mutation of the space and
 * is applied to vectors with ncme_space_compact_vector (deleteat!(p, dropids), :45). */
int ncme_space_prune_by_mass(ncme_space* space, const double* p_dev, double threshold, int strict, int64_t* dropcount);
/* out_dev[newidx(i)] = in_dev[i] for every state i kept by the last ncme_space_delete / ncme_space_prune_by_mass;
 * in_dev has the old length, out_dev the new one.  They must not alias. */
int ncme_space_compact_vector(ncme_space* space, const double* in_dev, double* out_dev);

/* ---------------------------------------------------------------- native integrator ------------ */
/* Device-resident replacement for `DE.init(ODEProblem(fsprhs!, u, (t0,t1), theta), CVODE_BDF(...); callback, saveat)`
 * + `DE.step!(integrator, t1 - t0, true)`      src/transientcme/sparse/fspsolve.jl:145-161
 * selected from the host by `ode_method = nothing`.  u stays in HBM; per step only the error norm and the
 * R sink entries of the stage vectors cross PCIe.
 *
 * coef_fn(t, coef, user) : host callback filling coef[r] = tfactor_r(t, theta) for separable reactions (and, for
 *                          joint reactions, calling ncme_matrix_set_joint_values); may be NULL for time-invariant A.
 * save_fn(t, u_host, user): called with a pinned host copy of the N-vector at every requested output time.
 * Event (fspsolve.jl:145-156): integration stops at the first t where sum(u[n+1..n+R]) - event_slope * t crosses
 * zero from below (event_slope = fsptol / tend), located on the dense output of the sink entries only. */
typedef void (*ncme_coef_fn)(double t, double* coef, void* user);
typedef void (*ncme_save_fn)(double t, const double* u_host, void* user);

typedef struct ncme_solve_opts {
    double rtol;          /* odertol */
    double atol;          /* odeatol */
    double event_slope;   /* fsptol / tend */
    int check_event;      /* 0: ignore the sinks (fixed-space solve, fspsolve.jl:10-41) */
    int save_every_step;  /* saveat = []: every accepted step (and t0) is handed to save_fn */
    int nsave;            /* saveat times (ascending); those inside [t0, t_final] are produced by dense output */
    const double* save_t;
    double h_init;        /* 0 = automatic */
    int64_t max_steps;    /* 0 = 10^8 */
    int method;           /* 0 = Dormand-Prince 5(4) explicit; 1 = BDF/GMRES (see ncme_solve_segment docs): one fused
                           * kernel per step attempt for unsharded matrices up to NCME_BDF_FUSED_MAX_ROWS (env,
                           * default 2e6) rows, launch-per-operation otherwise; 2 / 3 force the latter / the former */
} ncme_solve_opts;

typedef struct ncme_solve_stats {
    double t_final;
    double h_last;
    int event_hit;
    int nsaved;
    int64_t steps, rejected, rhs_evals, launches;
} ncme_solve_stats;

int ncme_solve_segment(ncme_matrix* mat, ncme_coef_fn coef_fn, ncme_save_fn save_fn, void* user, double t0, double t1,
                       double* u_dev, const ncme_solve_opts* opts, ncme_solve_stats* stats);

/* Callbacks return nothing and exceptions must not unwind through C frames (ctypes swallows them, a Julia @cfunction
 * would crash): a callback that failed stores its exception on the host side and calls ncme_request_abort(); the
 * running ncme_[sens_]solve_segment of this thread then stops before using the callback's output and returns
 * NCME_ERR_ABORTED (the reference propagates the exception out of DE.step!, fspsolve.jl:161).  The flag is
 * thread-local and cleared at the start of every segment. */
void ncme_request_abort(void);

/* Forward-sensitivity segment (src/forwardsenscme/sparse/forwardsenscmesparse.jl:142-166): the same integrator on
 * the block vector U = [p; s_1; ...; s_P] ((P+1)*(n+nr) doubles) with ncme_sens_matvec as right-hand side; the event
 * watches the sinks of the probability block.  coef_fn fills nr + nentries doubles: coef[r] as for ncme_matvec, then
 * dcoef[e] as for ncme_sens_matvec.  Single GPU. */
int ncme_sens_solve_segment(ncme_sensmatrix* smat, ncme_coef_fn coef_fn, ncme_save_fn save_fn, void* user, double t0,
                            double t1, double* U_dev, const ncme_solve_opts* opts, ncme_solve_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* NCME_H */
