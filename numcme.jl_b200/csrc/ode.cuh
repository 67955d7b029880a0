// Shared description of the linear ODE system du/dt = F(t) u handed to the native integrators (solve.cu: explicit
// Dormand-Prince; bdf.cu: variable-order BDF/NDF + matrix-free GMRES).
#pragma once
#include <functional>

#include "comm.cuh"
#include "common.cuh"

struct ncme_matrix;

namespace ncme {

struct OdeSystem {
    ncme_ctx* ctx = nullptr;
    ncme_comm* comm = nullptr;   // row-sharded FSP vectors only
    int64_t len = 0;             // local vector length
    int64_t sink_off = 0;        // the R event-sink entries are [sink_off, sink_off + R)
    int R = 0;
    int64_t hl = 0, hh = 0;      // halo margins every RHS input must carry
    int64_t len_global = 0;      // number of entries of the global vector (error norm)
    int64_t n_global = 0;        // sharded: global number of state rows (output gather)
    int peers[2] = {-1, -1};     // sharded: ranks whose vectors this rank reads (peer-memory halo)
    // y = F(t) x (sharded: sink entries of y are partial sums).  `rhs` may assume that consecutive calls use different
    // input buffers (explicit RK alternates them); `rhs_safe` makes no such assumption (BDF/GMRES).
    std::function<int(double, const double*, double*)> rhs;
    std::function<int(double, const double*, double*)> rhs_safe;
    // BDF only -------------------------------------------------------------------------------------------------
    // entries [0, n_impl) are solved implicitly on the device; for FSP systems the remaining R sink entries do not
    // feed back (their columns are empty) and are completed explicitly from rhs_sinks.
    int64_t n_impl = 0;
    std::function<int(double, double*)> jac_diag;                      // out[0..n_impl) = diag(F(t))
    std::function<int(double, const double*, double*)> rhs_sinks;       // only the R sink entries of F(t) x (cheap)
};

int cache_reserve(double** p, size_t* have, size_t want, bool pinned);
int solve_dp5(const OdeSystem& sys, ncme_save_fn save_fn, void* user, double t0, double t1, double* u,
              const ncme_solve_opts* o, ncme_solve_stats* st);
int solve_bdf(const OdeSystem& sys, ncme_save_fn save_fn, void* user, double t0, double t1, double* u,
              const ncme_solve_opts* o, ncme_solve_stats* st);

// bdf_fused.cu: the same BDF/GMRES algorithm with ONE kernel launch per step attempt (unsharded FSP matrices)
bool bdf_fused_eligible(const ncme_matrix* A);
int solve_bdf_fused(ncme_matrix* A, ncme_coef_fn coef_fn, ncme_save_fn save_fn, void* user, double t0, double t1,
                    double* u, const ncme_solve_opts* o, ncme_solve_stats* st);

// gathers a slice [all state rows | reduced sinks] on every rank and hands it to the host callback
struct SliceSaver {
    const OdeSystem* sys = nullptr;
    ncme_save_fn fn = nullptr;
    void* user = nullptr;
    ncme_solve_stats* st = nullptr;
    double* pinned = nullptr;
    double* full = nullptr;
    std::vector<int64_t> counts, displs;
    int init(const OdeSystem& s, ncme_save_fn f, void* u, ncme_solve_stats* stats);
    int save(double t, const double* v_dev);
};

}  // namespace ncme
