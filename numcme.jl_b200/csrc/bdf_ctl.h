// Step-size / order controller of the NDF (BDF) integrator, shared by the host loop (bdf_fused.cu, single-step launches)
// and the in-kernel multi-step driver (same file): one source of truth for
//   * the step about to be attempted (t_new with the landing on t1, c = h / alpha_k, pending rescaling of the
//     difference array D <- (R U)^T D composed over consecutive step-size changes),
//   * the reaction to a failed linear solve (halve h) and to a rejected step (h *= max(0.2, 0.9 err^(-1/(k+1)))),
//   * the order / step-size selection after order+1 equal steps (Shampine & Reichelt's NDF rules as in scipy's BDF).
// Plain C++ without library calls other than <math.h>, compiled for both sides.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define NCME_HD __host__ __device__ __forceinline__
#else
#define NCME_HD inline
#endif

namespace ncme {

constexpr int BDF_MAXO = 5;      // maximum order
constexpr int BDF_RING = 64;     // every-step output slices kept on the device per multi-step launch

struct StepDyn {                 // what changes from one step attempt to the next
    int order, have_change;
    double P[BDF_MAXO + 1][BDF_MAXO + 1];   // pending rescaling of the differences: D_r <- sum_j P[j][r] D_j
    double inv_alpha, c, err_const;
};

struct BdfConst {
    double gamma[BDF_MAXO + 1], alpha[BDF_MAXO + 1], error_const[BDF_MAXO + 2];
    double atol, rtol, Nglob;
};

enum { BDF_RUN = 0, BDF_HOST_STEP = 1, BDF_YIELD = 2, BDF_UNDERFLOW = 3, BDF_MAXSTEPS = 4 };

struct BdfCtl {
    double t, t1, tspan, h_abs, hmin, h_last;
    int order, n_equal_steps, have_change, have_g;
    double P[BDF_MAXO + 1][BDF_MAXO + 1];
    double g_prev, event_slope;
    long long steps, rejected, rhs_evals, max_steps;
    int check_event, save_every_step;
    // the step attempted last (BDF_HOST_STEP: accepted, its bookkeeping is left to the host)
    double t_new, h, error_norm;
    int status, ring_count;
    double ring_t[BDF_RING];
};

NCME_HD void bdf_constants(BdfConst& k) {
    const double kappa[BDF_MAXO + 1] = {0, -0.1850, -1.0 / 9, -0.0823, -0.0415, 0};
    k.gamma[0] = 0;
    for (int q = 1; q <= BDF_MAXO; ++q) k.gamma[q] = k.gamma[q - 1] + 1.0 / q;
    for (int q = 0; q <= BDF_MAXO; ++q) k.alpha[q] = (1 - kappa[q]) * k.gamma[q];
    for (int q = 0; q <= BDF_MAXO; ++q) k.error_const[q] = kappa[q] * k.gamma[q] + 1.0 / (q + 1);
    k.error_const[BDF_MAXO + 1] = 1.0 / (BDF_MAXO + 2);
}

constexpr int BDF_SCRATCH = 5 * (BDF_MAXO + 1) * (BDF_MAXO + 1);   // doubles of work space for bdf_queue_change

// R(factor) of the NDF step-size change, row-major 6 x 6 in R; M: work space of the same size
NCME_HD void bdf_compute_R(int order, double factor, double* R, double* M) {
    constexpr int W = BDF_MAXO + 1;
    for (int i = 0; i < W * W; ++i) M[i] = R[i] = 0.0;
    for (int j = 0; j <= order; ++j) M[j] = 1.0;
    for (int i = 1; i <= order; ++i)
        for (int j = 1; j <= order; ++j) M[i * W + j] = ((double)i - 1.0 - factor * j) / (double)i;
    for (int j = 0; j <= order; ++j) {
        double run = 1.0;
        for (int i = 0; i <= order; ++i) {
            run *= M[i * W + j];
            R[i * W + j] = run;
        }
    }
    R[0] = 1.0;
    for (int i = 1; i <= order; ++i) R[i * W] = 0.0;
}

// D <- (R(factor) U)^T D, queued: composed with a rescaling that is still pending.  `ws`: BDF_SCRATCH doubles (the
// device keeps them in shared memory: as locals they would cost every thread of the grid 1.4 KB of stack)
NCME_HD void bdf_queue_change(BdfCtl& c, int order, double factor, double* ws) {
    constexpr int W = BDF_MAXO + 1;
    double *Rm = ws, *Um = ws + W * W, *RU = ws + 2 * W * W, *T = ws + 3 * W * W, *M = ws + 4 * W * W;
    bdf_compute_R(order, factor, Rm, M);
    bdf_compute_R(order, 1.0, Um, M);
    for (int i = 0; i < W; ++i)
        for (int j = 0; j < W; ++j) {
            double v = 0.0;
            if (i <= order && j <= order)
                for (int q = 0; q <= order; ++q) v += Rm[i * W + q] * Um[q * W + j];
            RU[i * W + j] = v;
        }
    if (!c.have_change) {
        for (int i = 0; i < W; ++i)
            for (int j = 0; j < W; ++j) c.P[i][j] = RU[i * W + j];
    } else {
        for (int i = 0; i < W; ++i)
            for (int j = 0; j < W; ++j) {
                double v = 0.0;
                if (i <= order && j <= order)
                    for (int q = 0; q <= order; ++q) v += c.P[i][q] * RU[q * W + j];
                T[i * W + j] = v;
            }
        for (int i = 0; i < W; ++i)
            for (int j = 0; j < W; ++j) c.P[i][j] = T[i * W + j];
    }
    c.have_change = 1;
}

// the step about to be attempted: lands on t1 when the next step would pass (or nearly reach) it
NCME_HD void bdf_begin_step(BdfCtl& c, const BdfConst& k, StepDyn& dyn, double* ws) {
    double t_new = c.t + c.h_abs;
    if (t_new > c.t1 || c.t1 - t_new < 1e-12 * c.tspan) {
        t_new = c.t1;
        bdf_queue_change(c, c.order, fabs(t_new - c.t) / c.h_abs, ws);
        c.n_equal_steps = 0;
    }
    c.t_new = t_new;
    c.h = t_new - c.t;
    c.h_abs = fabs(c.h);
    dyn.order = c.order;
    dyn.have_change = c.have_change;
    for (int i = 0; i <= BDF_MAXO; ++i)
        for (int j = 0; j <= BDF_MAXO; ++j) dyn.P[i][j] = c.P[i][j];
    c.have_change = 0;
    dyn.inv_alpha = 1.0 / k.alpha[c.order];
    dyn.c = c.h / k.alpha[c.order];
    dyn.err_const = k.error_const[c.order];
}

// linear solver failed: halve the step (CVODE's reaction to a convergence failure)
NCME_HD void bdf_after_linfail(BdfCtl& c, double* ws) {
    c.rejected++;
    c.h_abs *= 0.5;
    bdf_queue_change(c, c.order, 0.5, ws);
    c.n_equal_steps = 0;
}

NCME_HD void bdf_after_reject(BdfCtl& c, double error_norm, double* ws) {
    c.rejected++;
    double factor = 0.2;
    if (error_norm < 1e299) {
        factor = 0.9 * pow(error_norm, -1.0 / (c.order + 1));
        if (factor < 0.2) factor = 0.2;
    }
    c.h_abs *= factor;
    bdf_queue_change(c, c.order, factor, ws);
    c.n_equal_steps = 0;
}

// after order+1 equal steps: sm / sp = sums over ALL entries of (D_order / scale)^2 and (D_{order+2} / scale)^2
NCME_HD void bdf_select_order(BdfCtl& c, const BdfConst& k, double error_norm, double sm, double sp, double* ws) {
    const double INF = 1e300;
    const int order = c.order;
    const double em = order > 1 ? k.error_const[order - 1] * sqrt(sm / k.Nglob) : INF;
    const double ep = order < BDF_MAXO ? k.error_const[order + 1] * sqrt(sp / k.Nglob) : INF;
    const double norms[3] = {em, error_norm, ep};
    double factors[3];
    for (int q = 0; q < 3; ++q)
        factors[q] = norms[q] >= INF ? 0.0 : (norms[q] > 0 ? pow(norms[q], -1.0 / (order + q)) : 1e9);
    int best = 1;
    for (int q = 0; q < 3; ++q)
        if (factors[q] > factors[best]) best = q;
    c.order += best - 1;
    double factor = 0.9 * factors[best];
    if (factor > 10.0) factor = 10.0;
    c.h_abs *= factor;
    bdf_queue_change(c, c.order, factor, ws);
    c.n_equal_steps = 0;
}

// dense output of the sink sum at tt inside the step that ended at t_new (differences nd[j][r], j <= order)
template <typename ND>
NCME_HD double bdf_sink_sum_at(const ND& nd, int R, int order, double t_new, double h, double tt) {
    double sum = 0.0;
    for (int r = 0; r < R; ++r) {
        double p = 1.0, y = nd(0, r);
        for (int j = 1; j <= order; ++j) {
            p *= (tt - (t_new - (j - 1) * h)) / (h * j);
            y += nd(j, r) * p;
        }
        sum += y;
    }
    return sum;
}

}  // namespace ncme
