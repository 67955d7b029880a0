// Host-side harness for the shared host/device BDF controller (numcme.jl_b200/csrc/bdf_ctl.h), built with g++ by
// tests/test_bdf_controller.py: the same functions run inside k_bdf_step on the device.
#include "bdf_ctl.h"

#include <string.h>

using namespace ncme;

extern "C" {

void h_constants(double* gamma, double* alpha, double* error_const) {
    BdfConst k;
    bdf_constants(k);
    memcpy(gamma, k.gamma, sizeof(k.gamma));
    memcpy(alpha, k.alpha, sizeof(k.alpha));
    memcpy(error_const, k.error_const, sizeof(k.error_const));
}

void h_compute_R(int order, double factor, double* R36) {
    double M[36];
    bdf_compute_R(order, factor, R36, M);
}

// queue `nfac` rescalings at the given order; returns the composed P (row-major 6 x 6)
void h_queue_changes(int order, int nfac, const double* factors, double* P36) {
    BdfCtl c;
    memset(&c, 0, sizeof(c));
    double ws[BDF_SCRATCH];
    for (int k = 0; k < nfac; ++k) bdf_queue_change(c, order, factors[k], ws);
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) P36[i * 6 + j] = c.P[i][j];
}

// one controller round: begin a step from (t, h_abs, order), then react; returns the new (h_abs, order, n_equal, have_change)
// mode 0: accepted with order selection (sm, sp), 1: rejected with error_norm, 2: linear solver failure
void h_round(double t, double t1, double h_abs, int order, int n_equal, int mode, double error_norm, double sm, double sp,
             double Nglob, double* out /* t_new, h, c, inv_alpha, err_const, h_abs', order', n_equal', have_change' */) {
    BdfCtl c;
    memset(&c, 0, sizeof(c));
    BdfConst k;
    bdf_constants(k);
    k.Nglob = Nglob;
    c.t = t;
    c.t1 = t1;
    c.tspan = t1;
    c.h_abs = h_abs;
    c.order = order;
    c.n_equal_steps = n_equal;
    StepDyn dyn;
    double ws[BDF_SCRATCH];
    bdf_begin_step(c, k, dyn, ws);
    out[0] = c.t_new;
    out[1] = c.h;
    out[2] = dyn.c;
    out[3] = dyn.inv_alpha;
    out[4] = dyn.err_const;
    if (mode == 0) {
        c.n_equal_steps++;
        if (c.n_equal_steps >= c.order + 1) bdf_select_order(c, k, error_norm, sm, sp, ws);
    } else if (mode == 1) {
        bdf_after_reject(c, error_norm, ws);
    } else {
        bdf_after_linfail(c, ws);
    }
    out[5] = c.h_abs;
    out[6] = c.order;
    out[7] = c.n_equal_steps;
    out[8] = c.have_change;
}

double h_sink_sum_at(int R, int order, const double* nd /* [order+1][R] */, double t_new, double h, double tt) {
    auto view = [&](int j, int r) { return nd[j * R + r]; };
    return bdf_sink_sum_at(view, R, order, t_new, h, tt);
}

int h_sizeof_ctl() { return (int)sizeof(BdfCtl); }
}
