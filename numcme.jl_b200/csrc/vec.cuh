// K7 internals shared with the native integrator.
#pragma once
#include <math.h>

#include "common.cuh"

namespace ncme {
int vec_sum(ncme_ctx* ctx, int64_t n, const double* x, double* out);
}
