/* CPU restatement of the reference matvec for the timed baseline (TEST INFRASTRUCTURE ONLY).
 *
 * The reference's FSP matvec (/root/reference/src/fspmatrix/sparse/fspsparsematrix.jl:196-217)
 * is 1 + n_sep + n_joint calls of Julia's stdlib SparseArrays.mul!(C, A, B, alpha, beta) on
 * SparseMatrixCSC{Float64,Int64} operands.  That stdlib routine is not under /root/reference
 * (Project.toml:24 `SparseArrays = "1"`); its published algorithm is restated here:
 *   beta == 0 ? fill C with 0 : scale C by beta (skipped when beta == 1), then
 *   for col in 1:n   axj = B[col]*alpha;  for k in nzrange(A,col)  C[rowval[k]] += nzval[k]*axj
 * It is serial: `cores = 1` is the faithful reference configuration.
 *
 * ncme_oracle_csr_omp is an additional "best CPU" arm (row-parallel CSR over all host cores);
 * it is not how the reference computes.
 */
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void ncme_oracle_csc_mul(int64_t ncols, int64_t nrows, const int64_t* colptr, const int64_t* rowval,
                         const double* nzval, const double* x, double alpha, double beta, double* y) {
    if (beta == 0.0) {
        for (int64_t i = 0; i < nrows; ++i) y[i] = 0.0;
    } else if (beta != 1.0) {
        for (int64_t i = 0; i < nrows; ++i) y[i] *= beta;
    }
    for (int64_t col = 0; col < ncols; ++col) {
        const double axj = x[col] * alpha;
        for (int64_t k = colptr[col]; k < colptr[col + 1]; ++k) y[rowval[k]] += nzval[k] * axj;
    }
}

/* matvec! over a list of terms: out = sum_k coef[k] * A_k * v   (first term beta = 0). */
void ncme_oracle_fsp_matvec(int nterms, int64_t n, const int64_t* const* colptr, const int64_t* const* rowval,
                            const double* const* nzval, const double* coef, const double* v, double* out) {
    if (nterms == 0) {
        for (int64_t i = 0; i < n; ++i) out[i] = 0.0;
        return;
    }
    for (int k = 0; k < nterms; ++k)
        ncme_oracle_csc_mul(n, n, colptr[k], rowval[k], nzval[k], v, coef[k], k == 0 ? 0.0 : 1.0, out);
}

/* Row-parallel CSR (one fused matrix), all host cores. */
void ncme_oracle_csr_omp(int64_t nrows, const int64_t* rowptr, const int32_t* colind, const double* val,
                         const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nrows; ++i) {
        double acc = 0.0;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) acc += val[k] * x[colind[k]];
        y[i] = acc;
    }
}

int ncme_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
