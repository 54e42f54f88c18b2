/*
 * dgoracle.c -- CPU restatement (plain C) of the third-party native routines
 * the reference hot path executes.  TEST INFRASTRUCTURE ONLY: nothing in the
 * product package links or loads this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * The routines restated here are NOT in /root/reference (un-vendored deps):
 *   - pyamg==5.0.1  pyamg/amg_core/relaxation.h : block_gauss_seidel
 *       reference call site: dgfem/pyamg_relaxation.py:252-255
 *                            (<- dgfem/relaxation.py:207)
 *   - scipy==1.11.3 scipy/sparse/sparsetools/bsr.h : bsr_matvec (+ gemv helper)
 *       reference call sites: dgfem/solver.py:117,119,150;
 *                             dgfem/relaxation.py:202,208; utils/helpers.py:39
 * Parity status: "parity unpinned" for block_gauss_seidel against a real pyamg
 * binary (pyamg cannot be installed offline); it is pinned indirectly against the
 * reference's own pure-NumPy forward block-GS (dgfem/relaxation.py:170-195),
 * see tests/test_oracle_vs_golden.py.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* y[0:R] += A[R x C] (row-major) * x[0:C]; accumulation order j ascending,
 * as scipy's generic gemv (scipy/sparse/sparsetools/dense.h). */
static void gemv_acc(int R, int C, const double *A, const double *x, double *y)
{
    for (int i = 0; i < R; ++i) {
        double dot = y[i];
        for (int j = 0; j < C; ++j) dot += A[(size_t)C * i + j] * x[j];
        y[i] = dot;
    }
}

/* scipy bsr_matvec: Y += A*X for a BSR matrix with square bs x bs blocks.
 * Caller zeroes Y first (scipy's _matmul_vector does). */
void orc_bsr_matvec(int32_t n_brow, int32_t bs, const int32_t *Ap,
                    const int32_t *Aj, const double *Ax, const double *X,
                    double *Y)
{
    const size_t B2 = (size_t)bs * bs;
    for (int32_t i = 0; i < n_brow; ++i) {
        double *y = Y + (size_t)bs * i;
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) {
            const int32_t j = Aj[jj];
            gemv_acc(bs, bs, Ax + B2 * jj, X + (size_t)bs * j, y);
        }
    }
}

/* r = b - A x ; returns sum(r^2) accumulated in index order. */
double orc_bsr_residual(int32_t n_brow, int32_t bs, const int32_t *Ap,
                        const int32_t *Aj, const double *Ax, const double *X,
                        const double *B, double *Rv)
{
    const size_t n = (size_t)n_brow * bs;
    memset(Rv, 0, n * sizeof(double));
    orc_bsr_matvec(n_brow, bs, Ap, Aj, Ax, X, Rv);
    double s = 0.0;
    for (size_t k = 0; k < n; ++k) {
        Rv[k] = B[k] - Rv[k];
        s += Rv[k] * Rv[k];
    }
    return s;
}

/* pyamg amg_core.block_gauss_seidel (relaxation.h), restated:
 *   for i in row_start:row_stop:row_step
 *     rsum = b_i
 *     for jj in Ap[i]..Ap[i+1]: if Aj[jj]==i skip; v = A_jj * x_j; rsum -= v
 *     x_i = Dinv_i * rsum
 * Blocks are row-major, products are plain dense gemv accumulated from zero,
 * block by block in stored (ascending column) order. */
void orc_block_gauss_seidel(const int32_t *Ap, const int32_t *Aj,
                            const double *Ax, double *x, const double *b,
                            const double *Dinv, int32_t row_start,
                            int32_t row_stop, int32_t row_step, int32_t bs)
{
    const size_t B2 = (size_t)bs * bs;
    double *rsum = (double *)malloc(sizeof(double) * bs);
    double *v = (double *)malloc(sizeof(double) * bs);
    for (int32_t i = row_start; i != row_stop; i += row_step) {
        for (int k = 0; k < bs; ++k) rsum[k] = b[(size_t)i * bs + k];
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) {
            const int32_t j = Aj[jj];
            if (j == i) continue;
            for (int k = 0; k < bs; ++k) v[k] = 0.0;
            gemv_acc(bs, bs, Ax + B2 * jj, x + (size_t)bs * j, v);
            for (int k = 0; k < bs; ++k) rsum[k] -= v[k];
        }
        for (int k = 0; k < bs; ++k) v[k] = 0.0;
        gemv_acc(bs, bs, Dinv + B2 * i, rsum, v);
        for (int k = 0; k < bs; ++k) x[(size_t)i * bs + k] = v[k];
    }
    free(rsum);
    free(v);
}

/* Same sweep restricted to one colour class of a structured Ni x Nj element
 * grid: rows with (i + j) % ncolours == colour, i = row % Ni, j = row / Ni.
 * Used only to check the product's multicolour (red-black) mode, which has no
 * counterpart in the reference. */
void orc_block_gauss_seidel_colour(const int32_t *Ap, const int32_t *Aj,
                                   const double *Ax, double *x, const double *b,
                                   const double *Dinv, int32_t Ni, int32_t Nj,
                                   int32_t ncolours, int32_t colour, int32_t bs)
{
    const size_t B2 = (size_t)bs * bs;
    const int32_t N = Ni * Nj;
    double *rsum = (double *)malloc(sizeof(double) * bs);
    double *v = (double *)malloc(sizeof(double) * bs);
    /* colour classes are independent sets for the 5-point block stencil when
     * ncolours == 2 and Ni is even (or non-periodic), so in-place == Jacobi
     * within the class */
    for (int32_t i = 0; i < N; ++i) {
        if (((i % Ni) + (i / Ni)) % ncolours != colour) continue;
        for (int k = 0; k < bs; ++k) rsum[k] = b[(size_t)i * bs + k];
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) {
            const int32_t j = Aj[jj];
            if (j == i) continue;
            for (int k = 0; k < bs; ++k) v[k] = 0.0;
            gemv_acc(bs, bs, Ax + B2 * jj, x + (size_t)bs * j, v);
            for (int k = 0; k < bs; ++k) rsum[k] -= v[k];
        }
        for (int k = 0; k < bs; ++k) v[k] = 0.0;
        gemv_acc(bs, bs, Dinv + B2 * i, rsum, v);
        for (int k = 0; k < bs; ++k) x[(size_t)i * bs + k] = v[k];
    }
    free(rsum);
    free(v);
}
