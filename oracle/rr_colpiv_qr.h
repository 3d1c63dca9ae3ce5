/*
 * TEST INFRASTRUCTURE — CPU oracle. Not part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use anything under oracle/.
 *
 * rr_colpiv_qr.h — plain-C restatement of Eigen 3.4's unblocked column-pivoted
 * Householder QR and its least-squares solve, the third-party arithmetic behind
 * `A.colPivHouseholderQr().solve(b)` (rils_rols_cpp/rils_rols_cpp.cpp:484).
 * Eigen's Core module is missing from the vendored tree, so the
 * real header cannot be compiled; the algorithm restated here is the one the
 * vendored tree still holds (paths relative to /root/reference/rils_rols_cpp/eigen/Eigen/src):
 *
 *   QR/ColPivHouseholderQR.h:482-581   computeInPlace(): norms, threshold, pivoting, downdate
 *   QR/ColPivHouseholderQR.h:587-607   _solve_impl(): Q^T b, back substitution, permutation
 *   Householder/Householder.h:67-98    makeHouseholder(): sign convention, tol = DBL_MIN
 *   Householder/Householder.h:116-135  applyHouseholderOnTheLeft()
 *
 * Reductions are sequential left-to-right sums (real Eigen uses packet-wise tree
 * reductions; the difference is at the ulp level and is part of the stated 1e-9
 * tolerance). Shared by the Eigen stand-in (eigen_shim/Core) that lets the
 * unmodified reference compile and by the C restatement (rr_oracle.c), so both
 * produce bit-identical coefficients when compiled with -ffp-contract=off.
 */
#ifndef RR_ORACLE_COLPIV_QR_H
#define RR_ORACLE_COLPIV_QR_H

#include <float.h>
#include <math.h>
#include <stddef.h>

/* A: rows x cols, column-major, leading dimension lda, factored in place.
 * hcoef[min(rows,cols)], perm[cols] (perm[i] = original column at position i),
 * work[cols], norms_upd[cols], norms_dir[cols]. Returns nonzero_pivots. */
static int rr_colpiv_qr_factor(double *A, ptrdiff_t rows, ptrdiff_t cols, ptrdiff_t lda,
                               double *hcoef, int *perm, double *work, double *norms_upd,
                               double *norms_dir)
{
    const ptrdiff_t size = rows < cols ? rows : cols;
    ptrdiff_t k, j, i;
    double maxnorm = 0.0;
    int nonzero_pivots;
    /* ColPivHouseholderQR.h:504-509 */
    for (k = 0; k < cols; ++k) {
        const double *c = A + k * lda;
        double s = 0.0;
        for (i = 0; i < rows; ++i) s += c[i] * c[i];
        norms_dir[k] = sqrt(s);
        norms_upd[k] = norms_dir[k];
        perm[k] = (int)k;
    }
    /* maxCoeff(): NaN never wins a `>` comparison, like Eigen's scalar visitor */
    if (cols > 0) maxnorm = norms_upd[0];
    for (k = 1; k < cols; ++k)
        if (norms_upd[k] > maxnorm) maxnorm = norms_upd[k];
    /* :511-512 */
    const double threshold_helper = (maxnorm * DBL_EPSILON) * (maxnorm * DBL_EPSILON) / (double)rows;
    const double norm_downdate_threshold = sqrt(DBL_EPSILON);
    nonzero_pivots = (int)size; /* :514 */

    for (k = 0; k < size; ++k) {
        /* :520-522 first maximum of the trailing updated norms */
        ptrdiff_t big = k;
        double bigv = norms_upd[k];
        for (j = k + 1; j < cols; ++j)
            if (norms_upd[j] > bigv) { bigv = norms_upd[j]; big = j; }
        const double biggest_col_sq_norm = bigv * bigv;
        /* :526-527 */
        if (nonzero_pivots == (int)size && biggest_col_sq_norm < threshold_helper * (double)(rows - k))
            nonzero_pivots = (int)k;
        /* :530-536 */
        if (k != big) {
            double *ck = A + k * lda, *cb = A + big * lda, t;
            int ti;
            for (i = 0; i < rows; ++i) { t = ck[i]; ck[i] = cb[i]; cb[i] = t; }
            t = norms_upd[k]; norms_upd[k] = norms_upd[big]; norms_upd[big] = t;
            t = norms_dir[k]; norms_dir[k] = norms_dir[big]; norms_dir[big] = t;
            ti = perm[k]; perm[k] = perm[big]; perm[big] = ti; /* == :576-578 */
        }
        /* :539-543 makeHouseholderInPlace on col k, rows k.. (Householder.h:67-98) */
        double *v = A + k * lda + k;
        const ptrdiff_t m = rows - k; /* vector length */
        double tail_sq = 0.0, beta, tau;
        for (i = 1; i < m; ++i) tail_sq += v[i] * v[i];
        const double c0 = v[0];
        if (tail_sq <= DBL_MIN) {
            tau = 0.0;
            beta = c0;
            for (i = 1; i < m; ++i) v[i] = 0.0;
        } else {
            beta = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0.0) beta = -beta;
            const double denom = c0 - beta;
            for (i = 1; i < m; ++i) v[i] = v[i] / denom;
            tau = (beta - c0) / beta;
        }
        hcoef[k] = tau;
        v[0] = beta;
        /* :549-550 apply H_k to the trailing columns (Householder.h:116-135) */
        if (cols - k - 1 > 0) {
            if (m == 1) {
                for (j = k + 1; j < cols; ++j) A[j * lda + k] *= (1.0 - tau);
            } else if (tau != 0.0) {
                for (j = k + 1; j < cols; ++j) {
                    double *cj = A + j * lda + k;
                    double s = 0.0;
                    for (i = 1; i < m; ++i) s += v[i] * cj[i];
                    work[j] = s + cj[0];
                }
                for (j = k + 1; j < cols; ++j) {
                    double *cj = A + j * lda + k;
                    cj[0] -= tau * work[j];
                    for (i = 1; i < m; ++i) cj[i] -= (tau * v[i]) * work[j];
                }
            }
        }
        /* :553-573 norm downdate */
        for (j = k + 1; j < cols; ++j) {
            if (norms_upd[j] != 0.0) {
                double temp = fabs(A[j * lda + k]) / norms_upd[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0.0 ? 0.0 : temp;
                const double ratio = norms_upd[j] / norms_dir[j];
                const double temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    const double *cj = A + j * lda;
                    double s = 0.0;
                    for (i = k + 1; i < rows; ++i) s += cj[i] * cj[i];
                    norms_dir[j] = sqrt(s);
                    norms_upd[j] = norms_dir[j];
                } else {
                    norms_upd[j] *= sqrt(temp);
                }
            }
        }
    }
    return nonzero_pivots;
}

/* ColPivHouseholderQR.h:587-607. b[rows] is overwritten (c = Q^T b); x[cols] receives
 * the solution in ORIGINAL column order; non-pivot entries are exactly 0. */
static void rr_colpiv_qr_solve(const double *A, ptrdiff_t rows, ptrdiff_t cols, ptrdiff_t lda,
                               const double *hcoef, const int *perm, int nonzero_pivots,
                               double *b, double *x)
{
    ptrdiff_t k, i, j;
    if (nonzero_pivots == 0) {
        for (j = 0; j < cols; ++j) x[j] = 0.0;
        return;
    }
    /* :599 c = H_{r-1} ... H_0 b, applied k ascending (HouseholderSequence adjoint, unblocked) */
    for (k = 0; k < nonzero_pivots; ++k) {
        const double *v = A + k * lda + k;
        const ptrdiff_t m = rows - k;
        const double tau = hcoef[k];
        if (m == 1) {
            b[k] *= (1.0 - tau);
        } else if (tau != 0.0) {
            double s = 0.0;
            for (i = 1; i < m; ++i) s += v[i] * b[k + i];
            const double tmp = s + b[k];
            b[k] -= tau * tmp;
            for (i = 1; i < m; ++i) b[k + i] -= (tau * v[i]) * tmp;
        }
    }
    /* :601-603 upper-triangular solve, column-oriented back substitution */
    for (i = nonzero_pivots - 1; i >= 0; --i) {
        b[i] = b[i] / A[i * lda + i];
        const double xi = b[i];
        const double *ci = A + i * lda;
        for (j = 0; j < i; ++j) b[j] -= xi * ci[j];
    }
    /* :605-606 */
    for (i = 0; i < nonzero_pivots; ++i) x[perm[i]] = b[i];
    for (i = nonzero_pivots; i < cols; ++i) x[perm[i]] = 0.0;
}

#endif /* RR_ORACLE_COLPIV_QR_H */
