// rr_exact.cuh — exact path: column-pivoted Householder QR on the materialised design matrix.
//
// For n <= exact_max_n (small data sets: BASELINE configs 1-3) every candidate takes this path;
// the work per neighbourhood is microseconds either way and the result follows the reference's
// own algorithm operation by operation instead of going through normal equations:
//   Eigen ColPivHouseholderQR::computeInPlace   eigen/Eigen/src/QR/ColPivHouseholderQR.h:482-581
//   ColPivHouseholderQR::_solve_impl            eigen/Eigen/src/QR/ColPivHouseholderQR.h:587-607
//   makeHouseholder / applyHouseholderOnTheLeft eigen/Eigen/src/Householder/Householder.h:67-98,116-135
// (paths under /root/reference/rils_rols_cpp). Sums run sequentially in row order and nothing is
// contracted into FMAs (--fmad=false), so a candidate whose terms use only + - * / sqrt and
// comparisons gets coefficients bit-identical to a sequential CPU evaluation of that algorithm;
// with sin/cos/log/exp the inputs differ by <= 1-2 ulp (libdevice vs libm).
//
// One thread per candidate. The distinct term columns were materialised once by the sweep
// (V[u][i]); each thread gathers its k columns into a workspace interleaved by candidate
// (element (i,j) of candidate slot q at ((j*n + i)*Q + q)), so that the threads of a warp, which
// walk their matrices in lock step, touch consecutive addresses.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/rr_b200.h"

namespace rr {

struct ExactArgs {
    const double *V;       // materialised distinct terms, column u at V + u*ldv
    int64_t ldv;
    const double *y;       // n targets
    int32_t n;
    const int32_t *term_ids;         // per term instance -> distinct term (batch order)
    const int32_t *cand_term_begin;  // batch offsets
    int32_t cand_lo, cand_hi;        // candidates [lo, hi) handled by this launch
    int32_t Q;                       // interleave factor (>= hi - lo)
    int32_t kmax;
    double *A;             // workspace n * kmax * Q
    double *rhs;           // workspace n * Q
    double *aux;           // per candidate slot 5*kmax doubles (hcoef, work, nu, nd, x)
    int32_t *perm;         // per candidate slot kmax
    double *coef;          // out (batch coef layout)
    int32_t *nzp;          // out
    uint32_t *flags;       // out
};

__global__ void rr_exact_qr(const ExactArgs a)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = a.cand_lo + q;
    if (c >= a.cand_hi) return;
    const int rows = a.n;
    const int m = a.cand_term_begin[c + 1] - a.cand_term_begin[c];
    const int cols = m + 1;
    const int64_t Q = a.Q;
#define AT(i, j) a.A[((int64_t)(j) * rows + (i)) * Q + q]
#define RHS(i) a.rhs[(int64_t)(i) * Q + q]
    double *hcoef = a.aux + (int64_t)q * 5 * a.kmax;
    double *work = hcoef + a.kmax, *nu = work + a.kmax, *nd = nu + a.kmax, *x = nd + a.kmax;
    int *perm = a.perm + (int64_t)q * a.kmax;

    // rils_rols_cpp.cpp:477-482: A.col(i) = factors[i]->evaluate_all(X), last column = ones
    for (int j = 0; j < m; ++j) {
        const double *src = a.V + (int64_t)a.term_ids[a.cand_term_begin[c] + j] * a.ldv;
        for (int i = 0; i < rows; ++i) AT(i, j) = src[i];
    }
    for (int i = 0; i < rows; ++i) AT(i, m) = 1.0;
    for (int i = 0; i < rows; ++i) RHS(i) = a.y[i];

    const int size = rows < cols ? rows : cols;
    // ColPivHouseholderQR.h:504-509
    for (int k = 0; k < cols; ++k) {
        double s = 0.0;
        for (int i = 0; i < rows; ++i) s += AT(i, k) * AT(i, k);
        nd[k] = sqrt(s);
        nu[k] = nd[k];
        perm[k] = k;
    }
    double maxnorm = nu[0];
    for (int k = 1; k < cols; ++k)
        if (nu[k] > maxnorm) maxnorm = nu[k];
    const double threshold_helper = (maxnorm * DBL_EPSILON) * (maxnorm * DBL_EPSILON) / (double)rows;  // :511
    const double norm_downdate_threshold = sqrt(DBL_EPSILON);
    int nonzero_pivots = size;

    for (int k = 0; k < size; ++k) {
        int big = k;
        double bigv = nu[k];
        for (int j = k + 1; j < cols; ++j)
            if (nu[j] > bigv) { bigv = nu[j]; big = j; }
        if (nonzero_pivots == size && bigv * bigv < threshold_helper * (double)(rows - k)) nonzero_pivots = k;  // :526
        if (k != big) {  // :530-536
            for (int i = 0; i < rows; ++i) { const double t = AT(i, k); AT(i, k) = AT(i, big); AT(i, big) = t; }
            double t = nu[k]; nu[k] = nu[big]; nu[big] = t;
            t = nd[k]; nd[k] = nd[big]; nd[big] = t;
            const int ti = perm[k]; perm[k] = perm[big]; perm[big] = ti;
        }
        // makeHouseholderInPlace, Householder.h:67-98
        const int mlen = rows - k;
        double tail_sq = 0.0;
        for (int i = 1; i < mlen; ++i) tail_sq += AT(k + i, k) * AT(k + i, k);
        const double c0 = AT(k, k);
        double beta, tau;
        if (tail_sq <= DBL_MIN) {
            tau = 0.0;
            beta = c0;
            for (int i = 1; i < mlen; ++i) AT(k + i, k) = 0.0;
        } else {
            beta = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0.0) beta = -beta;
            const double denom = c0 - beta;
            for (int i = 1; i < mlen; ++i) AT(k + i, k) = AT(k + i, k) / denom;
            tau = (beta - c0) / beta;
        }
        hcoef[k] = tau;
        AT(k, k) = beta;
        // applyHouseholderOnTheLeft to the trailing columns, Householder.h:116-135
        if (cols - k - 1 > 0) {
            if (mlen == 1) {
                for (int j = k + 1; j < cols; ++j) AT(k, j) *= (1.0 - tau);
            } else if (tau != 0.0) {
                for (int j = k + 1; j < cols; ++j) {
                    double s = 0.0;
                    for (int i = 1; i < mlen; ++i) s += AT(k + i, k) * AT(k + i, j);
                    work[j] = s + AT(k, j);
                }
                for (int j = k + 1; j < cols; ++j) {
                    AT(k, j) -= tau * work[j];
                    for (int i = 1; i < mlen; ++i) AT(k + i, j) -= (tau * AT(k + i, k)) * work[j];
                }
            }
        }
        // norm downdate, ColPivHouseholderQR.h:553-573
        for (int j = k + 1; j < cols; ++j) {
            if (nu[j] != 0.0) {
                double temp = fabs(AT(k, j)) / nu[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0.0 ? 0.0 : temp;
                const double ratio = nu[j] / nd[j];
                const double temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    double s = 0.0;
                    for (int i = k + 1; i < rows; ++i) s += AT(i, j) * AT(i, j);
                    nd[j] = sqrt(s);
                    nu[j] = nd[j];
                } else {
                    nu[j] *= sqrt(temp);
                }
            }
        }
    }

    // _solve_impl, ColPivHouseholderQR.h:587-607
    double *coef = a.coef + a.cand_term_begin[c] + c;
    if (nonzero_pivots == 0) {
        for (int j = 0; j < cols; ++j) coef[j] = 0.0;
    } else {
        for (int k = 0; k < nonzero_pivots; ++k) {
            const int mlen = rows - k;
            const double tau = hcoef[k];
            if (mlen == 1) {
                RHS(k) *= (1.0 - tau);
            } else if (tau != 0.0) {
                double s = 0.0;
                for (int i = 1; i < mlen; ++i) s += AT(k + i, k) * RHS(k + i);
                const double tmp = s + RHS(k);
                RHS(k) -= tau * tmp;
                for (int i = 1; i < mlen; ++i) RHS(k + i) -= (tau * AT(k + i, k)) * tmp;
            }
        }
        for (int i = nonzero_pivots - 1; i >= 0; --i) {
            RHS(i) = RHS(i) / AT(i, i);
            const double xi = RHS(i);
            for (int j = 0; j < i; ++j) RHS(j) -= xi * AT(j, i);
        }
        for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = RHS(i);
        for (int i = nonzero_pivots; i < cols; ++i) x[perm[i]] = 0.0;
        for (int j = 0; j < cols; ++j) coef[j] = x[j];
    }
    a.nzp[c] = nonzero_pivots;
    a.flags[c] = RR_RES_EXACT | (nonzero_pivots < size ? RR_RES_RANKDEF : 0u);
#undef AT
#undef RHS
}

// ---------------------------------------------------------------------------------------------
// The same algorithm, one WARP per candidate, the matrix in shared memory.
//
// One thread per candidate in global memory (above) leaves a neighbourhood of 800 candidates with 25 warps on 148
// SMs, each waiting on memory at every step (9 ms per neighbourhood at n = 331: all the GPU time of a fit() on the
// small BASELINE configs). Here a warp owns a candidate and keeps A (rows x cols, column-major) and y in shared
// memory. Parity with the sequential algorithm is kept operation by operation:
//   * every SUM (column norms, the reflector's tail norm, v'A_j, the norm recomputation, v'b) is still one
//     sequential chain in row order - the chains of different columns run in different lanes, a chain that is needed
//     by the whole warp is computed by every lane redundantly (same operands, same order, same bits);
//   * everything elementwise (column swap, scaling by 1/(c0 - beta), the rank-1 update, the updates of b) is spread
//     over the lanes: each element is produced by the same expression as in the sequential code.
// So the results are bit-identical to rr_exact_qr's (tests/test_gpu_golden.py compares both with the oracle).
struct ExactWarpArgs {
    ExactArgs a;
    const int32_t *list;  // candidates of this launch
    int32_t n_list;
    int32_t kcap;         // column capacity of a warp's shared-memory matrix (>= cols of every listed candidate)
};

__device__ __forceinline__ size_t exact_warp_smem_doubles(int rows, int kcap) { return (size_t)rows * (kcap + 1) + 5 * (size_t)kcap; }

__global__ void rr_exact_qr_warp(const ExactWarpArgs w)
{
    extern __shared__ __align__(16) double rr_ex_smem[];
    const ExactArgs &a = w.a;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int li = blockIdx.x * wpb + warp;
    if (li >= w.n_list) return;
    const int c = w.list[li];
    const int rows = a.n;
    const int m = a.cand_term_begin[c + 1] - a.cand_term_begin[c];
    const int cols = m + 1;
    const size_t per_warp = exact_warp_smem_doubles(rows, w.kcap) + (size_t)((w.kcap + 1) / 2 + 1);
    double *A = rr_ex_smem + (size_t)warp * per_warp;  // column j at A + j * rows
    double *rhs = A + (size_t)rows * w.kcap;
    double *hcoef = rhs + rows, *work = hcoef + w.kcap, *nu = work + w.kcap, *nd = nu + w.kcap, *x = nd + w.kcap;
    int *perm = reinterpret_cast<int *>(x + w.kcap);
#define AT(i, j) A[(size_t)(j) * rows + (i)]

    // rils_rols_cpp.cpp:477-482: A.col(i) = factors[i]->evaluate_all(X), last column = ones
    for (int j = 0; j < m; ++j) {
        const double *src = a.V + (int64_t)a.term_ids[a.cand_term_begin[c] + j] * a.ldv;
        for (int i = lane; i < rows; i += 32) AT(i, j) = src[i];
    }
    for (int i = lane; i < rows; i += 32) {
        AT(i, m) = 1.0;
        rhs[i] = a.y[i];
    }
    __syncwarp();

    const int size = rows < cols ? rows : cols;
    // ColPivHouseholderQR.h:504-509: one column per lane, each sum sequential in row order
    for (int k = lane; k < cols; k += 32) {
        double s = 0.0;
        for (int i = 0; i < rows; ++i) s += AT(i, k) * AT(i, k);
        nd[k] = sqrt(s);
        nu[k] = nd[k];
        perm[k] = k;
    }
    __syncwarp();
    double maxnorm = nu[0];
    for (int k = 1; k < cols; ++k)
        if (nu[k] > maxnorm) maxnorm = nu[k];
    const double threshold_helper = (maxnorm * DBL_EPSILON) * (maxnorm * DBL_EPSILON) / (double)rows;  // :511
    const double norm_downdate_threshold = sqrt(DBL_EPSILON);
    int nonzero_pivots = size;

    for (int k = 0; k < size; ++k) {
        // pivot search: every lane, same data (broadcast reads)
        int big = k;
        double bigv = nu[k];
        for (int j = k + 1; j < cols; ++j)
            if (nu[j] > bigv) { bigv = nu[j]; big = j; }
        if (nonzero_pivots == size && bigv * bigv < threshold_helper * (double)(rows - k)) nonzero_pivots = k;  // :526
        __syncwarp();
        if (k != big) {  // :530-536
            for (int i = lane; i < rows; i += 32) { const double t = AT(i, k); AT(i, k) = AT(i, big); AT(i, big) = t; }
            if (lane == 0) {
                double t = nu[k]; nu[k] = nu[big]; nu[big] = t;
                t = nd[k]; nd[k] = nd[big]; nd[big] = t;
                const int ti = perm[k]; perm[k] = perm[big]; perm[big] = ti;
            }
            __syncwarp();
        }
        // makeHouseholderInPlace, Householder.h:67-98: the tail norm is one chain, computed by every lane
        const int mlen = rows - k;
        double tail_sq = 0.0;
        for (int i = 1; i < mlen; ++i) tail_sq += AT(k + i, k) * AT(k + i, k);
        const double c0 = AT(k, k);
        double beta, tau;
        __syncwarp();
        if (tail_sq <= DBL_MIN) {
            tau = 0.0;
            beta = c0;
            for (int i = 1 + lane; i < mlen; i += 32) AT(k + i, k) = 0.0;
        } else {
            beta = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0.0) beta = -beta;
            const double denom = c0 - beta;
            for (int i = 1 + lane; i < mlen; i += 32) AT(k + i, k) = AT(k + i, k) / denom;
            tau = (beta - c0) / beta;
        }
        if (lane == 0) {
            hcoef[k] = tau;
            AT(k, k) = beta;
        }
        __syncwarp();
        // applyHouseholderOnTheLeft to the trailing columns, Householder.h:116-135
        if (cols - k - 1 > 0) {
            if (mlen == 1) {
                for (int j = k + 1 + lane; j < cols; j += 32) AT(k, j) *= (1.0 - tau);
            } else if (tau != 0.0) {
                for (int j = k + 1 + lane; j < cols; j += 32) {  // v'A_j: one chain per column, one column per lane
                    double s = 0.0;
                    for (int i = 1; i < mlen; ++i) s += AT(k + i, k) * AT(k + i, j);
                    work[j] = s + AT(k, j);
                }
                __syncwarp();
                for (int j = k + 1; j < cols; ++j) {
                    const double wj = work[j];
                    for (int i = lane; i < mlen; i += 32) {
                        if (i == 0) AT(k, j) -= tau * wj;
                        else AT(k + i, j) -= (tau * AT(k + i, k)) * wj;
                    }
                }
            }
            __syncwarp();
        }
        // norm downdate, ColPivHouseholderQR.h:553-573: one column per lane
        for (int j = k + 1 + lane; j < cols; j += 32) {
            if (nu[j] != 0.0) {
                double temp = fabs(AT(k, j)) / nu[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0.0 ? 0.0 : temp;
                const double ratio = nu[j] / nd[j];
                const double temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    double s = 0.0;
                    for (int i = k + 1; i < rows; ++i) s += AT(i, j) * AT(i, j);
                    nd[j] = sqrt(s);
                    nu[j] = nd[j];
                } else {
                    nu[j] *= sqrt(temp);
                }
            }
        }
        __syncwarp();
    }

    // _solve_impl, ColPivHouseholderQR.h:587-607
    double *coef = a.coef + a.cand_term_begin[c] + c;
    if (nonzero_pivots == 0) {
        for (int j = lane; j < cols; j += 32) coef[j] = 0.0;
    } else {
        for (int k = 0; k < nonzero_pivots; ++k) {
            const int mlen = rows - k;
            const double tau = hcoef[k];
            if (mlen == 1) {
                if (lane == 0) rhs[k] *= (1.0 - tau);
            } else if (tau != 0.0) {
                double s = 0.0;  // v'b: one chain, every lane
                for (int i = 1; i < mlen; ++i) s += AT(k + i, k) * rhs[k + i];
                const double tmp = s + rhs[k];
                __syncwarp();
                for (int i = lane; i < mlen; i += 32) {
                    if (i == 0) rhs[k] -= tau * tmp;
                    else rhs[k + i] -= (tau * AT(k + i, k)) * tmp;
                }
            }
            __syncwarp();
        }
        for (int i = nonzero_pivots - 1; i >= 0; --i) {
            const double xi = rhs[i] / AT(i, i);
            __syncwarp();
            if (lane == 0) rhs[i] = xi;
            for (int j = lane; j < i; j += 32) rhs[j] -= xi * AT(j, i);
            __syncwarp();
        }
        for (int i = lane; i < cols; i += 32) x[perm[i]] = i < nonzero_pivots ? rhs[i] : 0.0;
        __syncwarp();
        for (int j = lane; j < cols; j += 32) coef[j] = x[j];
    }
    if (lane == 0) {
        a.nzp[c] = nonzero_pivots;
        a.flags[c] = RR_RES_EXACT | (nonzero_pivots < size ? RR_RES_RANKDEF : 0u);
    }
#undef AT
}

struct ResidColsArgs {
    const double *V;
    int64_t ldv;
    const double *y;
    int32_t n;
    const int32_t *term_ids;
    const int32_t *cand_term_begin;
    int32_t n_cand;
    const double *coef;  // raw coefficients (batch layout); snapped here
    double *ssr;         // out
    uint32_t *flags;     // in/out: NONFINITE is added
};

// SSR of the rebuilt model for the exact path: one warp per candidate, yhat in the association
// order of rils_rols_cpp.cpp:488-515 ((c0*t0 + c1*t1) + ...) + c_free with snapped coefficients.
__global__ void rr_resid_cols(const ResidColsArgs a)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= a.n_cand) return;
    const int t0 = a.cand_term_begin[c];
    const int m = a.cand_term_begin[c + 1] - t0;
    const double *coef = a.coef + t0 + c;
    double part = 0.0;
    for (int i = lane; i < a.n; i += 32) {
        double yh = 0.0;
        bool first = true;
        for (int j = 0; j < m; ++j) {
            const double cj = coef[j];
            if (fabs(cj) < 1e-12) continue;  // value_zero
            const double v = a.V[(int64_t)a.term_ids[t0 + j] * a.ldv + i];
            const double term = fabs(cj - 1.0) < 1e-12 ? v : __dmul_rn(cj, v);  // value_one
            yh = first ? term : __dadd_rn(yh, term);
            first = false;
        }
        const double cf = coef[m];
        if (!(fabs(cf) < 1e-12)) {
            const double cv = __dmul_rn(cf, 1.0);
            yh = first ? cv : __dadd_rn(yh, cv);
            first = false;
        }
        const double r = __dsub_rn(a.y[i], yh);
        part = __dadd_rn(part, __dmul_rn(r, r));
    }
#pragma unroll
    for (int mk = 16; mk > 0; mk >>= 1) part += __shfl_xor_sync(0xffffffffu, part, mk);
    if (lane == 0) {
        a.ssr[c] = part;
        if (!isfinite(part)) a.flags[c] |= RR_RES_NONFINITE;
    }
}

}  // namespace rr
