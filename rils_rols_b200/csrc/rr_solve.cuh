// rr_solve.cuh — per-candidate small dense kernels (one thread per candidate):
//   rr_gram_solve     Gram path: pivoted Cholesky of A^T A, coefficients, Gram-derived SSR, and the
//                     decision whether the candidate needs the explicit-residual refinement or the
//                     double-double escalation.
//   rr_refine_update  after the residual sweep: one step of iterative refinement
//                     (corrected semi-normal equations) + SSR update.
//   rr_gram_solve_dd  the same factorisation in double-double arithmetic with the reference's own
//                     rank rule, for candidates whose Gram matrix is numerically singular in fp64.
//
// What they replace: `A.colPivHouseholderQr().solve(b)`, /root/reference/rils_rols_cpp/rils_rols_cpp.cpp:484
// (Eigen ColPivHouseholderQR.h:482-607). Column-pivoted QR picks, at every step, the column with
// the largest updated norm; diagonal-pivoted Cholesky of the Gram matrix picks the largest Schur
// diagonal, which is the same quantity squared, so both choose the same pivot order in exact
// arithmetic and drop the same columns. The rank rule (ColPivHouseholderQR.h:511,526-527):
//   updated_norm^2 < (eps * max_col_norm)^2 * (rows - step) / rows   =>  nonzero_pivots = step.
// The coefficient snapping of rils_rols_cpp.cpp:492-505 (value_zero / value_one, node.h:333-339)
// is applied here so that the SSR belongs to the model fitness() would score.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/rr_b200.h"

namespace rr {

struct SolveConsts {
    double n_total;   // rows of A over all ranks
    double y_mean;
    double sum_yc;    // sum of centred y (~0)
    double sst;       // yc . yc
    double rho_accurate;  // min Schur/orig ratio for "no refinement needed"
    double rho_escalate;  // below this the fp64 Gram matrix is treated as singular
    double ssr_rel_tol;   // acceptable relative error bound of the Gram-derived SSR
};

// per-candidate status written by rr_gram_solve (bit set)
enum : uint32_t {
    ST_DONE = 0,
    ST_NEED_REFINE = 1u << 8,
    ST_NEED_ESCALATE = 1u << 9,
};

#define RR_EPS_SNAP 1e-12  // pow(10, -PRECISION), node.h:13-14

__device__ __forceinline__ double snap_coef(double c, bool free_term)
{
    if (fabs(c) < RR_EPS_SNAP) return 0.0;                    // value_zero -> term dropped
    if (!free_term && fabs(c - 1.0) < RR_EPS_SNAP) return 1.0;  // value_one -> bare term
    return c;
}

// Gathers the k x k Gram matrix (lower triangle used) and rhs of candidate c.
// idx: m(m+1)/2 upper-triangle ids (row-major), m ids (t_i . yc), m ids (t_i . 1).
__device__ inline void gather_gram(const double *dots, const int32_t *idx, int m, const SolveConsts &k,
                                   double *G, int ldg, double *b)
{
    int p = 0;
    for (int i = 0; i < m; ++i)
        for (int j = i; j < m; ++j) {
            const double v = dots[idx[p++]];
            G[i * ldg + j] = v;
            G[j * ldg + i] = v;
        }
    for (int i = 0; i < m; ++i) b[i] = dots[idx[p++]];
    for (int i = 0; i < m; ++i) {
        const double v = dots[idx[p++]];
        G[i * ldg + m] = v;
        G[m * ldg + i] = v;
    }
    G[m * ldg + m] = k.n_total;
    b[m] = k.sum_yc;
}

// Columns that take part in the factorisation: all but the LATER copies of a term that a candidate lists
// twice (neighbours that rebuild an existing term do). The planner reduces each distinct (term, term) pair
// once, so copies share their reduction ids and their Gram rows are bit-identical: the second copy's
// updated norm is exactly zero after the first has been eliminated, column-pivoted QR drops it
// (ColPivHouseholderQR.h:517-527: the larger index loses the tie and ends below the threshold) and its
// coefficient is 0. Deciding that from the ids instead of from rounding noise keeps such candidates out
// of the double-double escalation. amap: alive column indices, the free term (index m) last.
// The same for terms that are constant BY CONSTRUCTION (cmask, BatchPlanner::cand_const_mask: sin(c), t / t, ...):
// their columns are multiples c_j of the free term's column of ones (exact, or up to the rounding of (c t) / t). In
// exact arithmetic - the reference's floating-point QR sometimes keeps both with coefficients of +-1e15 instead, by
// rounding residue: SURVEY App. B.6 - column-pivoted QR picks the longest of
// these parallel columns when its turn comes (squared norms c_j^2 n against n, unchanged in proportion by the
// reflectors before it; ties go to the lower index: ColPivHouseholderQR.h:517-522) and finds every other one at
// exactly zero afterwards, below the threshold: coefficient 0. So: of the constant terms and the free term only the
// one with the largest |c| stays (c_j = (t_j . 1) / n, which is exact for c = 1: t / t), ties to the lower index.
__device__ inline int alive_columns(const int32_t *idx, int m, int *amap, uint32_t cmask, const double *G, int ldg, double n_total)
{
    int keep_const = m;  // the free term, c = 1
    if (cmask && m <= 32) {
        double best = 1.0;
        for (int j = m - 1; j >= 0; --j)
            if ((cmask >> j) & 1u) {
                const double cj = fabs(G[j * ldg + m] / n_total);
                if (cj >= best) { best = cj; keep_const = j; }  // descending j: on a tie the lower index wins
            }
    } else {
        cmask = 0u;
    }
    int na = 0;
    for (int j = 0; j < m; ++j) {
        if (((cmask >> j) & 1u) && j != keep_const) continue;
        const int32_t dj = idx[j * m - j * (j - 1) / 2];  // id of t_j . t_j in the row-major upper triangle
        bool dup = false;
        for (int i = 0; i < j && !dup; ++i) dup = idx[i * m - i * (i - 1) / 2] == dj;
        if (!dup) amap[na++] = j;
    }
    if (keep_const == m) amap[na++] = m;
    return na;
}

// In-place diagonal-pivoted Cholesky of the symmetric W (kk x kk, both triangles valid on entry;
// on exit the lower triangle of the leading rank x rank block holds L in pivoted order).
// Returns the numerical rank by the reference's rule; rho_min = min over pivots of
// (Schur diagonal / original diagonal of that column).
__device__ inline int pivoted_cholesky(double *W, int ldw, int kk, int *perm, double *orig, double n_rows,
                                       double *rho_min_out)
{
    double maxd = 0.0;
    for (int i = 0; i < kk; ++i) {
        perm[i] = i;
        orig[i] = W[i * ldw + i];
        if (orig[i] > maxd) maxd = orig[i];
    }
    // (eps * maxnorm)^2 / rows, ColPivHouseholderQR.h:511
    const double thr = (DBL_EPSILON * DBL_EPSILON) * maxd / n_rows;
    double rho_min = 1.0;
    int rank = kk;
    for (int s = 0; s < kk; ++s) {
        int p = s;
        double best = W[s * ldw + s];
        for (int j = s + 1; j < kk; ++j)
            if (W[j * ldw + j] > best) { best = W[j * ldw + j]; p = j; }
        if (!(best > 0.0) || best < thr * (n_rows - s)) {  // :526-527
            if (best < 0.0) rho_min = 0.0;  // rounding garbage, not a verdict: the caller escalates
            rank = s;
            break;
        }
        if (p != s) {  // symmetric row/column swap of the trailing working matrix and of L's rows
            for (int j = 0; j < kk; ++j) { const double t = W[s * ldw + j]; W[s * ldw + j] = W[p * ldw + j]; W[p * ldw + j] = t; }
            for (int j = 0; j < kk; ++j) { const double t = W[j * ldw + s]; W[j * ldw + s] = W[j * ldw + p]; W[j * ldw + p] = t; }
            const int ti = perm[s]; perm[s] = perm[p]; perm[p] = ti;
            const double to = orig[s]; orig[s] = orig[p]; orig[p] = to;
        }
        const double rho = best / orig[s];
        if (rho < rho_min) rho_min = rho;
        const double l = sqrt(best);
        W[s * ldw + s] = l;
        for (int j = s + 1; j < kk; ++j) W[j * ldw + s] = W[j * ldw + s] / l;
        for (int j = s + 1; j < kk; ++j) {
            const double ljs = W[j * ldw + s];
            for (int i = s + 1; i <= j; ++i) {
                const double v = fma(-ljs, W[i * ldw + s], W[j * ldw + i]);
                W[j * ldw + i] = v;
                W[i * ldw + j] = v;
            }
        }
    }
    *rho_min_out = rho_min;
    return rank;
}

// x (original order) = solution of (L L^T) z = rhs[perm] on the leading `rank` pivots, 0 elsewhere
__device__ inline void cholesky_solve(const double *W, int ldw, int kk, int rank, const int *perm,
                                      const double *rhs, double *z, double *x)
{
    for (int i = 0; i < rank; ++i) {
        double s = rhs[perm[i]];
        for (int j = 0; j < i; ++j) s = fma(-W[i * ldw + j], z[j], s);
        z[i] = s / W[i * ldw + i];
    }
    for (int i = rank - 1; i >= 0; --i) {
        double s = z[i];
        for (int j = i + 1; j < rank; ++j) s = fma(-W[j * ldw + i], z[j], s);
        z[i] = s / W[i * ldw + i];
    }
    for (int i = 0; i < kk; ++i) x[i] = 0.0;
    for (int i = 0; i < rank; ++i) x[perm[i]] = z[i];
}

// quadratic form ssr = sst - 2 c.b + c.G.c with an a-priori rounding bound
__device__ inline double gram_ssr(const double *G, int ldg, const double *b, const double *c, int kk, double sst,
                                  double *bound)
{
    double lin = 0.0, quad = 0.0, alin = 0.0, aquad = 0.0;
    for (int i = 0; i < kk; ++i) {
        lin = fma(c[i], b[i], lin);
        alin += fabs(c[i] * b[i]);
        double row = 0.0, arow = 0.0;
        for (int j = 0; j < kk; ++j) {
            row = fma(G[i * ldg + j], c[j], row);
            arow += fabs(G[i * ldg + j] * c[j]);
        }
        quad = fma(c[i], row, quad);
        aquad += fabs(c[i]) * arow;
    }
    *bound = DBL_EPSILON * (kk + 4.0) * (fabs(sst) + 2.0 * alin + aquad);
    return sst - 2.0 * lin + quad;
}

struct GramArgs {
    const double *dots;
    const int32_t *cand_dot;        // index tables
    const int32_t *cand_dot_begin;  // [n_list + 1]
    const int32_t *list;            // candidate ids (nullptr: identity)
    const int32_t *cand_term_begin; // batch offsets (k = terms + 1, coef offset = begin + c)
    const uint32_t *cand_const;     // per candidate id: terms that are constant by construction (nullptr: none)
    int32_t n_list;
    double *ws;                     // workspace: per listed candidate 2*kk*kk + 10*kk doubles
    const int64_t *ws_begin;        // [n_list + 1] offsets into ws (doubles)
    SolveConsts k;
    // outputs, indexed by candidate id / coef offset
    double *coef;
    double *coef_snapped;
    int32_t *nzp;
    double *ssr;
    uint32_t *flags;
    uint32_t *status;
};

__global__ void rr_gram_solve(const GramArgs a)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= a.n_list) return;
    const int c = a.list ? a.list[li] : li;
    const int m = a.cand_term_begin[c + 1] - a.cand_term_begin[c];
    const int kk = m + 1;
    const int32_t *idx = a.cand_dot + a.cand_dot_begin[li];
    double *G = a.ws + a.ws_begin[li];
    double *W = G + kk * kk;  // factor workspace (copy of G)
    double *b = W + kk * kk, *rhs = b + kk, *z = rhs + kk, *x = z + kk, *orig = x + kk;
    int *perm = reinterpret_cast<int *>(orig + kk);
    double *coef = a.coef + a.cand_term_begin[c] + c;
    double *cs = a.coef_snapped + a.cand_term_begin[c] + c;

    gather_gram(a.dots, idx, m, a.k, G, kk, b);
    // b = A^T yc is what the sweep reduces (centred: no cancellation against sum(y)^2 in the SSR);
    // the system is solved for the uncentred target, A^T y = A^T yc + mean * A^T 1, so pivoting and
    // rank decisions see the same columns as the reference
    for (int i = 0; i < kk; ++i) rhs[i] = fma(a.k.y_mean, G[i * kk + m], b[i]);
    bool finite = true;
    for (int i = 0; i < kk; ++i) finite = finite && isfinite(G[i * kk + i]) && isfinite(b[i]);
    if (!finite) {
        // a term is NaN/inf somewhere: the reference's QR turns everything into NaN and fitness()
        // returns the sentinel (rils_rols_cpp.cpp:529-530,536-537)
        for (int i = 0; i < kk; ++i) { coef[i] = nan(""); cs[i] = nan(""); }
        a.nzp[c] = kk;
        a.ssr[c] = nan("");
        a.flags[c] = RR_RES_NONFINITE;
        a.status[c] = ST_DONE;
        return;
    }
    // workspace tail (4.5 kk doubles are free behind perm): alive map, compacted right-hand side and solution
    int *amap = perm + kk;
    double *rhs2 = reinterpret_cast<double *>(amap + kk), *x2 = rhs2 + kk;
    const int na = alive_columns(idx, m, amap, a.cand_const ? a.cand_const[c] : 0u, G, kk, a.k.n_total);
    for (int p = 0; p < na; ++p) {
        for (int q = 0; q < na; ++q) W[p * na + q] = G[amap[p] * kk + amap[q]];
        rhs2[p] = rhs[amap[p]];
    }
    double rho_min;
    const int rank = pivoted_cholesky(W, na, na, perm, orig, a.k.n_total, &rho_min);
    cholesky_solve(W, na, na, rank, perm, rhs2, z, x2);
    for (int i = 0; i < kk; ++i) x[i] = 0.0;
    for (int p = 0; p < na; ++p) x[amap[p]] = x2[p];

    uint32_t flags = rank < kk ? RR_RES_RANKDEF : 0u, status = ST_DONE;
    double cmax = 0.0;
    for (int i = 0; i < kk; ++i) {
        coef[i] = x[i];
        cmax = fmax(cmax, fabs(coef[i]));
    }
    // snapping and its ambiguity: a coefficient whose distance to a snap threshold is inside the
    // error estimate cannot be decided from the fp64 Gram solution. The error of c_i scales with
    // |y| / |t_i| (not with |c|): re-fitted models have exact coefficients 1 / free term 0.
    const double ynorm2 = a.k.sst + a.k.n_total * a.k.y_mean * a.k.y_mean;
    (void)cmax;
    bool ambiguous = false;
    uint64_t alive_bits = 0;  // (kk <= 64 wherever a column is dropped structurally; wider candidates have all alive)
    for (int p = 0; p < na; ++p)
        if (amap[p] < 64) alive_bits |= 1ull << amap[p];
    for (int i = 0; i < kk; ++i) {
        cs[i] = snap_coef(coef[i], i == m);
        // a column dropped by construction (a duplicate, a constant term) has the coefficient 0 exactly, as in the
        // reference: nothing to decide
        if (na < kk && i < 64 && !((alive_bits >> i) & 1ull)) continue;
        const double gii = G[i * kk + i];
        const double cerr = 64.0 * (DBL_EPSILON / fmax(rho_min, DBL_MIN)) * sqrt(ynorm2 / fmax(gii, DBL_MIN));
        const double d0 = fabs(fabs(coef[i]) - RR_EPS_SNAP);
        const double d1 = fabs(fabs(coef[i] - 1.0) - RR_EPS_SNAP);
        if (d0 < cerr || (i != m && d1 < cerr)) ambiguous = true;
    }
    // SSR of the snapped model from the Gram quantities: r = yc - sum cs_i t_i - (cs_free - mean)
    for (int i = 0; i < kk; ++i) z[i] = i == m ? cs[i] - a.k.y_mean : cs[i];
    double bound;
    double ssr = gram_ssr(G, kk, b, z, kk, a.k.sst, &bound);
    if (rho_min < a.k.rho_escalate) status |= ST_NEED_ESCALATE;
    else if (rho_min < a.k.rho_accurate || ambiguous || !(bound <= a.k.ssr_rel_tol * ssr)) status |= ST_NEED_REFINE;
    if (ssr < 0.0) ssr = 0.0;
    a.nzp[c] = rank;
    a.ssr[c] = ssr;
    a.flags[c] = flags;
    a.status[c] = status;
}

struct RefineArgs {
    GramArgs g;                       // same tables as the Gram solve of these candidates
    const double *rdots;              // residual-sweep outputs
    const int32_t *rcand_dot;         // per listed candidate: r.r, r.t_i (m), r.1
    const int32_t *rcand_dot_begin;
    double *delta_rel;                // [n_cand] |delta|_inf / |c|_inf of this step
};

// One refinement step. The residual sweep evaluated r0 = y - A cs0 explicitly with the snapped
// coefficients cs0 (in a.g.coef_snapped). delta = G^-1 A^T r0 on the pivot columns; c1 = cs0 + delta;
// ssr(c1s) = ssr0 - 2 d.g + d.G.d with d = c1s - cs0 (exact identity; every term is O(|d|)).
__global__ void rr_refine_update(const RefineArgs a)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= a.g.n_list) return;
    const int c = a.g.list ? a.g.list[li] : li;
    const int m = a.g.cand_term_begin[c + 1] - a.g.cand_term_begin[c];
    const int kk = m + 1;
    double *G = a.g.ws + a.g.ws_begin[li];
    double *W = G + kk * kk;
    double *b = W + kk * kk, *rhs = b + kk, *z = rhs + kk, *x = z + kk, *orig = x + kk;
    (void)rhs;
    int *perm = reinterpret_cast<int *>(orig + kk);
    double *coef = a.g.coef + a.g.cand_term_begin[c] + c;
    double *cs = a.g.coef_snapped + a.g.cand_term_begin[c] + c;
    const int32_t *ridx = a.rcand_dot + a.rcand_dot_begin[li];

    const int32_t *idx = a.g.cand_dot + a.g.cand_dot_begin[li];
    gather_gram(a.g.dots, idx, m, a.g.k, G, kk, b);
    int *amap = perm + kk;
    double *g2 = reinterpret_cast<double *>(amap + kk), *x2 = g2 + kk;
    const int na = alive_columns(idx, m, amap, a.g.cand_const ? a.g.cand_const[c] : 0u, G, kk, a.g.k.n_total);  // see rr_gram_solve
    for (int p = 0; p < na; ++p)
        for (int q = 0; q < na; ++q) W[p * na + q] = G[amap[p] * kk + amap[q]];
    double rho_min;
    const int rank = pivoted_cholesky(W, na, na, perm, orig, a.g.k.n_total, &rho_min);
    const double ssr0 = a.rdots[ridx[0]];
    double *g = b;  // reuse: g = A^T r0
    for (int i = 0; i < m; ++i) g[i] = a.rdots[ridx[1 + i]];
    g[m] = a.rdots[ridx[1 + m]];
    if (!isfinite(ssr0)) {
        a.g.ssr[c] = ssr0;
        a.g.flags[c] |= RR_RES_NONFINITE | RR_RES_REFINED;
        a.delta_rel[c] = 0.0;
        return;
    }
    for (int p = 0; p < na; ++p) g2[p] = g[amap[p]];
    cholesky_solve(W, na, na, rank, perm, g2, z, x2);  // delta on the alive columns
    for (int i = 0; i < kk; ++i) x[i] = 0.0;
    for (int p = 0; p < na; ++p) x[amap[p]] = x2[p];
    double dmax = 0.0, cmax = 0.0;
    for (int i = 0; i < kk; ++i) {
        const double c1 = cs[i] + x[i];
        coef[i] = c1;
        dmax = fmax(dmax, fabs(x[i]));
        cmax = fmax(cmax, fabs(c1));
    }
    // d = snapped(c1) - cs0
    for (int i = 0; i < kk; ++i) {
        const double c1s = snap_coef(coef[i], i == m);
        z[i] = c1s - cs[i];
        cs[i] = c1s;
    }
    double lin = 0.0, quad = 0.0;
    for (int i = 0; i < kk; ++i) {
        lin = fma(z[i], g[i], lin);
        double row = 0.0;
        for (int j = 0; j < kk; ++j) row = fma(G[i * kk + j], z[j], row);
        quad = fma(z[i], row, quad);
    }
    double ssr1 = ssr0 - 2.0 * lin + quad;
    if (!(ssr1 >= 0.0)) ssr1 = ssr1 < 0.0 ? 0.0 : ssr1;
    a.g.ssr[c] = ssr1;
    a.g.flags[c] |= RR_RES_REFINED;
    a.delta_rel[c] = cmax > 0.0 ? dmax / cmax : 0.0;
}

// ---------------------------------------------------------------------------------------------
// double-double arithmetic for the escalation path
// ---------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd dd_make(double h, double l)
{
    const double s = h + l;
    dd r;
    r.hi = s;
    r.lo = l - (s - h);
    return r;
}
__device__ __forceinline__ dd dd_sum(dd a, dd b)
{
    const double s = a.hi + b.hi;
    const double bb = s - a.hi;
    const double e = (a.hi - (s - bb)) + (b.hi - bb);
    return dd_make(s, e + a.lo + b.lo);
}
__device__ __forceinline__ dd dd_neg(dd a)
{
    dd r;
    r.hi = -a.hi;
    r.lo = -a.lo;
    return r;
}
__device__ __forceinline__ dd dd_mul(dd a, dd b)
{
    const double p = a.hi * b.hi;
    const double e = fma(a.hi, b.hi, -p);
    return dd_make(p, e + (a.hi * b.lo + a.lo * b.hi));
}
__device__ __forceinline__ dd dd_div(dd a, dd b)
{
    const double q1 = a.hi / b.hi;
    dd r = dd_sum(a, dd_neg(dd_mul(b, dd_make(q1, 0.0))));
    const double q2 = r.hi / b.hi;
    r = dd_sum(r, dd_neg(dd_mul(b, dd_make(q2, 0.0))));
    const double q3 = r.hi / b.hi;
    return dd_sum(dd_make(q1, q2), dd_make(q3, 0.0));
}
__device__ __forceinline__ dd dd_sqrt(dd a)
{
    if (a.hi <= 0.0) return dd_make(0.0, 0.0);
    const double x = 1.0 / sqrt(a.hi);
    const double ax = a.hi * x;
    const dd ax2 = dd_mul(dd_make(ax, 0.0), dd_make(ax, 0.0));
    const double corr = dd_sum(a, dd_neg(ax2)).hi * (x * 0.5);
    return dd_make(ax, corr);
}

// Same as rr_gram_solve with every quantity in double-double. The dots buffer holds (hi, lo)
// pairs at ids idx and idx+1. Workspace: per candidate 2*kk*kk + 10*kk doubles.
__global__ void rr_gram_solve_dd(const GramArgs a)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= a.n_list) return;
    const int c = a.list ? a.list[li] : li;
    const int m = a.cand_term_begin[c + 1] - a.cand_term_begin[c];
    const int kk = m + 1;
    const int32_t *idx = a.cand_dot + a.cand_dot_begin[li];
    dd *W = reinterpret_cast<dd *>(a.ws + a.ws_begin[li]);
    dd *b = W + kk * kk, *z = b + kk, *orig = z + kk;
    int *perm = reinterpret_cast<int *>(orig + kk);
    double *coef = a.coef + a.cand_term_begin[c] + c;
    double *cs = a.coef_snapped + a.cand_term_begin[c] + c;

    int p = 0;
    for (int i = 0; i < m; ++i)
        for (int j = i; j < m; ++j) {
            const dd v = dd_make(a.dots[idx[p]], a.dots[idx[p] + 1]);
            ++p;
            W[i * kk + j] = v;
            W[j * kk + i] = v;
        }
    for (int i = 0; i < m; ++i) { b[i] = dd_make(a.dots[idx[p]], a.dots[idx[p] + 1]); ++p; }
    for (int i = 0; i < m; ++i) {
        const dd v = dd_make(a.dots[idx[p]], a.dots[idx[p] + 1]);
        ++p;
        W[i * kk + m] = v;
        W[m * kk + i] = v;
    }
    W[m * kk + m] = dd_make(a.k.n_total, 0.0);
    b[m] = dd_make(a.k.sum_yc, 0.0);
    // uncentred right-hand side A^T y = A^T yc + mean * A^T 1 (see rr_gram_solve)
    for (int i = 0; i < kk; ++i) b[i] = dd_sum(b[i], dd_mul(dd_make(a.k.y_mean, 0.0), W[i * kk + m]));
    bool finite = true;
    for (int i = 0; i < kk; ++i) finite = finite && isfinite(W[i * kk + i].hi) && isfinite(b[i].hi);
    if (!finite) {
        for (int i = 0; i < kk; ++i) { coef[i] = nan(""); cs[i] = nan(""); }
        a.nzp[c] = kk;
        a.ssr[c] = nan("");
        a.flags[c] = RR_RES_NONFINITE | RR_RES_DD;
        a.status[c] = ST_DONE;
        return;
    }
    double maxd = 0.0;
    for (int i = 0; i < kk; ++i) {
        perm[i] = i;
        orig[i] = W[i * kk + i];
        if (orig[i].hi > maxd) maxd = orig[i].hi;
    }
    const double thr = (DBL_EPSILON * DBL_EPSILON) * maxd / a.k.n_total;
    int rank = kk;
    for (int s = 0; s < kk; ++s) {
        int pv = s;
        dd best = W[s * kk + s];
        for (int j = s + 1; j < kk; ++j) {
            const dd v = W[j * kk + j];
            if (v.hi > best.hi || (v.hi == best.hi && v.lo > best.lo)) { best = v; pv = j; }
        }
        if (!(best.hi > 0.0) || best.hi < thr * (a.k.n_total - s)) { rank = s; break; }
        if (pv != s) {
            for (int j = 0; j < kk; ++j) { const dd t = W[s * kk + j]; W[s * kk + j] = W[pv * kk + j]; W[pv * kk + j] = t; }
            for (int j = 0; j < kk; ++j) { const dd t = W[j * kk + s]; W[j * kk + s] = W[j * kk + pv]; W[j * kk + pv] = t; }
            const int ti = perm[s]; perm[s] = perm[pv]; perm[pv] = ti;
            const dd to = orig[s]; orig[s] = orig[pv]; orig[pv] = to;
        }
        const dd l = dd_sqrt(best);
        W[s * kk + s] = l;
        for (int j = s + 1; j < kk; ++j) W[j * kk + s] = dd_div(W[j * kk + s], l);
        for (int j = s + 1; j < kk; ++j) {
            const dd ljs = W[j * kk + s];
            for (int i = s + 1; i <= j; ++i) {
                const dd v = dd_sum(W[j * kk + i], dd_neg(dd_mul(ljs, W[i * kk + s])));
                W[j * kk + i] = v;
                W[i * kk + j] = v;
            }
        }
    }
    for (int i = 0; i < rank; ++i) {
        dd s = b[perm[i]];
        for (int j = 0; j < i; ++j) s = dd_sum(s, dd_neg(dd_mul(W[i * kk + j], z[j])));
        z[i] = dd_div(s, W[i * kk + i]);
    }
    for (int i = rank - 1; i >= 0; --i) {
        dd s = z[i];
        for (int j = i + 1; j < rank; ++j) s = dd_sum(s, dd_neg(dd_mul(W[j * kk + i], z[j])));
        z[i] = dd_div(s, W[i * kk + i]);
    }
    for (int i = 0; i < kk; ++i) coef[i] = 0.0;
    for (int i = 0; i < rank; ++i) coef[perm[i]] = z[i].hi;
    for (int i = 0; i < kk; ++i) cs[i] = snap_coef(coef[i], i == m);
    a.nzp[c] = rank;
    a.flags[c] = (rank < kk ? RR_RES_RANKDEF : 0u) | RR_RES_DD;
    a.status[c] = ST_DONE;  // ssr comes from the explicit residual sweep that follows
}

}  // namespace rr
