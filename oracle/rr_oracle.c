/*
 * TEST INFRASTRUCTURE — CPU oracle. Not part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.
 *
 * rr_oracle.c — plain-C restatement of the RILS-ROLS scoring hot path, taking the
 * same rr_batch the product's C ABI takes (include/rr_b200.h). Paths relative to
 * /root/reference/rils_rols_cpp:
 *
 *   eval_program()      node::evaluate_all / evaluate_inner          node.cpp:5-95
 *   score OLS_FIT       tune_constants(): design matrix + QR solve   rils_rols_cpp.cpp:474-484
 *                       coefficient snapping / tree rebuild          rils_rols_cpp.cpp:488-517, node.h:333-339
 *   fitness tuple       fitness(), R2(), RMSE()                      rils_rols_cpp.cpp:520-541, :40-49
 *   classifier metrics  classification_accuracy/average_log_loss/average_loss  rils_rols_cpp.cpp:51-86
 *   QR                  rr_colpiv_qr.h (Eigen ColPivHouseholderQR restated)
 *
 * Parity pin: this file is checked against the unmodified reference built in
 * oracle/_ref (tests/test_oracle_vs_reference.py, run where /root/reference exists)
 * and against the golden fixtures that build produced (tests/golden/). The reference
 * itself ships no golden vectors for this path.
 *
 * Build with -ffp-contract=off (see Makefile): separately rounded * and +, like the
 * reference's array-at-a-time evaluation.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rr_b200.h"
#include "rr_colpiv_qr.h"

#define RR_ORACLE_API __attribute__((visibility("default")))

static int op_arity(uint32_t op) /* node.h:40-56 */
{
    switch (op) {
    case RR_OP_CONST:
    case RR_OP_VAR:
        return 0;
    case RR_OP_SIN:
    case RR_OP_COS:
    case RR_OP_LN:
    case RR_OP_EXP:
    case RR_OP_SQRT:
    case RR_OP_SQR:
        return 1;
    default:
        return 2;
    }
}

/* node.cpp:23-95, one n-vector per stack entry. out[n]. Returns 0, or -1 if malformed. */
static int eval_program(const uint32_t *code, int32_t len, const double *consts, int32_t n_consts,
                        const double *X, int64_t n, int32_t d, double *out)
{
    int depth = 0, maxdepth = 0, i;
    int64_t s;
    for (i = 0; i < len; ++i) {
        const uint32_t op = RR_INS_OP(code[i]);
        if (op == RR_OP_NONE || op >= RR_OP_COUNT) return -1;
        const int ar = op_arity(op);
        if (depth < ar) return -1;
        if (op == RR_OP_CONST && (int32_t)RR_INS_ARG(code[i]) >= n_consts) return -1;
        if (op == RR_OP_VAR && (int32_t)RR_INS_ARG(code[i]) >= d) return -1;
        depth += 1 - ar;
        if (depth > maxdepth) maxdepth = depth;
    }
    if (depth != 1) return -1;
    double *stack = (double *)malloc(sizeof(double) * (size_t)n * (size_t)maxdepth);
    if (!stack) return -2;
    int sp = 0;
    for (i = 0; i < len; ++i) {
        const uint32_t op = RR_INS_OP(code[i]), arg = RR_INS_ARG(code[i]);
        const int ar = op_arity(op);
        double *a = ar >= 1 ? stack + (size_t)(sp - ar) * n : NULL; /* left operand, also result */
        double *b = ar == 2 ? stack + (size_t)(sp - 1) * n : NULL;  /* right operand */
        double *r = ar == 0 ? stack + (size_t)sp * n : a;
        switch (op) {
        case RR_OP_CONST: for (s = 0; s < n; ++s) r[s] = consts[arg]; break;
        case RR_OP_VAR: memcpy(r, X + (size_t)arg * n, sizeof(double) * (size_t)n); break;
        case RR_OP_PLUS: for (s = 0; s < n; ++s) r[s] = a[s] + b[s]; break;
        case RR_OP_MINUS: for (s = 0; s < n; ++s) r[s] = a[s] - b[s]; break;
        case RR_OP_MULTIPLY: for (s = 0; s < n; ++s) r[s] = a[s] * b[s]; break;
        case RR_OP_DIVIDE: for (s = 0; s < n; ++s) r[s] = a[s] / b[s]; break;
        case RR_OP_SIN: for (s = 0; s < n; ++s) r[s] = sin(a[s]); break;
        case RR_OP_COS: for (s = 0; s < n; ++s) r[s] = cos(a[s]); break;
        case RR_OP_LN: for (s = 0; s < n; ++s) r[s] = log(a[s]); break;
        case RR_OP_EXP: for (s = 0; s < n; ++s) r[s] = exp(a[s]); break;
        case RR_OP_SQRT: for (s = 0; s < n; ++s) r[s] = sqrt(a[s]); break;
        case RR_OP_SQR: for (s = 0; s < n; ++s) r[s] = a[s] * a[s]; break;
        case RR_OP_POW: for (s = 0; s < n; ++s) r[s] = pow(a[s], b[s]); break;
        case RR_OP_LESS_THAN: for (s = 0; s < n; ++s) r[s] = a[s] < b[s] ? 1 : 0; break;
        case RR_OP_GREATER_THAN: for (s = 0; s < n; ++s) r[s] = a[s] > b[s] ? 1 : 0; break;
        case RR_OP_EQUAL: for (s = 0; s < n; ++s) r[s] = a[s] == b[s] ? 1 : 0; break;
        case RR_OP_NOT_EQUAL: for (s = 0; s < n; ++s) r[s] = a[s] != b[s] ? 1 : 0; break;
        case RR_OP_MIN: for (s = 0; s < n; ++s) r[s] = a[s] < b[s] ? a[s] : b[s]; break;
        case RR_OP_MAX: for (s = 0; s < n; ++s) r[s] = a[s] > b[s] ? a[s] : b[s]; break;
        default: free(stack); return -1;
        }
        sp += 1 - ar;
    }
    memcpy(out, stack, sizeof(double) * (size_t)n);
    free(stack);
    return 0;
}

RR_ORACLE_API int rr_oracle_eval(const uint32_t *code, int32_t len, const double *consts,
                                 int32_t n_consts, const double *X, int64_t n, int32_t d, double *out)
{
    return eval_program(code, len, consts, n_consts, X, n, d, out);
}

/* rils_rols_cpp.cpp:40-49 + :526-538 given yhat. f0 = 1 - R2 exactly as the reference
 * forms it (1 - (1 - ssr/sst)), f1 = RMSE; NaN in either -> sentinel 1000/1000/1000. */
static void fitness_from_yhat(const double *y, const double *yp, int64_t n, int32_t size, double *ssr_out,
                              double *f0, double *f1, int32_t *fsize)
{
    int64_t s;
    double ysum = 0.0, ssr = 0.0, sst = 0.0;
    for (s = 0; s < n; ++s) ysum += y[s];
    const double y_avg = ysum / (double)n;
    for (s = 0; s < n; ++s) ssr += (y[s] - yp[s]) * (y[s] - yp[s]);
    for (s = 0; s < n; ++s) sst += (y[s] - y_avg) * (y[s] - y_avg);
    const double r2 = 1 - ssr / sst;
    const double rmse = sqrt(ssr / (double)n);
    if (ssr_out) *ssr_out = ssr;
    if (r2 != r2 || rmse != rmse) {
        if (f0) *f0 = 1000;
        if (f1) *f1 = 1000;
        if (fsize) *fsize = 1000;
    } else {
        if (f0) *f0 = 1 - r2;
        if (f1) *f1 = rmse;
        if (fsize) *fsize = size;
    }
}

/* Scores a batch like rr_score_batch. Extra outputs (may be NULL): f0/f1/fsize [n_cand] =
 * the fitness tuple of rils_rols_cpp.cpp:520-541. X is feature-major. */
RR_ORACLE_API int rr_oracle_score_batch(const double *X, const double *y, int64_t n, int32_t d,
                                        const rr_batch *b, rr_result *res, double *f0, double *f1,
                                        int32_t *fsize)
{
    const double EPS = pow(10, -12); /* node.h:13-14 */
    int32_t c;
    int64_t s;
    double *yp = (double *)malloc(sizeof(double) * (size_t)n);
    if (!yp) return -2;
    for (c = 0; c < b->n_cand; ++c) {
        const int32_t t0 = b->cand_term_begin[c], t1 = b->cand_term_begin[c + 1];
        uint32_t flags = 0;
        int32_t size = 0;
        if (b->mode == RR_MODE_EVAL_ONLY) {
            if (t1 - t0 != 1) { free(yp); return -1; }
            const int32_t c0 = b->term_code_begin[t0], c1 = b->term_code_begin[t0 + 1];
            if (eval_program(b->code + c0, c1 - c0, b->consts, b->n_consts, X, n, d, yp)) { free(yp); return -1; }
            size = c1 - c0; /* node count == node::size(), node.h:311-322 */
            if (res && res->nonzero_pivots) res->nonzero_pivots[c] = 0;
        } else {
            const int32_t k = t1 - t0 + 1;
            double *A = (double *)malloc(sizeof(double) * (size_t)n * (size_t)k * 2);
            double *coef = (double *)malloc(sizeof(double) * (size_t)k * 5 + sizeof(int) * (size_t)k +
                                            sizeof(double) * (size_t)n);
            if (!A || !coef) { free(A); free(coef); free(yp); return -2; }
            double *Aqr = A + (size_t)n * k; /* QR works on a copy, ColPivHouseholderQR.h:476 */
            double *hcoef = coef + k, *work = hcoef + k, *nu = work + k, *nd = nu + k;
            double *rhs = nd + k;
            int *perm = (int *)(rhs + n);
            int32_t j;
            /* rils_rols_cpp.cpp:474-482: A.col(i) = factors[i]->evaluate_all(X); last = ones */
            for (j = 0; j < k - 1; ++j) {
                const int32_t c0 = b->term_code_begin[t0 + j], c1 = b->term_code_begin[t0 + j + 1];
                if (eval_program(b->code + c0, c1 - c0, b->consts, b->n_consts, X, n, d, A + (size_t)j * n)) {
                    free(A); free(coef); free(yp); return -1;
                }
            }
            for (s = 0; s < n; ++s) A[(size_t)(k - 1) * n + s] = 1.0;
            memcpy(Aqr, A, sizeof(double) * (size_t)n * k);
            memcpy(rhs, y, sizeof(double) * (size_t)n);
            /* :484 */
            const int nzp = rr_colpiv_qr_factor(Aqr, n, k, n, hcoef, perm, work, nu, nd);
            rr_colpiv_qr_solve(Aqr, n, k, n, hcoef, perm, nzp, rhs, coef);
            if (nzp < (k < n ? k : (int32_t)n)) flags |= RR_RES_RANKDEF;
            /* :488-517 rebuild + evaluate the rebuilt tree in its association order */
            int first = 1;
            for (j = 0; j < k; ++j) {
                const double cf = coef[j];
                if (fabs(cf) < EPS) continue; /* value_zero, :492 */
                const double *col = A + (size_t)j * n;
                if (j == k - 1) { /* free term: node(coef * 1.0), :497-498 */
                    const double cv = cf * 1.0;
                    if (first) for (s = 0; s < n; ++s) yp[s] = cv;
                    else for (s = 0; s < n; ++s) yp[s] = yp[s] + cv;
                    size += 1;
                } else {
                    const int32_t tsize = b->term_code_begin[t0 + j + 1] - b->term_code_begin[t0 + j];
                    if (fabs(cf - 1) < EPS) { /* value_one, :500-501 */
                        if (first) for (s = 0; s < n; ++s) yp[s] = col[s];
                        else for (s = 0; s < n; ++s) yp[s] = yp[s] + col[s];
                        size += tsize;
                    } else { /* :503-504 MULTIPLY(const, term) */
                        if (first) for (s = 0; s < n; ++s) yp[s] = cf * col[s];
                        else for (s = 0; s < n; ++s) yp[s] = yp[s] + cf * col[s];
                        size += tsize + 2;
                    }
                }
                if (!first) size += 1; /* PLUS node, :510 */
                first = 0;
            }
            if (first) { /* :514-515 */
                for (s = 0; s < n; ++s) yp[s] = 0.0;
                size = 1;
            }
            if (res && res->coef) memcpy(res->coef + t0 + c, coef, sizeof(double) * (size_t)k);
            if (res && res->nonzero_pivots) res->nonzero_pivots[c] = nzp;
            free(A);
            free(coef);
        }
        double ssr;
        fitness_from_yhat(y, yp, n, size, &ssr, f0 ? f0 + c : NULL, f1 ? f1 + c : NULL, fsize ? fsize + c : NULL);
        if (!(ssr - ssr == 0.0)) flags |= RR_RES_NONFINITE;
        if (res && res->ssr) res->ssr[c] = ssr;
        if (res && res->flags) res->flags[c] = flags;
    }
    free(yp);
    return 0;
}

/* rils_rols_cpp.cpp:51-86 for EVAL_ONLY programs */
RR_ORACLE_API int rr_oracle_classifier_metrics(const double *X, const double *y, int64_t n, int32_t d,
                                               const rr_batch *b, double *accuracy, double *log_loss,
                                               double *abs_loss)
{
    int32_t c;
    int64_t i;
    double *yp = (double *)malloc(sizeof(double) * (size_t)n);
    if (!yp) return -2;
    for (c = 0; c < b->n_cand; ++c) {
        const int32_t t0 = b->cand_term_begin[c];
        const int32_t c0 = b->term_code_begin[t0], c1 = b->term_code_begin[t0 + 1];
        if (eval_program(b->code + c0, c1 - c0, b->consts, b->n_consts, X, n, d, yp)) { free(yp); return -1; }
        double acc = 0, ll = 0, al = 0;
        for (i = 0; i < n; ++i) {
            const double ypib = yp[i] >= 0.5 ? 1.0 : 0.0;
            const double yib = y[i] >= 0.5 ? 1.0 : 0.0;
            if (ypib == yib) acc += 1.0;
            const double prob = 1.0 / (1.0 + exp(-2.0 * (yp[i] - 0.5)));
            const double lli = (1.0 - yib) * log(1.0 - prob) + yib * log(prob);
            ll -= lli;
            al += fabs(yib - yp[i]);
        }
        if (accuracy) accuracy[c] = acc / (double)n;
        if (log_loss) log_loss[c] = ll / (double)n;
        if (abs_loss) abs_loss[c] = al / (double)n;
    }
    free(yp);
    return 0;
}
