/*
 * rr_b200.h — C ABI of the B200 scoring engine for the RILS-ROLS hot path.
 *
 * The reference has no FFI for this path: the replaced interface is three C++
 * member functions of the ILS driver (paths relative to /root/reference):
 *
 *   shared_ptr<node> rils_rols::tune_constants(shared_ptr<node>, X, y)   rils_rols_cpp/rils_rols_cpp.cpp:445-518
 *   tuple<double,double,int> rils_rols::fitness(shared_ptr<node>, X, y)  rils_rols_cpp/rils_rols_cpp.cpp:520-541
 *   Eigen::ArrayXd node::evaluate_all(const vector<ArrayXd>& X)          rils_rols_cpp/node.h:305, node.cpp:5-95
 *
 * called once per candidate from the local-search loop (rils_rols_cpp.cpp:611-616)
 * and the perturbation loop (rils_rols_cpp.cpp:819-829).  This ABI scores a whole
 * neighbourhood per call instead.  Plain C: pointers and sizes only, caller-owned
 * host buffers, int return codes, no exceptions and no exit() across the boundary.
 * One engine per rils_rols object; an engine is not thread-safe.
 *
 * Trees cross the boundary as postfix bytecode (left operand first), one 32-bit
 * word per node: low 8 bits = opcode (the node_type enumerator value,
 * rils_rols_cpp/node.h:16-38), high 24 bits = feature index (VAR) or index into
 * the batch constant pool (CONST).
 */
#ifndef RR_B200_H
#define RR_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_ABI_VERSION 3

#if defined(__GNUC__)
#define RR_API __attribute__((visibility("default")))
#else
#define RR_API
#endif

/* opcodes == enum class node_type, rils_rols_cpp/node.h:16-38 */
enum rr_opcode {
    RR_OP_NONE = 0,
    RR_OP_CONST = 1,
    RR_OP_VAR = 2,
    RR_OP_PLUS = 3,
    RR_OP_MINUS = 4,
    RR_OP_MULTIPLY = 5,
    RR_OP_DIVIDE = 6,
    RR_OP_SIN = 7,
    RR_OP_COS = 8,
    RR_OP_LN = 9,
    RR_OP_EXP = 10,
    RR_OP_SQRT = 11,
    RR_OP_SQR = 12,
    RR_OP_POW = 13,
    RR_OP_LESS_THAN = 14,
    RR_OP_GREATER_THAN = 15,
    RR_OP_EQUAL = 16,
    RR_OP_NOT_EQUAL = 17,
    RR_OP_MIN = 18,
    RR_OP_MAX = 19,
    RR_OP_COUNT = 20
};

#define RR_INS(op, arg) ((uint32_t)(op) | ((uint32_t)(arg) << 8))
#define RR_INS_OP(w) ((uint32_t)(w) & 0xffu)
#define RR_INS_ARG(w) ((uint32_t)(w) >> 8)

/* return codes */
enum rr_status {
    RR_OK = 0,
    RR_ERR_INVALID = 1,   /* bad argument / malformed batch */
    RR_ERR_CUDA = 2,      /* CUDA runtime failure (see rr_last_error) */
    RR_ERR_NO_DEVICE = 3, /* no usable sm_100 device: there is no CPU fallback */
    RR_ERR_NOMEM = 4,
    RR_ERR_COLLECTIVE = 5 /* the all-reduce hook reported failure */
};

/* batch modes */
enum rr_mode {
    /* one program per candidate, scored as-is: fitness() without tune_constants()
     * (rils_rols_cpp.cpp:800, :828, :842, :631). cand_term_begin must give exactly
     * one term per candidate. */
    RR_MODE_EVAL_ONLY = 0,
    /* candidate = list of additive terms (the `factors[]` of rils_rols_cpp.cpp:452-473,
     * in extraction order, WITHOUT the free term: the engine appends the ones column
     * like :474-475); fit coefficients by least squares (:477-484), snap them
     * (:492-505), score the rebuilt model (:520-541). */
    RR_MODE_OLS_FIT = 1
};

/* engine creation flags */
enum rr_engine_flags {
    RR_FLAG_DEFAULT = 0,
    /* OLS path selection (default: exact Householder when n <= RR_B200_EXACT_MAX_N
     * (env, default 4096), Gram otherwise) */
    RR_FLAG_FORCE_GRAM = 1u << 0,  /* always Gram + Cholesky (+ refinement / dd escalation) */
    RR_FLAG_FORCE_EXACT = 1u << 1, /* always materialise A and run column-pivoted Householder QR */
    RR_FLAG_NO_CSE = 1u << 2,      /* plan every candidate independently (no cross-candidate sharing) */
    RR_FLAG_X_DEVICE = 1u << 3,    /* X / y pointers passed to rr_engine_create are device pointers */
    RR_FLAG_X_ROWMAJOR = 1u << 4   /* rr_engine_create_sharded: X is row-major n x d (the numpy layout of the pybind boundary) */
};

/* per-candidate result flags */
enum rr_result_flags {
    RR_RES_NONFINITE = 1u << 0, /* a term (or the model) evaluated to NaN/inf somewhere: ssr is NaN/inf */
    RR_RES_RANKDEF = 1u << 1,   /* nonzero_pivots < k: trailing pivot columns got coefficient 0 */
    RR_RES_REFINED = 1u << 2,   /* Gram path: an explicit-residual refinement pass was used */
    RR_RES_EXACT = 1u << 3,     /* exact column-pivoted Householder QR on the materialised A was used */
    RR_RES_DD = 1u << 4,        /* Gram path: double-double Gram + Cholesky escalation was used */
    RR_RES_SLOWPATH = 1u << 5   /* candidate exceeded fast-path limits (k, program length, stack) */
};

typedef struct rr_engine rr_engine;

typedef struct rr_batch {
    int32_t mode;                   /* enum rr_mode */
    int32_t n_cand;
    const int32_t *cand_term_begin; /* [n_cand+1] candidate -> first term; terms of c = [b[c], b[c+1]) */
    const int32_t *term_code_begin; /* [n_terms+1] term -> first code word */
    const uint32_t *code;           /* postfix words, RR_INS(op,arg) */
    const double *consts;           /* constant pool */
    int32_t n_consts;
    int32_t n_code;                 /* words in code[] (ABI 2; 0 = unknown: offsets are then only checked for monotony) */
} rr_batch;

/*
 * Caller-allocated result arrays (any pointer except ssr may be NULL).
 *   coef: OLS_FIT only. Candidate c has k_c = (#terms of c) + 1 coefficients, stored at
 *         coef[cand_term_begin[c] + c ... + k_c), in factor order, free term last — the
 *         vector `coefs` of rils_rols_cpp.cpp:484 (raw, before the :492-505 snapping;
 *         non-pivot columns are exactly 0.0 like ColPivHouseholderQR.h:606).
 *   nonzero_pivots: OLS_FIT only, ColPivHouseholderQR::nonzeroPivots() (ColPivHouseholderQR.h:526-527).
 *   ssr:  sum_i (y_i - yhat_i)^2 of the model fitness() would score: for OLS_FIT the
 *         rebuilt tree with snapped coefficients in the reference association order
 *         ((c0*t0 + c1*t1) + ...) + c_free. NaN/inf propagate like rils_rols_cpp.cpp:40-49.
 *         Host derives 1-R2 = ssr/sst, RMSE = sqrt(ssr/n) and the NaN sentinel (:529-530,:536-537).
 *   flags: enum rr_result_flags.
 */
typedef struct rr_result {
    double *coef;
    int32_t *nonzero_pivots;
    double *ssr;
    uint32_t *flags;
} rr_result;

typedef struct rr_engine_info {
    int64_t n;      /* samples held by THIS engine (its shard; all shards of a single-process multi-GPU engine) */
    int64_t n_total;/* samples over all ranks (== n without an all-reduce hook) */
    int32_t d;
    int32_t device; /* CUDA device ordinal */
    double y_mean;  /* mean of y over n_total, rils_rols_cpp.cpp:41 */
    double sst;     /* sum (y - y_mean)^2 over n_total, rils_rols_cpp.cpp:43 */
    int32_t sm_count;
    int32_t exact_max_n;
    int32_t n_gpus; /* devices this engine object drives itself (rr_engine_create_sharded), else 1 */
    int32_t world;  /* ranks the rows are sharded over (n_gpus, or the size of the installed communicator / hook) */
} rr_engine_info;

typedef struct rr_stats {
    uint64_t batches;
    uint64_t candidates;
    uint64_t sweep_launches;   /* interpreter kernel launches */
    uint64_t kernel_launches;  /* all engine kernel launches */
    uint64_t refined;          /* candidates that took the refinement pass */
    uint64_t exact;            /* candidates that took the exact QR */
    uint64_t dd;               /* candidates that took the double-double escalation */
    uint64_t nonfinite;
    uint64_t distinct_terms;   /* after cross-candidate CSE, summed over batches */
    uint64_t term_instances;   /* before CSE */
    uint64_t distinct_dots;    /* distinct Gram / A^T y / residual reductions */
    uint64_t dot_instances;
    double last_sweep_ms;      /* device time of the interpreter launches of the last batch */
    double last_batch_ms;      /* device time of the whole last batch */
    double w_contract;         /* last batch: SURVEY 8(d) no-sharing FP64 thread-instructions per sample */
    double w_shared;           /* last batch: same count for the work actually issued (after CSE, all passes) */
    uint64_t h2d_bytes;        /* last batch */
    uint64_t d2h_bytes;        /* last batch */
    double last_host_ms;       /* last batch: host wall time of rr_score_batch (analyse + planning + waits) */
    double ingest_ms;          /* engine creation: upload + device-side transpose / gather + target statistics */
    uint64_t collectives;      /* NCCL collectives issued by the engine itself (0 with the callback hook) */
    uint64_t row_groups;       /* ABI 3: groups reduced by the row machine (R8 plans), summed over batches */
    uint64_t row_group_rows;   /* ABI 3: rows in those groups */
} rr_stats;

/*
 * All-reduce hook for sample-sharded multi-GPU runs (one process per GPU): called
 * with a DEVICE buffer of `count` doubles that must be summed element-wise over all
 * ranks in place, ordered on `cuda_stream` (a cudaStream_t). Return 0 on success.
 * The Python host installs a torch.distributed (NCCL) implementation.
 */
typedef int (*rr_allreduce_fn)(void *dev_buf, size_t count, void *cuda_stream, void *user);

/* Replaces the data hand-off of fit(): X is FEATURE-major (d columns of n contiguous
 * doubles, the vector<ArrayXd> layout of rils_rols_cpp.cpp:675-698), y has n doubles.
 * device < 0 selects the current device. Data is copied; nothing is retained. */
RR_API int rr_engine_create(const double *X_feature_major, const double *y, int64_t n, int32_t d,
                     int32_t device, uint32_t flags, rr_engine **out);
/* Same, from the row-major numpy layout the pybind boundary receives
 * (rils_rols_cpp.cpp:690-696): the transpose runs on the device. */
RR_API int rr_engine_create_rowmajor(const double *X_row_major, const double *y, int64_t n, int32_t d,
                              int32_t device, uint32_t flags, rr_engine **out);
/*
 * ONE engine object over n_gpus devices of this process (devices 0 .. n_gpus-1; n_gpus <= 0 = all
 * visible). Rows are sharded in contiguous blocks, every device sweeps its block with the same plan, and the
 * per-candidate partial reductions are summed by NCCL (ncclCommInitAll, one ncclReduce per sweep on the engines'
 * streams, no host synchronisation in between); the solves run on device 0. X is feature-major unless
 * RR_FLAG_X_ROWMAJOR is set. row_index (may be NULL) selects and orders the rows: row i of the engine is row
 * row_index[i] of X / y, n_rows of them - the shuffled sub-sample of rils_rols_cpp.cpp:774-795, gathered on the
 * device(s) instead of on the host (n_rows = n and row_index = NULL: all rows in order). With n_gpus == 1 this is
 * rr_engine_create plus the device-side gather. Always takes the Gram path.
 */
RR_API int rr_engine_create_sharded(const double *X, const double *y, int64_t n, int32_t d, const int32_t *row_index,
                                    int64_t n_rows, int32_t n_gpus, uint32_t flags, rr_engine **out);
RR_API void rr_engine_destroy(rr_engine *e);

/* One process per GPU (torchrun): NCCL inside the engine instead of the callback below. Rank 0 obtains a unique id
 * (128 bytes, ncclGetUniqueId), the caller broadcasts it by whatever means it has, and every rank installs it:
 * the engine then all-reduces its partial reductions itself, on its own stream, with no host involvement.
 * Collective: all ranks must call rr_engine_comm_init together. */
RR_API int rr_comm_unique_id(void *id128);
RR_API int rr_engine_comm_init(rr_engine *e, const void *id128, int32_t rank, int32_t world);

/* Install the all-reduce hook (this engine holds shard `rank` of `world`) and recompute
 * y_mean / sst over all ranks. fn == NULL removes it. */
RR_API int rr_engine_set_allreduce(rr_engine *e, rr_allreduce_fn fn, void *user, int32_t rank, int32_t world);

RR_API int rr_engine_get_info(const rr_engine *e, rr_engine_info *info);
RR_API int rr_get_stats(const rr_engine *e, rr_stats *stats);

/* Score one neighbourhood. Synchronous: results are valid on return. */
RR_API int rr_score_batch(rr_engine *e, const rr_batch *batch, rr_result *result);

/* Classifier metrics of rils_rols_cpp.cpp:51-86 for EVAL_ONLY programs (not used by
 * the search, :527): out arrays [n_cand], any may be NULL. */
RR_API int rr_classifier_metrics(rr_engine *e, const rr_batch *batch, double *accuracy, double *log_loss,
                          double *abs_loss);

/* predict(): evaluate one program over a caller-supplied FEATURE-major matrix
 * (rils_rols_cpp.cpp:730-750 without the 0.5 threshold). out has n doubles. */
RR_API int rr_predict(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts,
               int32_t n_consts, const double *X_feature_major, int64_t n, int32_t d, double *out);
RR_API int rr_predict_rowmajor(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts,
                        int32_t n_consts, const double *X_row_major, int64_t n, int32_t d,
                        double *out);
/* predict_proba() of the classifier through the same interpreter: out[2 i] = 1 - p_i, out[2 i + 1] = p_i with
 * p = 1 / (1 + exp(-2 (yhat - 0.5))), the logistic of average_log_loss (rils_rols_cpp.cpp:69). (The reference's
 * Python predict_proba applies utils.logistic to the already thresholded prediction, rils_rols.py:177-179; that
 * wrapper keeps working on predict().) out has 2 n doubles. */
RR_API int rr_predict_proba_rowmajor(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts,
                                     int32_t n_consts, const double *X_row_major, int64_t n, int32_t d, double *out);

/* relevant_features(), rils_rols_cpp.cpp:753-770: r2[j] = R2(X[j], y) with the reference's (truth, prediction) =
 * (feature, target) argument order, i.e. 1 - sum (x_j - y)^2 / sum (x_j - mean x_j)^2, for every feature column of
 * the engine (one device reduction; the host keeps the 200 best when d > 200). r2 has d doubles. */
RR_API int rr_feature_r2(rr_engine *e, double *r2);

/* Copies rows [row0, row0 + rows) of the engine's resident matrix back: Xout feature-major rows x d (column stride
 * `rows`), yout rows values; either may be NULL. Test hook for the device-side ingest (same row order and bits as the
 * reference's host loops). */
RR_API int rr_engine_read_rows(rr_engine *e, int64_t row0, int64_t rows, double *Xout_feature_major, double *yout);

/* Measured FP64-pipe peak (thread-instructions / s) of the engine's device from a
 * DFMA-only microkernel: the roofline denominator of SURVEY 8(d). */
RR_API int rr_measure_fp64_peak(rr_engine *e, double *dfma_per_second);

/* Host-only tooling (no device work): runs the planner on a batch and returns the engine's internal
 * instruction stream (rils_rols_b200/csrc/rr_isa.h) so that tests can execute it on the CPU and
 * check the compiler half of the engine without a GPU.
 *   kind: 0 Gram, 1 Gram in double-double, 2 EVAL_ONLY, 3 EVAL_ONLY + classifier metrics,
 *         4 materialise distinct terms, 5 residual (coef_snapped = batch coefficient layout).
 * All arrays are malloc'ed by the call and released by rr_debug_plan_free. */
typedef struct rr_debug_plan {
    int64_t n_ins, n_chunks, n_cols, n_tab, n_tab_begin, n_term_ids;
    void *ins;           /* RRIns[n_ins], 16 bytes each: u32 w0, u32 w1, f64 imm */
    void *chunks;        /* RRChunk[n_chunks], 8 x i32: pc_begin, n_ins, dot_base, n_dots, col_begin, n_cols, 0, 0 */
    int32_t *cols;       /* staged global column ids (features 0..d-1, d = y, d+1 = y - mean) */
    int32_t *tab;        /* per-candidate dot-id table (layout depends on kind) */
    int32_t *tab_begin;  /* [n_cand + 1] (kinds 0, 1, 5) */
    int32_t *term_ids;   /* per term instance -> distinct term */
    int32_t n_dots, max_tile_cols, n_terms_distinct, reserved;
    double w_issued, w_contract;
    char error[256];
} rr_debug_plan;
RR_API int rr_debug_plan_batch(const rr_batch *batch, int32_t d, int32_t kind, int32_t tile_cols,
                               int32_t max_slots, int32_t target_chunks, int32_t no_cse, int32_t n_pins,
                               const double *coef_snapped, rr_debug_plan *out);
RR_API void rr_debug_plan_free(rr_debug_plan *p);
/* Host-only test hook: the two halves of a batch planned on two threads at once must equal the halves planned
 * one after the other (rr_score_batch plans the second half of a large neighbourhood on a helper thread).
 * 0 = identical, 1 = different, RR_ERR_INVALID = malformed batch. */
RR_API int rr_debug_plan_concurrency_check(const rr_batch *batch, int32_t d, int32_t tile_cols);
/* Host-only test hook: per candidate, bit i set = term i is constant by construction (no variable in it, t / t,
 * t - t, (c t) / t, 0 * t, ...). The Gram-path solver keeps such a column or the free term - the longer of the two,
 * the reference's pivot rule - instead of escalating a singular Gram matrix. out: n_cand words. */
RR_API int rr_debug_const_terms(const rr_batch *batch, int32_t d, uint32_t *out);

/* Last error text: of the engine, or of the calling thread when e == NULL. */
RR_API const char *rr_last_error(const rr_engine *e);
RR_API int rr_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RR_B200_H */
