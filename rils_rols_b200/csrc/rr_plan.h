// rr_plan.h — host-side planner: rr_batch (postfix trees) -> sweep programs.
//
// This is the "tree -> compact bytecode batch" compiler of the north star plus the
// cross-candidate sharing SURVEY.md App. B.9 measures (only ~10 % of the term instances
// of a local-search neighbourhood are distinct): terms are hashed, evaluated once per
// sample and kept in tile slots; Gram / A^T y entries are keyed by the (term, term) pair
// and reduced once.
#ifndef RR_PLAN_H
#define RR_PLAN_H

#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rr_b200.h"
#include "rr_isa.h"

namespace rr {

// global column ids of the engine's sample-major matrix: features 0..d-1, then
struct ColIds {
    int y;   // d     : y
    int yc;  // d + 1 : y - mean(y)
};

struct TermNode {
    uint8_t op;
    int32_t left = -1, right = -1;  // node indices
    int32_t var = -1;
    double cval = 0.0;
    int32_t need = 0;  // Sethi-Ullman spill need
    int32_t first = 0;     // postfix index of the first node of this subtree
    int32_t sub_term = -1; // distinct term whose whole program equals this subtree (subtree-level sharing)
    int32_t sub_id = -1;   // id of this subtree among the expensive inner subtrees that several distinct terms share

    bool leaf() const { return op == RR_OP_CONST || op == RR_OP_VAR; }
};

struct Term {
    int32_t code_begin = 0, code_len = 0;  // first instance in the batch code
    std::vector<TermNode> nodes;           // postfix order, root = back()
    double w = 0.0;                        // SURVEY 8(d) contract weight of one evaluation
    bool is_const_one = false;
    // the term has the same value at every sample BY CONSTRUCTION: no variable in it, or every variable sits under
    // S / S, S - S or (c S) / S of identical subtrees (a local-search neighbourhood holds a few dozen of these: sin(c),
    // t / t, (c t) / t, ...; the last kind is constant up to the rounding of its products). Its column is a multiple
    // of the free term's; the solver is told, so that it keeps one of the two by the reference's pivot rule (what
    // column-pivoted QR does in exact arithmetic, and what the double-double escalation arrives at a sweep later)
    // instead of finding a singular Gram matrix and escalating.
    bool exact_const = false;
};

struct SweepPlan {
    std::vector<RRIns> ins;
    std::vector<RRChunk> chunks;
    std::vector<int32_t> cols;  // staged global column ids, per chunk [col_begin, col_begin+n_cols)
    int32_t n_dots = 0;
    int32_t max_tile_cols = 0;  // max over chunks of staged columns + slots used
    int32_t n_stg_cols = 0;     // columns written by RI_STG
    double w_issued = 0.0;      // contract-weighted FP64 work actually issued, per sample
    uint64_t n_term_evals = 0;
    uint64_t n_dot_ins = 0;
    uint64_t n_gram_groups = 0, n_gram_rows = 0;  // G8 plans: RI_GRAM8 instructions and the rows they reduce
    bool r8 = false;  // an R8 plan (rr_isa.h RQ_*): runs in rr_sweep_r8_kernel only
    uint64_t n_stored_evals = 0;  // R8 plans: evaluations of stored sub-expressions
    bool empty() const { return chunks.empty(); }
};

struct PlanLimits {
    int32_t tile_cols = 56;  // columns that fit the shared-memory tile for the chosen tile height
    int32_t max_slots = 1 << 20;  // cap on value slots (cached terms + temporaries) per chunk: bounds the tile
    int32_t target_chunks = 1;
    int32_t n_pins = RR_NPIN;     // pins the Gram plans may use (rr_isa.h); 0 = tile slots only
    int32_t n_cache = RR_NREG - RR_NPIN;  // cache registers for shared sub-expressions; 0 = off
    int32_t transient_horizon = 4;  // a new term no candidate lists again within this many units is not stored
    bool no_cse = false;
    bool fuse = true;  // super-instruction peephole (rr_isa.h)
    bool g8 = false;         // G8 plan (rr_isa.h RI_GRAM8): fresh terms in tile slots, reductions by DMMA against the pins
    int32_t ins_window = RR_INS_WINDOW;  // instructions per shared-memory window of the kernel that runs the plan
    bool mdot_rows = false;  // data slot with the ring rows behind every instruction that ends in RI_MDOT (rr_isa.h)
};

// dot-id sentinels used in the per-candidate index tables
enum : int32_t { DOT_NONE = -1 };

class BatchPlanner {
public:
    BatchPlanner(const rr_batch *b, int32_t d);
    // returns "" or an error description (malformed batch)
    std::string analyse(bool no_cse);

    int32_t n_cand() const { return b_->n_cand; }
    int32_t n_terms_distinct() const { return (int32_t)terms_.size(); }
    int32_t n_term_instances() const { return (int32_t)term_id_.size(); }
    int32_t k_of(int32_t c) const { return b_->cand_term_begin[c + 1] - b_->cand_term_begin[c] + 1; }
    int32_t max_k() const { return max_k_; }
    const std::vector<int32_t> &term_ids() const { return term_id_; }  // per term instance -> distinct id
    const Term &term(int32_t u) const { return terms_[u]; }
    double w_contract() const { return w_contract_; }
    // bit i set: term i of candidate c is exact_const (candidates with more than 32 terms report none)
    uint32_t cand_const_mask(int32_t c) const
    {
        const int32_t t0 = b_->cand_term_begin[c], t1 = b_->cand_term_begin[c + 1];
        if (t1 - t0 > 32) return 0u;
        uint32_t m = 0;
        for (int32_t t = t0; t < t1; ++t)
            if (terms_[term_id_[t]].exact_const) m |= 1u << (t - t0);
        return m;
    }
    const std::vector<double> &cand_contract_w() const { return cand_w_; }
    // distinct terms (ascending ids) that contain shared sub-expression `sub`
    const std::vector<int32_t> &sub_occurrences(int32_t sub) const { return sub_occ_[sub]; }
    int32_t sub_size(int32_t sub) const { return sub_size_[sub]; }  // nodes of the sub-expression

    // OLS_FIT, Gram path: Gram + A^T yc + column sums for the candidates in `subset`
    // (nullptr = all). cand_dot: per listed candidate m(m+1)/2 (upper triangle, row-major)
    // + m (with yc) + m (with ones) dot ids; cand_dot_begin has subset size + 1 entries.
    // dd = accumulate in double-double (each id then addresses a (hi,lo) pair: id, id+1).
    std::string plan_gram(const PlanLimits &lim, const ColIds &cols, const std::vector<int32_t> *subset,
                          bool dd, SweepPlan &out, std::vector<int32_t> &cand_dot,
                          std::vector<int32_t> &cand_dot_begin);

    // the same for a G8 plan (candidates of at most RR_NPIN terms)
    std::string plan_gram_g8(const PlanLimits &lim, const ColIds &cols, const std::vector<int32_t> *subset, SweepPlan &out,
                             std::vector<int32_t> &cand_dot, std::vector<int32_t> &cand_dot_begin);

    // the same for an R8 plan (rr_isa.h: the row machine); fails with a message when the neighbourhood does not fit
    // one chunk or a tree needs more tile slots than there are - the caller then plans a G8 piece instead
    std::string plan_gram_r8(const PlanLimits &lim, const ColIds &cols, const std::vector<int32_t> *subset, SweepPlan &out,
                             std::vector<int32_t> &cand_dot, std::vector<int32_t> &cand_dot_begin) const;

    // explicit residual of the model sum_i cs_i t_i + cs_free (snapped coefficients, reference
    // association order) for the listed candidates: per candidate 1 (r.r) + m (r.t_i) + 1 (r.1) ids.
    std::string plan_residual(const PlanLimits &lim, const ColIds &cols, const std::vector<int32_t> &subset,
                              const double *coef_snapped, SweepPlan &out, std::vector<int32_t> &cand_dot,
                              std::vector<int32_t> &cand_dot_begin);

    // EVAL_ONLY: ssr of each program as-is; one dot id per candidate. metrics = also the three
    // classifier reductions (ids cand_dot[c] + 1..3 are then accuracy count, log-loss sum, abs-loss sum).
    std::string plan_eval(const PlanLimits &lim, const ColIds &cols, bool metrics, SweepPlan &out,
                          std::vector<int32_t> &cand_dot);

    // materialise every distinct term as a global column (RI_STG u)
    std::string plan_materialise(const PlanLimits &lim, const ColIds &cols, SweepPlan &out);

private:
    struct Chunk;
    const rr_batch *b_;
    int32_t d_;
    int32_t max_k_ = 0;
    std::vector<Term> terms_;
    std::vector<int32_t> term_id_;
    std::vector<double> cand_w_;
    std::vector<std::vector<int32_t>> sub_occ_;
    std::vector<int32_t> sub_size_;
    double w_contract_ = 0.0;

    std::string build_term(int32_t code_begin, int32_t code_len, Term &t) const;
};

}  // namespace rr
#endif
