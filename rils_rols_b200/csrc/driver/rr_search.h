// rr_search.h — the iterated-local-search driver on top of the C ABI.
//
// Same control flow, budget accounting and random streams as class rils_rols
// (/root/reference/rils_rols_cpp/rils_rols_cpp.cpp:88-881); what changed is subsystem (b) of the
// north star: the two candidate loops submit a WHOLE neighbourhood to the engine
// (rr_score_batch) and then replay the reference's sequential decisions on the returned numbers
// (SURVEY.md 3.4):
//   local search      :608-640  -> one OLS_FIT batch per neighbourhood, first-improvement replay
//   perturbation pass :815-831  -> one EVAL_ONLY batch, then the :831 sort
// Single fitness() calls (:800, :606-607, :631, :842) are one-candidate batches.
#pragma once

#include <chrono>
#include <cstdio>
#include <string>
#include <tuple>
#include <unordered_set>
#include <vector>

#include "../../../include/rr_b200.h"
#include "rr_expr.h"

namespace rrd {

using Fitness = std::tuple<double, double, int>;  // (1-R2, RMSE, size), rils_rols_cpp.cpp:520-541

struct SearchParams {
    bool classification = false;
    int max_fit_calls = 100000;
    int max_seconds = 100;
    double complexity_penalty = 0.001;
    int max_complexity = 50;
    double sample_size = 1.0;
    bool verbose = false;
    int random_state = 0;
};

// one scored neighbourhood, kept when tracing is on (RR_B200_TRACE=1 or set_trace(true)): the
// submitted rr_batch arrays, what the engine returned and what the replay decided
struct TraceBatch {
    int mode = 0;
    std::vector<int32_t> cand_term_begin, term_code_begin;
    std::vector<uint32_t> code;
    std::vector<double> consts, coef, ssr;
    std::vector<int32_t> size;       // size of the tree that was scored (tuned tree for OLS_FIT)
    std::vector<int32_t> accepted;   // indices accepted by the local-search replay
    std::vector<double> accepted_fit;  // (f0, f1, size) of each accepted candidate after its :631 re-score
    std::vector<int32_t> consumed;   // 1 if the candidate consumed a fit call in the replay
    double curr_f0 = 0, curr_f1 = 0; // current fitness when the neighbourhood was generated
    int curr_size = 0;
    int fit_calls_before = 0;
};

class Search {
public:
    explicit Search(const SearchParams &p);
    ~Search();
    Search(const Search &) = delete;
    Search &operator=(const Search &) = delete;

    // X row-major n x d (the numpy layout of the pybind boundary), y n values
    void fit(const double *X_rowmajor, const double *y, int64_t n, int32_t d);
    void predict(const double *X_rowmajor, int64_t n, int32_t d, double *out) const;
    void predict_proba(const double *X_rowmajor, int64_t n, int32_t d, double *out) const;  // n x 2: (1 - p, p)
    std::string model_string() const;
    double best_time() const { return best_time_; }
    double total_time() const { return total_time_; }
    int fit_calls() const { return fit_calls_; }
    const Expr *model() const { return final_.get(); }

    // tooling / tests (host only)
    void setup_nodes_for(int32_t d);  // allowed node set for d features without data
    std::vector<Expr> all_candidates(const Expr &solution, bool local_search) const;
    void set_trace(bool on) { trace_ = on; }
    // classification only: score with (1 - accuracy, log-loss, size) instead of (1 - R2, RMSE, size); default off = reference
    void set_classifier_objective(bool on) { classifier_objective_ = on; }
    bool classifier_objective() const { return p_.classification && classifier_objective_; }
    const std::vector<TraceBatch> &trace() const { return trace_log_; }
    rr_stats engine_stats() const { return stats_; }

private:
    struct BatchBuilder;
    SearchParams p_;
    rr_engine *eng_ = nullptr;
    rr_stats stats_{};
    int64_t n_ = 0;
    int32_t d_ = 0;
    double sst_ = 0.0;
    // search state, rils_rols_cpp.cpp:97-104
    int main_it_ = 0, fit_calls_ = 0, ls_calls_ = 0, skipped_perts_ = 0, total_perts_ = 0;
    std::unordered_set<std::string> checked_perts_;
    std::chrono::time_point<std::chrono::high_resolution_clock> start_;
    ExprP final_;
    Fitness final_fit_{0, 0, 0};
    double best_time_ = 0.0, total_time_ = 0.0;
    std::vector<Expr> allowed_;
    bool trace_ = false;
    bool classifier_objective_ = false;
    std::vector<TraceBatch> trace_log_;

    void reset();
    bool finished() const;
    bool check_skip(const std::string &s);
    void setup_nodes(const std::vector<int> &rel_feat);
    std::vector<Expr> change_candidates(const Expr &old_node) const;
    std::vector<Expr> perturb_candidates(const Expr &old_node) const;

    Fitness fitness_from(double ssr, int size) const;
    Fitness classifier_fitness_from(double accuracy, double log_loss, int size) const;
    std::vector<Fitness> score_trees(const std::vector<const Expr *> &trees, std::vector<double> *ssr_out);
    double fitness_value(const Fitness &f) const;
    int compare_fitness(const Fitness &a, const Fitness &b) const;
    Fitness score_single(const Expr &tree);                       // fitness(), one EVAL_ONLY candidate
    ExprP tune_single(const Expr &tree, Fitness *fit);            // tune_constants() + fitness()
    ExprP local_search(const Expr &start);
    void print_state(const Fitness &curr) const;
    void engine_check(int rc, const char *what) const;
};

// tune_constants()'s tree rebuild from coefficients, rils_rols_cpp.cpp:488-517
ExprP rebuild_from_coefficients(const std::vector<const Expr *> &factors, const double *coef);

}  // namespace rrd
