// rr_search.cpp — see rr_search.h. Line numbers refer to /root/reference/rils_rols_cpp/rils_rols_cpp.cpp.
#include "rr_search.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <random>
#include <stdexcept>

namespace rrd {

using namespace std::chrono;

namespace {

bool dominates(const Fitness &p, const Fitness &f)  // :581-583
{
    return std::get<0>(p) <= std::get<0>(f) && std::get<1>(p) <= std::get<1>(f) && std::get<2>(p) <= std::get<2>(f);
}
bool is_dominated(const std::vector<Fitness> &pareto, const Fitness &f)  // :585-590
{
    for (const auto &p : pareto)
        if (dominates(p, f)) return true;
    return false;
}
void add_to_pareto(std::vector<Fitness> &pareto, const Fitness &f)  // :592-599
{
    if (is_dominated(pareto, f)) return;
    for (int i = (int)pareto.size() - 1; i >= 0; --i)
        if (dominates(f, pareto[i])) pareto.erase(pareto.begin() + i);
    pareto.push_back(f);
}

// size of the tree tune_constants() would rebuild, without building it (:488-517, node.h:311-322)
int rebuilt_size(const std::vector<const Expr *> &factors, const double *coef)
{
    int size = 0, kept = 0;
    const size_t m = factors.size();
    for (size_t i = 0; i < m; ++i) {
        if (value_zero(coef[i])) continue;
        size += size_of(*factors[i]) + (value_one(coef[i]) ? 0 : 2);
        ++kept;
    }
    if (!value_zero(coef[m])) {
        size += 1;
        ++kept;
    }
    return kept ? size + kept - 1 : 1;
}

}  // namespace

ExprP rebuild_from_coefficients(const std::vector<const Expr *> &factors, const double *coef)
{
    ExprP ols;
    auto append = [&](ExprP f) {
        if (!ols) ols = std::move(f);
        else ols = std::make_unique<Expr>(Op::PLUS, std::move(ols), std::move(f));
    };
    const size_t m = factors.size();
    for (size_t i = 0; i < m; ++i) {
        const double c = coef[i];
        if (value_zero(c)) continue;
        if (value_one(c)) append(clone(*factors[i]));
        else append(std::make_unique<Expr>(Op::MULTIPLY, std::make_unique<Expr>(c), clone(*factors[i])));
    }
    if (!value_zero(coef[m])) append(std::make_unique<Expr>(coef[m] * 1.0));  // free term: node(coef * 1.0), :497-498
    if (!ols) ols = std::make_unique<Expr>(0.0);
    return ols;
}

// ---------------------------------------------------------------------------------------------
// batch assembly
// ---------------------------------------------------------------------------------------------
struct Search::BatchBuilder {
    int mode;
    std::vector<int32_t> cand_term_begin{0}, term_code_begin{0};
    std::vector<uint32_t> code;
    std::vector<double> consts;
    std::vector<ExprP> work;                           // expanded + simplified copies (OLS_FIT)
    std::vector<std::vector<const Expr *>> factors;    // pointers into work[i]
    std::vector<int32_t> eval_size;

    explicit BatchBuilder(int m) : mode(m) {}
    size_t n_cand() const { return cand_term_begin.size() - 1; }

    void add_program(const Expr &e)
    {
        compile_postfix(e, code, consts);
        term_code_begin.push_back((int32_t)code.size());
    }
    void add_ols(const Expr &cand)
    {
        ExprP w = clone(cand);
        expand(*w);    // :448
        simplify(*w);  // :449
        std::vector<const Expr *> f = select_factors(*w);
        for (const Expr *t : f) add_program(*t);
        cand_term_begin.push_back((int32_t)term_code_begin.size() - 1);
        factors.push_back(std::move(f));
        work.push_back(std::move(w));
    }
    void add_eval(const Expr &tree)
    {
        add_program(tree);
        cand_term_begin.push_back((int32_t)term_code_begin.size() - 1);
        eval_size.push_back(size_of(tree));
    }
    rr_batch view() const
    {
        rr_batch b;
        b.mode = mode;
        b.n_cand = (int32_t)n_cand();
        b.cand_term_begin = cand_term_begin.data();
        b.term_code_begin = term_code_begin.data();
        b.code = code.data();
        b.consts = consts.data();
        b.n_consts = (int32_t)consts.size();
        b.n_code = (int32_t)code.size();
        return b;
    }
};

// ---------------------------------------------------------------------------------------------
Search::Search(const SearchParams &p) : p_(p)
{
    const char *t = std::getenv("RR_B200_TRACE");
    trace_ = t && *t && *t != '0';
    const char *co = std::getenv("RR_B200_CLASSIFIER_OBJECTIVE");
    classifier_objective_ = co && *co && *co != '0';
    reset();
}

Search::~Search()
{
    if (eng_) rr_engine_destroy(eng_);
}

void Search::reset()  // :106-117
{
    main_it_ = 0;
    fit_calls_ = 0;
    ls_calls_ = 0;
    start_ = high_resolution_clock::now();
    checked_perts_.clear();
    skipped_perts_ = 0;
    total_perts_ = 0;
    srand(p_.random_state);
}

bool Search::finished() const  // :659-661
{
    return fit_calls_ >= p_.max_fit_calls ||
           duration_cast<seconds>(high_resolution_clock::now() - start_).count() > p_.max_seconds;
}

bool Search::check_skip(const std::string &s)  // :663-672
{
    total_perts_++;
    if (checked_perts_.find(s) != checked_perts_.end()) {
        skipped_perts_++;
        return true;
    }
    checked_perts_.insert(s);
    return false;
}

void Search::engine_check(int rc, const char *what) const
{
    if (rc != RR_OK) throw std::runtime_error(std::string(what) + " failed: " + rr_last_error(eng_));
}

void Search::setup_nodes(const std::vector<int> &rel_feat)  // :121-148
{
    allowed_.clear();
    for (Op t : {Op::PLUS, Op::MINUS, Op::MULTIPLY, Op::DIVIDE, Op::SIN, Op::COS, Op::LN, Op::EXP, Op::SQRT, Op::SQR}) {
        Expr e;
        e.type = t;
        allowed_.push_back(std::move(e));
    }
    const double constants[] = {-1., 0., 0.5, 1., 2., 3.14159265358979323846, 10.};
    for (double c : constants) allowed_.emplace_back(c);
    for (int j : rel_feat) allowed_.push_back(Expr::variable(j));
    if (p_.classification) {
        for (Op t : {Op::LESS_THAN, Op::GREATER_THAN, Op::EQUAL, Op::NOT_EQUAL, Op::MIN, Op::MAX}) {
            Expr e;
            e.type = t;
            allowed_.push_back(std::move(e));
        }
    }
    if (p_.verbose) std::cout << "Finished creating allowed nodes" << std::endl;
}

void Search::setup_nodes_for(int32_t d)
{
    std::vector<int> rel(d);
    std::iota(rel.begin(), rel.end(), 0);
    setup_nodes(rel);
}

// ---- candidate generators (:163-346) -----------------------------------------------------------
namespace {

void gen_const_finetune(const Expr &old, std::vector<Expr> &out)  // :163-179
{
    if (!old.is(Op::CONST)) return;
    if (old.value == 0.0) {
        out.emplace_back(-1.0);
        out.emplace_back(1.0);
    } else {
        const double mult[] = {-1., 0.01, 0.1, 0.2, 0.5, 0.8, 0.9, 0., 1, 1.1, 1.2, 2., 3.14159265358979323846, 5., 10., 20., 50., 100.};
        for (double m : mult) out.emplace_back(old.value * m);
    }
}

void gen_to_subtree(const Expr &old, std::vector<Expr> &out)  // :200-209
{
    std::vector<const Expr *> sub;
    if (old.arity() >= 1) all_subtrees(*old.left, sub);
    if (old.arity() >= 2) all_subtrees(*old.right, sub);
    for (const Expr *s : sub) out.push_back(*s);
}

void gen_to_var_or_1(const Expr &old, const std::vector<Expr> &allowed, std::vector<Expr> &out)  // :222-232
{
    for (const Expr &n : allowed) {
        if (!n.is(Op::VAR)) continue;
        if (old.is(Op::VAR) && old.var == n.var) continue;
        out.push_back(n);
    }
    out.emplace_back(1.0);
}

void gen_unary_applied(const Expr &old, const std::vector<Expr> &allowed, std::vector<Expr> &out)  // :245-252
{
    for (const Expr &u : allowed) {
        if (u.arity() != 1 || !allowed_left(u.type, old)) continue;
        out.emplace_back(u.type, clone(old), nullptr);
    }
}

void gen_unary_to_another(const Expr &old, const std::vector<Expr> &allowed, std::vector<Expr> &out)  // :258-270
{
    if (old.arity() != 1) return;
    for (const Expr &u : allowed) {
        if (u.arity() != 1 || u.type == old.type) continue;
        if (!allowed_left(u.type, *old.left)) continue;
        out.emplace_back(u.type, clone(*old.left), nullptr);
    }
}

void gen_binary_applied(const Expr &old, const std::vector<Expr> &allowed, std::vector<Expr> &out)  // :272-296
{
    std::vector<const Expr *> args;
    all_subtrees(old, args);
    for (const Expr &n : allowed)
        if (n.is(Op::VAR) || n.is(Op::CONST)) args.push_back(&n);
    for (const Expr &b : allowed) {
        if (b.arity() != 2) continue;
        for (const Expr *a : args) {
            if (allowed_left(b.type, old)) out.emplace_back(b.type, clone(old), clone(*a));
            if (!symmetric_of(b.type) && allowed_left(b.type, *a)) out.emplace_back(b.type, clone(*a), clone(old));
        }
    }
}

void gen_binary_to_another(const Expr &old, const std::vector<Expr> &allowed, std::vector<Expr> &out)  // :303-316
{
    if (old.arity() != 2) return;
    for (const Expr &b : allowed) {
        if (b.arity() != 2 || b.type == old.type) continue;
        if (allowed_left(b.type, *old.left)) out.emplace_back(b.type, clone(*old.left), clone(*old.right));
        if (!symmetric_of(b.type) && allowed_left(b.type, *old.right))
            out.emplace_back(b.type, clone(*old.right), clone(*old.left));
    }
}

}  // namespace

std::vector<Expr> Search::perturb_candidates(const Expr &old) const  // :318-330
{
    std::vector<Expr> c;
    c.reserve(1000);
    gen_to_subtree(old, c);
    gen_to_var_or_1(old, allowed_, c);
    if (old.is(Op::VAR)) gen_unary_applied(old, allowed_, c);                     // :253-256
    gen_unary_to_another(old, allowed_, c);
    if (old.is(Op::VAR) || old.is(Op::CONST)) gen_binary_applied(old, allowed_, c);  // :298-301
    gen_binary_to_another(old, allowed_, c);
    return c;
}

std::vector<Expr> Search::change_candidates(const Expr &old) const  // :332-346
{
    std::vector<Expr> c;
    c.reserve(1000);
    gen_const_finetune(old, c);
    gen_to_subtree(old, c);
    gen_to_var_or_1(old, allowed_, c);
    gen_unary_applied(old, allowed_, c);
    gen_unary_to_another(old, allowed_, c);
    gen_binary_applied(old, allowed_, c);
    gen_binary_to_another(old, allowed_, c);
    return c;
}

std::vector<Expr> Search::all_candidates(const Expr &passed, bool local_search) const  // :348-443
{
    ExprP solution = clone(passed);
    std::vector<Expr> all;
    all.reserve(3000);
    std::unordered_set<std::string> seen;
    seen.reserve(3000);
    auto push_unique = [&](Expr cand) {
        std::string s = to_string(cand);
        if (seen.insert(std::move(s)).second) all.push_back(std::move(cand));
    };
    // the whole solution with one subtree replaced: the string decides first, the deep copy is only made for
    // a tree that has not been seen
    auto push_unique_copy = [&](const Expr &whole_tree) {
        std::string s = to_string(whole_tree);
        if (seen.insert(std::move(s)).second) all.push_back(Expr(whole_tree));
    };
    auto candidates = [&](const Expr &t) { return local_search ? change_candidates(t) : perturb_candidates(t); };
    const int whole = size_of(*solution);
    std::vector<Expr *> queue{solution.get()};
    for (size_t i = 0; i < queue.size(); ++i) {
        Expr &sub = *queue[i];
        if (size_of(sub) == whole) {  // the whole tree is replaced
            for (Expr &c : candidates(sub)) push_unique(std::move(c));
        }
        if (sub.arity() >= 1) {
            std::vector<Expr> cands = candidates(*sub.left);
            ExprP keep = std::move(sub.left);
            for (Expr &c : cands) {
                sub.left = std::make_unique<Expr>(std::move(c));
                push_unique_copy(*solution);
            }
            sub.left = std::move(keep);
            queue.push_back(sub.left.get());
        }
        if (sub.arity() >= 2) {
            std::vector<Expr> cands = candidates(*sub.right);
            ExprP keep = std::move(sub.right);
            for (Expr &c : cands) {
                sub.right = std::make_unique<Expr>(std::move(c));
                push_unique_copy(*solution);
            }
            sub.right = std::move(keep);
            queue.push_back(sub.right.get());
        }
    }
    for (Expr &c : all) {  // :413-425
        int it_max = 5;
        while (it_max > 0) {
            const int sz = size_of(c);
            expand(c);
            normalize_factor_constants(c, Op::NONE, false);
            simplify(c);
            if (size_of(c) == sz) break;
            it_max--;
        }
    }
    std::unordered_set<std::string> filtered_strings;  // :427-442
    filtered_strings.reserve(all.size());
    std::vector<Expr> filtered;
    filtered.reserve(all.size());
    for (Expr &c : all) {
        if (!filtered_strings.insert(to_string(c)).second) continue;
        filtered.push_back(std::move(c));
    }
    return filtered;
}

// ---- fitness -----------------------------------------------------------------------------------
Fitness Search::fitness_from(double ssr, int size) const  // :520-541 with R2()/RMSE() :40-49
{
    const double r2 = 1 - ssr / sst_;
    const double rmse = std::sqrt(ssr / (double)n_);
    if (r2 != r2 || rmse != rmse) return Fitness{1000, 1000, 1000};
    return Fitness{1 - r2, rmse, size};
}

double Search::fitness_value(const Fitness &f) const  // :543-545
{
    return (1 + std::get<0>(f)) * (1 + std::get<1>(f)) * (1 + std::get<2>(f) * p_.complexity_penalty);
}

int Search::compare_fitness(const Fitness &a, const Fitness &b) const  // :547-561
{
    const int s1 = std::get<2>(a), s2 = std::get<2>(b);
    if ((s1 > p_.max_complexity || s2 > p_.max_complexity) && s1 != s2) return s1 - s2;
    const double f1 = fitness_value(a), f2 = fitness_value(b);
    if (f1 < f2) return -1;
    if (f1 > f2) return 1;
    return 0;
}

// RR_FLAG_CLASSIFIER_OBJECTIVE of the driver (default off): the objective the reference's authors
// left commented out at :527 - loss = 1 - classification_accuracy - with average_log_loss (:63-76) in the second slot,
// both from the engine's RI_CLSMET reduction of the scored model. NaN in either -> sentinel, like :529-530.
Fitness Search::classifier_fitness_from(double accuracy, double log_loss, int size) const
{
    const double loss = 1 - accuracy;
    if (loss != loss || log_loss != log_loss) return Fitness{1000, 1000, 1000};
    return Fitness{loss, log_loss, size};
}

// fitness of fully specified trees (no OLS): one EVAL_ONLY batch; with the classifier objective the metrics call
std::vector<Fitness> Search::score_trees(const std::vector<const Expr *> &trees, std::vector<double> *ssr_out)
{
    BatchBuilder bb(RR_MODE_EVAL_ONLY);
    for (const Expr *t : trees) bb.add_eval(*t);
    rr_batch b = bb.view();
    const size_t m = trees.size();
    std::vector<Fitness> f(m);
    if (classifier_objective()) {
        std::vector<double> acc(m), ll(m);
        engine_check(rr_classifier_metrics(eng_, &b, acc.data(), ll.data(), nullptr), "rr_classifier_metrics");
        for (size_t i = 0; i < m; ++i) f[i] = classifier_fitness_from(acc[i], ll[i], bb.eval_size[i]);
        if (ssr_out) ssr_out->assign(m, 0.0);
        return f;
    }
    std::vector<double> ssr(m);
    rr_result r{nullptr, nullptr, ssr.data(), nullptr};
    engine_check(rr_score_batch(eng_, &b, &r), "rr_score_batch");
    for (size_t i = 0; i < m; ++i) f[i] = fitness_from(ssr[i], bb.eval_size[i]);
    if (ssr_out) *ssr_out = std::move(ssr);
    return f;
}

Fitness Search::score_single(const Expr &tree)
{
    std::vector<const Expr *> one{&tree};
    const Fitness f = score_trees(one, nullptr)[0];
    fit_calls_++;
    return f;
}

ExprP Search::tune_single(const Expr &tree, Fitness *fit)
{
    BatchBuilder bb(RR_MODE_OLS_FIT);
    bb.add_ols(tree);
    rr_batch b = bb.view();
    std::vector<double> coef(bb.factors[0].size() + 1);
    double ssr = 0.0;
    rr_result r{coef.data(), nullptr, &ssr, nullptr};
    engine_check(rr_score_batch(eng_, &b, &r), "rr_score_batch");
    ExprP tuned = rebuild_from_coefficients(bb.factors[0], coef.data());
    if (classifier_objective()) {
        *fit = score_single(*tuned);  // counts the fit call
        return tuned;
    }
    fit_calls_++;
    *fit = fitness_from(ssr, size_of(*tuned));
    return tuned;
}

void Search::print_state(const Fitness &curr) const  // :563-579
{
    std::cout << "it=" << main_it_ << "\tfit_calls=" << fit_calls_ << "\tls_calls=" << ls_calls_;
    if (p_.classification) {
        std::cout << "\tcurr_LOSS=" << std::get<0>(curr) << "\tcurr_size=" << std::get<2>(curr);
        std::cout << "\tfinal_LOSS=" << std::get<0>(final_fit_) << "\tfinal_size=" << std::get<2>(final_fit_);
    } else {
        std::cout << "\tcurr_R2=" << (1 - std::get<0>(curr)) << "\tcurr_RMSE=" << std::get<1>(curr)
                  << "\tcurr_size=" << std::get<2>(curr);
        std::cout << "\tfinal_R2=" << (1 - std::get<0>(final_fit_)) << "\tfinal_RMSE=" << std::get<1>(final_fit_)
                  << "\tfinal_size=" << std::get<2>(final_fit_);
    }
    std::cout << "\tchecks_skip=" << skipped_perts_ << "/" << total_perts_ << "\tsol=" << to_string(*final_) << std::endl
              << std::endl;
}

// ---- local search (:601-643): whole neighbourhood in one OLS_FIT batch, sequential replay --------
ExprP Search::local_search(const Expr &start)
{
    std::vector<Fitness> pareto;
    ls_calls_++;
    bool improved = true;
    Fitness curr_fit;
    ExprP curr = tune_single(start, &curr_fit);  // :606-607
    while (improved && !finished()) {
        improved = false;
        std::vector<Expr> perts = all_candidates(*curr, true);
        // every candidate costs at least one fit call, so at most `remaining` of them can be replayed
        const size_t remaining = (size_t)std::max(0, p_.max_fit_calls - fit_calls_);
        const size_t m = std::min(perts.size(), remaining);
        if (m == 0) break;
        BatchBuilder bb(RR_MODE_OLS_FIT);
        for (size_t j = 0; j < m; ++j) bb.add_ols(perts[j]);
        rr_batch b = bb.view();
        std::vector<double> coef(bb.term_code_begin.size() - 1 + m), ssr(m);
        rr_result r{coef.data(), nullptr, ssr.data(), nullptr};
        engine_check(rr_score_batch(eng_, &b, &r), "rr_score_batch");
        TraceBatch *tb = nullptr;
        if (trace_) {
            trace_log_.emplace_back();
            tb = &trace_log_.back();
            tb->mode = RR_MODE_OLS_FIT;
            tb->cand_term_begin = bb.cand_term_begin;
            tb->term_code_begin = bb.term_code_begin;
            tb->code = bb.code;
            tb->consts = bb.consts;
            tb->coef = coef;
            tb->ssr = ssr;
            tb->size.assign(m, 0);
            tb->consumed.assign(m, 0);
            tb->curr_f0 = std::get<0>(curr_fit);
            tb->curr_f1 = std::get<1>(curr_fit);
            tb->curr_size = std::get<2>(curr_fit);
            tb->fit_calls_before = fit_calls_;
        }
        std::vector<Fitness> cls_fit;
        if (classifier_objective()) {
            // the tuned trees themselves (coefficients from the least-squares fit as in the reference), scored with the
            // classifier metrics in one batch
            std::vector<ExprP> tuned(m);
            std::vector<const Expr *> ptrs(m);
            for (size_t j = 0; j < m; ++j) {
                tuned[j] = rebuild_from_coefficients(bb.factors[j], coef.data() + bb.cand_term_begin[j] + j);
                ptrs[j] = tuned[j].get();
            }
            cls_fit = score_trees(ptrs, nullptr);
        }
        for (size_t j = 0; j < m; ++j) {  // :611-639 replay
            if (finished()) break;
            const double *cj = coef.data() + bb.cand_term_begin[j] + j;
            const int size = rebuilt_size(bb.factors[j], cj);
            fit_calls_++;  // the fitness() call of :616
            Fitness f = classifier_objective() ? cls_fit[j] : fitness_from(ssr[j], size);
            if (tb) {
                tb->size[j] = size;
                tb->consumed[j] = 1;
            }
            if (p_.verbose && fit_calls_ % 10000 == 0) print_state(curr_fit);
            if (!is_dominated(pareto, f) && compare_fitness(f, curr_fit) < 0) {
                improved = true;
                ExprP tuned = rebuild_from_coefficients(bb.factors[j], cj);
                int it_max = 5;
                while (it_max > 0) {  // :622-630
                    const int sz = size_of(*tuned);
                    expand(*tuned);
                    simplify(*tuned);
                    if (sz == size_of(*tuned)) break;
                    it_max--;
                }
                f = score_single(*tuned);  // :631
                curr = std::move(tuned);
                curr_fit = f;
                add_to_pareto(pareto, f);
                if (tb) {
                    tb->accepted.push_back((int32_t)j);
                    tb->accepted_fit.push_back(std::get<0>(f));
                    tb->accepted_fit.push_back(std::get<1>(f));
                    tb->accepted_fit.push_back((double)std::get<2>(f));
                }
            }
        }
    }
    return curr;
}

// ---- fit (:717-728, :772-859) -------------------------------------------------------------------
void Search::fit(const double *Xr, const double *y, int64_t n_all, int32_t d)
{
    if (!Xr || !y || n_all <= 0 || d <= 0) throw std::invalid_argument("fit: empty data");
    reset();
    trace_log_.clear();
    const int64_t sample_cnt = (int64_t)(int)(p_.sample_size * n_all);  // :774
    if (sample_cnt <= 0) throw std::invalid_argument("fit: sample_size selects no rows");
    std::vector<int> selected(n_all);
    std::iota(selected.begin(), selected.end(), 0);
    std::shuffle(selected.begin(), selected.end(), std::default_random_engine(p_.random_state));  // :778
    // :788-795 without the host loops: the engine gathers rows selected[0 .. sample_cnt) straight from the caller's
    // row-major matrix on the device(s) (same row order, same bits). Large data sets are sharded by sample inside the
    // engine: one GPU per 2^23 rows, as many as are visible (RR_B200_GPUS overrides: a count, 0 = all). A search step
    // on few rows per GPU is host-bound - measured on 2^24 x 20 rows, 6000 fitness calls: 4.3 s on one GPU, 4.2 s on
    // two, 12.1 s on eight (eight contexts and communicators to set up, eight launches per sweep from one thread).
    if (eng_) {
        rr_engine_destroy(eng_);
        eng_ = nullptr;
    }
    int n_gpus = 1;
    {
        const char *g = std::getenv("RR_B200_GPUS");
        if (g && *g) n_gpus = std::atoi(g);
        else n_gpus = (int)std::max<int64_t>(1, sample_cnt >> 23);  // the engine clamps to the visible devices
    }
    static_assert(sizeof(int) == sizeof(int32_t), "row index type");
    const int rc = rr_engine_create_sharded(Xr, y, n_all, d, reinterpret_cast<const int32_t *>(selected.data()), sample_cnt, n_gpus,
                                            RR_FLAG_X_ROWMAJOR, &eng_);
    if (rc != RR_OK) throw std::runtime_error(std::string("rr_engine_create failed: ") + rr_last_error(nullptr));
    rr_engine_info info;
    engine_check(rr_engine_get_info(eng_, &info), "rr_engine_get_info");
    n_ = sample_cnt;
    d_ = d;
    sst_ = info.sst;
    // relevant_features, :753-770 (only active beyond 200 features): R2(X[j], y) per feature is one device reduction
    std::vector<int> rel;
    const int max_feat = 200;
    if (d <= max_feat) {
        rel.resize(d);
        std::iota(rel.begin(), rel.end(), 0);
    } else {
        std::vector<double> r2(d);
        engine_check(rr_feature_r2(eng_, r2.data()), "rr_feature_r2");
        std::vector<std::tuple<double, int>> by_r2;
        for (int j = 0; j < d; ++j) by_r2.emplace_back(r2[j], j);
        std::sort(by_r2.begin(), by_r2.end(), std::greater<>());
        for (int i = 0; i < max_feat; ++i) rel.push_back(std::get<1>(by_r2[i]));
    }
    setup_nodes(rel);

    final_ = std::make_unique<Expr>(0.0);  // :799-800
    final_fit_ = score_single(*final_);
    bool improved = true;
    while (!finished()) {  // :803
        main_it_ += 1;
        ExprP start = clone(*final_);
        if (!improved) {  // :806-813: two random perturbations of the best solution
            std::vector<Expr> p1 = all_candidates(*final_, false);
            std::vector<Expr> p2 = all_candidates(p1[rand() % p1.size()], false);
            start = clone(p2[rand() % p2.size()]);
            if (p_.verbose) std::cout << "Randomized to " << to_string(*start) << std::endl;
        }
        improved = false;
        std::vector<Expr> perts = all_candidates(*start, false);  // :815
        if (p_.verbose) std::cout << "Checking " << perts.size() << " perturbations of starting solution." << std::endl;
        // :819-830 — check_skip sequentially (it mutates the set), then one EVAL_ONLY batch
        std::vector<size_t> picked;
        {
            int sim_calls = fit_calls_;
            for (size_t i = 0; i < perts.size(); ++i) {
                if (sim_calls >= p_.max_fit_calls ||
                    duration_cast<seconds>(high_resolution_clock::now() - start_).count() > p_.max_seconds)
                    break;
                if (check_skip(to_string(perts[i]))) continue;
                picked.push_back(i);
                ++sim_calls;
            }
        }
        std::vector<std::pair<double, size_t>> by_r2;
        if (!picked.empty()) {
            BatchBuilder bb(RR_MODE_EVAL_ONLY);
            for (size_t i : picked) bb.add_eval(perts[i]);
            std::vector<const Expr *> ptrs;
            for (size_t i : picked) ptrs.push_back(&perts[i]);
            std::vector<double> ssr;
            const std::vector<Fitness> pf = score_trees(ptrs, &ssr);
            fit_calls_ += (int)picked.size();
            for (size_t k = 0; k < picked.size(); ++k) by_r2.emplace_back(std::get<0>(pf[k]), picked[k]);
            if (trace_) {
                trace_log_.emplace_back();
                TraceBatch &tb = trace_log_.back();
                tb.mode = RR_MODE_EVAL_ONLY;
                tb.cand_term_begin = bb.cand_term_begin;
                tb.term_code_begin = bb.term_code_begin;
                tb.code = bb.code;
                tb.consts = bb.consts;
                tb.ssr = ssr;
                tb.size = bb.eval_size;
                tb.consumed.assign(picked.size(), 1);
                tb.fit_calls_before = fit_calls_ - (int)picked.size();
            }
        }
        // :831 — std::sort on the double only (ties: whatever introsort does with this sequence)
        std::sort(by_r2.begin(), by_r2.end(),
                  [](const std::pair<double, size_t> &a, const std::pair<double, size_t> &b) { return a.first < b.first; });
        for (size_t k = 0; k < by_r2.size(); ++k) {  // :833-855
            if (finished()) break;
            const Expr &pert = perts[by_r2[k].second];
            checked_perts_.insert(to_string(pert));
            ExprP ls = local_search(pert);
            const Fitness f = score_single(*ls);  // :842
            if (compare_fitness(f, final_fit_) < 0) {
                improved = true;
                final_ = clone(*ls);
                final_fit_ = f;
                if (p_.verbose) print_state(final_fit_);
                best_time_ = duration_cast<milliseconds>(high_resolution_clock::now() - start_).count() / 1000.0;
            }
        }
    }
    total_time_ = duration_cast<milliseconds>(high_resolution_clock::now() - start_).count() / 1000.0;
    rr_get_stats(eng_, &stats_);
}

void Search::predict(const double *Xr, int64_t n, int32_t d, double *out) const  // :730-750
{
    if (!final_ || !eng_) throw std::runtime_error("predict before fit");
    std::vector<uint32_t> code;
    std::vector<double> consts;
    compile_postfix(*final_, code, consts);
    const double dummy = 0.0;
    engine_check(rr_predict_rowmajor(eng_, code.data(), (int32_t)code.size(), consts.empty() ? &dummy : consts.data(),
                                     (int32_t)consts.size(), Xr, n, d, out),
                 "rr_predict");
    if (p_.classification)
        for (int64_t i = 0; i < n; ++i) out[i] = out[i] >= 0.5 ? 1.0 : 0.0;
}

void Search::predict_proba(const double *Xr, int64_t n, int32_t d, double *out) const
{
    if (!final_ || !eng_) throw std::runtime_error("predict_proba before fit");
    std::vector<uint32_t> code;
    std::vector<double> consts;
    compile_postfix(*final_, code, consts);
    const double dummy = 0.0;
    engine_check(rr_predict_proba_rowmajor(eng_, code.data(), (int32_t)code.size(), consts.empty() ? &dummy : consts.data(),
                                           (int32_t)consts.size(), Xr, n, d, out),
                 "rr_predict_proba");
}

std::string Search::model_string() const
{
    if (!final_) throw std::runtime_error("model_string before fit");
    return to_string(*final_);
}

}  // namespace rrd
