// TEST INFRASTRUCTURE — CPU oracle. Not part of the product.
//
// ref_harness.cpp — builds the UNMODIFIED reference sources, where they lie under
// /root/reference, into a pybind11 module `rils_rols_cpp_ref`:
//   * class rils_rols exactly as the reference binds it (rils_rols_cpp.cpp:998-1007) —
//     the reference's PYBIND11_MODULE body is captured and re-used verbatim;
//   * class RefHarness that calls the reference's private hot-path members
//     (tune_constants :445, fitness :520, all_candidates :348) on trees passed as the
//     postfix bytecode of include/rr_b200.h, and reports what the reference computed
//     (coefficients, nonzero_pivots, fitness tuple, tuned tree) for golden fixtures,
//     parity tests and the CPU baseline timing.
// The reference translation unit is #included, not copied; `private` is opened for
// that include only (the standard headers are included before it, with normal access).
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cstdint>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <numeric>
#include <ostream>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "Core" // the stand-in, parsed before `private` is redefined

#include "../include/rr_b200.h"

#ifndef RR_REFERENCE_MAIN_CPP
#error "define RR_REFERENCE_MAIN_CPP to the path of the reference's rils_rols_cpp.cpp"
#endif

// The real module is declared first, with pybind11's own macro; the reference's
// PYBIND11_MODULE(rils_rols_cpp, m) block (rils_rols_cpp.cpp:998-1007) is then captured as
// the body of rr_reference_bindings().
static void rr_reference_bindings(pybind11::module_ &m);
static void rr_harness_bindings(pybind11::module_ &m);
PYBIND11_MODULE(rils_rols_cpp_ref, m)
{
    m.doc() = "unmodified kartelj/rils-rols reference (Eigen stand-in) + hot-path harness; test oracle";
    rr_reference_bindings(m);
    rr_harness_bindings(m);
}

#undef PYBIND11_MODULE
#define PYBIND11_MODULE(name, var) static void rr_reference_bindings(pybind11::module_ &var)
#define main rr_reference_main_unused
#define private public
#include RR_REFERENCE_MAIN_CPP
#undef private
#undef main
#undef PYBIND11_MODULE

namespace {

using NodeP = std::shared_ptr<node>;

struct Postfix {
    std::vector<uint32_t> code;
    std::vector<double> consts;
};

void emit_postfix(const node *t, Postfix &out)
{
    if (t->get_arity() >= 1) emit_postfix(t->get_left(), out);
    if (t->get_arity() >= 2) emit_postfix(t->get_right(), out);
    const auto op = static_cast<uint32_t>(t->get_type());
    if (t->is<node_type::CONST>()) {
        out.code.push_back(RR_INS(op, out.consts.size()));
        out.consts.push_back(t->get_const_value());
    } else if (t->is<node_type::VAR>()) {
        out.code.push_back(RR_INS(op, t->get_var_index()));
    } else {
        out.code.push_back(RR_INS(op, 0));
    }
}

NodeP build_tree(const uint32_t *code, size_t len, const double *consts, size_t n_consts)
{
    std::vector<NodeP> st;
    for (size_t i = 0; i < len; ++i) {
        const uint32_t op = RR_INS_OP(code[i]), arg = RR_INS_ARG(code[i]);
        if (op == RR_OP_CONST) {
            if (arg >= n_consts) throw std::runtime_error("constant index out of range");
            st.push_back(std::make_shared<node>(consts[arg]));
        } else if (op == RR_OP_VAR) {
            st.push_back(std::make_shared<node>(static_cast<int>(arg)));
        } else if (op > RR_OP_VAR && op < RR_OP_COUNT) {
            const auto type = static_cast<node_type>(op);
            const int ar = get_arity(type);
            if (static_cast<int>(st.size()) < ar) throw std::runtime_error("malformed postfix");
            NodeP r, l;
            if (ar == 2) { r = st.back(); st.pop_back(); }
            l = st.back(); st.pop_back();
            st.push_back(std::make_shared<node>(type, l, r));
        } else {
            throw std::runtime_error("bad opcode");
        }
    }
    if (st.size() != 1) throw std::runtime_error("postfix does not reduce to one tree");
    return st[0];
}

// Restates the factor selection of tune_constants(), rils_rols_cpp.cpp:450-473, on a tree
// that has already been through expand(); simplify() (:448-449).
std::vector<node *> select_factors(node *solution)
{
    std::vector<node *> all_factors, factors;
    solution->extract_non_constant_factors(all_factors); // node.cpp:140-147
    for (auto f : all_factors) {
        if (f->is<node_type::CONST>()) continue;
        if (f->get_arity() == 2 && f->get_left()->is<node_type::CONST>() &&
            f->get_right()->is<node_type::CONST>())
            continue;
        if (f->is<node_type::MULTIPLY>() || f->is<node_type::PLUS>() || f->is<node_type::MINUS>()) {
            if (f->get_left()->is<node_type::CONST>()) { factors.push_back(f->get_right()); continue; }
            if (f->get_right()->is<node_type::CONST>()) { factors.push_back(f->get_left()); continue; }
        }
        if (f->is<node_type::DIVIDE>() && f->get_right()->is<node_type::CONST>()) {
            factors.push_back(f->get_left());
            continue;
        }
        factors.push_back(f);
    }
    return factors;
}

py::array_t<double> to_np(const std::vector<double> &v)
{
    py::array_t<double> a(v.size());
    std::copy(v.begin(), v.end(), a.mutable_data());
    return a;
}
template <typename T> py::array_t<T> to_np_t(const std::vector<T> &v)
{
    py::array_t<T> a(v.size());
    std::copy(v.begin(), v.end(), a.mutable_data());
    return a;
}

class RefHarness {
    bool classification_;
    double penalty_;
    int max_complexity_, random_state_;
    std::unique_ptr<rils_rols> rr_;
    std::vector<Eigen::ArrayXd> X_;
    Eigen::ArrayXd y_;

    std::unique_ptr<rils_rols> make_rr() const
    {
        return std::make_unique<rils_rols>(classification_, 2000000000, 2000000000, penalty_,
                                           max_complexity_, 1.0, false, random_state_);
    }

public:
    RefHarness(bool classification, double complexity_penalty, int max_complexity, int random_state)
        : classification_(classification), penalty_(complexity_penalty),
          max_complexity_(max_complexity), random_state_(random_state), rr_(make_rr())
    {
    }

    // X row-major n x d as the pybind boundary receives it (rils_rols_cpp.cpp:690-696)
    void set_data(py::array_t<double, py::array::c_style | py::array::forcecast> X,
                  py::array_t<double, py::array::c_style | py::array::forcecast> y)
    {
        if (X.ndim() != 2 || y.ndim() != 1 || X.shape(0) != y.shape(0))
            throw std::runtime_error("set_data: X must be (n,d), y (n,)");
        const auto n = X.shape(0), d = X.shape(1);
        X_.assign(d, Eigen::ArrayXd());
        for (py::ssize_t j = 0; j < d; ++j) X_[j].resize(n);
        const double *px = X.data();
        for (py::ssize_t i = 0; i < n; ++i)
            for (py::ssize_t j = 0; j < d; ++j) X_[j][i] = px[i * d + j];
        y_.resize(n);
        for (py::ssize_t i = 0; i < n; ++i) y_[i] = y.data()[i];
        rr_ = make_rr();
        rr_->reset();
        rr_->setup_nodes(rr_->relevant_features(X_, y_)); // rils_rols_cpp.cpp:797-798
    }

    std::string to_string(py::array_t<uint32_t> code, py::array_t<double> consts) const
    {
        return build_tree(code.data(), code.size(), consts.data(), consts.size())->to_string();
    }

    // fitness() of the tree as-is (rils_rols_cpp.cpp:520-541) -> (1-R2, RMSE, size)
    std::tuple<double, double, int> fitness(py::array_t<uint32_t> code, py::array_t<double> consts)
    {
        auto t = build_tree(code.data(), code.size(), consts.data(), consts.size());
        return rr_->fitness(t, X_, y_);
    }

    py::array_t<double> evaluate(py::array_t<uint32_t> code, py::array_t<double> consts)
    {
        auto t = build_tree(code.data(), code.size(), consts.data(), consts.size());
        Eigen::ArrayXd v = t->evaluate_all(X_);
        py::array_t<double> out(v.size());
        std::copy(v.data(), v.data() + v.size(), out.mutable_data());
        return out;
    }

    // tune_constants() + fitness() of one tree (rils_rols_cpp.cpp:615-616) with everything the
    // reference computed on the way.
    py::dict tune(py::array_t<uint32_t> code, py::array_t<double> consts, bool keep_matrix)
    {
        py::dict out;
        auto t = build_tree(code.data(), code.size(), consts.data(), consts.size());
        // the factor list the reference will build, from a private copy
        auto t2 = node::node_copy(*t);
        t2->expand();
        t2->simplify();
        py::list term_codes, term_consts, term_strs;
        for (node *f : select_factors(t2.get())) {
            Postfix p;
            emit_postfix(f, p);
            term_codes.append(to_np_t<uint32_t>(p.code));
            term_consts.append(to_np(p.consts));
            term_strs.append(f->to_string());
        }
        auto &rec = Eigen::shim_last_qr();
        rec.keep_matrix = keep_matrix;
        const auto calls0 = rec.calls;
        NodeP tuned = rr_->tune_constants(t, X_, y_);
        if (rec.calls != calls0 + 1) throw std::runtime_error("tune_constants did not run exactly one QR");
        if (static_cast<size_t>(rec.cols) != py::len(term_codes) + 1)
            throw std::runtime_error("factor selection restatement disagrees with the reference (k mismatch)");
        const auto fit = rr_->fitness(tuned, X_, y_);
        Postfix tp;
        emit_postfix(tuned.get(), tp);
        out["term_code"] = term_codes;
        out["term_consts"] = term_consts;
        out["term_str"] = term_strs;
        out["coef"] = to_np(rec.coefs);
        out["nonzero_pivots"] = rec.nonzero_pivots;
        out["perm"] = to_np_t<int>(rec.perm);
        out["fitness"] = py::make_tuple(std::get<0>(fit), std::get<1>(fit), std::get<2>(fit));
        out["tuned_str"] = tuned->to_string();
        out["tuned_code"] = to_np_t<uint32_t>(tp.code);
        out["tuned_consts"] = to_np(tp.consts);
        if (keep_matrix) {
            py::array_t<double> A({rec.cols, rec.rows}); // A[j] = column j
            std::copy(rec.A.begin(), rec.A.end(), A.mutable_data());
            out["A_cols"] = A;
        }
        rec.keep_matrix = false;
        return out;
    }

    // all_candidates() (rils_rols_cpp.cpp:348-443) as a list of (code, consts, string)
    py::list all_candidates(py::array_t<uint32_t> code, py::array_t<double> consts, bool local_search)
    {
        auto t = build_tree(code.data(), code.size(), consts.data(), consts.size());
        std::vector<node> cands = rr_->all_candidates(t, X_, y_, local_search);
        py::list out;
        for (auto &c : cands) {
            Postfix p;
            emit_postfix(&c, p);
            out.append(py::make_tuple(to_np_t<uint32_t>(p.code), to_np(p.consts), c.to_string()));
        }
        return out;
    }

    // Score a whole list of trees the way the LS loop does (tune + fitness, :615-616) or the
    // perturbation loop does (fitness only, :828) and return the rr_batch arrays together with
    // the reference's results, ready to be stored as a golden fixture.
    py::dict score_list(py::list trees, bool ols_fit)
    {
        std::vector<int32_t> cand_term_begin{0}, term_code_begin{0};
        std::vector<uint32_t> code;
        std::vector<double> consts, coef, f0, f1;
        std::vector<int32_t> nzp, size;
        py::list strs;
        auto &rec = Eigen::shim_last_qr();
        rec.keep_matrix = false;
        auto append_term = [&](const node *f) {
            Postfix p;
            emit_postfix(f, p);
            for (uint32_t w : p.code) {
                if (RR_INS_OP(w) == RR_OP_CONST)
                    code.push_back(RR_INS(RR_OP_CONST, consts.size() + RR_INS_ARG(w)));
                else
                    code.push_back(w);
            }
            consts.insert(consts.end(), p.consts.begin(), p.consts.end());
            term_code_begin.push_back(static_cast<int32_t>(code.size()));
        };
        for (auto item : trees) {
            auto tup = item.cast<py::tuple>();
            auto c = tup[0].cast<py::array_t<uint32_t>>();
            auto k = tup[1].cast<py::array_t<double>>();
            auto t = build_tree(c.data(), c.size(), k.data(), k.size());
            std::tuple<double, double, int> fit;
            if (ols_fit) {
                auto t2 = node::node_copy(*t);
                t2->expand();
                t2->simplify();
                size_t nf = 0;
                for (node *f : select_factors(t2.get())) { append_term(f); ++nf; }
                const auto calls0 = rec.calls;
                NodeP tuned = rr_->tune_constants(t, X_, y_);
                if (rec.calls != calls0 + 1 || static_cast<size_t>(rec.cols) != nf + 1)
                    throw std::runtime_error("factor selection restatement disagrees with the reference");
                coef.insert(coef.end(), rec.coefs.begin(), rec.coefs.end());
                nzp.push_back(rec.nonzero_pivots);
                fit = rr_->fitness(tuned, X_, y_);
                strs.append(tuned->to_string());
            } else {
                append_term(t.get());
                fit = rr_->fitness(t, X_, y_);
                strs.append(t->to_string());
            }
            cand_term_begin.push_back(static_cast<int32_t>(term_code_begin.size() - 1));
            f0.push_back(std::get<0>(fit));
            f1.push_back(std::get<1>(fit));
            size.push_back(std::get<2>(fit));
        }
        py::dict out;
        out["mode"] = ols_fit ? (int)RR_MODE_OLS_FIT : (int)RR_MODE_EVAL_ONLY;
        out["cand_term_begin"] = to_np_t<int32_t>(cand_term_begin);
        out["term_code_begin"] = to_np_t<int32_t>(term_code_begin);
        out["code"] = to_np_t<uint32_t>(code);
        out["consts"] = to_np(consts);
        out["ref_coef"] = to_np(coef);
        out["ref_nonzero_pivots"] = to_np_t<int32_t>(nzp);
        out["ref_f0"] = to_np(f0);
        out["ref_f1"] = to_np(f1);
        out["ref_size"] = to_np_t<int32_t>(size);
        out["ref_str"] = strs;
        return out;
    }

    // CPU baseline: wall seconds for the reference to score trees[lo:hi) (tune_constants +
    // fitness when ols_fit, else fitness), on n_threads independent replicas of the
    // single-threaded reference (candidates split round-robin).
    double time_list(py::list trees, bool ols_fit, int n_threads)
    {
        std::vector<NodeP> ts;
        for (auto item : trees) {
            auto tup = item.cast<py::tuple>();
            auto c = tup[0].cast<py::array_t<uint32_t>>();
            auto k = tup[1].cast<py::array_t<double>>();
            ts.push_back(build_tree(c.data(), c.size(), k.data(), k.size()));
        }
        if (n_threads < 1) n_threads = 1;
        std::vector<std::unique_ptr<rils_rols>> reps;
        for (int t = 0; t < n_threads; ++t) reps.push_back(make_rr());
        py::gil_scoped_release nogil;
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) {
            th.emplace_back([&, t]() {
                for (size_t i = t; i < ts.size(); i += n_threads) {
                    NodeP cand = node::node_copy(*ts[i]);
                    if (ols_fit) {
                        NodeP tuned = reps[t]->tune_constants(cand, X_, y_);
                        (void)reps[t]->fitness(tuned, X_, y_);
                    } else {
                        (void)reps[t]->fitness(cand, X_, y_);
                    }
                }
            });
        }
        for (auto &x : th) x.join();
        const auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
};

} // namespace

static void rr_harness_bindings(pybind11::module_ &m)
{
    py::class_<RefHarness>(m, "RefHarness")
        .def(py::init<bool, double, int, int>(), py::arg("classification") = false,
             py::arg("complexity_penalty") = 0.001, py::arg("max_complexity") = 50,
             py::arg("random_state") = 0)
        .def("set_data", &RefHarness::set_data)
        .def("to_string", &RefHarness::to_string)
        .def("fitness", &RefHarness::fitness)
        .def("evaluate", &RefHarness::evaluate)
        .def("tune", &RefHarness::tune, py::arg("code"), py::arg("consts"), py::arg("keep_matrix") = false)
        .def("all_candidates", &RefHarness::all_candidates)
        .def("score_list", &RefHarness::score_list)
        .def("time_list", &RefHarness::time_list);
}
