// rr_pymodule.cpp — pybind11 module `rils_rols_cpp`: the boundary the Python front end binds
// (/root/reference/rils_rols_cpp/rils_rols_cpp.cpp:998-1007), kept signature for signature:
//   rils_rols(bool classification, int max_fit_calls, int max_seconds, double complexity_penalty,
//             int max_complexity, double sample_size, bool verbose, int random_state)
//   .fit(X flat row-major, y, data_cnt, feat_cnt) .predict(X, data_cnt, feat_cnt)
//   .get_model_string() .get_best_time() .get_fit_calls() .get_total_time()
// Behind it: the host driver (rr_search.cpp) and, through the C ABI, the B200 engine.
// Differences on purpose: a size mismatch raises ValueError instead of exit(1) (:722-725), and a
// few extra, optional entry points expose engine statistics and the decision trace for tests.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <numeric>
#include <random>
#include <sstream>

#include "rr_search.h"

namespace py = pybind11;
using rrd::Expr;
using rrd::ExprP;
using rrd::Search;
using rrd::SearchParams;

namespace {

using DArr = py::array_t<double, py::array::c_style | py::array::forcecast>;
using UArr = py::array_t<uint32_t, py::array::c_style | py::array::forcecast>;

template <typename T> py::array_t<T> to_np(const std::vector<T> &v)
{
    py::array_t<T> a(v.size());
    std::copy(v.begin(), v.end(), a.mutable_data());
    return a;
}

class PyRilsRols {
    Search s_;
    bool classification_;

public:
    PyRilsRols(bool classification, int max_fit_calls, int max_seconds, double complexity_penalty, int max_complexity,
               double sample_size, bool verbose, int random_state)
        : s_(SearchParams{classification, max_fit_calls, max_seconds, complexity_penalty, max_complexity, sample_size,
                          verbose, random_state}),
          classification_(classification)
    {
    }

    void fit(DArr X, DArr y, int data_cnt, int feat_cnt)
    {
        if (X.size() != (py::ssize_t)data_cnt * feat_cnt) {
            std::ostringstream m;
            m << "Size of X " << X.size() << " is not the same as the product of data count and feature count "
              << (long long)data_cnt * feat_cnt;
            throw py::value_error(m.str());
        }
        if (y.size() != data_cnt) {
            std::ostringstream m;
            m << "Size of y " << y.size() << " is not the same as the data count " << data_cnt;
            throw py::value_error(m.str());
        }
        const double *px = X.data(), *pyv = y.data();
        py::gil_scoped_release nogil;
        s_.fit(px, pyv, data_cnt, feat_cnt);
    }

    py::array_t<double> predict(DArr X, int data_cnt, int feat_cnt)
    {
        if (X.size() != (py::ssize_t)data_cnt * feat_cnt) throw py::value_error("Size of X does not match data_cnt * feat_cnt");
        py::array_t<double> out(data_cnt);
        s_.predict(X.data(), data_cnt, feat_cnt, out.mutable_data());
        return out;
    }

    py::array_t<double> predict_proba(DArr X, int data_cnt, int feat_cnt)
    {
        if (X.size() != (py::ssize_t)data_cnt * feat_cnt) throw py::value_error("Size of X does not match data_cnt * feat_cnt");
        py::array_t<double> out({(py::ssize_t)data_cnt, (py::ssize_t)2});
        s_.predict_proba(X.data(), data_cnt, feat_cnt, out.mutable_data());
        return out;
    }
    void set_classifier_objective(bool on) { s_.set_classifier_objective(on); }

    std::string get_model_string() { return s_.model_string(); }
    double get_best_time() const { return s_.best_time(); }
    double get_total_time() const { return s_.total_time(); }
    int get_fit_calls() const { return s_.fit_calls(); }

    void set_trace(bool on) { s_.set_trace(on); }
    py::list get_trace() const
    {
        py::list out;
        for (const rrd::TraceBatch &t : s_.trace()) {
            py::dict d;
            d["mode"] = t.mode;
            d["cand_term_begin"] = to_np(t.cand_term_begin);
            d["term_code_begin"] = to_np(t.term_code_begin);
            d["code"] = to_np(t.code);
            d["consts"] = to_np(t.consts);
            d["coef"] = to_np(t.coef);
            d["ssr"] = to_np(t.ssr);
            d["size"] = to_np(t.size);
            d["accepted"] = to_np(t.accepted);
            d["accepted_fit"] = to_np(t.accepted_fit);
            d["consumed"] = to_np(t.consumed);
            d["curr"] = py::make_tuple(t.curr_f0, t.curr_f1, t.curr_size);
            d["fit_calls_before"] = t.fit_calls_before;
            out.append(d);
        }
        return out;
    }
    py::dict get_engine_stats() const
    {
        const rr_stats st = s_.engine_stats();
        py::dict d;
        d["batches"] = st.batches;
        d["candidates"] = st.candidates;
        d["sweep_launches"] = st.sweep_launches;
        d["kernel_launches"] = st.kernel_launches;
        d["refined"] = st.refined;
        d["exact"] = st.exact;
        d["dd"] = st.dd;
        d["nonfinite"] = st.nonfinite;
        d["distinct_terms"] = st.distinct_terms;
        d["term_instances"] = st.term_instances;
        d["ingest_ms"] = st.ingest_ms;
        d["collectives"] = st.collectives;
        return d;
    }
    py::tuple get_model_program() const
    {
        std::vector<uint32_t> code;
        std::vector<double> consts;
        if (!s_.model()) throw std::runtime_error("no model yet");
        rrd::compile_postfix(*s_.model(), code, consts);
        return py::make_tuple(to_np(code), to_np(consts));
    }
};

py::list debug_all_candidates(UArr code, DArr consts, int d, bool classification, bool local_search)
{
    SearchParams p;
    p.classification = classification;
    Search s(p);
    s.setup_nodes_for(d);
    ExprP t = rrd::from_postfix(code.data(), code.size(), consts.data(), consts.size());
    py::list out;
    for (const Expr &c : s.all_candidates(*t, local_search)) {
        std::vector<uint32_t> cc;
        std::vector<double> kk;
        rrd::compile_postfix(c, cc, kk);
        out.append(py::make_tuple(to_np(cc), to_np(kk), rrd::to_string(c)));
    }
    return out;
}

// the OLS_FIT rr_batch the driver submits for a list of candidate trees (expand, simplify, factor
// selection: rils_rols_cpp.cpp:448-475)
py::dict debug_term_batch(py::list trees)
{
    std::vector<int32_t> ctb{0}, tcb{0};
    std::vector<uint32_t> code;
    std::vector<double> consts;
    for (auto item : trees) {
        auto tup = item.cast<py::tuple>();
        auto c = tup[0].cast<UArr>();
        auto k = tup[1].cast<DArr>();
        ExprP t = rrd::from_postfix(c.data(), c.size(), k.data(), k.size());
        rrd::expand(*t);
        rrd::simplify(*t);
        for (const Expr *f : rrd::select_factors(*t)) {
            rrd::compile_postfix(*f, code, consts);
            tcb.push_back((int32_t)code.size());
        }
        ctb.push_back((int32_t)tcb.size() - 1);
    }
    py::dict d;
    d["mode"] = (int)RR_MODE_OLS_FIT;
    d["cand_term_begin"] = to_np(ctb);
    d["term_code_begin"] = to_np(tcb);
    d["code"] = to_np(code);
    d["consts"] = to_np(consts);
    return d;
}

// the row order fit() works on: selected[0 .. count) of std::shuffle(iota(n), default_random_engine(seed)),
// rils_rols_cpp.cpp:774-795 (libstdc++'s shuffle: tests that feed the oracle the same rows need it)
py::array_t<int32_t> debug_shuffle_index(int n, int seed, int count)
{
    std::vector<int> selected(n);
    std::iota(selected.begin(), selected.end(), 0);
    std::shuffle(selected.begin(), selected.end(), std::default_random_engine(seed));
    py::array_t<int32_t> out(std::max(0, std::min(count, n)));
    std::copy(selected.begin(), selected.begin() + out.size(), out.mutable_data());
    return out;
}

std::string debug_to_string(UArr code, DArr consts)
{
    return rrd::to_string(*rrd::from_postfix(code.data(), code.size(), consts.data(), consts.size()));
}

py::tuple debug_rebuild(UArr code, DArr consts, DArr coef)
{
    ExprP t = rrd::from_postfix(code.data(), code.size(), consts.data(), consts.size());
    rrd::expand(*t);
    rrd::simplify(*t);
    auto f = rrd::select_factors(*t);
    if ((size_t)coef.size() != f.size() + 1) throw py::value_error("coef must have one entry per factor plus the free term");
    ExprP r = rrd::rebuild_from_coefficients(f, coef.data());
    return py::make_tuple(rrd::to_string(*r), rrd::size_of(*r));
}

}  // namespace

PYBIND11_MODULE(rils_rols_cpp, m)
{
    m.doc() = "RILS-ROLS driver on the B200 scoring engine (drop-in for the reference's rils_rols_cpp)";
    py::class_<PyRilsRols>(m, "rils_rols", py::module_local())
        .def(py::init<bool, int, int, double, int, double, bool, int>())
        .def("fit", &PyRilsRols::fit)
        .def("predict", &PyRilsRols::predict)
        .def("get_model_string", &PyRilsRols::get_model_string)
        .def("get_best_time", &PyRilsRols::get_best_time)
        .def("get_fit_calls", &PyRilsRols::get_fit_calls)
        .def("get_total_time", &PyRilsRols::get_total_time)
        // extras (not in the reference)
        .def("predict_proba", &PyRilsRols::predict_proba)
        .def("set_classifier_objective", &PyRilsRols::set_classifier_objective)
        .def("set_trace", &PyRilsRols::set_trace)
        .def("get_trace", &PyRilsRols::get_trace)
        .def("get_engine_stats", &PyRilsRols::get_engine_stats)
        .def("get_model_program", &PyRilsRols::get_model_program);
    m.def("debug_all_candidates", &debug_all_candidates);
    m.def("debug_term_batch", &debug_term_batch);
    m.def("debug_to_string", &debug_to_string);
    m.def("debug_shuffle_index", &debug_shuffle_index);
    m.def("debug_rebuild", &debug_rebuild);
}
