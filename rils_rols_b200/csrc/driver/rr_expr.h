// rr_expr.h — host-side expression trees of the ILS driver.
//
// A from-scratch value-semantics implementation (unique ownership, free functions) of the tree
// algebra the reference keeps in class `node` (/root/reference/rils_rols_cpp/node.h, node.cpp).
// Which candidates exist, and in which order, is defined by these routines and by their string
// form (the dedupe key everywhere), so every rule — including the quirks SURVEY.md App. C lists —
// is reproduced exactly; each function cites the lines it mirrors. Nothing numeric happens here:
// evaluation is compiled to postfix bytecode (compile_postfix) and runs on the GPU.
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/rr_b200.h"

namespace rrd {

// same enumerator values as enum class node_type (node.h:16-38) == enum rr_opcode
enum class Op : uint8_t {
    NONE = RR_OP_NONE, CONST = RR_OP_CONST, VAR = RR_OP_VAR, PLUS = RR_OP_PLUS, MINUS = RR_OP_MINUS,
    MULTIPLY = RR_OP_MULTIPLY, DIVIDE = RR_OP_DIVIDE, SIN = RR_OP_SIN, COS = RR_OP_COS, LN = RR_OP_LN,
    EXP = RR_OP_EXP, SQRT = RR_OP_SQRT, SQR = RR_OP_SQR, POW = RR_OP_POW, LESS_THAN = RR_OP_LESS_THAN,
    GREATER_THAN = RR_OP_GREATER_THAN, EQUAL = RR_OP_EQUAL, NOT_EQUAL = RR_OP_NOT_EQUAL, MIN = RR_OP_MIN,
    MAX = RR_OP_MAX
};

int arity_of(Op t);       // node.h:40-56
bool symmetric_of(Op t);  // node.h:58-68

struct Expr;
using ExprP = std::unique_ptr<Expr>;

struct Expr {
    Op type = Op::NONE;
    int var = -1;
    double value = 0.0;
    ExprP left, right;

    Expr() = default;
    explicit Expr(double c) : type(Op::CONST), value(c) {}
    static Expr variable(int j)
    {
        Expr e;
        e.type = Op::VAR;
        e.var = j;
        return e;
    }
    Expr(Op t, ExprP l, ExprP r) : type(t), left(std::move(l)), right(std::move(r)) {}
    Expr(const Expr &o);  // deep copy
    Expr &operator=(const Expr &o);
    Expr(Expr &&) noexcept = default;
    Expr &operator=(Expr &&) noexcept = default;

    int arity() const { return arity_of(type); }
    bool is(Op t) const { return type == t; }
};

ExprP clone(const Expr &e);
ExprP make(Op t, const Expr &l);
ExprP make(Op t, const Expr &l, const Expr &r);

int size_of(const Expr &e);             // node.h:311-322
std::string to_string(const Expr &e);   // node.h:255-301
bool allowed_left(Op parent, const Expr &child);  // node.cpp:97-118

void simplify(Expr &e);                 // node.cpp:152-296
void expand(Expr &e);                   // node.cpp:329-374
void normalize_factor_constants(Expr &e, Op parent, bool inside_factor);  // node.cpp:312-327

// additive terms of a tree that went through expand(); simplify() — the `factors` of
// tune_constants(), rils_rols_cpp.cpp:450-473 (pointers into `e`)
std::vector<const Expr *> select_factors(const Expr &e);

// breadth-first list of all subtrees, node.h:351-377
void all_subtrees(const Expr &root, std::vector<const Expr *> &out);

// postfix bytecode of include/rr_b200.h; constants are appended to `consts`
void compile_postfix(const Expr &e, std::vector<uint32_t> &code, std::vector<double> &consts);
ExprP from_postfix(const uint32_t *code, size_t len, const double *consts, size_t n_consts);

inline bool value_zero(double v) { return std::fabs(v) < 1e-12; }      // node.h:333-335, EPS = 10^-12
inline bool value_one(double v) { return std::fabs(v - 1) < 1e-12; }   // node.h:337-339

}  // namespace rrd
