// rr_expr.cpp — see rr_expr.h. Paths in comments are relative to /root/reference/rils_rols_cpp.
#include "rr_expr.h"

#include <algorithm>
#include <charconv>
#include <cstdio>

#include <cmath>
#include <stdexcept>

namespace rrd {

namespace {
const double kEps = std::pow(10, -12);  // node.h:13-14
}

int arity_of(Op t)
{
    switch (t) {
    case Op::CONST:
    case Op::VAR:
        return 0;
    case Op::SIN:
    case Op::COS:
    case Op::LN:
    case Op::EXP:
    case Op::SQRT:
    case Op::SQR:
        return 1;
    default:
        return 2;
    }
}

bool symmetric_of(Op t) { return !(t == Op::MINUS || t == Op::DIVIDE || t == Op::POW); }

Expr::Expr(const Expr &o) : type(o.type), var(o.var), value(o.value)
{
    if (o.left) left = clone(*o.left);
    if (o.right) right = clone(*o.right);
}

Expr &Expr::operator=(const Expr &o)
{
    if (this != &o) {
        Expr tmp(o);
        *this = std::move(tmp);
    }
    return *this;
}

ExprP clone(const Expr &e) { return std::make_unique<Expr>(e); }
ExprP make(Op t, const Expr &l) { return std::make_unique<Expr>(t, clone(l), nullptr); }
ExprP make(Op t, const Expr &l, const Expr &r) { return std::make_unique<Expr>(t, clone(l), clone(r)); }

int size_of(const Expr &e)
{
    const int a = e.arity();
    if (a == 0) return 1;
    if (a == 1) return 1 + size_of(*e.left);
    return 1 + size_of(*e.left) + size_of(*e.right);
}

namespace {

void append_int(std::string &s, long long v)
{
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);
    s.append(buf, r.ptr);
}

// std::to_string(double) == printf("%f"): 6 decimals, correctly rounded from the exact binary value. The key of
// every dedupe set is built from it, so the fast path below must give the same characters: |v| < 10^6 is
// scaled by 10^6 in fp64 (error < 1.2e-4 of a unit in the last printed digit) and rounded; when the scaled
// value is within 1e-3 of a rounding boundary - where that error could matter, ties included - and for
// everything else (large, nan, inf) snprintf decides.
void append_fixed6(std::string &s, double v)
{
    const double a = std::fabs(v);
    if (a < 1e6) {  // false for nan
        const double scaled = a * 1e6;
        const double fl = std::floor(scaled);
        const double frac = scaled - fl;  // exact
        if (std::fabs(frac - 0.5) > 1e-3) {
            const unsigned long long q = (unsigned long long)fl + (frac > 0.5 ? 1u : 0u);
            if (std::signbit(v)) s.push_back('-');
            append_int(s, (long long)(q / 1000000u));
            char d[7];
            unsigned r = (unsigned)(q % 1000000u);
            for (int i = 5; i >= 0; --i) {
                d[i] = (char)('0' + r % 10);
                r /= 10;
            }
            s.push_back('.');
            s.append(d, 6);
            return;
        }
    }
    char buf[400];
    const int n = std::snprintf(buf, sizeof(buf), "%f", v);
    s.append(buf, (size_t)std::max(0, std::min(n, (int)sizeof(buf) - 1)));
}

void append_string(const Expr &e, std::string &s)
{
    auto bin = [&](const char *open, const char *mid, const char *close) {
        s += open;
        append_string(*e.left, s);
        s += mid;
        append_string(*e.right, s);
        s += close;
    };
    auto un = [&](const char *open, const char *close) {
        s += open;
        append_string(*e.left, s);
        s += close;
    };
    switch (e.type) {
    case Op::CONST:
        // near-integers print as ints, everything else with std::to_string's 6 decimals (node.h:257-261)
        if (std::abs(std::round(e.value) - e.value) < kEps) append_int(s, (int)(std::round(e.value)));
        else append_fixed6(s, e.value);
        return;
    case Op::VAR:
        s.push_back('x');
        append_int(s, e.var);
        return;
    case Op::PLUS: return bin("(", "+", ")");
    case Op::MINUS: return bin("(", "-", ")");
    case Op::MULTIPLY: return bin("(", "*", ")");
    case Op::DIVIDE: return bin("(", "/", ")");
    case Op::SIN: return un("sin(", ")");
    case Op::COS: return un("cos(", ")");
    case Op::LN: return un("ln(", ")");
    case Op::EXP: return un("exp(", ")");
    case Op::SQRT: return un("sqrt(", ")");
    case Op::SQR: return un("((", ")**2)");
    case Op::POW: return bin("pow(", ",", ")");
    // all four comparisons print as '<' (node.h:286-293): they collide in every string-keyed set
    case Op::LESS_THAN:
    case Op::GREATER_THAN:
    case Op::EQUAL:
    case Op::NOT_EQUAL: return bin("(", "<", ")");
    case Op::MIN: return bin("MIN(", ", ", ")");
    case Op::MAX: return bin("MAX(", ", ", ")");
    default: s += "*****UNKNOWN*****"; return;
    }
}

}  // namespace

// one buffer, appended to in place (the recursive concatenation of temporaries this replaces was 40 % of
// all_candidates)
std::string to_string(const Expr &e)
{
    std::string s;
    s.reserve(160);
    append_string(e, s);
    return s;
}

bool allowed_left(Op parent, const Expr &child)
{
    const Op t = child.type;
    switch (parent) {
    case Op::EXP:
    case Op::LN: return !(t == Op::EXP || t == Op::LN);
    case Op::POW: return t != Op::POW;
    case Op::COS:
    case Op::SIN: return !(t == Op::COS || t == Op::SIN);
    default: return true;
    }
}

namespace {

void become(Expr &dst, ExprP src) { dst = std::move(*src); }  // node::update_with, node.h:130-139
void become_const(Expr &e, double v)                            // node::set_const_value, node.h:222-228
{
    e.type = Op::CONST;
    e.value = v;
    e.left.reset();
    e.right.reset();
}
bool additive(const Expr &e) { return e.type == Op::PLUS || e.type == Op::MINUS; }

}  // namespace

void simplify(Expr &e)
{
    const int ar = e.arity();
    if (ar == 0) return;
    if (ar == 1) {
        simplify(*e.left);
        return;
    }
    simplify(*e.left);
    simplify(*e.right);
    Expr &L = *e.left, &R = *e.right;
    if (L.is(Op::CONST) && R.is(Op::CONST)) {  // node.cpp:162-196: fold
        const double a = L.value, b = R.value;
        double v;
        switch (e.type) {
        case Op::PLUS: v = a + b; break;
        case Op::MINUS: v = a - b; break;
        case Op::MULTIPLY: v = a * b; break;
        case Op::DIVIDE: v = a / b; break;
        case Op::POW: v = std::pow(a, b); break;
        case Op::LESS_THAN: v = a < b ? 1 : 0; break;
        case Op::GREATER_THAN: v = a > b ? 1 : 0; break;
        case Op::EQUAL: v = a == b ? 1 : 0; break;
        case Op::NOT_EQUAL: v = a != b ? 1 : 0; break;
        case Op::MIN: v = a < b ? a : b; break;
        case Op::MAX: v = a > b ? a : b; break;
        default: throw std::runtime_error("Simplification is not supported for this binary operator!");
        }
        become_const(e, v);
    } else if (L.is(Op::CONST)) {  // node.cpp:197-249
        if (additive(e)) {
            if (value_zero(L.value)) {
                become(e, std::move(e.right));  // 0+t = t and (sic) 0-t = t
            } else if (additive(R)) {
                if (R.left->is(Op::CONST)) {  // c1 +- (c2 +- t): the inner operator is dropped (sic)
                    if (e.type == Op::PLUS) L.value += R.left->value;
                    else L.value -= R.left->value;
                    become(R, std::move(R.right));
                } else if (R.right->is(Op::CONST)) {  // c1 +- (t +- c2)
                    if (e.type == Op::PLUS) {
                        if (R.type == Op::PLUS) L.value += R.right->value;
                        else L.value -= R.right->value;
                    } else {
                        if (R.type == Op::PLUS) L.value -= R.right->value;
                        else L.value += R.right->value;
                    }
                    become(R, std::move(R.left));
                }
            }
        } else if (e.type == Op::MULTIPLY) {
            if (value_zero(L.value)) become_const(e, 0.0);
            else if (value_one(L.value)) become(e, std::move(e.right));
            else if (R.is(Op::MULTIPLY)) {
                if (R.left->is(Op::CONST)) {  // c1*(c2*t)
                    L.value *= R.left->value;
                    become(R, std::move(R.right));
                } else if (R.right->is(Op::CONST)) {  // c1*(t*c2)
                    L.value *= R.right->value;
                    become(R, std::move(R.left));
                }
            }
        } else if (e.type == Op::DIVIDE && value_zero(L.value)) {
            become_const(e, 0.0);
        }
    } else if (R.is(Op::CONST)) {  // node.cpp:250-294
        if (additive(e)) {
            if (value_zero(R.value)) {
                become(e, std::move(e.left));
            } else if (additive(L)) {
                if (L.left->is(Op::CONST)) {  // (c1 +- t) +- c2 -> (c3 +- t) +- 0, the 0 goes next round
                    if (e.type == Op::PLUS) L.left->value += R.value;
                    else L.left->value -= R.value;
                    R.value = 0;
                } else if (L.right->is(Op::CONST)) {  // (t +- c1) +- c2 -> (t +- 0) +- c3
                    if (L.type == Op::PLUS) R.value += L.right->value;
                    else R.value -= L.right->value;
                    L.right->value = 0;
                }
            }
        } else if (e.type == Op::MULTIPLY) {
            if (value_one(R.value)) become(e, std::move(e.left));
            else if (value_zero(R.value)) become_const(e, 0.0);
            else if (L.is(Op::MULTIPLY)) {
                if (L.left->is(Op::CONST)) {  // (c1*t)*c2
                    R.value *= L.left->value;
                    become(L, std::move(L.right));
                } else if (L.right->is(Op::CONST)) {  // (t*c1)*c2
                    R.value *= L.right->value;
                    become(L, std::move(L.left));
                }
            }
        }
    }
}

void expand(Expr &e)
{
    const int ar = e.arity();
    if (ar == 0) return;
    if (ar == 1) {
        expand(*e.left);
        return;
    }
    expand(*e.left);
    expand(*e.right);
    if (e.type != Op::MULTIPLY) return;
    if (additive(*e.left)) {
        if (additive(*e.right)) {
            // (t1+-t2)*(t3+-t4), binomial_mult node.cpp:329-338
            const Expr &l = *e.left, &r = *e.right;
            ExprP f1 = make(Op::MULTIPLY, *l.left, *r.left), f2 = make(Op::MULTIPLY, *l.left, *r.right);
            ExprP f3 = make(Op::MULTIPLY, *l.right, *r.left), f4 = make(Op::MULTIPLY, *l.right, *r.right);
            auto nl = std::make_unique<Expr>(r.type, std::move(f1), std::move(f2));
            auto nr = std::make_unique<Expr>(r.type, std::move(f3), std::move(f4));
            Expr res(l.type, std::move(nl), std::move(nr));
            e = std::move(res);
        } else {
            // node.cpp:359-364. The reference overwrites `left` before reading it again, so
            // (t1+-t2)*t3 becomes (t1*t3)*(t3*t3) with type MULTIPLY; preserved.
            ExprP nl = make(Op::MULTIPLY, *e.left->left, *e.right);
            e.left = std::move(nl);
            ExprP nr = make(Op::MULTIPLY, *e.left->right, *e.right);
            e.right = std::move(nr);
            e.type = e.left->type;
        }
    } else if (additive(*e.right)) {
        // node.cpp:366-371, same pattern: t1*(t2+-t3) becomes (t1*t2)*((t1*t2)*t3)
        ExprP nl = make(Op::MULTIPLY, *e.left, *e.right->left);
        e.left = std::move(nl);
        ExprP nr = make(Op::MULTIPLY, *e.left, *e.right->right);
        e.right = std::move(nr);
        e.type = e.right->type;
    }
}

void normalize_factor_constants(Expr &e, Op parent, bool inside_factor)
{
    (void)parent;
    if (e.type == Op::CONST) {
        e.value = 1;
    } else if (!inside_factor && additive(e)) {
        normalize_factor_constants(*e.left, e.type, false);
        normalize_factor_constants(*e.right, e.type, false);
    } else if (!inside_factor) {
        if (e.type == Op::MULTIPLY) {
            normalize_factor_constants(*e.left, e.type, true);
            normalize_factor_constants(*e.right, e.type, true);
        } else if (e.type == Op::DIVIDE) {
            normalize_factor_constants(*e.right, e.type, true);
        }
    }
}

namespace {
void non_constant_factors(const Expr &e, std::vector<const Expr *> &out)  // node.cpp:140-147
{
    if (additive(e)) {
        non_constant_factors(*e.left, out);
        non_constant_factors(*e.right, out);
    } else if (e.type != Op::CONST) {
        out.push_back(&e);
    }
}
}  // namespace

std::vector<const Expr *> select_factors(const Expr &e)
{
    std::vector<const Expr *> all, out;
    non_constant_factors(e, all);
    for (const Expr *f : all) {
        if (f->is(Op::CONST)) continue;
        if (f->arity() == 2 && f->left->is(Op::CONST) && f->right->is(Op::CONST)) continue;
        if (f->is(Op::MULTIPLY) || f->is(Op::PLUS) || f->is(Op::MINUS)) {
            // exactly one operand is a constant: it goes to the coefficient / free term
            if (f->left->is(Op::CONST)) { out.push_back(f->right.get()); continue; }
            if (f->right->is(Op::CONST)) { out.push_back(f->left.get()); continue; }
        }
        if (f->is(Op::DIVIDE) && f->right->is(Op::CONST)) { out.push_back(f->left.get()); continue; }
        out.push_back(f);
    }
    return out;
}

void all_subtrees(const Expr &root, std::vector<const Expr *> &out)
{
    size_t pos = out.size();
    out.push_back(&root);
    while (pos < out.size()) {
        const Expr *cur = out[pos];
        if (cur->left) out.push_back(cur->left.get());
        if (cur->right) out.push_back(cur->right.get());
        ++pos;
    }
}

void compile_postfix(const Expr &e, std::vector<uint32_t> &code, std::vector<double> &consts)
{
    if (e.arity() >= 1) compile_postfix(*e.left, code, consts);
    if (e.arity() >= 2) compile_postfix(*e.right, code, consts);
    if (e.is(Op::CONST)) {
        // the argument field of a code word has 24 bits (include/rr_b200.h): a batch with more constants must be split
        if (consts.size() >= (1u << 24)) throw std::length_error("more than 2^24 constants in one batch");
        code.push_back(RR_INS(RR_OP_CONST, consts.size()));
        consts.push_back(e.value);
    } else if (e.is(Op::VAR)) {
        if (e.var < 0 || e.var >= (1 << 24)) throw std::length_error("feature index does not fit the 24-bit argument field");
        code.push_back(RR_INS(RR_OP_VAR, e.var));
    } else {
        code.push_back(RR_INS((uint32_t)e.type, 0));
    }
}

ExprP from_postfix(const uint32_t *code, size_t len, const double *consts, size_t n_consts)
{
    std::vector<ExprP> st;
    for (size_t i = 0; i < len; ++i) {
        const uint32_t op = RR_INS_OP(code[i]), arg = RR_INS_ARG(code[i]);
        if (op == RR_OP_CONST) {
            if (arg >= n_consts) throw std::runtime_error("constant index out of range");
            st.push_back(std::make_unique<Expr>(consts[arg]));
        } else if (op == RR_OP_VAR) {
            st.push_back(std::make_unique<Expr>(Expr::variable((int)arg)));
        } else if (op > RR_OP_VAR && op < RR_OP_COUNT) {
            const Op t = (Op)op;
            const int ar = arity_of(t);
            if ((int)st.size() < ar) throw std::runtime_error("malformed postfix");
            ExprP r, l;
            if (ar == 2) { r = std::move(st.back()); st.pop_back(); }
            l = std::move(st.back());
            st.pop_back();
            st.push_back(std::make_unique<Expr>(t, std::move(l), std::move(r)));
        } else {
            throw std::runtime_error("bad opcode");
        }
    }
    if (st.size() != 1) throw std::runtime_error("postfix does not reduce to one tree");
    return std::move(st[0]);
}

}  // namespace rrd
