// rr_plan.cpp — see rr_plan.h.
#include "rr_plan.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace rr {

namespace {

// SURVEY.md 8(d) contract weights (FP64-pipe thread-instructions per node)
const double kW[RR_OP_COUNT] = {0, 0, 0, 1, 1, 1, 10, 16, 16, 28, 18, 10, 1, 90, 1, 1, 1, 1, 1, 1};

int arity(uint32_t op)  // == get_arity, /root/reference/rils_rols_cpp/node.h:40-56
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

// binary node -> specialised opcode for operand kind (konst) and order (swap: t = src op t)
uint32_t bin_ins(uint32_t op, bool konst, bool swap)
{
    switch (op) {
    case RR_OP_PLUS: return RR_W0(konst ? RI_ADD_C : RI_ADD_M, 0);
    case RR_OP_MULTIPLY: return RR_W0(konst ? RI_MUL_C : RI_MUL_M, 0);
    case RR_OP_MINUS: return RR_W0(swap ? (konst ? RI_RSUB_C : RI_RSUB_M) : (konst ? RI_SUB_C : RI_SUB_M), 0);
    case RR_OP_DIVIDE: return RR_W0(swap ? (konst ? RI_RDIV_C : RI_RDIV_M) : (konst ? RI_DIV_C : RI_DIV_M), 0);
    }
    uint32_t rare = 0;
    switch (op) {
    case RR_OP_POW: rare = RR_POW; break;
    case RR_OP_LESS_THAN: rare = RR_LT; break;
    case RR_OP_GREATER_THAN: rare = RR_GT; break;
    case RR_OP_EQUAL: rare = RR_EQ; break;
    case RR_OP_NOT_EQUAL: rare = RR_NE; break;
    case RR_OP_MIN: rare = RR_MIN; break;
    case RR_OP_MAX: rare = RR_MAX; break;
    }
    return RR_W0(RI_RARE, rare | (konst ? RB_CONST : 0u) | (swap ? RB_SWAP : 0u));
}
bool is_rare(uint32_t op)
{
    return !(op == RR_OP_PLUS || op == RR_OP_MULTIPLY || op == RR_OP_MINUS || op == RR_OP_DIVIDE);
}

uint32_t un_ins(uint32_t op)
{
    switch (op) {
    case RR_OP_SIN: return RI_SIN;
    case RR_OP_COS: return RI_COS;
    case RR_OP_LN: return RI_LN;
    case RR_OP_EXP: return RI_EXP;
    case RR_OP_SQRT: return RI_SQRT;
    case RR_OP_SQR: return RI_SQR;
    }
    return RI_END;
}

// (term, term) -> reduction id: open addressing, linear probing; the planner does a few hundred lookups per
// candidate, which made std::unordered_map the largest item of the plan time
class DotMap {
public:
    explicit DotMap(size_t expect)
    {
        size_t cap = 64;
        while (cap < expect * 2) cap <<= 1;
        keys_.assign(cap, kEmpty);
        vals_.assign(cap, 0);
        mask_ = cap - 1;
    }
    const int32_t *find(uint64_t k) const
    {
        for (size_t i = hash(k) & mask_;; i = (i + 1) & mask_) {
            if (keys_[i] == k) return &vals_[i];
            if (keys_[i] == kEmpty) return nullptr;
        }
    }
    void set(uint64_t k, int32_t v)
    {
        if ((n_ + 1) * 2 > keys_.size()) grow();
        for (size_t i = hash(k) & mask_;; i = (i + 1) & mask_) {
            if (keys_[i] == k) { vals_[i] = v; return; }
            if (keys_[i] == kEmpty) { keys_[i] = k; vals_[i] = v; ++n_; return; }
        }
    }

private:
    static constexpr uint64_t kEmpty = 0x7fffffff7fffffffull;  // not a key: term ids are small, sentinels are -1, -2
    static size_t hash(uint64_t k)
    {
        k ^= k >> 33;
        k *= 0xff51afd7ed558ccdull;
        k ^= k >> 33;
        return (size_t)k;
    }
    void grow()
    {
        std::vector<uint64_t> ok;
        std::vector<int32_t> ov;
        ok.swap(keys_);
        ov.swap(vals_);
        keys_.assign(ok.size() * 2, kEmpty);
        vals_.assign(ok.size() * 2, 0);
        mask_ = keys_.size() - 1;
        n_ = 0;
        for (size_t i = 0; i < ok.size(); ++i)
            if (ok[i] != kEmpty) set(ok[i], ov[i]);
    }
    std::vector<uint64_t> keys_;
    std::vector<int32_t> vals_;
    size_t mask_ = 0, n_ = 0;
};

// Postfix code ranges of a batch as hash-table keys (terms and sub-expressions that are equal as programs:
// same operators, same variable indices, bit-identical constants). Open addressing on a 64-bit hash of the
// range; a hit is verified word by word against the stored representative, so equal hashes never merge
// different programs.
class CodeTable {
public:
    CodeTable(const rr_batch *b, size_t expect) : b_(b)
    {
        size_t cap = 64;
        while (cap < expect * 2) cap <<= 1;
        slot_.assign(cap, Slot{0, -1, 0, 0});
        mask_ = cap - 1;
    }
    uint64_t hash_range(int32_t c0, int32_t c1) const
    {
        uint64_t h = 0x9e3779b97f4a7c15ull ^ (uint64_t)(c1 - c0);
        for (int32_t i = c0; i < c1; ++i) {
            const uint32_t w = b_->code[i];
            const uint32_t op = RR_INS_OP(w);
            uint64_t v = op;
            if (op == RR_OP_CONST) {
                uint64_t bits;
                std::memcpy(&bits, &b_->consts[RR_INS_ARG(w)], 8);
                v ^= bits * 0xff51afd7ed558ccdull;
            } else if (op == RR_OP_VAR) {
                v ^= (uint64_t)RR_INS_ARG(w) << 8;
            }
            h = (h ^ v) * 0xc4ceb9fe1a85ec53ull;
            h ^= h >> 29;
        }
        return h;
    }
    // id of the stored range equal to [c0, c1), or -1
    int32_t find(int32_t c0, int32_t c1) const
    {
        const uint64_t h = hash_range(c0, c1);
        for (size_t i = (size_t)h & mask_;; i = (i + 1) & mask_) {
            const Slot &e = slot_[i];
            if (e.id < 0) return -1;
            if (e.h == h && equal(e.c0, e.c1, c0, c1)) return e.id;
        }
    }
    // id of the stored range equal to [c0, c1); stores it under new_id when there is none (returns new_id)
    int32_t find_or_insert(int32_t c0, int32_t c1, int32_t new_id)
    {
        if ((n_ + 1) * 2 > slot_.size()) grow();
        const uint64_t h = hash_range(c0, c1);
        for (size_t i = (size_t)h & mask_;; i = (i + 1) & mask_) {
            Slot &e = slot_[i];
            if (e.id < 0) {
                e = Slot{h, new_id, c0, c1};
                ++n_;
                return new_id;
            }
            if (e.h == h && equal(e.c0, e.c1, c0, c1)) return e.id;
        }
    }

private:
    struct Slot {
        uint64_t h;
        int32_t id, c0, c1;
    };
    bool equal(int32_t a0, int32_t a1, int32_t b0, int32_t b1) const
    {
        if (a1 - a0 != b1 - b0) return false;
        for (int32_t i = 0; i < a1 - a0; ++i) {
            const uint32_t wa = b_->code[a0 + i], wb = b_->code[b0 + i];
            const uint32_t op = RR_INS_OP(wa);
            if (op != RR_INS_OP(wb)) return false;
            if (op == RR_OP_CONST) {
                if (std::memcmp(&b_->consts[RR_INS_ARG(wa)], &b_->consts[RR_INS_ARG(wb)], 8) != 0) return false;
            } else if (op == RR_OP_VAR) {
                if (RR_INS_ARG(wa) != RR_INS_ARG(wb)) return false;
            }
        }
        return true;
    }
    void grow()
    {
        std::vector<Slot> old;
        old.swap(slot_);
        slot_.assign(old.size() * 2, Slot{0, -1, 0, 0});
        mask_ = slot_.size() - 1;
        for (const Slot &e : old)
            if (e.id >= 0) {
                size_t i = (size_t)e.h & mask_;
                while (slot_[i].id >= 0) i = (i + 1) & mask_;
                slot_[i] = e;
            }
    }
    const rr_batch *b_;
    std::vector<Slot> slot_;
    size_t mask_ = 0, n_ = 0;
};

// pre-patch value locations: a tile slot index, a staged column (STAGED | index) or a pin (PINREF | j)
constexpr uint32_t STAGED = 0x8000u;
constexpr uint32_t PINREF = 0x4000u;
constexpr int32_t PIN_RESERVED = -3;  // pin owner: held for the whole chunk (the centred target)

}  // namespace

BatchPlanner::BatchPlanner(const rr_batch *b, int32_t d) : b_(b), d_(d) {}

std::string BatchPlanner::build_term(int32_t code_begin, int32_t code_len, Term &t) const
{
    t.code_begin = code_begin;
    t.code_len = code_len;
    t.nodes.clear();
    t.nodes.reserve(code_len);
    t.w = 0.0;
    std::vector<int32_t> st;
    for (int32_t i = 0; i < code_len; ++i) {
        const uint32_t w = b_->code[code_begin + i];
        const uint32_t op = RR_INS_OP(w), arg = RR_INS_ARG(w);
        if (op == RR_OP_NONE || op >= RR_OP_COUNT) return "bad opcode";
        TermNode n;
        n.op = (uint8_t)op;
        n.first = i;
        const int ar = arity(op);
        if ((int)st.size() < ar) return "malformed postfix (stack underflow)";
        if (op == RR_OP_CONST) {
            if ((int32_t)arg >= b_->n_consts) return "constant index out of range";
            n.cval = b_->consts[arg];
        } else if (op == RR_OP_VAR) {
            if ((int32_t)arg >= d_) return "feature index out of range";
            n.var = (int32_t)arg;
        } else {
            if (ar == 2) { n.right = st.back(); st.pop_back(); }
            n.left = st.back();
            st.pop_back();
            n.first = t.nodes[n.left].first;
            const TermNode &L = t.nodes[n.left];
            if (ar == 1) {
                n.need = L.need;
            } else {
                const TermNode &R = t.nodes[n.right];
                if (R.leaf()) n.need = L.need;
                else if (L.leaf()) n.need = R.need;
                else n.need = L.need == R.need ? L.need + 1 : std::max(L.need, R.need);
            }
        }
        t.w += kW[op];
        st.push_back((int32_t)t.nodes.size());
        t.nodes.push_back(n);
    }
    if (st.size() != 1) return "postfix does not reduce to one tree";
    t.is_const_one = code_len == 1 && t.nodes[0].op == RR_OP_CONST && t.nodes[0].cval == 1.0;
    {
        // exact_const (rr_plan.h): bottom-up; two subtrees are identical when their postfix ranges are (constants by bits)
        std::vector<char> kc(t.nodes.size(), 0), kz(t.nodes.size(), 0);  // constant / zero at every sample
        auto same = [&](int32_t a, int32_t b) {
            const int32_t la = a - t.nodes[a].first + 1, lb = b - t.nodes[b].first + 1;
            if (la != lb) return false;
            for (int32_t i = 0; i < la; ++i) {
                const TermNode &x = t.nodes[t.nodes[a].first + i], &y = t.nodes[t.nodes[b].first + i];
                if (x.op != y.op || x.var != y.var || std::memcmp(&x.cval, &y.cval, 8) != 0) return false;
            }
            return true;
        };
        for (size_t x = 0; x < t.nodes.size(); ++x) {
            const TermNode &n = t.nodes[x];
            if (n.op == RR_OP_CONST) {
                kc[x] = 1;
                kz[x] = n.cval == 0.0;
            } else if (n.op == RR_OP_VAR) {
                kc[x] = 0;
            } else if (n.right < 0) {
                kc[x] = kc[n.left];
                kz[x] = kz[n.left] && (n.op == RR_OP_SQRT || n.op == RR_OP_SQR || n.op == RR_OP_SIN);
            } else {
                // zero columns: S - S, 0 * S, 0 / S (where the other side is not finite the result is NaN at that
                // sample, the reductions see it and the candidate gets the sentinel before anything is dropped)
                kz[x] = (n.op == RR_OP_MINUS && same(n.left, n.right)) || (n.op == RR_OP_MULTIPLY && (kz[n.left] || kz[n.right])) ||
                        (n.op == RR_OP_DIVIDE && kz[n.left]);
                // (c * S) / S, S / (c * S), (c * S) / (c' * S): constant up to the rounding of the products
                auto core = [&](int32_t y) -> int32_t {
                    const TermNode &q = t.nodes[y];
                    if (q.op == RR_OP_MULTIPLY && t.nodes[q.left].op == RR_OP_CONST) return q.right;
                    if ((q.op == RR_OP_MULTIPLY || q.op == RR_OP_DIVIDE) && t.nodes[q.right].op == RR_OP_CONST) return q.left;
                    return y;
                };
                kc[x] = kz[x] || (kc[n.left] && kc[n.right]) || ((n.op == RR_OP_DIVIDE || n.op == RR_OP_MINUS) && same(n.left, n.right)) ||
                        (n.op == RR_OP_DIVIDE && same(core(n.left), core(n.right)));
            }
        }
        t.exact_const = kc.back() != 0;
    }
    return "";
}

std::string BatchPlanner::analyse(bool no_cse)
{
    if (!b_ || b_->n_cand < 0) return "null batch";
    if (b_->n_cand == 0) return "";
    if (!b_->cand_term_begin || !b_->term_code_begin) return "null batch arrays";
    const int32_t n_terms = b_->cand_term_begin[b_->n_cand];
    if (n_terms < 0) return "negative term count";
    if (n_terms > 0 && !b_->code) return "null batch arrays";  // a batch of term-less candidates has no code
    // (term offsets are absolute positions in code[]: a contiguous run of a larger batch's candidates is a batch too)
    if (b_->cand_term_begin[0] != 0 || b_->term_code_begin[0] < 0) return "offset arrays must start at 0";
    for (int32_t c = 0; c < b_->n_cand; ++c)
        if (b_->cand_term_begin[c + 1] < b_->cand_term_begin[c]) return "candidate offsets must not decrease";
    if (b_->n_consts < 0 || (b_->n_consts > 0 && !b_->consts)) return "bad constant pool";
    if (b_->n_code < 0) return "negative code length";
    // with the code length known (ABI 2) every term's range is checked against it; without it only monotony can be
    if (b_->n_code > 0 && n_terms > 0 && b_->term_code_begin[n_terms] > b_->n_code) return "term offsets exceed the code length";
    term_id_.assign(n_terms, -1);
    terms_.clear();
    CodeTable seen(b_, (size_t)n_terms / 4 + 64);
    for (int32_t t = 0; t < n_terms; ++t) {
        const int32_t c0 = b_->term_code_begin[t], c1 = b_->term_code_begin[t + 1];
        if (c1 <= c0) return "empty term program";
        for (int32_t i = c0; i < c1; ++i) {
            const uint32_t w = b_->code[i];
            if (RR_INS_OP(w) == RR_OP_CONST && (int32_t)RR_INS_ARG(w) >= b_->n_consts) return "constant index out of range";
        }
        const int32_t id = (int32_t)terms_.size();
        if (!no_cse) {
            const int32_t hit = seen.find_or_insert(c0, c1, id);
            if (hit != id) { term_id_[t] = hit; continue; }
        }
        Term tm;
        std::string err = build_term(c0, c1 - c0, tm);
        if (!err.empty()) return err;
        terms_.push_back(std::move(tm));
        term_id_[t] = id;
    }
    // subtree-level sharing: an inner subtree that is itself one of the batch's distinct terms can be
    // read from that term's slot when it is resident (neighbours are built by wrapping or combining
    // existing terms: `t * x3`, `cos(t)`, ...)
    if (!no_cse) {
        for (Term &tm : terms_) {
            const int32_t root = (int32_t)tm.nodes.size() - 1;
            for (int32_t x = 0; x < root; ++x) {
                TermNode &nd = tm.nodes[x];
                if (nd.leaf()) continue;
                const int32_t hit = seen.find(tm.code_begin + nd.first, tm.code_begin + x + 1);
                if (hit >= 0) nd.sub_term = hit;
            }
        }
    }
    // expensive inner subtrees shared by several distinct terms (the neighbours of one tree node keep its
    // sibling subtrees): candidates for the cache registers. Distinct terms are numbered in order of first
    // appearance, which is the order the planner evaluates them in, so the occurrence lists double as
    // next-use information.
    sub_occ_.clear();
    sub_size_.clear();
    if (!no_cse) {
        const double kMinW = 8.0;  // at least a division, a square root or a transcendental inside
        CodeTable sub_seen(b_, terms_.size() * 2 + 64);
        std::vector<std::vector<int32_t>> occ;
        std::vector<int32_t> osize;
        std::vector<double> wsub;
        for (size_t u = 0; u < terms_.size(); ++u) {
            Term &tm = terms_[u];
            const int32_t root = (int32_t)tm.nodes.size() - 1;
            wsub.assign(tm.nodes.size(), 0.0);
            for (int32_t x = 0; x <= root; ++x) {
                TermNode &nd = tm.nodes[x];
                if (nd.leaf()) continue;
                wsub[x] = kW[nd.op] + wsub[nd.left] + (nd.right >= 0 ? wsub[nd.right] : 0.0);
                if (x == root || wsub[x] < kMinW) continue;
                const int32_t id = sub_seen.find_or_insert(tm.code_begin + nd.first, tm.code_begin + x + 1, (int32_t)occ.size());
                if (id == (int32_t)occ.size()) {
                    occ.emplace_back();
                    osize.push_back(x - nd.first + 1);
                }
                nd.sub_id = id;
                if (occ[id].empty() || occ[id].back() != (int32_t)u) occ[id].push_back((int32_t)u);
            }
        }
        // keep the shared ones only
        std::vector<int32_t> remap(occ.size(), -1);
        for (size_t i = 0; i < occ.size(); ++i)
            if (occ[i].size() >= 2) {
                remap[i] = (int32_t)sub_occ_.size();
                sub_occ_.push_back(std::move(occ[i]));
                sub_size_.push_back(osize[i]);
            }
        for (Term &tm : terms_)
            for (TermNode &nd : tm.nodes)
                if (nd.sub_id >= 0) nd.sub_id = remap[nd.sub_id];
    }
    // contract weights, SURVEY.md 8(d): W(c) = sum w(op) + k(k+1)/2 + k + (k + 2)
    cand_w_.assign(b_->n_cand, 0.0);
    w_contract_ = 0.0;
    max_k_ = 0;
    for (int32_t c = 0; c < b_->n_cand; ++c) {
        const int32_t t0 = b_->cand_term_begin[c], t1 = b_->cand_term_begin[c + 1];
        if (t1 < t0) return "cand_term_begin not monotone";
        if (b_->mode == RR_MODE_EVAL_ONLY && t1 - t0 != 1) return "EVAL_ONLY needs exactly one program per candidate";
        double w = 0.0;
        for (int32_t t = t0; t < t1; ++t) w += terms_[term_id_[t]].w;
        if (b_->mode == RR_MODE_OLS_FIT) {
            const double k = t1 - t0 + 1;
            w += k * (k + 1) / 2 + k + k + 2;
            max_k_ = std::max(max_k_, (int32_t)k);
        } else {
            w += 2;
        }
        cand_w_[c] = w;
        w_contract_ += w;
    }
    return "";
}

// ---------------------------------------------------------------------------------------------
// chunk builder
// ---------------------------------------------------------------------------------------------
struct BatchPlanner::Chunk {
    const BatchPlanner &bp;
    SweepPlan &P;
    const PlanLimits &lim;
    int32_t pc_begin, dot_base, col_begin;
    std::unordered_map<int32_t, int32_t> colmap;  // global column -> staged index
    int32_t pool_cap = 0;                          // tile slots available in this chunk
    std::vector<int32_t> slot_term;                // slot -> cached term (-1 free, -2 temp)
    std::vector<uint64_t> slot_stamp;
    std::vector<uint64_t> slot_pin;
    // pins (rr_isa.h): cached terms live there first, tile slots take the overflow and the temporaries
    int32_t n_pins = 0;
    int32_t pin_term[RR_NPIN];
    uint64_t pin_stamp[RR_NPIN], pin_hold[RR_NPIN];
    std::unordered_map<int32_t, uint32_t> term_loc;  // resident term -> slot index or PINREF | j
    // cache registers RR_NPIN.. (rr_isa.h): shared sub-expression held there (-1 free), users in flight
    int32_t n_cache = 0;
    int32_t creg_sub[RR_NREG - RR_NPIN];
    int32_t creg_busy[RR_NREG - RR_NPIN];
    int32_t cur_term = -1;  // distinct term being generated (next-use horizon of the cache)
    uint64_t clock = 1, epoch = 1;
    std::string err;
    // G8 plans (rr_isa.h RI_GRAM8): fresh terms live in tile slots and are reduced eight at a time against the pins
    bool g8 = false;
    struct GRow {
        uint32_t col;      // pre-patch location: slot index or STAGED | index
        uint32_t want;     // bits 0-7 pins, 8 self, 9 one
        bool free_after;   // the slot is a temporary of this row: released by the flush
    };
    std::vector<GRow> rows;          // the open group
    std::vector<char> slot_rowheld;  // slot -> an open row reads it
    uint64_t n_gram_groups = 0, n_gram_rows = 0;

    Chunk(const BatchPlanner &bp_, SweepPlan &P_, const PlanLimits &lim_, const std::vector<int32_t> &cols, int32_t pins)
        : bp(bp_), P(P_), lim(lim_)
    {
        pc_begin = (int32_t)P.ins.size();
        dot_base = P.n_dots;
        col_begin = (int32_t)P.cols.size();
        for (int32_t g : cols) {
            colmap.emplace(g, (int32_t)colmap.size());
            P.cols.push_back(g);
        }
        pool_cap = std::min(lim.tile_cols - (int32_t)cols.size(), lim.max_slots);
        n_pins = std::min<int32_t>(std::max(pins, 0), RR_NPIN);
        n_cache = lim.no_cse ? 0 : std::min<int32_t>(std::max(lim.n_cache, 0), RR_NREG - RR_NPIN);
        for (int r = 0; r < RR_NREG - RR_NPIN; ++r) {
            creg_sub[r] = -1;
            creg_busy[r] = 0;
        }
        for (int j = 0; j < RR_NPIN; ++j) {
            pin_term[j] = -1;
            pin_stamp[j] = pin_hold[j] = 0;
        }
    }

    uint32_t staged(int32_t gcol)
    {
        auto it = colmap.find(gcol);
        if (it == colmap.end()) { err = "internal: column not staged"; return STAGED; }
        return STAGED | (uint32_t)it->second;
    }

    void emit(uint32_t w0, uint32_t w1, double imm, double w)
    {
        RRIns i;
        i.w0 = w0;
        i.w1 = w1;
        i.imm = imm;
        P.ins.push_back(i);
        P.w_issued += w;
    }

    // pin 0 <- engine column gcol (read from global memory, not staged), held until the chunk ends
    uint32_t reserve_pin_global(int32_t gcol)
    {
        if (n_pins < 1) { err = "internal: no pin to reserve"; return PINREF; }
        if (g8) {
            emit(RR_W0(RI_PINBG, 0), (uint32_t)gcol, 0.0, 0);
            pin_term[0] = PIN_RESERVED;
            return PINREF | 0u;
        }
        emit(RI_LDG, (uint32_t)gcol, 0.0, 0);
        emit(RI_PIN0, 0, 0.0, 0);
        pin_term[0] = PIN_RESERVED;
        return PINREF | 0u;
    }
    // G8 plans: the last pin <- a column of ones, held until the chunk ends. A row's sum(t) is then one more of the
    // products RI_GRAM8 forms anyway (t . pin), and the kernel keeps no running sum of its own.
    void reserve_pin_ones()
    {
        emit(RI_LOAD_C, 0, 1.0, 0);
        const int32_t sl = alloc_slot(-2);
        if (!err.empty()) return;
        emit(RI_ST, (uint32_t)sl, 0.0, 0);
        emit(RI_PINB0 + (uint32_t)(RR_NPIN - 1), (uint32_t)sl, 0.0, 0);
        free_slot(sl);
        pin_term[RR_NPIN - 1] = PIN_RESERVED;
    }
    int32_t free_pins() const
    {
        int32_t f = 0;
        for (int j = 0; j < n_pins; ++j)
            if (pin_term[j] != PIN_RESERVED) ++f;
        return f;
    }

    // a free tile slot, growing the pool or evicting the least recently used unpinned cached term
    int32_t alloc_slot(int32_t owner)
    {
        int32_t s = -1;
        for (size_t i = 0; i < slot_term.size(); ++i)
            if (slot_term[i] == -1) { s = (int32_t)i; break; }
        if (s < 0 && (int32_t)slot_term.size() < pool_cap) {
            s = (int32_t)slot_term.size();
            slot_term.push_back(-1);
            slot_stamp.push_back(0);
            slot_pin.push_back(0);
            slot_rowheld.push_back(0);
        }
        if (s < 0) {
            auto victim = [&]() {
                int32_t v = -1;
                uint64_t best = ~0ull;
                for (size_t i = 0; i < slot_term.size(); ++i)
                    if (slot_term[i] >= 0 && slot_pin[i] != epoch && !slot_rowheld[i] && slot_stamp[i] < best) {
                        best = slot_stamp[i];
                        v = (int32_t)i;
                    }
                return v;
            };
            s = victim();
            if (s < 0 && g8 && !rows.empty()) {
                // every slot is read by a row of the open group (or holds an operand of this unit): reduce the
                // group now, which releases its temporaries
                flush_rows();
                for (size_t i = 0; i < slot_term.size() && s < 0; ++i)
                    if (slot_term[i] == -1) s = (int32_t)i;
                if (s < 0) s = victim();
                if (s >= 0 && slot_term[s] == -1) {
                    slot_term[s] = owner;
                    slot_stamp[s] = clock++;
                    slot_pin[s] = epoch;
                    if (owner >= 0) term_loc[owner] = (uint32_t)s;
                    return s;
                }
            }
            if (s < 0) { err = "tile slots exhausted"; return 0; }
            term_loc.erase(slot_term[s]);
        }
        slot_term[s] = owner;
        slot_stamp[s] = clock++;
        slot_pin[s] = epoch;
        if (owner >= 0) term_loc[owner] = (uint32_t)s;
        return s;
    }
    void free_slot(int32_t s)
    {
        if (slot_term[s] >= 0) term_loc.erase(slot_term[s]);
        slot_term[s] = -1;
    }
    // a home for cached term `owner`: a free pin, else the least recently used pin that the current
    // unit does not hold, else a tile slot
    uint32_t alloc_loc(int32_t owner)
    {
        if (g8) return (uint32_t)alloc_slot(owner);  // pins are filled explicitly (pin_partner): they are reduction partners
        int j = -1;
        for (int i = 0; i < n_pins; ++i)
            if (pin_term[i] == -1) { j = i; break; }
        if (j < 0) {
            uint64_t best = ~0ull;
            for (int i = 0; i < n_pins; ++i)
                if (pin_term[i] >= 0 && pin_hold[i] != epoch && pin_stamp[i] < best) {
                    best = pin_stamp[i];
                    j = i;
                }
            if (j >= 0) term_loc.erase(pin_term[j]);
        }
        if (j >= 0) {
            pin_term[j] = owner;
            pin_stamp[j] = clock++;
            pin_hold[j] = epoch;
            term_loc[owner] = PINREF | (uint32_t)j;
            return PINREF | (uint32_t)j;
        }
        return (uint32_t)alloc_slot(owner);
    }
    void touch(uint32_t loc)
    {
        if (loc & PINREF) {
            pin_stamp[loc & 0xff] = clock++;
            pin_hold[loc & 0xff] = epoch;
        } else {
            slot_stamp[loc] = clock++;
            slot_pin[loc] = epoch;
        }
    }
    // location of a resident term or -1
    int64_t lookup(int32_t term)
    {
        auto it = term_loc.find(term);
        if (it == term_loc.end()) return -1;
        touch(it->second);
        return (int64_t)it->second;
    }
    bool resident(int32_t term) const { return term_loc.find(term) != term_loc.end(); }
    void unpin_all() { ++epoch; }

    void emit_store(uint32_t loc)
    {
        if (loc & PINREF) emit(RI_PIN0 + (loc & 0xff), 0, 0.0, 0);
        else emit(RI_ST, loc, 0.0, 0);
    }
    void emit_load(uint32_t loc)
    {
        if (loc & PINREF) emit(RI_LDP0 + (loc & 0xff), 0, 0.0, 0);
        else emit(RI_LOAD_M, loc, 0.0, 0);
    }
    // a tile-column operand form whose operand may sit in a pin: USEP j redirects the operand
    void emit_m(uint32_t w0, uint32_t loc, double imm, double w)
    {
        if (loc & PINREF) {
            emit(RI_USEP0 + (loc & 0xff), 0, 0.0, 0);
            emit(w0, 0, imm, w);
        } else {
            emit(w0, loc, imm, w);
        }
    }

    // a node that can be used as an operand without evaluation: a constant, a staged feature, or an
    // inner subtree equal to a distinct term that is resident right now. For the latter the
    // location is held only until the consuming instruction has been emitted (release()).
    struct Operand {
        bool ok = false, konst = false;
        uint32_t col = 0;
        double imm = 0.0;
        bool held = false;       // resident sub-term
        uint64_t saved_hold = 0;
        int cache_reg = -1;      // shared sub-expression in a cache register
    };
    Operand operand_of(const TermNode &n, bool allow_pin = true)
    {
        Operand o;
        if (n.op == RR_OP_CONST) { o.ok = true; o.konst = true; o.imm = n.cval; return o; }
        if (n.op == RR_OP_VAR) { o.ok = true; o.col = staged(n.var); return o; }
        if (n.sub_term >= 0) {
            auto it = term_loc.find(n.sub_term);
            if (it != term_loc.end() && (allow_pin || !(it->second & PINREF))) {
                o.ok = true;
                o.held = true;
                o.col = it->second;
                if (o.col & PINREF) {
                    o.saved_hold = pin_hold[o.col & 0xff];
                    pin_hold[o.col & 0xff] = epoch;
                    pin_stamp[o.col & 0xff] = clock++;
                } else {
                    o.saved_hold = slot_pin[o.col];
                    slot_pin[o.col] = epoch;
                    slot_stamp[o.col] = clock++;
                }
                return o;
            }
        }
        if (n.sub_id >= 0 && allow_pin) {
            const int r = cached(n.sub_id);
            if (r >= 0) {
                o.ok = true;
                o.cache_reg = r;
                o.col = PINREF | (uint32_t)(RR_NPIN + r);
                ++creg_busy[r];
            }
        }
        return o;
    }
    void release(const Operand &o)
    {
        if (o.cache_reg >= 0) --creg_busy[o.cache_reg];
        if (!o.held) return;
        if (o.col & PINREF) pin_hold[o.col & 0xff] = o.saved_hold;
        else slot_pin[o.col] = o.saved_hold;
    }
    int cached(int32_t sub) const
    {
        for (int r = 0; r < n_cache; ++r)
            if (creg_sub[r] == sub) return r;
        return -1;
    }
    // first distinct term after the one being generated that contains `sub` (INT32_MAX: none)
    int32_t next_use(int32_t sub) const
    {
        const std::vector<int32_t> &occ = bp.sub_occurrences(sub);
        auto it = std::upper_bound(occ.begin(), occ.end(), cur_term);
        return it == occ.end() ? INT32_MAX : *it;
    }
    // t holds shared sub-expression `sub`, just computed: park it in a cache register when it is needed
    // again sooner than what the register holds now (Belady on the planner's own evaluation order)
    void maybe_cache(int32_t sub)
    {
        if (n_cache == 0 || cached(sub) >= 0) return;
        const int32_t nu = next_use(sub);
        if (nu == INT32_MAX) return;
        // victim: farthest next use; among equals the smaller sub-expression (an enclosing one that is
        // needed just as soon makes the enclosed one redundant: the hit happens at the outer node)
        int victim = -1;
        int32_t victim_nu = -1, victim_size = 0;
        for (int r = 0; r < n_cache; ++r) {
            if (creg_busy[r]) continue;
            const int32_t v = creg_sub[r] < 0 ? INT32_MAX : next_use(creg_sub[r]);
            const int32_t sz = creg_sub[r] < 0 ? 0 : bp.sub_size(creg_sub[r]);
            if (v > victim_nu || (v == victim_nu && sz < victim_size)) { victim_nu = v; victim_size = sz; victim = r; }
        }
        if (victim < 0) return;
        if (victim_nu < nu || (victim_nu == nu && victim_size >= bp.sub_size(sub))) return;
        creg_sub[victim] = sub;
        emit(RI_PIN0 + RR_NPIN + victim, 0, 0.0, 0);
    }

    // leaves t = value(node x)
    void gen(const Term &T, int32_t x)
    {
        const TermNode &n = T.nodes[x];
        Operand o = operand_of(n);
        if (o.ok) {
            if (o.konst) emit(RI_LOAD_C, 0, o.imm, 0);
            else emit_load(o.col);
            release(o);
            return;
        }
        gen_compute(T, x);
        if (n.sub_id >= 0) maybe_cache(n.sub_id);
    }
    void gen_compute(const Term &T, int32_t x)
    {
        const TermNode &n = T.nodes[x];
        const int ar = arity(n.op);
        if (ar == 1) {
            gen(T, n.left);
            emit(un_ins(n.op), 0, 0.0, kW[n.op]);
            return;
        }
        const TermNode &L = T.nodes[n.left], &R = T.nodes[n.right];
        const bool pin_ok = !is_rare(n.op);  // RI_RARE reads its operand from the tile only
        Operand ro = operand_of(R, pin_ok);
        if (ro.ok) {
            gen(T, n.left);
            if (ro.konst) emit(bin_ins(n.op, true, false), 0, ro.imm, kW[n.op]);
            else emit_m(bin_ins(n.op, false, false), ro.col, 0.0, kW[n.op]);
            release(ro);
            return;
        }
        Operand lo = operand_of(L, pin_ok);
        if (lo.ok) {
            gen(T, n.right);
            if (lo.konst) emit(bin_ins(n.op, true, true), 0, lo.imm, kW[n.op]);
            else emit_m(bin_ins(n.op, false, true), lo.col, 0.0, kW[n.op]);
            release(lo);
        } else if (L.need >= R.need) {
            gen(T, n.left);
            const int32_t s = alloc_slot(-2);
            emit(RI_ST, (uint32_t)s, 0.0, 0);
            gen(T, n.right);
            emit(bin_ins(n.op, false, true), (uint32_t)s, 0.0, kW[n.op]);  // t = L(slot) op R(t)
            free_slot(s);
        } else {
            gen(T, n.right);
            const int32_t s = alloc_slot(-2);
            emit(RI_ST, (uint32_t)s, 0.0, 0);
            gen(T, n.left);
            emit(bin_ins(n.op, false, false), (uint32_t)s, 0.0, kW[n.op]);  // t = L(t) op R(slot)
            free_slot(s);
        }
    }

    void gen_term(int32_t u)
    {
        const Term &T = bp.term(u);
        cur_term = u;
        gen(T, (int32_t)T.nodes.size() - 1);
        P.n_term_evals++;
    }

    // make term u resident AND leave its value in t; returns its location
    uint32_t ensure_tos(int32_t u)
    {
        int64_t s = lookup(u);
        if (s >= 0) {
            emit_load((uint32_t)s);
            return (uint32_t)s;
        }
        gen_term(u);
        const uint32_t loc = alloc_loc(u);
        emit_store(loc);
        return loc;
    }
    // make term u resident (t is clobbered when it has to be evaluated)
    uint32_t ensure(int32_t u)
    {
        int64_t s = lookup(u);
        if (s >= 0) return (uint32_t)s;
        gen_term(u);
        const uint32_t loc = alloc_loc(u);
        emit_store(loc);
        return loc;
    }

    // Reductions of t against itself / ones / value locations (pre-patch refs: pins, tile slots,
    // staged columns). Appends the output ids in the order self, one, partners... (the caller's
    // partner order); dd outputs take two ids each and accept tile columns only.
    void mdot(bool self, bool one, const std::vector<uint32_t> &partners, bool dd, std::vector<int32_t> &ids)
    {
        const int step = dd ? 2 : 1;
        const double wdot = dd ? 10.0 : 1.0;
        // pinned partners ride in the mask of RI_MDOT (outputs: self, one, pins ascending; at most
        // RR_MDOT_MAX_OUT per instruction); every other partner is one RI_DOTM
        const size_t id0 = ids.size();
        ids.resize(id0 + (self ? 1 : 0) + (one ? 1 : 0) + partners.size(), DOT_NONE);
        const size_t part0 = id0 + (self ? 1 : 0) + (one ? 1 : 0);
        int32_t pin_pos[RR_NPIN];
        for (int j = 0; j < RR_NPIN; ++j) pin_pos[j] = -1;
        for (size_t i = 0; i < partners.size(); ++i)
            if (partners[i] & PINREF) {
                const int j = (int)(partners[i] & 0xff);
                if (pin_pos[j] >= 0) { err = "internal: duplicate pinned partner"; return; }
                pin_pos[j] = (int32_t)i;
            }
        bool want_self = self, want_one = one;
        int j = 0;
        for (;;) {
            uint32_t aux = 0;
            int n_out = 0;
            std::vector<size_t> slots_out;  // positions in ids, emission order
            if (want_self) { aux |= MD_SELF; slots_out.push_back(id0); ++n_out; want_self = false; }
            if (want_one) { aux |= MD_ONE; slots_out.push_back(id0 + (self ? 1 : 0)); ++n_out; want_one = false; }
            for (; j < RR_NPIN && n_out < (int)RR_MDOT_MAX_OUT; ++j)
                if (pin_pos[j] >= 0) {
                    aux |= 1u << (8 + j);
                    slots_out.push_back(part0 + (size_t)pin_pos[j]);
                    ++n_out;
                }
            if (n_out == 0) break;
            emit(RR_W0(dd ? RI_MDOTDD : RI_MDOT, aux), 0, 0.0, wdot * n_out);
            for (size_t q : slots_out) {
                ids[q] = P.n_dots;
                P.n_dots += step;
            }
            P.n_dot_ins += n_out;
            bool more = false;
            for (int jj = j; jj < RR_NPIN; ++jj) more = more || pin_pos[jj] >= 0;
            if (!more) break;
        }
        for (size_t i = 0; i < partners.size(); ++i)
            if (!(partners[i] & PINREF)) {
                emit(dd ? RI_DOTMDD : RI_DOTM, partners[i], 0.0, wdot);
                ids[part0 + i] = P.n_dots;
                P.n_dots += step;
                P.n_dot_ins += 1;
            }
    }
    // ---- G8 ------------------------------------------------------------------------------------------
    // reduce the open group: one RI_GRAM8 + its column slot
    void flush_rows()
    {
        if (rows.empty()) return;
        uint64_t lo = 0, hi = 0;  // 80 wanted bits
        int n_want = 0;
        uint8_t cols8[8];
        for (size_t g = 0; g < 8; ++g) {
            const GRow &r = rows[g < rows.size() ? g : 0];
            // byte encoding of a pre-patch location: bit 7 = staged column, low bits = index (patched in close())
            cols8[g] = (uint8_t)((r.col & STAGED) ? (0x80u | (r.col & 0x7fu)) : (r.col & 0x7fu));
            if (g >= rows.size()) continue;
            for (int o = 0; o < 10; ++o)
                if ((r.want >> o) & 1u) {
                    const int bit = (int)g * 10 + o;
                    if (bit < 64) lo |= 1ull << bit;
                    else hi |= 1ull << (bit - 64);
                    ++n_want;
                }
        }
        RRIns gi;
        gi.w0 = RR_W0(RI_GRAM8, (uint32_t)rows.size());
        gi.w1 = (uint32_t)(lo & 0xffffffffu);
        const uint64_t immbits = (lo >> 32) | (hi << 32);
        std::memcpy(&gi.imm, &immbits, 8);
        P.ins.push_back(gi);
        P.w_issued += n_want;
        RRIns d;
        std::memset(&d, 0, sizeof(d));
        d.w0 = RR_W0(RI_NOP, RR_GRAM_COLS);
        std::memcpy(&d.w1, cols8, 4);
        std::memcpy(&d.imm, cols8 + 4, 4);
        P.ins.push_back(d);
        for (const GRow &r : rows)
            if (!(r.col & STAGED)) {
                slot_rowheld[r.col] = 0;
                if (r.free_after) free_slot((int32_t)r.col);
            }
        n_gram_groups++;
        n_gram_rows += rows.size();
        rows.clear();
    }
    // term v becomes (or already is) a pin: returns j. Changing a pin first reduces the open group.
    int pin_partner(int32_t v)
    {
        auto it = term_loc.find(v);
        if (it != term_loc.end() && (it->second & PINREF)) {
            touch(it->second);
            return (int)(it->second & 0xff);
        }
        int j = -1;
        for (int i = 0; i < n_pins; ++i)
            if (pin_term[i] == -1) { j = i; break; }
        if (j < 0) {
            uint64_t best = ~0ull;
            for (int i = 0; i < n_pins; ++i)
                if (pin_term[i] >= 0 && pin_hold[i] != epoch && pin_stamp[i] < best) {
                    best = pin_stamp[i];
                    j = i;
                }
        }
        if (j < 0) { err = "internal: no pin available for a reduction partner"; return 0; }
        flush_rows();
        uint32_t slot;
        bool temp = false;
        it = term_loc.find(v);
        if (it != term_loc.end()) {
            slot = it->second;  // resident in a tile slot
        } else {
            gen_term(v);
            slot = (uint32_t)alloc_slot(-2);
            emit(RI_ST, slot, 0.0, 0);
            temp = true;
        }
        if (!err.empty()) return 0;
        if (pin_term[j] >= 0) term_loc.erase(pin_term[j]);
        emit(RI_PINB0 + (uint32_t)j, slot, 0.0, 0);
        // the pin is the term's home from now on (operand register and reduction partner): the slot is released
        if (!temp) term_loc.erase(v);
        free_slot((int32_t)slot);
        pin_term[j] = v;
        pin_stamp[j] = clock++;
        pin_hold[j] = epoch;
        term_loc[v] = PINREF | (uint32_t)j;
        return j;
    }
    // a row of the open group: term u against the pins in `pin_mask`, itself and ones. Appends the output ids in
    // the order pins ascending, self, one.
    void add_row(int32_t u, uint32_t pin_mask, bool self, bool one, bool keep, std::vector<int32_t> &ids)
    {
        if (rows.size() == 8) flush_rows();
        uint32_t col;
        bool free_after = false;
        const Term &T = bp.term(u);
        const TermNode &root = T.nodes.back();
        auto it = term_loc.find(u);
        if (it != term_loc.end() && !(it->second & PINREF)) {
            col = it->second;
            touch(col);
        } else if (it != term_loc.end()) {
            // resident as a pin only: a row needs the values in the tile
            const int32_t sl = alloc_slot(-2);
            emit(RI_LDP0 + (it->second & 0xff), 0, 0.0, 0);
            emit(RI_ST, (uint32_t)sl, 0.0, 0);
            col = (uint32_t)sl;
            free_after = true;
        } else if (root.op == RR_OP_VAR) {
            col = staged(root.var);  // a bare feature is its own row: nothing to evaluate, nothing to store
        } else {
            // the slot first: if finding one has to reduce the open group, that RI_GRAM8 goes in FRONT of the term's
            // code and the term's last operation can carry its store (rr_isa.h RR_THEN_ST)
            const int32_t sl = alloc_slot(-2);
            if (!err.empty()) return;
            slot_rowheld[sl] = 1;
            gen_term(u);
            emit(RI_ST, (uint32_t)sl, 0.0, 0);
            if (keep) {
                slot_term[sl] = u;
                term_loc[u] = (uint32_t)sl;
            }
            col = (uint32_t)sl;
            free_after = !keep;
        }
        if (!err.empty()) return;
        if (rows.size() == 8) {
            // evaluating u has filled the group's last place meanwhile (cannot happen: rows are only added here) - keep
            // the invariant anyway
            flush_rows();
        }
        if (!(col & STAGED)) slot_rowheld[col] = 1;
        GRow r;
        r.col = col;
        r.want = (pin_mask & 0xffu) | (self ? 1u << 8 : 0u) | (one ? 1u << 9 : 0u);
        r.free_after = free_after;
        rows.push_back(r);
        for (int o = 0; o < 10; ++o)
            if ((r.want >> o) & 1u) {
                ids.push_back(P.n_dots);
                P.n_dots += 1;
                P.n_dot_ins += 1;
            }
    }

    int32_t clsmet(uint32_t y_ref)
    {
        emit(RI_CLSMET, y_ref, 0.0, 60);
        const int32_t id = P.n_dots;
        P.n_dots += 3;
        P.n_dot_ins += 3;
        return id;
    }

    void close()
    {
        if (g8) flush_rows();
        emit(RI_END, 0, 0.0, 0);
        const int32_t n_cols = (int32_t)colmap.size();
        auto patch = [&](uint32_t v) -> uint32_t { return (v & STAGED) ? (v & 0x3fffu) : (uint32_t)n_cols + v; };
        for (size_t i = pc_begin; i < P.ins.size(); ++i) {
            RRIns &x = P.ins[i];
            const uint32_t op = RR_OP(x.w0);
            if (op == RI_GRAM8) {
                // the column slot behind it: pre-patch bytes -> tile columns
                RRIns &dsl = P.ins[i + 1];
                uint8_t c8[8];
                std::memcpy(c8, &dsl.w1, 4);
                std::memcpy(c8 + 4, &dsl.imm, 4);
                for (int g = 0; g < 8; ++g) c8[g] = (uint8_t)((c8[g] & 0x80u) ? (c8[g] & 0x7fu) : (uint32_t)n_cols + c8[g]);
                std::memcpy(&dsl.w1, c8, 4);
                std::memcpy(&dsl.imm, c8 + 4, 4);
                ++i;
                continue;
            }
            if (op == RI_PINBG) continue;  // w1 is an engine column
            bool has_col = op >= RI_FIRST_M || op == RI_ST || op == RI_CLSMET;
            if (op == RI_RARE) has_col = !(RR_AUX(x.w0) & RB_CONST);
            // the operand of the instruction after USEP comes from the pin: its column field stays 0
            if (has_col && i > (size_t)pc_begin && RR_OP(P.ins[i - 1].w0) >= RI_USEP0 && RR_OP(P.ins[i - 1].w0) < RI_USEP0 + RR_NREG)
                has_col = false;
            if (has_col) x.w1 = patch(x.w1);
        }
        // peephole: "PIN j; MDOT" (a fresh term is pinned, then reduced against its partners) becomes one
        // instruction, the pin riding in aux bits 16-19 of the MDOT: one dispatch less per new term
        {
            std::vector<RRIns> out;
            out.reserve(P.ins.size() - pc_begin);
            for (size_t i = pc_begin; i < P.ins.size(); ++i) {
                const RRIns &x = P.ins[i];
                const uint32_t op = RR_OP(x.w0);
                if (op >= RI_PIN0 && op < RI_PIN0 + RR_NPIN && i + 1 < P.ins.size() &&
                    ((RR_OP(P.ins[i + 1].w0) == RI_MDOT && !lim.mdot_rows) || RR_OP(P.ins[i + 1].w0) == RI_MDOTDD) &&
                    !(P.ins[i + 1].w0 >> 24) && !((P.ins[i + 1].w0 >> (16 + (op - RI_PIN0))) & 1u)) {
                    RRIns m = P.ins[i + 1];
                    m.w0 |= (op - RI_PIN0 + 1u) << 24;
                    out.push_back(m);
                    ++i;
                    continue;
                }
                out.push_back(x);
            }
            P.ins.resize(pc_begin);
            P.ins.insert(P.ins.end(), out.begin(), out.end());
        }
        // peephole: super-instructions (rr_isa.h). One left-to-right pass, longest pattern first; every fused
        // form performs the operations of the sequence it replaces in the same order, so results are
        // bit-identical. A local-search neighbourhood is dominated by "constant * base term", "base term
        // op variable" and short products of variables, so this removes about 30 % of all dispatches.
        if (lim.fuse) {
            auto is_reg = [](uint32_t op, uint32_t base) { return op >= base && op < base + RR_NREG; };
            auto with_col2 = [](RRIns x, uint32_t col2) {
                const uint64_t bits = col2;
                std::memcpy(&x.imm, &bits, 8);
                return x;
            };
            std::vector<RRIns> out;
            out.reserve(P.ins.size() - pc_begin);
            const size_t n = P.ins.size();
            for (size_t i = pc_begin; i < n;) {
                const RRIns &x = P.ins[i];
                const uint32_t op0 = RR_OP(x.w0);
                const uint32_t op1 = i + 1 < n ? RR_OP(P.ins[i + 1].w0) : (uint32_t)RI_END;
                const uint32_t op2 = i + 2 < n ? RR_OP(P.ins[i + 2].w0) : (uint32_t)RI_END;
                RRIns f;
                std::memset(&f, 0, sizeof(f));
                size_t used = 0;
                if (op0 == RI_LOAD_C) {
                    if (is_reg(op1, RI_USEP0) && (op2 == RI_MUL_M || op2 == RI_DIV_M)) {
                        f.w0 = (op2 == RI_MUL_M ? RI_CMULP0 : RI_CDIVP0) + (op1 - RI_USEP0);
                        f.imm = x.imm;
                        used = 3;
                    } else if (op1 == RI_MUL_M || op1 == RI_DIV_M) {
                        f.w0 = op1 == RI_MUL_M ? RI_CMUL_M : RI_CDIV_M;
                        f.w1 = P.ins[i + 1].w1;
                        f.imm = x.imm;
                        used = 2;
                    }
                } else if (op0 == RI_LOAD_M) {
                    if (is_reg(op1, RI_USEP0) && op2 == RI_DIV_M) {
                        f.w0 = RI_LDMDIVP0 + (op1 - RI_USEP0);
                        f.w1 = x.w1;
                        used = 3;
                    } else if (op1 == RI_MUL_M && op2 == RI_MUL_M && g8) {
                        // (tile[a] * tile[b]) * tile[c]: b and c ride in the two halves of imm
                        f.w0 = RI_MUL_MMM;
                        f.w1 = x.w1;
                        const uint64_t bits = (uint64_t)P.ins[i + 1].w1 | ((uint64_t)P.ins[i + 2].w1 << 32);
                        std::memcpy(&f.imm, &bits, 8);
                        used = 3;
                    } else if (op1 == RI_MUL_M) {
                        f.w0 = RI_MUL_MM;
                        f.w1 = x.w1;
                        f = with_col2(f, P.ins[i + 1].w1);
                        used = 2;
                    }
                } else if (is_reg(op0, RI_USEP0)) {
                    if (op1 == RI_MUL_M || op1 == RI_DIV_M || op1 == RI_RDIV_M) {
                        f.w0 = (op1 == RI_MUL_M ? RI_MULP0 : (op1 == RI_DIV_M ? RI_DIVP0 : RI_RDIVP0)) + (op0 - RI_USEP0);
                        used = 2;
                    } else {
                        // the consumer of a USEP that stays is never the head of a pattern
                        out.push_back(x);
                        out.push_back(P.ins[i + 1]);
                        i += 2;
                        continue;
                    }
                } else if (is_reg(op0, RI_LDP0)) {
                    if (op1 == RI_MUL_M || op1 == RI_DIV_M) {
                        f.w0 = (op1 == RI_MUL_M ? RI_LDPMUL_M0 : RI_LDPDIV_M0) + (op0 - RI_LDP0);
                        f.w1 = P.ins[i + 1].w1;
                        used = 2;
                    }
                } else if (op0 == RI_MUL_M && op1 == RI_ST) {
                    f.w0 = RI_MUL_M_ST;
                    f.w1 = x.w1;
                    f = with_col2(f, P.ins[i + 1].w1);
                    used = 2;
                }
                if (used) {
                    out.push_back(f);
                    i += used;
                } else {
                    out.push_back(x);
                    ++i;
                }
            }
            P.ins.resize(pc_begin);
            P.ins.insert(P.ins.end(), out.begin(), out.end());
            // second pass: "X; MDOT" -> X with RR_THEN_MDOT (the term's last operation runs into its reductions);
            // G8 plans: "X; ST c" -> X with RR_THEN_ST (the term's last operation stores the row it has computed)
            out.clear();
            for (size_t i = pc_begin; i < P.ins.size(); ++i) {
                RRIns x = P.ins[i];
                const bool consumer_of_usep = !out.empty() && is_reg(RR_OP(out.back().w0), RI_USEP0);
                const bool data_slot = !out.empty() && RR_OP(out.back().w0) == RI_GRAM8;
                if (!data_slot && !consumer_of_usep && rr_md_fusable(RR_OP(x.w0)) && RR_AUX(x.w0) == 0 && i + 1 < P.ins.size()) {
                    if (!g8 && RR_OP(P.ins[i + 1].w0) == RI_MDOT) {
                        x.w0 |= RR_THEN_MDOT | (P.ins[i + 1].w0 & ~0xffu);
                        ++i;
                    } else if (g8 && RR_OP(P.ins[i + 1].w0) == RI_ST && P.ins[i + 1].w1 < 256u) {
                        x.w0 |= RR_THEN_ST | (P.ins[i + 1].w1 << 16);
                        ++i;
                    }
                }
                out.push_back(x);
            }
            P.ins.resize(pc_begin);
            P.ins.insert(P.ins.end(), out.begin(), out.end());
        }
        // Ring rows of the reductions, decided here instead of in the kernel (rr_isa.h RR_MDOT_ROWS): the
        // stream is straight-line, so the number of reductions emitted before an instruction - and with it
        // the ring row (count & 15) of each of its outputs - is known statically. Every instruction that
        // ends in an RI_MDOT is followed by a data slot holding the row of each of the 10 potential outputs
        // (self, one, pins 0..7); an output that is not wanted gets the row of the next wanted one (which
        // overwrites it) or the first free row behind this instruction's outputs.
        auto carries_mdot = [](const RRIns &x) {
            const uint32_t op = RR_OP(x.w0);
            return op == RI_MDOT || ((x.w0 & RR_THEN_MDOT) && rr_md_fusable(op));
        };
        if (lim.mdot_rows && !g8) {
            std::vector<RRIns> out;
            out.reserve(P.ins.size() - pc_begin + 1024);
            uint32_t cnt = 0, fl = 0;  // reductions pushed / flushed, exactly as the kernel counts them
            RRIns comb;
            std::memset(&comb, 0, sizeof(comb));
            comb.w0 = RI_COMBINE;
            for (size_t i = pc_begin; i < P.ins.size(); ++i) {
                const RRIns &x = P.ins[i];
                const uint32_t op = RR_OP(x.w0);
                out.push_back(x);
                bool flushed = false;
                if (carries_mdot(x)) {
                    // RI_MDOT flushes 8 rows on entry when 8 or more are pending, then pushes
                    if (cnt - fl >= 8) { fl += 8; flushed = true; }
                    const uint32_t want = ((x.w0 >> 8) & 3u) | (((x.w0 >> 16) & 0xffu) << 2);
                    uint8_t rows[16] = {0};
                    uint32_t r = cnt;
                    for (int o = 0; o < 10; ++o)
                        if ((want >> o) & 1u) rows[o] = (uint8_t)(r++ & 15u);
                    uint8_t next = (uint8_t)(r & 15u);
                    for (int o = 9; o >= 0; --o) {
                        if ((want >> o) & 1u) next = rows[o];
                        else rows[o] = next;
                    }
                    cnt = r;
                    RRIns d;
                    std::memset(&d, 0, sizeof(d));
                    d.w0 = RR_W0(RI_NOP, RR_MDOT_ROWS);
                    std::memcpy(&d.w1, rows, 4);
                    std::memcpy(&d.imm, rows + 4, 8);
                    out.push_back(d);
                } else if (op == RI_DOTM) {
                    // RI_DOTM pushes, then flushes behind itself
                    cnt += 1;
                    if (cnt - fl >= 8) { fl += 8; flushed = true; }
                } else if (op == RI_CLSMET) {
                    cnt += 3;  // classifier plans never run in the core
                }
                if (flushed && (fl & 31u) == 0) out.push_back(comb);
            }
            P.ins.resize(pc_begin);
            P.ins.insert(P.ins.end(), out.begin(), out.end());
        }
        // USEP and its consumer, and an RI_MDOT carrier and its data slot, stay inside one instruction
        // window (rr_isa.h RR_INS_WINDOW)
        {
            std::vector<RRIns> out;
            out.reserve(P.ins.size() - pc_begin + 16);
            RRIns nop;
            std::memset(&nop, 0, sizeof(nop));
            nop.w0 = RI_NOP;
            for (size_t i = pc_begin; i < P.ins.size(); ++i) {
                const uint32_t op = RR_OP(P.ins[i].w0);
                const bool pair_head = (op >= RI_USEP0 && op < RI_USEP0 + RR_NREG) || (lim.mdot_rows && !g8 && carries_mdot(P.ins[i])) ||
                                       op == RI_GRAM8;
                if (pair_head && out.size() % (size_t)lim.ins_window == (size_t)lim.ins_window - 1)
                    out.push_back(nop);
                out.push_back(P.ins[i]);
            }
            P.ins.resize(pc_begin);
            P.ins.insert(P.ins.end(), out.begin(), out.end());
        }
        RRChunk c;
        std::memset(&c, 0, sizeof(c));
        c.pc_begin = pc_begin;
        c.n_ins = (int32_t)P.ins.size() - pc_begin;
        c.dot_base = dot_base;
        c.n_dots = P.n_dots - dot_base;
        c.col_begin = col_begin;
        c.n_cols = n_cols;
        P.chunks.push_back(c);
        P.max_tile_cols = std::max(P.max_tile_cols, n_cols + (int32_t)slot_term.size());
    }
};

namespace {

struct Unit {
    std::vector<int32_t> terms;  // distinct term ids this unit touches
    double w;
};

// cut the unit list into chunks: by contract weight (target_chunks) and by tile capacity
struct ChunkSpec {
    int32_t begin, end;
    std::vector<int32_t> cols;
};

void term_vars(const Term &t, std::vector<int32_t> &out)
{
    for (const TermNode &n : t.nodes)
        if (n.op == RR_OP_VAR) out.push_back(n.var);
}

}  // namespace

static std::string cut_chunks(const BatchPlanner &bp, const std::vector<Unit> &units, const PlanLimits &lim,
                              const std::vector<int32_t> &always_cols, int32_t min_slots,
                              std::vector<ChunkSpec> &out)
{
    double total = 0.0;
    for (const Unit &u : units) total += u.w;
    const double target = lim.target_chunks > 1 ? total / lim.target_chunks : 1e300;
    const int32_t cap_cols = lim.tile_cols - min_slots;
    if (cap_cols < (int32_t)always_cols.size() + 1) return "tile too small for the required slots";
    size_t i = 0;
    std::vector<int32_t> vars;
    while (i < units.size()) {
        ChunkSpec cs;
        cs.begin = (int32_t)i;
        std::vector<int32_t> cols(always_cols);
        std::vector<char> have;
        auto has = [&](int32_t v) { return v < (int32_t)have.size() && have[v]; };
        auto add = [&](int32_t v) {
            if (v >= (int32_t)have.size()) have.resize(v + 1, 0);
            have[v] = 1;
            cols.push_back(v);
        };
        for (int32_t v : always_cols) {
            if (v >= (int32_t)have.size()) have.resize(v + 1, 0);
            have[v] = 1;
        }
        double w = 0.0;
        while (i < units.size()) {
            vars.clear();
            for (int32_t t : units[i].terms) term_vars(bp.term(t), vars);
            std::sort(vars.begin(), vars.end());
            vars.erase(std::unique(vars.begin(), vars.end()), vars.end());
            int32_t fresh = 0;
            for (int32_t v : vars)
                if (!has(v)) ++fresh;
            if ((int32_t)cols.size() + fresh > cap_cols) {
                if ((int32_t)i == cs.begin) return "a candidate uses more feature columns than the tile can stage";
                break;
            }
            for (int32_t v : vars)
                if (!has(v)) add(v);
            w += units[i].w;
            ++i;
            if (w >= target) break;
        }
        cs.end = (int32_t)i;
        cs.cols = std::move(cols);
        out.push_back(std::move(cs));
    }
    return "";
}

static int32_t max_need(const BatchPlanner &bp, const std::vector<Unit> &units)
{
    int32_t m = 0;
    for (const Unit &u : units)
        for (int32_t t : u.terms) m = std::max(m, bp.term(t).nodes.back().need);
    return m;
}

std::string BatchPlanner::plan_gram(const PlanLimits &lim, const ColIds &cols, const std::vector<int32_t> *subset,
                                    bool dd, SweepPlan &P, std::vector<int32_t> &cand_dot,
                                    std::vector<int32_t> &cand_dot_begin)
{
    std::vector<int32_t> list;
    if (subset) list = *subset;
    else {
        list.resize(b_->n_cand);
        for (int32_t c = 0; c < b_->n_cand; ++c) list[c] = c;
    }
    std::vector<Unit> units(list.size());
    for (size_t i = 0; i < list.size(); ++i) {
        const int32_t c = list[i];
        for (int32_t t = b_->cand_term_begin[c]; t < b_->cand_term_begin[c + 1]; ++t)
            units[i].terms.push_back(term_id_[t]);
        units[i].w = cand_w_[c];
    }
    // units (ascending) that list each distinct term: look-ahead for "is this term used again soon"
    std::vector<std::vector<int32_t>> term_units(terms_.size());
    for (size_t i = 0; i < units.size(); ++i)
        for (int32_t t : units[i].terms)
            if (term_units[t].empty() || term_units[t].back() != (int32_t)i) term_units[t].push_back((int32_t)i);
    const int32_t need = max_need(*this, units);
    // cached terms and the centred target live in pins; the tile only holds the spill temporaries and
    // the overflow
    const int32_t pins = std::min<int32_t>(std::max(lim.n_pins, 0), RR_NPIN);
    // slots wanted: all terms of the widest candidate resident + spill temporaries; if the tile
    // cannot give that, pairs are scheduled in blocks (see below)
    const int32_t min_slots = pins > 1 ? need + 1 + (pins < 4 ? 3 : 0)
                                       : std::min(std::max(4, need + 3), std::max(4, lim.tile_cols / 2));
    std::vector<ChunkSpec> specs;
    std::vector<int32_t> always;
    if (pins == 0) always.push_back(cols.yc);
    std::string err = cut_chunks(*this, units, lim, always, min_slots, specs);
    if (!err.empty()) return err;

    cand_dot.clear();
    cand_dot_begin.assign(1, 0);
    // dot key -> id, global over the plan (a dot computed in an earlier chunk is simply reused)
    DotMap dots((size_t)units.size() * 8 + 64);
    const int64_t KEY_YC = -1, KEY_ONE = -2;
    auto key = [](int64_t a, int64_t b) -> uint64_t {
        if (a > b) std::swap(a, b);
        return ((uint64_t)(uint32_t)(int32_t)a << 32) | (uint64_t)(uint32_t)(int32_t)b;
    };
    std::vector<uint32_t> partners;
    std::vector<uint64_t> pkeys;
    std::vector<int32_t> ids;
    for (const ChunkSpec &cs : specs) {
        Chunk ch(*this, P, lim, cs.cols, pins);
        const uint32_t yc_loc = pins > 0 ? ch.reserve_pin_global(cols.yc) : ch.staged(cols.yc);
        for (int32_t ui = cs.begin; ui < cs.end; ++ui) {
            const std::vector<int32_t> &T = units[ui].terms;
            const int32_t m = (int32_t)T.size();
            auto missing = [&](int64_t a, int64_t b) { return dots.find(key(a, b)) == nullptr; };
            // distinct terms that take part in a reduction that is still missing; a missing pair has
            // both ends in N (missing() is symmetric), so processing N in order emits each pair once,
            // when its later term is in t
            std::vector<int32_t> N;
            for (int32_t i = 0; i < m; ++i) {
                bool miss = missing(T[i], KEY_YC) || missing(T[i], KEY_ONE);
                for (int32_t j = 0; j < m && !miss; ++j) miss = missing(T[i], T[j]);
                if (miss && std::find(N.begin(), N.end(), T[i]) == N.end()) N.push_back(T[i]);
            }
            // resident terms first: what is new in this candidate then meets all of its partners in ONE
            // reduction instruction while it is in t, instead of being loaded back once per partner
            std::stable_partition(N.begin(), N.end(), [&](int32_t u) { return ch.resident(u); });
            ch.unpin_all();
            const int32_t room = ch.free_pins() + ch.pool_cap - (need + 1);
            if (room < 2) return "tile too small";
            // t = u; reductions of u with itself, ones, yc and every partner already resident.
            // transient: u is the last term reduced for this candidate and no candidate close by lists it
            // again, so it is evaluated into t and never stored (one instruction less per new term)
            auto reduce_term = [&](int32_t u, const std::vector<int32_t> &done, bool transient = false) {
                const bool self = missing(u, u), one = missing(u, KEY_ONE), with_yc = missing(u, KEY_YC);
                bool any = self || one || with_yc;
                for (int32_t v : done)
                    if (v != u && missing(u, v)) { any = true; break; }
                if (!any) {
                    // all of u's missing pairs are with terms that come later: they are reduced when
                    // that term is in t; u only has to be resident by then
                    ch.ensure(u);
                    return;
                }
                if (transient && !ch.resident(u)) ch.gen_term(u);
                else ch.ensure_tos(u);
                partners.clear();
                pkeys.clear();
                if (with_yc) { partners.push_back(yc_loc); pkeys.push_back(key(u, KEY_YC)); }
                for (int32_t v : done) {
                    if (v == u || !missing(u, v)) continue;
                    const int64_t sv = ch.lookup(v);
                    if (sv < 0) { ch.err = "internal: partner not resident"; return; }
                    partners.push_back((uint32_t)sv);
                    pkeys.push_back(key(u, v));
                }
                ids.clear();
                ch.mdot(self, one, partners, dd, ids);
                if (!ch.err.empty()) return;
                size_t q = 0;
                if (self) dots.set(key(u, u), ids[q++]);
                if (one) dots.set(key(u, KEY_ONE), ids[q++]);
                for (uint64_t k : pkeys) dots.set(k, ids[q++]);
            };
            if ((int32_t)N.size() <= room) {
                std::vector<int32_t> done;
                for (size_t q = 0; q < N.size(); ++q) {
                    const int32_t u = N[q];
                    bool transient = false;
                    if (q + 1 == N.size() && lim.transient_horizon > 0) {
                        const std::vector<int32_t> &occ = term_units[u];
                        auto it = std::upper_bound(occ.begin(), occ.end(), ui);
                        transient = it == occ.end() || *it > ui + lim.transient_horizon;
                    }
                    reduce_term(u, done, transient);
                    done.push_back(u);
                    if (!ch.err.empty()) return ch.err;
                }
            } else {
                // blocked schedule: groups of g terms; for every pair of groups make both resident
                const int32_t g = std::max(1, room / 2);
                const int32_t ng = ((int32_t)N.size() + g - 1) / g;
                for (int32_t ga = 0; ga < ng; ++ga)
                    for (int32_t gb = ga; gb < ng; ++gb) {
                        ch.unpin_all();
                        std::vector<int32_t> done;
                        for (int32_t pass = 0; pass < (ga == gb ? 1 : 2); ++pass) {
                            const int32_t grp = pass == 0 ? ga : gb;
                            for (int32_t i = grp * g; i < std::min((grp + 1) * g, (int32_t)N.size()); ++i) {
                                reduce_term(N[i], done);
                                done.push_back(N[i]);
                                if (!ch.err.empty()) return ch.err;
                            }
                        }
                    }
            }
            // index table of this candidate
            for (int32_t i = 0; i < m; ++i)
                for (int32_t j = i; j < m; ++j) {
                    const int32_t *it = dots.find(key(T[i], T[j]));
                    if (!it) return "internal: missing Gram dot";
                    cand_dot.push_back(*it);
                }
            for (int32_t i = 0; i < m; ++i) {
                const int32_t *it = dots.find(key(T[i], KEY_YC));
                if (!it) return "internal: missing Gram dot";
                cand_dot.push_back(*it);
            }
            for (int32_t i = 0; i < m; ++i) {
                const int32_t *it = dots.find(key(T[i], KEY_ONE));
                if (!it) return "internal: missing Gram dot";
                cand_dot.push_back(*it);
            }
            cand_dot_begin.push_back((int32_t)cand_dot.size());
        }
        if (!ch.err.empty()) return ch.err;
        ch.close();
    }
    return "";
}

// G8 variant of plan_gram (rr_isa.h RI_GRAM8): every term that still has reductions missing is evaluated into a tile
// slot and becomes one ROW of the open group; its partners (the centred target and the candidate's other terms)
// are pins; one RI_GRAM8 reduces up to eight rows against all pins with DMMA. A local-search neighbourhood keeps
// its base terms pinned and contributes one row per candidate. Candidates with more than 8 terms need more partners
// than there are pins: the caller plans those with plan_gram.
std::string BatchPlanner::plan_gram_g8(const PlanLimits &lim, const ColIds &cols, const std::vector<int32_t> *subset, SweepPlan &P,
                                       std::vector<int32_t> &cand_dot, std::vector<int32_t> &cand_dot_begin)
{
    std::vector<int32_t> list;
    if (subset) list = *subset;
    else {
        list.resize(b_->n_cand);
        for (int32_t c = 0; c < b_->n_cand; ++c) list[c] = c;
    }
    std::vector<Unit> units(list.size());
    for (size_t i = 0; i < list.size(); ++i) {
        const int32_t c = list[i];
        for (int32_t t = b_->cand_term_begin[c]; t < b_->cand_term_begin[c + 1]; ++t)
            units[i].terms.push_back(term_id_[t]);
        units[i].w = cand_w_[c];
        // a row meets its partners as pins: the other terms, the centred target and the column of ones
        if ((int32_t)units[i].terms.size() > RR_NPIN - 1) return "candidate too wide for a G8 plan";
    }
    std::vector<std::vector<int32_t>> term_units(terms_.size());
    for (size_t i = 0; i < units.size(); ++i)
        for (int32_t t : units[i].terms)
            if (term_units[t].empty() || term_units[t].back() != (int32_t)i) term_units[t].push_back((int32_t)i);
    const int32_t need = max_need(*this, units);
    const int32_t min_slots = need + 1 + 2;  // spill temporaries + at least two rows
    std::vector<ChunkSpec> specs;
    std::vector<int32_t> always;
    std::string err = cut_chunks(*this, units, lim, always, min_slots, specs);
    if (!err.empty()) return err;

    cand_dot.clear();
    cand_dot_begin.assign(1, 0);
    DotMap dots((size_t)units.size() * 8 + 64);
    const int64_t KEY_YC = -1, KEY_ONE = -2;
    auto key = [](int64_t a, int64_t b) -> uint64_t {
        if (a > b) std::swap(a, b);
        return ((uint64_t)(uint32_t)(int32_t)a << 32) | (uint64_t)(uint32_t)(int32_t)b;
    };
    std::vector<int32_t> ids;
    int32_t prev_m = 0, prev_T[RR_NPIN], prev_pid[RR_NPIN][RR_NPIN + 2];
    (void)prev_T;
    for (const ChunkSpec &cs : specs) {
        Chunk ch(*this, P, lim, cs.cols, RR_NPIN);
        ch.g8 = true;
        ch.reserve_pin_global(cols.yc);  // pin 0, for the whole chunk
        ch.reserve_pin_ones();           // pin 7: sum(t) of every row comes out of the DMMA like its other products
        if (!ch.err.empty()) return ch.err;
        for (int32_t ui = cs.begin; ui < cs.end; ++ui) {
            const std::vector<int32_t> &T = units[ui].terms;
            const int32_t m = (int32_t)T.size();
            // every pair of this candidate is looked up ONCE: pid[i][j] (i <= j < m: Gram, j = m: with yc, j = m + 1:
            // with ones) holds the reduction id or -1; the rows added below fill in what was missing
            // (neighbours differ in one or two terms: what the previous candidate found is taken from its table)
            int32_t pid[RR_NPIN][RR_NPIN + 2];
            int32_t was[RR_NPIN];
            for (int32_t i = 0; i < m; ++i) {
                was[i] = -1;
                for (int32_t k = 0; k < prev_m; ++k)
                    if (prev_T[k] == T[i]) { was[i] = k; break; }
            }
            for (int32_t i = 0; i < m; ++i) {
                for (int32_t j = i; j < m; ++j) {
                    int32_t id = -1;
                    if (was[i] >= 0 && was[j] >= 0) id = was[i] <= was[j] ? prev_pid[was[i]][was[j]] : prev_pid[was[j]][was[i]];
                    if (id < 0) {
                        const int32_t *it = dots.find(key(T[i], T[j]));
                        id = it ? *it : -1;
                    }
                    pid[i][j] = id;
                }
                int32_t idy = was[i] >= 0 ? prev_pid[was[i]][prev_m] : -1, ido = was[i] >= 0 ? prev_pid[was[i]][prev_m + 1] : -1;
                if (idy < 0) {
                    const int32_t *iy = dots.find(key(T[i], KEY_YC));
                    idy = iy ? *iy : -1;
                }
                if (ido < 0) {
                    const int32_t *io = dots.find(key(T[i], KEY_ONE));
                    ido = io ? *io : -1;
                }
                pid[i][m] = idy;
                pid[i][m + 1] = ido;
            }
            auto index_of = [&](int64_t t) -> int32_t {
                for (int32_t i = 0; i < m; ++i)
                    if (T[i] == t) return i;
                return -1;
            };
            auto slot = [&](int64_t a, int64_t b) -> int32_t & {
                // a is a term of this candidate; b a term, KEY_YC or KEY_ONE
                const int32_t i = index_of(a);
                if (b == KEY_YC) return pid[i][m];
                if (b == KEY_ONE) return pid[i][m + 1];
                const int32_t j = index_of(b);
                return i <= j ? pid[i][j] : pid[j][i];
            };
            auto missing = [&](int64_t a, int64_t b) { return slot(a, b) < 0; };
            std::vector<int32_t> N;
            for (int32_t i = 0; i < m; ++i) {
                bool miss = pid[i][m] < 0 || pid[i][m + 1] < 0;
                for (int32_t j = 0; j < m && !miss; ++j) miss = (i <= j ? pid[i][j] : pid[j][i]) < 0;
                if (miss && std::find(N.begin(), N.end(), T[i]) == N.end()) N.push_back(T[i]);
            }
            // resident terms first: what is new in this candidate then meets all of its partners in its one row
            std::stable_partition(N.begin(), N.end(), [&](int32_t u) { return ch.resident(u); });
            ch.unpin_all();
            std::vector<int32_t> done;
            for (size_t q = 0; q < N.size(); ++q) {
                const int32_t u = N[q];
                const bool self = missing(u, u), one = missing(u, KEY_ONE), with_yc = missing(u, KEY_YC);
                bool any = self || one || with_yc;
                for (int32_t v : done)
                    if (v != u && missing(u, v)) { any = true; break; }
                if (!any) {
                    // all of u's missing pairs are with terms that come later: u only has to be resident by then
                    ch.ensure(u);
                    done.push_back(u);
                    if (!ch.err.empty()) return ch.err;
                    continue;
                }
                uint32_t pin_mask = with_yc ? 1u : 0u;
                uint64_t pin_key[RR_NPIN];
                int64_t pin_partner_term[RR_NPIN];
                if (with_yc) {
                    pin_key[0] = key(u, KEY_YC);
                    pin_partner_term[0] = KEY_YC;
                }
                if (one) {
                    pin_mask |= 1u << (RR_NPIN - 1);
                    pin_key[RR_NPIN - 1] = key(u, KEY_ONE);
                    pin_partner_term[RR_NPIN - 1] = KEY_ONE;
                }
                for (int32_t v : done) {
                    if (v == u || !missing(u, v)) continue;
                    const int j = ch.pin_partner(v);
                    if (!ch.err.empty()) return ch.err;
                    if ((pin_mask >> j) & 1u) return "internal: duplicate pinned partner";
                    pin_mask |= 1u << j;
                    pin_key[j] = key(u, v);
                    pin_partner_term[j] = v;
                }
                bool transient = false;
                if (q + 1 == N.size() && lim.transient_horizon > 0) {
                    const std::vector<int32_t> &occ = term_units[u];
                    auto it = std::upper_bound(occ.begin(), occ.end(), ui);
                    transient = it == occ.end() || *it > ui + lim.transient_horizon;
                }
                ids.clear();
                ch.add_row(u, pin_mask, self, false, !transient, ids);
                if (!ch.err.empty()) return ch.err;
                size_t k = 0;
                for (int j = 0; j < RR_NPIN; ++j)
                    if ((pin_mask >> j) & 1u) {
                        dots.set(pin_key[j], ids[k]);
                        slot(u, pin_partner_term[j]) = ids[k];
                        ++k;
                    }
                if (self) {
                    dots.set(key(u, u), ids[k]);
                    slot(u, u) = ids[k++];
                }
                done.push_back(u);
            }
            // (a term listed twice by one candidate shares its ids through the map, not through pid)
            auto take = [&](int32_t have, uint64_t k) -> bool {
                if (have < 0) {
                    const int32_t *it = dots.find(k);
                    if (!it) return false;
                    have = *it;
                }
                cand_dot.push_back(have);
                return true;
            };
            for (int32_t i = 0; i < m; ++i)
                for (int32_t j = i; j < m; ++j)
                    if (!take(pid[i][j], key(T[i], T[j]))) return "internal: missing Gram dot";
            for (int32_t i = 0; i < m; ++i)
                if (!take(pid[i][m], key(T[i], KEY_YC))) return "internal: missing Gram dot";
            for (int32_t i = 0; i < m; ++i)
                if (!take(pid[i][m + 1], key(T[i], KEY_ONE))) return "internal: missing Gram dot";
            cand_dot_begin.push_back((int32_t)cand_dot.size());
            prev_m = m;
            for (int32_t i = 0; i < m; ++i) {
                prev_T[i] = T[i];
                for (int32_t j = i; j < m + 2; ++j) prev_pid[i][j] = j < m ? pid[i][j] : pid[i][j];
            }
        }
        if (!ch.err.empty()) return ch.err;
        ch.close();
        P.n_gram_groups += ch.n_gram_groups;
        P.n_gram_rows += ch.n_gram_rows;
    }
    return "";
}

std::string BatchPlanner::plan_residual(const PlanLimits &lim, const ColIds &cols, const std::vector<int32_t> &subset,
                                        const double *coef, SweepPlan &P, std::vector<int32_t> &cand_dot,
                                        std::vector<int32_t> &cand_dot_begin)
{
    std::vector<Unit> units(subset.size());
    for (size_t i = 0; i < subset.size(); ++i) {
        const int32_t c = subset[i];
        for (int32_t t = b_->cand_term_begin[c]; t < b_->cand_term_begin[c + 1]; ++t)
            units[i].terms.push_back(term_id_[t]);
        units[i].w = cand_w_[c];
    }
    const int32_t need = max_need(*this, units);
    int32_t widest = 0;
    for (const Unit &u : units) widest = std::max(widest, (int32_t)u.terms.size());
    const int32_t pins = std::min<int32_t>(std::max(lim.n_pins, 0), RR_NPIN);
    // all terms resident (pins first, then tile slots) when they fit; wide candidates stream their terms
    // instead (below)
    const int32_t min_slots = pins > 1 ? need + 2
                                       : std::min(widest + need + 2, std::max(need + 4, lim.tile_cols / 2));
    std::vector<ChunkSpec> specs;
    std::string err = cut_chunks(*this, units, lim, {cols.y}, min_slots, specs);
    if (!err.empty()) return err + " (residual pass)";
    cand_dot.clear();
    cand_dot_begin.assign(1, 0);
    std::vector<uint32_t> partners;
    std::vector<int32_t> ids;
    for (const ChunkSpec &cs : specs) {
        Chunk ch(*this, P, lim, cs.cols, pins);
        const uint32_t y_col = ch.staged(cols.y);
        for (int32_t ui = cs.begin; ui < cs.end; ++ui) {
            const int32_t c = subset[ui];
            const std::vector<int32_t> &T = units[ui].terms;
            const int32_t m = (int32_t)T.size();
            const double *cf = coef + b_->cand_term_begin[c] + c;
            ch.unpin_all();
            if (m + need + 2 > ch.pool_cap + ch.free_pins()) {
                // wide candidate: not all terms fit at once. Stream them: yhat accumulates in one
                // slot in the reference's association order, then every term is evaluated a second time
                // against the residual.
                const int32_t acc = ch.alloc_slot(-2);
                bool first = true;
                for (int32_t i = 0; i < m; ++i) {
                    const double ci = cf[i];
                    if (ci == 0.0) continue;
                    ch.gen_term(T[i]);
                    if (ci != 1.0) ch.emit(RI_MUL_C, 0, ci, 1);
                    if (!first) ch.emit(RI_ADD_M, (uint32_t)acc, 0.0, 1);  // acc + c_i t_i (commutative)
                    ch.emit(RI_ST, (uint32_t)acc, 0.0, 0);
                    first = false;
                    if (!ch.err.empty()) return ch.err;
                }
                if (first) ch.emit(RI_LOAD_C, 0, cf[m], 0);
                else if (cf[m] != 0.0) ch.emit(RI_ADD_C, 0, cf[m], 1);
                ch.emit(RI_RSUB_M, y_col, 0.0, 1);  // t = y - yhat
                ch.emit(RI_ST, (uint32_t)acc, 0.0, 0);
                ids.clear();
                ch.mdot(true, true, {}, false, ids);  // r.r, r.1
                cand_dot.push_back(ids[0]);
                const int32_t one_id = ids[1];
                for (int32_t i = 0; i < m; ++i) {
                    ch.gen_term(T[i]);
                    ids.clear();
                    ch.mdot(false, false, {(uint32_t)acc}, false, ids);  // t_i . r
                    cand_dot.push_back(ids[0]);
                    if (!ch.err.empty()) return ch.err;
                }
                cand_dot.push_back(one_id);
                cand_dot_begin.push_back((int32_t)cand_dot.size());
                ch.free_slot(acc);
                continue;
            }
            std::vector<uint32_t> loc(m);
            for (int32_t i = 0; i < m; ++i) {
                loc[i] = ch.ensure(T[i]);
                if (!ch.err.empty()) return ch.err;
            }
            // a candidate may list one distinct term twice: reduce against each location once
            // yhat in the association order of rils_rols_cpp.cpp:488-515
            bool first = true;
            for (int32_t i = 0; i < m; ++i) {
                const double ci = cf[i];
                if (ci == 0.0) continue;  // snapped away (value_zero)
                if (first) {
                    ch.emit_load(loc[i]);
                    if (ci != 1.0) ch.emit(RI_MUL_C, 0, ci, 1);
                    first = false;
                } else if (ci == 1.0) {
                    ch.emit_m(RI_ADD_M, loc[i], 0.0, 1);
                } else {
                    ch.emit_m(RI_AXPY, loc[i], ci, 2);
                }
            }
            if (cf[m] != 0.0) {
                if (first) ch.emit(RI_LOAD_C, 0, cf[m], 0);
                else ch.emit(RI_ADD_C, 0, cf[m], 1);
                first = false;
            }
            if (first) ch.emit(RI_LOAD_C, 0, 0.0, 0);
            ch.emit(RI_RSUB_M, y_col, 0.0, 1);  // t = y - yhat
            // r.r, r.1, r.t_i; a distinct term listed twice is reduced once
            partners.clear();
            std::vector<int32_t> ppos(m, -1);
            for (int32_t i = 0; i < m; ++i) {
                for (int32_t j = 0; j < i; ++j)
                    if (loc[j] == loc[i]) { ppos[i] = ppos[j]; break; }
                if (ppos[i] < 0) {
                    ppos[i] = (int32_t)partners.size();
                    partners.push_back(loc[i]);
                }
            }
            ids.clear();
            ch.mdot(true, true, partners, false, ids);
            if (!ch.err.empty()) return ch.err;
            cand_dot.push_back(ids[0]);
            for (int32_t i = 0; i < m; ++i) cand_dot.push_back(ids[2 + ppos[i]]);
            cand_dot.push_back(ids[1]);
            cand_dot_begin.push_back((int32_t)cand_dot.size());
        }
        if (!ch.err.empty()) return ch.err;
        ch.close();
    }
    return "";
}

std::string BatchPlanner::plan_eval(const PlanLimits &lim, const ColIds &cols, bool metrics, SweepPlan &P,
                                    std::vector<int32_t> &cand_dot)
{
    std::vector<Unit> units(b_->n_cand);
    for (int32_t c = 0; c < b_->n_cand; ++c) {
        units[c].terms.push_back(term_id_[b_->cand_term_begin[c]]);
        units[c].w = cand_w_[c];
    }
    const int32_t need = max_need(*this, units);
    std::vector<ChunkSpec> specs;
    std::string err = cut_chunks(*this, units, lim, {cols.y}, need + 2, specs);
    if (!err.empty()) return err;
    cand_dot.assign(b_->n_cand, DOT_NONE);
    std::unordered_map<int32_t, int32_t> done;  // identical programs share their result
    std::vector<int32_t> ids;
    for (const ChunkSpec &cs : specs) {
        Chunk ch(*this, P, lim, cs.cols, 0);
        const uint32_t y_col = ch.staged(cols.y);
        for (int32_t c = cs.begin; c < cs.end; ++c) {
            const int32_t u = units[c].terms[0];
            auto it = done.find(u);
            if (it != done.end()) { cand_dot[c] = it->second; continue; }
            ch.unpin_all();
            ch.gen_term(u);
            int32_t id;
            if (metrics) {
                id = ch.clsmet(y_col);
            } else {
                ch.emit(RI_RSUB_M, y_col, 0.0, 1);  // t = y - yhat
                ids.clear();
                ch.mdot(true, false, {}, false, ids);
                id = ids[0];
            }
            done.emplace(u, id);
            cand_dot[c] = id;
            if (!ch.err.empty()) return ch.err;
        }
        ch.close();
    }
    return "";
}

std::string BatchPlanner::plan_materialise(const PlanLimits &lim, const ColIds &cols, SweepPlan &P)
{
    (void)cols;
    std::vector<Unit> units(terms_.size());
    for (size_t u = 0; u < terms_.size(); ++u) {
        units[u].terms.push_back((int32_t)u);
        units[u].w = terms_[u].w + 1;
    }
    const int32_t need = max_need(*this, units);
    std::vector<ChunkSpec> specs;
    std::string err = cut_chunks(*this, units, lim, {}, need + 2, specs);
    if (!err.empty()) return err;
    for (const ChunkSpec &cs : specs) {
        Chunk ch(*this, P, lim, cs.cols, 0);
        for (int32_t u = cs.begin; u < cs.end; ++u) {
            ch.unpin_all();
            ch.gen_term(u);
            ch.emit(RI_STG, (uint32_t)u, 0.0, 0);
            if (!ch.err.empty()) return ch.err;
        }
        ch.close();
    }
    P.n_stg_cols = (int32_t)terms_.size();
    return "";
}

}  // namespace rr

// ---------------------------------------------------------------------------------------------
// R8 plans: the row machine (rr_isa.h RQ_*, rr_sweep_r8.cuh)
// ---------------------------------------------------------------------------------------------
namespace rr {
namespace {

struct RqOp {
    uint8_t op, mode, rare;  // rare: RRRareOp | swap << 4
    int32_t ref;             // RQ_M: >= 0 engine feature column, < 0 stored sub-expression -1 - sid
    double k;                // RQ_K
};
struct RqProg {
    std::vector<RqOp> ops;
    std::vector<int32_t> stored;  // distinct stored sub-expressions read, in program order
    double w = 0.0;               // contract weight of the operations
    uint64_t shape = 0;           // equal shapes: same operations, modes and constants; operands may differ
    bool done = false;
    // programs of stored sub-expressions only: a segment reads at most kSegStored stored operands; between two
    // segments the value waits in the sub-expression's own slot (so that its operands need not be resident at once)
    std::vector<uint32_t> seg_begin;
};
constexpr int kSegStored = 2;

// Builder of the one chunk of an R8 plan. Terms become ROWS; rows wait in a pool until a pin has to change (or the
// plan ends), then the pool is sorted by (first stored operand, shape) and cut into groups of up to eight rows of one
// shape. Sub-expressions with more than one user are STORED: evaluated once as a uniform group, kept in a tile slot
// (LRU over the slots the tile has left), read by the rows as an RQ_M operand.
struct R8Builder {
    const BatchPlanner &bp;
    const rr_batch *b;
    SweepPlan &P;
    DotMap &dots;
    std::unordered_map<int32_t, int32_t> colmap;
    int32_t n_staged = 0, slot_cap = 0, slots_used = 0;
    int32_t pc_begin = 0, dot_base = 0, col_begin = 0;
    std::string err;

    struct Sub {
        int32_t term, node;
        std::vector<int64_t> parents;  // distinct (parent, side) uses, capped
        int uses = 0;
        double w = 0.0;  // contract weight of one evaluation
        mutable double inc_w = -1.0;
        bool forced = false;
        int32_t slot = -1;
        RqProg prog;
    };
    CodeTable subs;
    std::vector<Sub> sub;
    std::vector<std::vector<int32_t>> node_sid;  // [distinct term][node]; empty until registered
    std::vector<RqProg> term_prog;

    std::vector<int32_t> slot_sid, slot_lock;
    std::vector<uint64_t> slot_stamp;
    uint64_t clock = 1, epoch = 1;

    int32_t pin_term[RR_NPIN];
    uint64_t pin_stamp[RR_NPIN], pin_hold[RR_NPIN];
    std::vector<int8_t> term_pin;

    struct Row {
        int32_t term;
        uint32_t want;      // bits 0-7 pins, 8 self, 9 one
        uint64_t key[10];   // dot key of each wanted output
        int32_t primary;
        uint64_t shape;
    };
    std::vector<Row> pool;
    uint64_t n_groups = 0, n_rows = 0, n_stored_evals = 0;

    R8Builder(const BatchPlanner &bp_, const rr_batch *b_, SweepPlan &P_, DotMap &dots_, const PlanLimits &lim, const std::vector<int32_t> &cols)
        : bp(bp_), b(b_), P(P_), dots(dots_), subs(b_, 4096)
    {
        pc_begin = (int32_t)P.ins.size();
        dot_base = P.n_dots;
        col_begin = (int32_t)P.cols.size();
        for (int32_t g : cols) {
            colmap.emplace(g, (int32_t)colmap.size());
            P.cols.push_back(g);
        }
        n_staged = (int32_t)cols.size();
        slot_cap = std::min(std::min(lim.tile_cols - n_staged, lim.max_slots), 200 - n_staged);
        node_sid.resize(bp.n_terms_distinct());
        term_prog.resize(bp.n_terms_distinct());
        term_pin.assign(bp.n_terms_distinct(), -1);
        for (int j = 0; j < RR_NPIN; ++j) {
            pin_term[j] = -1;
            pin_stamp[j] = pin_hold[j] = 0;
        }
    }

    // ---- sub-expression table ----
    void register_term(int32_t u)
    {
        if (!node_sid[u].empty()) return;
        const Term &T = bp.term(u);
        const int32_t n = (int32_t)T.nodes.size();
        std::vector<int32_t> &sid = node_sid[u];
        sid.assign(n, -1);
        std::vector<double> wsub(n, 0.0);
        for (int32_t x = 0; x < n; ++x) {
            const TermNode &nd = T.nodes[x];
            if (nd.leaf()) continue;
            wsub[x] = kW[nd.op] + wsub[nd.left] + (nd.right >= 0 ? wsub[nd.right] : 0.0);
            const int32_t id = subs.find_or_insert(T.code_begin + nd.first, T.code_begin + x + 1, (int32_t)sub.size());
            if (id == (int32_t)sub.size()) {
                Sub s;
                s.term = u;
                s.node = x;
                s.w = wsub[x];
                sub.push_back(std::move(s));
            }
            sid[x] = id;
        }
        auto use = [&](int32_t x, int64_t parent) {
            if (x < 0 || sid[x] < 0) return;
            Sub &s = sub[sid[x]];
            if (s.parents.size() >= 16 || std::find(s.parents.begin(), s.parents.end(), parent) != s.parents.end()) return;
            s.parents.push_back(parent);
            s.uses = (int)s.parents.size();
        };
        // a use = (parent sub-expression, side); the root's parent is the term itself
        use(n - 1, -(int64_t)(1 + u) * 4);
        for (int32_t x = 0; x < n; ++x) {
            const TermNode &nd = T.nodes[x];
            if (nd.leaf()) continue;
            use(nd.left, (int64_t)sid[x] * 4 + 1);
            if (nd.right >= 0) use(nd.right, (int64_t)sid[x] * 4 + 2);
        }
    }
    // Stored: evaluated once per tile as a uniform group and read from a tile slot. Worth it for an expensive
    // sub-expression with two users, or a cheap one with many (evaluating it inline costs a group the same whether
    // one row or eight use it).
    int min_uses_cheap = 6;
    // contract weight of evaluating sub-expression `sid` when its stored parts are read from their slots
    double inc_w(int32_t sid) const
    {
        const Sub &x = sub[sid];
        if (x.inc_w >= 0.0) return x.inc_w;
        const TermNode &nd = bp.term(x.term).nodes[x.node];
        const std::vector<int32_t> &ids = node_sid[x.term];
        double w = kW[nd.op];
        const int32_t kids[2] = {nd.left, nd.right};
        for (int32_t c : kids)
            if (c >= 0 && ids[c] >= 0 && !is_stored(ids[c])) w += inc_w(ids[c]);
        x.inc_w = w;
        return w;
    }
    bool is_stored(int32_t sid) const
    {
        if (sid < 0) return false;
        const Sub &x = sub[sid];
        return x.forced || (x.uses >= 2 && (x.uses >= min_uses_cheap || inc_w(sid) >= 8.0));
    }

    // ---- compilation of a tree into a row program ----
    struct Gen {
        const Term &T;
        const std::vector<int32_t> &sid;
        int32_t top;
        RqProg &pg;
        int32_t self = -1;            // stored sub-expression being compiled (segmented), -1 for a row
        bool u_live = false;
        std::vector<int32_t> seg_refs;  // stored operands of the open segment
    };
    bool simple(const Gen &g, int32_t y) const { return g.T.nodes[y].leaf() || is_stored(g.sid[y]); }
    bool needs_u(const Gen &g, int32_t y) const
    {
        const TermNode &n = g.T.nodes[y];
        if (simple(g, y)) return false;
        if (n.right < 0) return needs_u(g, n.left);
        if (simple(g, n.right)) return needs_u(g, n.left);
        if (simple(g, n.left)) return needs_u(g, n.right);
        return true;
    }
    void push_op(Gen &g, uint8_t op, uint8_t mode, uint8_t rare, int32_t ref, double k)
    {
        RqProg &pg = g.pg;
        if (op == RQ_TU) g.u_live = true;
        if (mode == RQ_U && op >= RQ_LD && op <= RQ_RARE) g.u_live = false;
        if (g.self >= 0 && mode == RQ_M && ref < 0 && op >= RQ_LD && op <= RQ_RARE) {
            const int32_t s = -1 - ref;
            if (std::find(g.seg_refs.begin(), g.seg_refs.end(), s) == g.seg_refs.end()) {
                // a new segment: the value so far waits in the sub-expression's own slot (not while u is live or t is
                // about to be overwritten: the segment then simply reads one more stored operand)
                if ((int)g.seg_refs.size() >= kSegStored && !g.u_live && op != RQ_LD) {
                    pg.seg_begin.push_back((uint32_t)pg.ops.size());
                    g.seg_refs.clear();
                    RqOp l;
                    l.op = RQ_LD;
                    l.mode = RQ_M;
                    l.rare = 0;
                    l.ref = -1 - g.self;
                    l.k = 0.0;
                    pg.ops.push_back(l);
                }
                g.seg_refs.push_back(s);
            }
        }
        RqOp o;
        o.op = op;
        o.mode = mode;
        o.rare = rare;
        o.ref = ref;
        o.k = k;
        pg.ops.push_back(o);
        if (mode == RQ_M && ref < 0 && op != RQ_ST) {
            const int32_t s = -1 - ref;
            if (s != g.self && std::find(pg.stored.begin(), pg.stored.end(), s) == pg.stored.end()) pg.stored.push_back(s);
        }
    }
    // operand of a binary operation / LD: node y is simple
    void operand(const Gen &g, int32_t y, uint8_t &mode, int32_t &ref, double &k) const
    {
        const TermNode &n = g.T.nodes[y];
        mode = RQ_M;
        ref = 0;
        k = 0.0;
        if (n.op == RR_OP_CONST) { mode = RQ_K; k = n.cval; }
        else if (n.op == RR_OP_VAR) ref = n.var;
        else ref = -1 - g.sid[y];
    }
    void bin(Gen &g, uint32_t op, uint8_t mode, int32_t ref, double k, bool swap)
    {
        switch (op) {
        case RR_OP_PLUS: push_op(g, RQ_ADD, mode, 0, ref, k); return;
        case RR_OP_MULTIPLY: push_op(g, RQ_MUL, mode, 0, ref, k); return;
        case RR_OP_MINUS: push_op(g, swap ? RQ_RSUB : RQ_SUB, mode, 0, ref, k); return;
        case RR_OP_DIVIDE: push_op(g, swap ? RQ_RDIV : RQ_DIV, mode, 0, ref, k); return;
        }
        uint8_t rare = 0;
        switch (op) {
        case RR_OP_POW: rare = RR_POW; break;
        case RR_OP_LESS_THAN: rare = RR_LT; break;
        case RR_OP_GREATER_THAN: rare = RR_GT; break;
        case RR_OP_EQUAL: rare = RR_EQ; break;
        case RR_OP_NOT_EQUAL: rare = RR_NE; break;
        case RR_OP_MIN: rare = RR_MIN; break;
        case RR_OP_MAX: rare = RR_MAX; break;
        }
        push_op(g, RQ_RARE, mode, (uint8_t)(rare | (swap ? 16u : 0u)), ref, k);
    }
    void gen(Gen &g, int32_t x)  // leaves t = value(node x)
    {
        const TermNode &n = g.T.nodes[x];
        if (n.leaf() || (x != g.top && is_stored(g.sid[x]))) {
            uint8_t mode;
            int32_t ref;
            double k;
            operand(g, x, mode, ref, k);
            push_op(g, RQ_LD, mode, 0, ref, k);
            return;
        }
        if (n.right < 0) {
            gen(g, n.left);
            uint8_t op = RQ_NOP;
            switch (n.op) {
            case RR_OP_SIN: op = RQ_SIN; break;
            case RR_OP_COS: op = RQ_COS; break;
            case RR_OP_LN: op = RQ_LN; break;
            case RR_OP_EXP: op = RQ_EXP; break;
            case RR_OP_SQRT: op = RQ_SQRT; break;
            case RR_OP_SQR: op = RQ_SQR; break;
            default: err = "r8: unknown unary operator"; break;
            }
            push_op(g, op, 0, 0, 0, 0.0);
            g.pg.w += kW[n.op];
            return;
        }
        uint8_t mode;
        int32_t ref;
        double k;
        if (simple(g, n.right)) {
            gen(g, n.left);
            operand(g, n.right, mode, ref, k);
            bin(g, n.op, mode, ref, k, false);
        } else if (simple(g, n.left)) {
            gen(g, n.right);
            operand(g, n.left, mode, ref, k);
            bin(g, n.op, mode, ref, k, true);
        } else {
            const bool ul = needs_u(g, n.left), ur = needs_u(g, n.right);
            if (ul && ur) {
                // both sides need the second register: the right one becomes a stored sub-expression
                sub[g.sid[n.right]].forced = true;
                gen(g, n.left);
                operand(g, n.right, mode, ref, k);
                bin(g, n.op, mode, ref, k, false);
            } else if (ul) {
                gen(g, n.left);
                push_op(g, RQ_TU, 0, 0, 0, 0.0);
                gen(g, n.right);
                bin(g, n.op, RQ_U, 0, 0.0, true);  // t = u op t
            } else {
                gen(g, n.right);
                push_op(g, RQ_TU, 0, 0, 0, 0.0);
                gen(g, n.left);
                bin(g, n.op, RQ_U, 0, 0.0, false);  // t = t op u
            }
        }
        g.pg.w += kW[n.op];
    }
    void finish_prog(RqProg &pg)
    {
        uint64_t h = 0x9e3779b97f4a7c15ull;
        for (const RqOp &o : pg.ops) {
            uint64_t v = (uint64_t)o.op | ((uint64_t)o.mode << 8) | ((uint64_t)o.rare << 16);
            h = (h ^ v) * 0xc4ceb9fe1a85ec53ull;
            h ^= h >> 29;
        }
        pg.shape = h;
        pg.done = true;
    }
    const RqProg &row_prog(int32_t u)
    {
        RqProg &pg = term_prog[u];
        if (pg.done) return pg;
        const Term &T = bp.term(u);
        const int32_t root = (int32_t)T.nodes.size() - 1;
        Gen g{T, node_sid[u], -1, pg};  // top = -1: a stored root is read, not re-evaluated
        gen(g, root);
        if ((int)pg.stored.size() > kSegStored && err.empty()) {
            // too many stored operands for a row (a row has no slot to wait in): the whole term becomes a stored
            // sub-expression and the row reads it
            sub[node_sid[u][root]].forced = true;
            pg = RqProg();
            Gen g2{T, node_sid[u], -1, pg};
            gen(g2, root);
        }
        finish_prog(pg);
        return pg;
    }
    const RqProg &sub_prog(int32_t s)
    {
        RqProg &pg = sub[s].prog;
        if (pg.done) return pg;
        const Term &T = bp.term(sub[s].term);
        Gen g{T, node_sid[sub[s].term], sub[s].node, pg};
        g.self = s;
        pg.seg_begin.assign(1, 0u);
        gen(g, sub[s].node);
        finish_prog(pg);
        return pg;
    }

    // ---- emission ----
    void push(const RRIns &x) { P.ins.push_back(x); }
    void push_nop()
    {
        RRIns x;
        std::memset(&x, 0, sizeof(x));
        x.w0 = RQ_NOP;
        push(x);
    }
    uint8_t col_of(int32_t ref)
    {
        if (ref >= 0) {
            auto it = colmap.find(ref);
            if (it == colmap.end()) { err = "internal: r8 column not staged"; return 0; }
            return (uint8_t)it->second;
        }
        const int32_t sl = sub[-1 - ref].slot;
        if (sl < 0) { err = "internal: r8 stored operand not resident"; return 0; }
        return (uint8_t)(n_staged + sl);
    }
    // the operations of up to eight programs of one shape, for one half
    void emit_ops(const RqProg *const *pg, int n, size_t op_begin = 0, size_t op_end = ~(size_t)0)
    {
        const size_t n_ops = std::min(pg[0]->ops.size(), op_end);
        for (size_t i = op_begin; i < n_ops; ++i) {
            const RqOp &o = pg[0]->ops[i];
            RRIns x;
            std::memset(&x, 0, sizeof(x));
            x.w0 = RQ_W0(o.op, o.mode);
            if (o.op == RQ_RARE) x.w0 |= ((uint32_t)(o.rare & 15u) << RQ_RARE_SHIFT) | ((o.rare & 16u) ? RQ_SWAP : 0u);
            const bool has_operand = o.op >= RQ_LD && o.op <= RQ_RARE;
            if (has_operand && o.mode == RQ_M) {
                uint8_t c8[8];
                for (int g = 0; g < 8; ++g) c8[g] = col_of(pg[g < n ? g : 0]->ops[i].ref);
                std::memcpy(&x.imm, c8, 8);
            } else if (has_operand && o.mode == RQ_K) {
                bool same = true;
                for (int g = 1; g < n; ++g) same = same && std::memcmp(&pg[g]->ops[i].k, &o.k, 8) == 0;
                if (same) {
                    x.imm = o.k;
                } else {
                    // one constant per row: four data slots behind the instruction, all five in one window
                    while ((P.ins.size() - (size_t)pc_begin) % RR_INS_WINDOW > RR_INS_WINDOW - 5) push_nop();
                    x.w0 = RQ_W0(o.op, RQ_C) | (x.w0 & ~0x3ffu);
                    push(x);
                    for (int d = 0; d < 4; ++d) {
                        RRIns c;
                        double kk[2];
                        for (int e = 0; e < 2; ++e) {
                            const int g = 2 * d + e;
                            kk[e] = pg[g < n ? g : 0]->ops[i].k;
                        }
                        std::memcpy(&c, kk, 16);
                        push(c);
                    }
                    continue;
                }
            }
            push(x);
        }
    }
    int32_t alloc_slot()
    {
        for (int32_t s = 0; s < slots_used; ++s)
            if (slot_sid[s] < 0 && slot_lock[s] == 0) return s;
        if (slots_used < slot_cap) {
            slot_sid.push_back(-1);
            slot_lock.push_back(0);
            slot_stamp.push_back(0);
            return slots_used++;
        }
        int32_t best = -1;
        for (int32_t s = 0; s < slots_used; ++s)
            if (slot_lock[s] == 0 && (best < 0 || slot_stamp[s] < slot_stamp[best])) best = s;
        if (best < 0) { err = "r8: out of tile slots"; return 0; }
        if (slot_sid[best] >= 0) sub[slot_sid[best]].slot = -1;
        slot_sid[best] = -1;
        return best;
    }
    // make stored sub-expression s resident (evaluating it as a uniform group when it is not) and lock its slot
    void ensure_locked(int32_t s)
    {
        if (!err.empty()) return;
        if (sub[s].slot >= 0) {
            slot_stamp[sub[s].slot] = clock++;
            slot_lock[sub[s].slot]++;
            return;
        }
        const RqProg &pg = sub_prog(s);
        if (!err.empty()) return;
        const RqProg *one[1] = {&pg};
        int32_t sl = -1;
        std::vector<int32_t> refs;
        for (size_t k = 0; k < pg.seg_begin.size() && err.empty(); ++k) {
            const size_t o0 = pg.seg_begin[k], o1 = k + 1 < pg.seg_begin.size() ? pg.seg_begin[k + 1] : pg.ops.size();
            refs.clear();
            for (size_t i = o0; i < o1; ++i) {
                const RqOp &o = pg.ops[i];
                if (o.mode == RQ_M && o.ref < 0 && o.op >= RQ_LD && o.op <= RQ_RARE && -1 - o.ref != s &&
                    std::find(refs.begin(), refs.end(), -1 - o.ref) == refs.end())
                    refs.push_back(-1 - o.ref);
            }
            for (int32_t c : refs) ensure_locked(c);
            if (!err.empty()) return;
            if (sl < 0) {
                sl = alloc_slot();
                if (!err.empty()) return;
                sub[s].slot = sl;
                slot_sid[sl] = s;
                slot_lock[sl] = 1;
            }
            {
                emit_ops(one, 1, o0, o1);
                RRIns x;
                std::memset(&x, 0, sizeof(x));
                x.w0 = RQ_W0(RQ_ST, RQ_M);
                uint8_t c8[8];
                for (int g = 0; g < 8; ++g) c8[g] = (uint8_t)(n_staged + sl);
                std::memcpy(&x.imm, c8, 8);
                push(x);
            }
            for (int32_t c : refs) slot_lock[sub[c].slot]--;
        }
        P.w_issued += pg.w;
        P.n_term_evals++;
        n_stored_evals++;
        slot_stamp[sl] = clock++;
    }
    void unlock(int32_t s) { slot_lock[sub[s].slot]--; }

    void emit_group(const Row *rows, int n)
    {
        const RqProg *pg[8];
        std::vector<int32_t> need;
        for (int g = 0; g < n; ++g) {
            pg[g] = &term_prog[rows[g].term];
            for (int32_t s : pg[g]->stored)
                if (std::find(need.begin(), need.end(), s) == need.end()) need.push_back(s);
        }
        for (int32_t s : need) ensure_locked(s);
        if (!err.empty()) return;
        uint64_t lo = 0, hi = 0;
        int n_want = 0;
        const int32_t id0 = P.n_dots;
        for (int g = 0; g < n; ++g)
            for (int o = 0; o < 10; ++o)
                if ((rows[g].want >> o) & 1u) {
                    const int bit = g * 10 + o;
                    if (bit < 64) lo |= 1ull << bit;
                    else hi |= 1ull << (bit - 64);
                    dots.set(rows[g].key[o], P.n_dots);
                    P.n_dots += 1;
                    P.n_dot_ins += 1;
                    ++n_want;
                }
        {
            emit_ops(pg, n);
            RRIns x;
            std::memset(&x, 0, sizeof(x));
            x.w0 = RQ_W0(RQ_GRAM, 0) | ((uint32_t)n << RQ_AUX_SHIFT);
            if ((P.ins.size() - (size_t)pc_begin) % RR_INS_WINDOW == RR_INS_WINDOW - 1) push_nop();
            x.w1 = (uint32_t)(id0 - dot_base);
            std::memcpy(&x.imm, &lo, 8);
            push(x);
            RRIns d;
            std::memset(&d, 0, sizeof(d));
            d.w0 = RQ_NOP;
            d.w1 = (uint32_t)hi;
            push(d);
        }
        for (int32_t s : need) unlock(s);
        for (int g = 0; g < n; ++g) P.w_issued += pg[g]->w;
        P.w_issued += n_want;
        P.n_term_evals += n;
        n_groups++;
        n_rows += n;
    }

    static bool same_shape(const RqProg &a, const RqProg &b)
    {
        if (a.ops.size() != b.ops.size()) return false;
        for (size_t i = 0; i < a.ops.size(); ++i) {
            const RqOp &x = a.ops[i], &y = b.ops[i];
            if (x.op != y.op || x.mode != y.mode || x.rare != y.rare) return false;
        }
        return true;
    }
    void flush_pool()
    {
        if (pool.empty() || !err.empty()) { pool.clear(); return; }
        const bool shape_first = std::getenv("RR_B200_R8_PRIMARY_FIRST") == nullptr;
        std::stable_sort(pool.begin(), pool.end(), [shape_first](const Row &a, const Row &b) {
            if (shape_first) {
                if (a.shape != b.shape) return a.shape < b.shape;
                return a.primary < b.primary;
            }
            if (a.primary != b.primary) return a.primary < b.primary;
            return a.shape < b.shape;
        });
        const int cap = std::max(1, slot_cap - 4);
        size_t i = 0;
        std::vector<int32_t> uni;
        while (i < pool.size() && err.empty()) {
            uni = term_prog[pool[i].term].stored;
            if ((int)uni.size() > cap) { err = "r8: a row reads more stored sub-expressions than the tile has slots"; break; }
            size_t j = i + 1;
            while (j < pool.size() && j - i < 8 && pool[j].shape == pool[i].shape &&
                   same_shape(term_prog[pool[i].term], term_prog[pool[j].term])) {
                size_t extra = 0;
                for (int32_t s : term_prog[pool[j].term].stored)
                    if (std::find(uni.begin(), uni.end(), s) == uni.end()) ++extra;
                if ((int)(uni.size() + extra) > cap) break;
                for (int32_t s : term_prog[pool[j].term].stored)
                    if (std::find(uni.begin(), uni.end(), s) == uni.end()) uni.push_back(s);
                ++j;
            }
            emit_group(&pool[i], (int)(j - i));
            i = j;
        }
        pool.clear();
    }

    void add_row(int32_t u, uint32_t want, const uint64_t *key)
    {
        const RqProg &pg = row_prog(u);
        if (!err.empty()) return;
        Row r;
        r.term = u;
        r.want = want;
        std::memcpy(r.key, key, sizeof(r.key));
        r.primary = pg.stored.empty() ? -1 : pg.stored[0];
        r.shape = pg.shape;
        pool.push_back(r);
        for (int o = 0; o < 10; ++o)
            if ((want >> o) & 1u) dots.set(key[o], -2);  // planned; the id follows when the group is emitted
    }

    void pin_global(int j, int32_t gcol)
    {
        RRIns x;
        std::memset(&x, 0, sizeof(x));
        x.w0 = RQ_W0(RQ_PINB, 0) | ((uint32_t)j << RQ_AUX_SHIFT) | RQ_PIN_GLOBAL;
        x.w1 = (uint32_t)gcol;
        push(x);
        pin_term[j] = PIN_RESERVED;
    }
    void unpin_all() { ++epoch; }
    int pin_partner(int32_t v)
    {
        if (term_pin[v] >= 0) {
            const int j = term_pin[v];
            pin_stamp[j] = clock++;
            pin_hold[j] = epoch;
            return j;
        }
        int j = -1;
        for (int i = 0; i < RR_NPIN; ++i)
            if (pin_term[i] == -1) { j = i; break; }
        if (j < 0) {
            uint64_t best = ~0ull;
            for (int i = 0; i < RR_NPIN; ++i)
                if (pin_term[i] >= 0 && pin_hold[i] != epoch && pin_stamp[i] < best) {
                    best = pin_stamp[i];
                    j = i;
                }
        }
        if (j < 0) { err = "internal: no pin available for a reduction partner"; return 0; }
        flush_pool();
        if (!err.empty()) return 0;
        const Term &T = bp.term(v);
        const TermNode &root = T.nodes.back();
        uint32_t col;
        int32_t locked = -1;
        if (root.op == RR_OP_VAR) {
            col = col_of(root.var);
        } else if (root.op == RR_OP_CONST) {
            // a bare constant as a partner: through a temporary slot
            const int32_t sl = alloc_slot();
            if (!err.empty()) return 0;
            RRIns l, st;
            std::memset(&l, 0, sizeof(l));
            std::memset(&st, 0, sizeof(st));
            l.w0 = RQ_W0(RQ_LD, RQ_K);
            l.imm = root.cval;
            push(l);
            st.w0 = RQ_W0(RQ_ST, RQ_M);
            uint8_t c8[8];
            for (int g = 0; g < 8; ++g) c8[g] = (uint8_t)(n_staged + sl);
            std::memcpy(&st.imm, c8, 8);
            push(st);
            col = (uint32_t)(n_staged + sl);  // the slot stays free: nothing is emitted between here and the RQ_PINB
        } else {
            locked = node_sid[v][T.nodes.size() - 1];
            ensure_locked(locked);
            if (!err.empty()) return 0;
            col = (uint32_t)(n_staged + sub[locked].slot);
        }
        RRIns x;
        std::memset(&x, 0, sizeof(x));
        x.w0 = RQ_W0(RQ_PINB, 0) | ((uint32_t)j << RQ_AUX_SHIFT);
        x.w1 = col;
        push(x);
        if (locked >= 0) unlock(locked);
        if (pin_term[j] >= 0) term_pin[pin_term[j]] = -1;
        pin_term[j] = v;
        term_pin[v] = (int8_t)j;
        pin_stamp[j] = clock++;
        pin_hold[j] = epoch;
        return j;
    }

    void close()
    {
        flush_pool();
        RRIns x;
        std::memset(&x, 0, sizeof(x));
        x.w0 = RQ_END;
        push(x);
        RRChunk c;
        std::memset(&c, 0, sizeof(c));
        c.pc_begin = pc_begin;
        c.n_ins = (int32_t)P.ins.size() - pc_begin;
        c.dot_base = dot_base;
        c.n_dots = P.n_dots - dot_base;
        c.col_begin = col_begin;
        c.n_cols = n_staged;
        P.chunks.push_back(c);
        P.max_tile_cols = std::max(P.max_tile_cols, n_staged + std::max(slots_used, 1));
        P.n_gram_groups += n_groups;
        P.n_gram_rows += n_rows;
        P.n_stored_evals += n_stored_evals;
        P.r8 = true;
        if (std::getenv("RR_B200_R8_DEBUG")) {
            size_t n_st = 0, n_ev = 0;
            int hist[17] = {0};
            for (size_t i = 0; i < sub.size(); ++i) {
                const Sub &x = sub[i];
                if (is_stored((int32_t)i)) ++n_st;
                if (x.prog.done) { ++n_ev; hist[std::min(x.uses, 16)]++; }
            }
            for (int i = 0; i <= 16; ++i)
                if (hist[i]) std::fprintf(stderr, "r8:   evaluated stored sub-expressions with %d uses: %d\n", i, hist[i]);
            std::fprintf(stderr, "r8: %zu sub-expressions, %zu stored, %zu of them evaluated at least once, %llu evaluations, %d slots used\n",
                         sub.size(), n_st, n_ev, (unsigned long long)n_stored_evals, slots_used);
        }
    }
};

}  // namespace

std::string BatchPlanner::plan_gram_r8(const PlanLimits &lim, const ColIds &cols, const std::vector<int32_t> *subset, SweepPlan &P,
                                       std::vector<int32_t> &cand_dot, std::vector<int32_t> &cand_dot_begin) const
{
    std::vector<int32_t> list;
    if (subset) list = *subset;
    else {
        list.resize(b_->n_cand);
        for (int32_t c = 0; c < b_->n_cand; ++c) list[c] = c;
    }
    std::vector<Unit> units(list.size());
    for (size_t i = 0; i < list.size(); ++i) {
        const int32_t c = list[i];
        for (int32_t t = b_->cand_term_begin[c]; t < b_->cand_term_begin[c + 1]; ++t)
            units[i].terms.push_back(term_id_[t]);
        units[i].w = cand_w_[c];
        if ((int32_t)units[i].terms.size() > RR_NPIN) return "candidate too wide for an R8 plan";
    }
    std::vector<ChunkSpec> specs;
    std::vector<int32_t> always;
    PlanLimits one = lim;
    one.target_chunks = 1;
    std::string err = cut_chunks(*this, units, one, always, 3, specs);
    if (!err.empty()) return err;
    if (specs.size() != 1) return "r8: the neighbourhood does not fit one chunk";

    cand_dot.clear();
    cand_dot_begin.assign(1, 0);
    DotMap dots((size_t)units.size() * 8 + 64);
    const int64_t KEY_YC = -1, KEY_ONE = -2;
    auto key = [](int64_t a, int64_t b) -> uint64_t {
        if (a > b) std::swap(a, b);
        return ((uint64_t)(uint32_t)(int32_t)a << 32) | (uint64_t)(uint32_t)(int32_t)b;
    };
    R8Builder ch(*this, b_, P, dots, lim, specs[0].cols);
    if (ch.slot_cap < 3) return "r8: tile too small";
    for (const Unit &u : units)
        for (int32_t t : u.terms) ch.register_term(t);
    ch.pin_global(0, cols.yc);  // pin 0 = the centred target, for the whole chunk
    auto missing = [&](int64_t a, int64_t b) { return dots.find(key(a, b)) == nullptr; };
    std::vector<int32_t> N, done;
    for (size_t ui = 0; ui < units.size(); ++ui) {
        const std::vector<int32_t> &T = units[ui].terms;
        const int32_t m = (int32_t)T.size();
        N.clear();
        for (int32_t i = 0; i < m; ++i) {
            bool miss = missing(T[i], KEY_YC) || missing(T[i], KEY_ONE);
            for (int32_t j = 0; j < m && !miss; ++j) miss = missing(T[i], T[j]);
            if (miss && std::find(N.begin(), N.end(), T[i]) == N.end()) N.push_back(T[i]);
        }
        // pinned terms first: what is new in this candidate then meets all of its partners in its one row
        // ... then terms whose own reductions exist already (they only lack pairs with what is new here: as partners
        // they are pinned once for the whole family of candidates that share them), the new terms last
        std::stable_partition(N.begin(), N.end(), [&](int32_t u) { return !missing(u, u); });
        std::stable_partition(N.begin(), N.end(), [&](int32_t u) { return ch.term_pin[u] >= 0; });
        ch.unpin_all();
        done.clear();
        for (size_t q = 0; q < N.size(); ++q) {
            const int32_t u = N[q];
            const bool self = missing(u, u), one_ = missing(u, KEY_ONE), with_yc = missing(u, KEY_YC);
            bool any = self || one_ || with_yc;
            for (int32_t v : done)
                if (v != u && missing(u, v)) { any = true; break; }
            if (!any) { done.push_back(u); continue; }
            uint32_t want = with_yc ? 1u : 0u;
            uint64_t k10[10] = {0};
            if (with_yc) k10[0] = key(u, KEY_YC);
            for (int32_t v : done) {
                if (v == u || !missing(u, v)) continue;
                const int j = ch.pin_partner(v);
                if (!ch.err.empty()) return ch.err;
                if ((want >> j) & 1u) return "internal: duplicate pinned partner";
                want |= 1u << j;
                k10[j] = key(u, v);
            }
            if (self) { want |= 1u << 8; k10[8] = key(u, u); }
            if (one_) { want |= 1u << 9; k10[9] = key(u, KEY_ONE); }
            ch.add_row(u, want, k10);
            if (!ch.err.empty()) return ch.err;
            done.push_back(u);
        }
    }
    ch.close();
    if (!ch.err.empty()) return ch.err;
    for (size_t ui = 0; ui < units.size(); ++ui) {
        const std::vector<int32_t> &T = units[ui].terms;
        const int32_t m = (int32_t)T.size();
        auto put = [&](uint64_t k) -> bool {
            const int32_t *it = dots.find(k);
            if (!it || *it < 0) return false;
            cand_dot.push_back(*it);
            return true;
        };
        for (int32_t i = 0; i < m; ++i)
            for (int32_t j = i; j < m; ++j)
                if (!put(key(T[i], T[j]))) return "internal: missing Gram dot";
        for (int32_t i = 0; i < m; ++i)
            if (!put(key(T[i], KEY_YC))) return "internal: missing Gram dot";
        for (int32_t i = 0; i < m; ++i)
            if (!put(key(T[i], KEY_ONE))) return "internal: missing Gram dot";
        cand_dot_begin.push_back((int32_t)cand_dot.size());
    }
    return "";
}

}  // namespace rr

// ---------------------------------------------------------------------------------------------
// host-only tooling entry points (include/rr_b200.h)
// ---------------------------------------------------------------------------------------------
#include <chrono>
#include <cstdlib>
#include <thread>

namespace {
template <typename T> T *dup_vec(const std::vector<T> &v)
{
    T *p = (T *)std::malloc(sizeof(T) * std::max<size_t>(v.size(), 1));
    if (p && !v.empty()) std::memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}
}  // namespace

extern "C" int rr_debug_plan_batch(const rr_batch *batch, int32_t d, int32_t kind, int32_t tile_cols,
                                   int32_t max_slots, int32_t target_chunks, int32_t no_cse, int32_t n_pins,
                                   const double *coef_snapped, rr_debug_plan *out)
{
    if (!batch || !out) return RR_ERR_INVALID;
    std::memset(out, 0, sizeof(*out));
    rr::BatchPlanner bp(batch, d);
    const auto tm0 = std::chrono::steady_clock::now();
    std::string err = bp.analyse(no_cse != 0);
    const auto tm1 = std::chrono::steady_clock::now();
    struct Rep {
        std::chrono::steady_clock::time_point a, b;
        ~Rep()
        {
            if (std::getenv("RR_B200_PLAN_TIMING"))
                std::fprintf(stderr, "analyse %.3f ms, plan %.3f ms\n", std::chrono::duration<double, std::milli>(b - a).count(),
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - b).count());
        }
    } rep{tm0, tm1};
    rr::SweepPlan P;
    std::vector<int32_t> tab, tab_begin;
    if (err.empty()) {
        rr::PlanLimits lim;
        lim.tile_cols = tile_cols;
        if (max_slots > 0) lim.max_slots = max_slots;
        lim.target_chunks = std::max(1, target_chunks);
        lim.no_cse = no_cse != 0;
        lim.n_pins = n_pins;
        lim.mdot_rows = std::getenv("RR_B200_DEBUG_NO_MDROWS") == nullptr;  // as for the 4-samples-per-thread core
        lim.fuse = std::getenv("RR_B200_DEBUG_NO_FUSE") == nullptr;
        rr::ColIds cols{d, d + 1};
        switch (kind) {
        case 0: err = bp.plan_gram(lim, cols, nullptr, false, P, tab, tab_begin); break;
        case 1: err = bp.plan_gram(lim, cols, nullptr, true, P, tab, tab_begin); break;
        case 2: err = bp.plan_eval(lim, cols, false, P, tab); break;
        case 3: err = bp.plan_eval(lim, cols, true, P, tab); break;
        case 4: err = bp.plan_materialise(lim, cols, P); break;
        case 5: {
            if (!coef_snapped) { err = "residual plan needs coefficients"; break; }
            std::vector<int32_t> all(batch->n_cand);
            for (int32_t c = 0; c < batch->n_cand; ++c) all[c] = c;
            err = bp.plan_residual(lim, cols, all, coef_snapped, P, tab, tab_begin);
            break;
        }
        case 6:
            lim.g8 = true;
            lim.ins_window = RR_G8_INS_WINDOW;
            err = bp.plan_gram_g8(lim, cols, nullptr, P, tab, tab_begin);
            break;
        case 7: err = bp.plan_gram_r8(lim, cols, nullptr, P, tab, tab_begin); break;
        default: err = "bad kind";
        }
    }
    if (!err.empty()) {
        std::snprintf(out->error, sizeof(out->error), "%s", err.c_str());
        return RR_ERR_INVALID;
    }
    out->n_ins = (int64_t)P.ins.size();
    out->n_chunks = (int64_t)P.chunks.size();
    out->n_cols = (int64_t)P.cols.size();
    out->n_tab = (int64_t)tab.size();
    out->n_tab_begin = (int64_t)tab_begin.size();
    out->n_term_ids = (int64_t)bp.term_ids().size();
    out->ins = dup_vec(P.ins);
    out->chunks = dup_vec(P.chunks);
    out->cols = dup_vec(P.cols);
    out->tab = dup_vec(tab);
    out->tab_begin = dup_vec(tab_begin);
    out->term_ids = dup_vec(bp.term_ids());
    out->n_dots = P.n_dots;
    out->max_tile_cols = P.max_tile_cols;
    out->n_terms_distinct = bp.n_terms_distinct();
    out->w_issued = P.w_issued;
    out->w_contract = bp.w_contract();
    return RR_OK;
}

// Test hook (host only): plan_gram of the two halves of a batch on two threads at once and one after the
// other must give the same instruction streams and tables (run_gram plans its second half on a helper thread).
// Returns 0 when they agree, 1 when they differ, RR_ERR_INVALID on a malformed batch.
extern "C" int rr_debug_plan_concurrency_check(const rr_batch *batch, int32_t d, int32_t tile_cols)
{
    if (!batch || batch->n_cand < 2) return RR_ERR_INVALID;
    rr::BatchPlanner bp(batch, d);
    if (!bp.analyse(false).empty()) return RR_ERR_INVALID;
    rr::PlanLimits lim;
    lim.tile_cols = tile_cols;
    lim.mdot_rows = true;
    rr::ColIds cols{d, d + 1};
    const int32_t half = batch->n_cand / 2;
    std::vector<int32_t> la(half), lb(batch->n_cand - half);
    for (int32_t c = 0; c < half; ++c) la[c] = c;
    for (int32_t c = half; c < batch->n_cand; ++c) lb[c - half] = c;
    struct Out {
        rr::SweepPlan P;
        std::vector<int32_t> tab, tab_begin;
        std::string err;
    } seq[2], par[2];
    seq[0].err = bp.plan_gram(lim, cols, &la, false, seq[0].P, seq[0].tab, seq[0].tab_begin);
    seq[1].err = bp.plan_gram(lim, cols, &lb, false, seq[1].P, seq[1].tab, seq[1].tab_begin);
    std::thread t([&]() { par[1].err = bp.plan_gram(lim, cols, &lb, false, par[1].P, par[1].tab, par[1].tab_begin); });
    par[0].err = bp.plan_gram(lim, cols, &la, false, par[0].P, par[0].tab, par[0].tab_begin);
    t.join();
    for (int h = 0; h < 2; ++h) {
        if (!seq[h].err.empty() || !par[h].err.empty()) return RR_ERR_INVALID;
        if (seq[h].P.ins.size() != par[h].P.ins.size() || seq[h].tab != par[h].tab || seq[h].tab_begin != par[h].tab_begin ||
            seq[h].P.n_dots != par[h].P.n_dots)
            return 1;
        if (!seq[h].P.ins.empty() && std::memcmp(seq[h].P.ins.data(), par[h].P.ins.data(), seq[h].P.ins.size() * sizeof(RRIns)) != 0)
            return 1;
    }
    return 0;
}

extern "C" int rr_debug_const_terms(const rr_batch *batch, int32_t d, uint32_t *out)
{
    if (!batch || !out) return RR_ERR_INVALID;
    try {
        rr::BatchPlanner bp(batch, d);
        if (!bp.analyse(false).empty()) return RR_ERR_INVALID;
        for (int32_t c = 0; c < batch->n_cand; ++c) out[c] = bp.cand_const_mask(c);
    } catch (...) {
        return RR_ERR_NOMEM;
    }
    return RR_OK;
}

extern "C" void rr_debug_plan_free(rr_debug_plan *p)
{
    if (!p) return;
    std::free(p->ins);
    std::free(p->chunks);
    std::free(p->cols);
    std::free(p->tab);
    std::free(p->tab_begin);
    std::free(p->term_ids);
    std::memset(p, 0, sizeof(*p));
}
