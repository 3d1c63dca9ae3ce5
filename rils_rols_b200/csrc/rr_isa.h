// rr_isa.h — the engine's internal instruction set.
//
// The C ABI carries plain postfix trees (include/rr_b200.h). The planner (rr_plan.cpp)
// recompiles a whole neighbourhood into ONE instruction stream for the sweep kernel
// (rr_sweep.cuh): an accumulator machine whose per-sample state is a single fp64
// register `t` (S samples per thread, so S registers) plus columns of a shared-memory
// tile. Tile columns hold the staged feature columns of X, y, the centred y, and value
// slots (cached term values, spill temporaries, residuals). There is no dynamic stack:
// the planner allocates spill slots statically.
//
// Each node of node::evaluate_inner (/root/reference/rils_rols_cpp/node.cpp:23-95) maps
// to one instruction with IEEE-identical operand order; leaves are folded into their
// parent as an operand (COL or CONST), so about half of the reference's node evaluations
// cost no instruction at all.
#ifndef RR_ISA_H
#define RR_ISA_H

#include <stdint.h>

struct RRIns {
    uint32_t w0;  // opcode | flags
    uint32_t w1;  // tile column index (operand / destination) or packed DOT operands
    double imm;   // constant operand / AXPY coefficient
};
static_assert(sizeof(RRIns) == 16, "RRIns must be 16 bytes");

enum RRInsOp : uint32_t {
    RI_END = 0,
    RI_LOAD,   // t = src
    RI_ST,     // tile[w1] = t
    RI_STG,    // out[w1][sample] = t   (materialise a column in global memory)
    // binary: t = t op src   (RF_SWAP: t = src op t)
    RI_ADD,
    RI_SUB,
    RI_MUL,
    RI_DIV,
    RI_POW,
    RI_LT,
    RI_GT,
    RI_EQ,
    RI_NE,
    RI_MIN,
    RI_MAX,
    RI_AXPY,   // t = t + imm * tile[w1]   (product rounded, then sum: the c*term + ... chain of
               //                           rils_rols_cpp.cpp:503-510)
    // unary: t = f(t)
    RI_SIN,
    RI_COS,
    RI_LN,
    RI_EXP,
    RI_SQRT,
    RI_SQR,
    // reductions over the samples: out[dot_id] += sum_s a_s * b_s ; dot ids are implicit,
    // consecutive in program order from the chunk's dot_base
    RI_DOT,
    RI_DOTDD,  // same, accumulated in double-double (two outputs: hi, lo)
    // classifier metrics of t (rils_rols_cpp.cpp:51-86): three consecutive dot outputs
    RI_CLSMET,
    RI_OPCOUNT
};

// w0 layout: bits 0-7 opcode, bit 8 RF_CONST (operand is imm, else tile column w1),
// bit 9 RF_SWAP. DOT: bits 8-9 = kind of a, bits 10-11 = kind of b;
// w1 = a column | b column << 16.
enum : uint32_t {
    RF_CONST = 1u << 8,
    RF_SWAP = 1u << 9,
};
enum RRDotKind : uint32_t { RD_COL = 0, RD_ONE = 1, RD_TOS = 2 };
#define RR_DOT_W0(op, ka, kb) ((uint32_t)(op) | ((uint32_t)(ka) << 8) | ((uint32_t)(kb) << 10))
#define RR_DOT_KA(w0) (((w0) >> 8) & 3u)
#define RR_DOT_KB(w0) (((w0) >> 10) & 3u)

// One independently schedulable piece of a sweep: its own staged columns, slot state and
// dot range. Large-n sweeps use one chunk (maximal sharing); small-n sweeps are cut into
// many chunks so that every SM has work.
struct RRChunk {
    int32_t pc_begin;   // first instruction (the chunk ends with RI_END)
    int32_t dot_base;   // first dot output id
    int32_t n_dots;     // dot outputs of this chunk (DOTDD counts 2, CLSMET counts 3)
    int32_t col_begin;  // into the plan's staged-column list
    int32_t n_cols;     // staged global columns; slots follow them in the tile
    int32_t reserved[3];
};
static_assert(sizeof(RRChunk) == 32, "RRChunk must be 32 bytes");

#endif  // RR_ISA_H
