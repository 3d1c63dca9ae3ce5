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
// parent as an operand (tile column or immediate), so about half of the reference's node
// evaluations cost no instruction at all.
//
// Opcodes are dense (one jump table in the kernel) and specialised by operand kind
// (_C immediate, _M tile column) and operand order (R* = reversed: t = src op t), so a
// dispatched case does no further flag tests. All tile-column forms sit at the top of the
// opcode range: the dispatcher starts their operand load before it branches.
#ifndef RR_ISA_H
#define RR_ISA_H

#include <stdint.h>

struct RRIns {
    uint32_t w0;  // opcode | aux << 8
    uint32_t w1;  // tile column index (operand / destination); 0 when unused
    double imm;   // constant operand / AXPY coefficient
};
static_assert(sizeof(RRIns) == 16, "RRIns must be 16 bytes");

// Besides tile columns the machine has RR_NPIN *pinned value registers* per sample ("pins"): the
// planner keeps the terms that are reduction partners of many others there (for a local-search
// neighbourhood: the base solution's terms and the centred target), so that the reductions of a
// freshly evaluated term against them read no shared memory at all.
#define RR_NPIN 8
// ... plus RR_NREG - RR_NPIN *cache registers* of the same kind that PIN / LDP / USEP can address but
// reductions cannot: the planner parks expensive sub-expressions there that the next few terms share
// (neighbours of one tree node keep its sibling subtrees: exp(x8)*x9 -> exp(x8)*x3, exp(x8)*sqrt(x9), ...)
#define RR_NREG 10
// The kernel streams a chunk's instructions through shared memory in windows of RR_INS_WINDOW, counted
// from the chunk's first instruction. USEP and its consumer must sit in the same window (the redirected
// operand lives in registers that do not survive a window switch): the planner pads with RI_NOP.
#define RR_INS_WINDOW 64
#define RR_G8_INS_WINDOW 32  // G8 plans (rr_sweep_g8.cuh)

enum RRInsOp : uint32_t {
    RI_END = 0,
    RI_WINEND,  // never planned: the kernel's sentinel behind each instruction window
    RI_LOAD_C,  // t = imm
    RI_ST,      // tile[w1] = t
    RI_STG,     // out[w1][sample] = t   (materialise a column in global memory)
    RI_LDG,     // t = X[w1][sample]     (engine column straight from global memory, not staged)
    RI_NOP,     // padding (see RR_INS_WINDOW)
    RI_COMBINE, // plans for the 4-samples-per-thread core: the block's warps add up the 32 reductions parked in their
                // staging rows and RED them to the block's accumulator row. The planner places it behind the
                // instruction whose flush filled the fourth staging slot (flushes happen at statically known
                // points), which keeps the block barrier out of the reduction handlers. Every other interpreter
                // combines inside its flush and executes this as a NOP.
    RI_ADD_C, RI_SUB_C, RI_RSUB_C, RI_MUL_C, RI_DIV_C, RI_RDIV_C,  // t = t op imm / imm op t (R*)
    RI_SIN, RI_COS, RI_LN, RI_EXP, RI_SQRT, RI_SQR,  // t = f(t)
    // rarely generated operators share one case: aux = RRRareOp | RB_CONST | RB_SWAP
    RI_RARE,
    // Reductions over the samples with a = t: aux bit 0 = t.t, bit 1 = sum(t), aux bits 8-15 = mask
    // of pins to reduce against. Outputs in that order (pins ascending), ids implicit and
    // consecutive from the chunk's dot_base; at most RR_MDOT_MAX_OUT outputs per instruction.
    // aux bits 16-19 = 1 + j: afterwards pin[j] = t (the planner's "PIN j; MDOT" pair in one dispatch;
    // j is never in the mask).
    RI_MDOT,
    // ---- G8 plans (Gram reductions on the FP64 tensor-core path, rr_sweep_g8.cuh) ----
    // Up to 8 freshly evaluated terms sit in tile slots ("rows"); ONE instruction reduces all of them against the 8
    // pins with DMMA.8x8x4 (a warp-level 8 terms x 8 pins x 4 samples product-accumulate per instruction: the
    // accumulator fragment IS the warp total, nothing is transposed) plus each row's t.t and sum(t).
    // w0 = RI_GRAM8 | n_rows << 8; w1, lo32(imm), hi32(imm) = 80 "wanted" bits, bit 10 g + o of row g: o = 0..7 pin o,
    // 8 = t.t, 9 = unused (a row's sum(t) is its product with pin 7, which G8 plans hold at a column of ones: the
    // kernel forms no sum of its own and the planner never sets the bit); output ids are consecutive from the chunk's running count in bit order. A data slot
    // (RI_NOP, aux RR_GRAM_COLS) follows: bytes 4..11 = the tile column of each row (unused rows repeat row 0).
    RI_GRAM8,
    // pin j (aux = j) <- engine column w1 read from global memory, as a reduction partner only (B fragment)
    RI_PINBG,
    // double-double reductions (escalation plans): same encoding as RI_MDOT, each output accumulated in
    // double-double and taking two ids (hi, lo)
    RI_MDOTDD,
    // classifier metrics of t against y = tile[w1] (rils_rols_cpp.cpp:51-86): three outputs
    RI_CLSMET,
    // value registers j < RR_NREG: PIN j: reg[j] = t;  LDP j: t = reg[j];  USEP j: the NEXT instruction (a tile-column
    // operand form) takes pin[j] as its operand instead of the tile column
    RI_PIN0,
    RI_LDP0 = RI_PIN0 + RR_NREG,
    RI_USEP0 = RI_LDP0 + RR_NREG,
    // ---- super-instructions (planner peephole, rr_plan.cpp close()): frequent sequences in one dispatch.
    // Each is bit-identical to the sequence it replaces (same operations, same operand order).
    RI_MULP0 = RI_USEP0 + RR_NREG,   // t = t * reg[j]        (USEP j; MUL_M)
    RI_DIVP0 = RI_MULP0 + RR_NREG,   // t = t / reg[j]        (USEP j; DIV_M)
    RI_RDIVP0 = RI_DIVP0 + RR_NREG,  // t = reg[j] / t        (USEP j; RDIV_M)
    RI_CMULP0 = RI_RDIVP0 + RR_NREG, // t = imm * reg[j]      (LOAD_C; USEP j; MUL_M)
    RI_CDIVP0 = RI_CMULP0 + RR_NREG, // t = imm / reg[j]      (LOAD_C; USEP j; DIV_M)
    // ---- forms with a tile-column operand tile[w1]: everything from RI_FIRST_M on ----
    RI_FIRST_M = RI_CDIVP0 + RR_NREG,
    RI_LOAD_M = RI_FIRST_M,  // t = tile[w1]
    RI_ADD_M, RI_SUB_M, RI_RSUB_M, RI_MUL_M, RI_DIV_M, RI_RDIV_M,  // t = t op tile[w1] / tile[w1] op t (R*)
    RI_AXPY,    // t = t + imm * tile[w1]  (product rounded, then sum: the c*term + ... chain of
                //                          rils_rols_cpp.cpp:503-510)
    RI_DOTM,    // one reduction: t . tile[w1]
    RI_DOTMDD,  // the same in double-double (two ids)
    // super-instructions with a tile-column operand; a second column index rides in the low word of imm
    RI_CMUL_M,    // t = imm * tile[w1]                          (LOAD_C; MUL_M)
    RI_CDIV_M,    // t = imm / tile[w1]                          (LOAD_C; DIV_M)
    RI_MUL_MM,    // t = tile[w1] * tile[lo32(imm)]              (LOAD_M; MUL_M)
    RI_MUL_M_ST,  // t = t * tile[w1]; tile[lo32(imm)] = t       (MUL_M; ST)
    RI_MUL_MMM,   // t = (tile[w1] * tile[lo32(imm)]) * tile[hi32(imm)]   (LOAD_M; MUL_M; MUL_M), G8 plans only
    RI_LDPMUL_M0,                             // t = reg[j] * tile[w1]   (LDP j; MUL_M)
    RI_LDPDIV_M0 = RI_LDPMUL_M0 + RR_NREG,    // t = reg[j] / tile[w1]   (LDP j; DIV_M)
    RI_LDMDIVP0 = RI_LDPDIV_M0 + RR_NREG,     // t = tile[w1] / reg[j]   (LOAD_M; USEP j; DIV_M)
    // G8 plans: pin j <- tile[w1], both as an operand register (like RI_PIN) and as a reduction partner (the lanes that
    // own pin j in the DMMA B fragment reload their 32 samples from the column)
    RI_PINB0 = RI_LDMDIVP0 + RR_NREG,
    RI_LAST_M = RI_PINB0 + RR_NPIN - 1,
    RI_OPCOUNT = RI_LAST_M + 1
};

enum RRRareOp : uint32_t { RR_POW = 0, RR_LT, RR_GT, RR_EQ, RR_NE, RR_MIN, RR_MAX };
enum : uint32_t {
    RB_CONST = 1u << 4,  // operand is imm, else tile[w1]
    RB_SWAP = 1u << 5,   // t = src op t
    MD_SELF = 1u << 0,
    MD_ONE = 1u << 1,
    RR_MDOT_MAX_OUT = 8,  // outputs of one RI_MDOT (the kernel's reduction ring drains in groups of 8)
};
// "X; MDOT" in one dispatch: the instruction is X (aux otherwise unused) with RR_THEN_MDOT set and the
// MDOT's aux bits OR-ed into w0; X's handler ends in the MDOT handler instead of a dispatch. Only for the
// multiplication / division forms below (what a term's last operation is in a local-search neighbourhood).
// Data slot behind an instruction that ends in an RI_MDOT (plans for the 4-samples-per-thread core):
// opcode RI_NOP with aux = RR_MDOT_ROWS; bytes 4..13 of the slot (w1, then imm) hold the ring row (0..15)
// of the 10 potential outputs self, one, pins 0..7. The core reads them from its prefetch registers and
// skips the slot; every other interpreter executes the slot as the NOP it is and derives the rows from
// its running count (the two agree by construction; tests/isa_emu.py checks it).
#define RR_MDOT_ROWS 1u
#define RR_GRAM_COLS 2u
#define RR_THEN_MDOT 0x8000u
// G8 plans never hold an RI_MDOT; there the same bit means "X; ST c": the handler of X stores t to tile column c
// (bits 16-23 of w0) before it dispatches. Same set of carriers (rr_md_fusable).
#define RR_THEN_ST 0x8000u
#define RR_THEN_ST_COL(w0) (((w0) >> 16) & 0xffu)
#ifdef __CUDACC__
#define RR_HD __host__ __device__
#else
#define RR_HD
#endif
RR_HD static inline bool rr_md_fusable(uint32_t op)
{
    return op == RI_MUL_M || op == RI_DIV_M || op == RI_RDIV_M || op == RI_DIV_C || op == RI_RDIV_C ||
           (op >= RI_MULP0 && op < RI_FIRST_M) || op == RI_CMUL_M || op == RI_CDIV_M || op == RI_MUL_MM || op == RI_MUL_MMM ||
           (op >= RI_LDPMUL_M0 && op < RI_PINB0);
}
#define RR_W0(op, aux) ((uint32_t)(op) | ((uint32_t)(aux) << 8))
#define RR_OP(w0) ((w0) & 0xffu)
#define RR_AUX(w0) ((w0) >> 8)
#define RR_MDOT_MASK(w0) (((w0) >> 16) & 0xffu)  /* pins to reduce against */

// ---- R8 plans: the row machine (rr_sweep_r8.cuh, BatchPlanner::plan_gram_r8) -------------------------------------
// A second, independent instruction set for the Gram pass of large-n neighbourhoods. Where the accumulator machine
// above gives every THREAD four samples and evaluates one term at a time, the row machine gives every LANE one of
// eight terms ("rows") of the same shape: lane (g, q) = (lane >> 2, lane & 3) of a warp evaluates row g at the
// samples 4 s + q, s = 0..15, of the warp's 64 samples - which is exactly the A fragment of
// mma.m8n8k4.f64 (rows x samples), so a freshly evaluated group of rows is reduced against the eight pins (the B
// fragment, as in G8 plans) straight from registers: no row is ever stored. One dispatched operation works on 16
// values per lane, so the interpreter's overhead per FP64 operation is a quarter of the accumulator machine's, and
// eight different rows share it. A block of four warps sweeps tiles of 256 samples, which leaves room for twice as
// many tile columns (stored sub-expressions) as the 512-sample tiles of the other kernels.
// Registers per lane: t[16] (accumulator), u[16] (second operand of trees whose two sides both need evaluating),
// pb[16] (pin g at the lane's 16 samples). Operand modes of LD and the binary operations:
//   RQ_M  imm = eight tile-column indices, byte g for row g (rows of one group differ in their operands only)
//   RQ_K  imm = one constant for all rows
//   RQ_C  one constant per row: four data slots follow, row g's constant at byte 16 + 8 g from the instruction
//   RQ_U  the register u
// Shared sub-expressions (and pins) are evaluated once as a "uniform" group - all eight rows identical - and stored
// to a tile slot with RQ_ST; lanes that read the same column are served by one shared-memory broadcast.
enum RQOp : uint32_t {
    RQ_END = 0,      // == RI_END
    RQ_WINEND = 1,   // == RI_WINEND
    RQ_NOP = 2,
    RQ_LD,           // t = operand
    RQ_ADD, RQ_SUB, RQ_RSUB, RQ_MUL, RQ_DIV, RQ_RDIV,  // t = t op operand / operand op t (R*)
    RQ_RARE,         // bits 12-15 = RRRareOp, bit 11 = swap (t = operand op t)
    RQ_SIN, RQ_COS, RQ_LN, RQ_EXP, RQ_SQRT, RQ_SQR,    // t = f(t)
    RQ_TU,           // u = t
    RQ_ST,           // tile[byte g of imm] = t
    // reduce the group in t against the 8 pins, itself and ones (DMMA), bits 16-23 = rows in the group:
    // w1 = first output id (relative to the chunk's dot_base), imm = wanted bits 0-63, bit 10 g + o of row g
    // (o = 0..7 pin o, 8 = t.t, 9 = sum t), and a data slot (RQ_NOP) follows whose w1 holds wanted bits 64-79;
    // the wanted outputs take consecutive ids in bit order
    RQ_GRAM,
    // pin (bits 16-23) <- tile column w1 (RQ_PIN_GLOBAL: engine column w1 read from global memory): the lanes
    // g == pin reload their 16 B-fragment values
    RQ_PINB,
    RQ_OPCOUNT
};
enum : uint32_t {
    RQ_M = 0, RQ_K = 1, RQ_U = 2, RQ_C = 3,
    RQ_MODE_SHIFT = 8,
    RQ_SWAP = 1u << 11,
    RQ_RARE_SHIFT = 12,
    RQ_AUX_SHIFT = 16,
    RQ_PIN_GLOBAL = 1u << 24,
};
#define RQ_W0(op, mode) ((uint32_t)(op) | ((uint32_t)(mode) << RQ_MODE_SHIFT))
#define RQ_MODE(w0) (((w0) >> RQ_MODE_SHIFT) & 3u)

// One independently schedulable piece of a sweep: its own staged columns, slot state and
// dot range. Large-n sweeps use one chunk (maximal sharing); small-n sweeps are cut into
// many chunks so that every SM has work.
struct RRChunk {
    int32_t pc_begin;   // first instruction; the chunk ends with RI_END
    int32_t n_ins;      // instructions including the RI_END
    int32_t dot_base;   // first dot output id
    int32_t n_dots;     // dot outputs of this chunk (MDOTDD outputs count 2, CLSMET 3)
    int32_t col_begin;  // into the plan's staged-column list
    int32_t n_cols;     // staged global columns; slots follow them in the tile
    int32_t reserved[2];
};
static_assert(sizeof(RRChunk) == 32, "RRChunk must be 32 bytes");

#endif  // RR_ISA_H
