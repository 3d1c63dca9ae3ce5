// rr_sweep_core_g8.cuh — hot loop of the interpreter for G8 plans (rr_isa.h RI_GRAM8), inline PTX, 4 samples per thread.
//
// Same machine as rr_sweep_core.cuh (accumulator t0..t3, 8 pins + 2 cache registers per sample, threaded dispatch
// through one brx.idx table, interleaved division / square root / sin / cos / exp / log bodies: those macros are
// shared) with ONE difference: where the classic core reduces a freshly evaluated term against the pins with 40 DFMA
// per thread and carries the per-lane partials through a shared-memory ring (transpose, shuffles, staging), this core
// leaves the term in a tile slot and reduces EIGHT such rows at once on the FP64 tensor-core path:
//
//   mma.sync.aligned.m8n8k4.row.col.f64  D[8 terms][8 pins] += A[8 terms][4 samples] * B[4 samples][8 pins]
//
// Lane (g, q) = (lane >> 2, lane & 3) supplies A[g][q] = term g at sample 4 step + q (one LDS.64 from the row's tile
// slot) and B[q][g] = pin g at the same sample, which it holds in 32 registers (its "B fragment": the warp's 128
// samples, 32 steps); after 32 steps it holds D[g][2q], D[g][2q+1]: the WARP totals of two reductions. Nothing is
// transposed and nothing is summed across lanes except a row's t.t and sum(t) (two butterfly steps over q).
// One DMMA issues 256 multiply-adds in one slot at the full FP64 rate (measured: tools/micro/fp64_pipes.cu, 18.5 T
// FMA/s against 17.0 for DFMA), so the reductions cost the pipe what they cost before but a twelfth of the
// instructions; the pins a row does not want are computed and dropped.
// The warp totals (80 per group: 8 rows x (8 pins, t.t, sum t)) are parked in the warp's staging row; after a block
// barrier thread i < 80 adds the four warps' totals of output i in fixed order and, if the instruction wants that
// output, issues one RED.ADD.F64 to the block's accumulator row (one writer per address, fixed order:
// bit-deterministic).
//
// Tile columns are padded by 32 bytes (column stride 4128): the eight rows of a group then sit in different banks
// and an A-fragment load is the two wavefronts 256 bytes need.
// Partial tiles: the three store handlers write zeros for the samples beyond n, so that rows and pins (both are
// loaded from tile slots) contribute nothing; the centred target and bare feature rows are zero padded in memory.
//
// rr_core_g8 returns like rr_core_s4: 0 window sentinel, 1 RI_END, 2 an instruction for the C++ caller.
#pragma once

#include <stdint.h>

#include "rr_isa.h"
#include "rr_sweep_core.cuh"

// ---- asm operand map ----
//  %0-%3 t0..t3   %4-%43 value registers   %44 cnt (outputs so far in this chunk and tile)   %45 vbits (valid samples
//  of this thread, 4 bits; 15 in a full tile)   %46 ibp   %47 exit code   %48 w0   %49 w1   %50 imm
//  %51-%82 pb0..pb31: the B fragment (pin g at the lane's 32 samples)                               (all outputs)
//  %83 tile_sh   %84 frag_sh (byte address of this lane's first fragment sample in tile column 0)   %85 acc_row
//  %86 gsel (prmt selector of this lane's row byte in the column slot)   %87 g   %88 q   %89 stage_w (staging address
//  of D[g][2q])   %90 xg   %91 ld_bytes   %92 stage_s (staging address of row g's t.t)   %93 comb_rd (staging address
//  of output tid in warp 0's row)   %94 comb_word (32-bit word of the wanted mask that holds output tid; 3 = none)
//  %95 comb_bit (1 << (tid & 31))   %96 xg_frag (address of this lane's first fragment sample in engine column 0)
// RR_MDCHK (the tail of every handler that may carry RR_THEN_ST) and RR_G8_STORE are defined per variant in
// rr_sweep_core_g8_body.inc: a predicated store INLINE, so that a handler and its dispatch stay one basic block
// the shared macros of rr_sweep_core.cuh name these operands symbolically; this core's numbering (and its
// compile-time tile geometry: 512-sample tiles, 4128-byte columns) replaces the classic core's from here on
#undef RR_O_TILE
#undef RR_O_COLB
#undef RR_O_HALFB
#define RR_O_TILE "%83"
#define RR_O_COLB "4128"
#define RR_O_HALFB "2048"
#define G_TILE "%83"
#define G_FRAG "%84"
#define G_ACC "%85"
#define G_GSEL "%86"
#define G_G "%87"
#define G_Q "%88"
#define G_STW "%89"
#define G_XG "%90"
#define G_LD "%91"
#define G_STS "%92"
#define G_CRD "%93"
#define G_CWORD "%94"
#define G_CBIT "%95"
#define G_XGF "%96"
#define RR_G8_PINB0_VALUE 151
static_assert(RR_G8_PINB0_VALUE == RI_PINB0, "update RR_G8_PINB0_VALUE");

#define RR_G8_FIRST_M_VALUE 106
static_assert(RR_G8_FIRST_M_VALUE == RI_FIRST_M, "update RR_G8_FIRST_M_VALUE and the jump table");
static_assert(RI_OPCOUNT == 159, "update the jump table of rr_core_g8");

#define RR_PB(s) RR_PB_(s)
#define RR_PB_(s) RR_PBREG_##s
#define RR_PBREG_0 "%51"
#define RR_PBREG_1 "%52"
#define RR_PBREG_2 "%53"
#define RR_PBREG_3 "%54"
#define RR_PBREG_4 "%55"
#define RR_PBREG_5 "%56"
#define RR_PBREG_6 "%57"
#define RR_PBREG_7 "%58"
#define RR_PBREG_8 "%59"
#define RR_PBREG_9 "%60"
#define RR_PBREG_10 "%61"
#define RR_PBREG_11 "%62"
#define RR_PBREG_12 "%63"
#define RR_PBREG_13 "%64"
#define RR_PBREG_14 "%65"
#define RR_PBREG_15 "%66"
#define RR_PBREG_16 "%67"
#define RR_PBREG_17 "%68"
#define RR_PBREG_18 "%69"
#define RR_PBREG_19 "%70"
#define RR_PBREG_20 "%71"
#define RR_PBREG_21 "%72"
#define RR_PBREG_22 "%73"
#define RR_PBREG_23 "%74"
#define RR_PBREG_24 "%75"
#define RR_PBREG_25 "%76"
#define RR_PBREG_26 "%77"
#define RR_PBREG_27 "%78"
#define RR_PBREG_28 "%79"
#define RR_PBREG_29 "%80"
#define RR_PBREG_30 "%81"
#define RR_PBREG_31 "%82"

// byte offset of fragment step s from the lane's first fragment sample: steps 0-15 walk the first half of the tile
// (32 bytes = 4 samples per step), steps 16-31 the second half (2048 = byte offset of the second half)
#define RR_G8_OFF_LO(s) RR_G8_OFF_LO_(s)
#define RR_G8_OFF_LO_(s) RR_G8_OFFV_##s
#define RR_G8_OFFV_0 "0"
#define RR_G8_OFFV_1 "32"
#define RR_G8_OFFV_2 "64"
#define RR_G8_OFFV_3 "96"
#define RR_G8_OFFV_4 "128"
#define RR_G8_OFFV_5 "160"
#define RR_G8_OFFV_6 "192"
#define RR_G8_OFFV_7 "224"
#define RR_G8_OFFV_8 "256"
#define RR_G8_OFFV_9 "288"
#define RR_G8_OFFV_10 "320"
#define RR_G8_OFFV_11 "352"
#define RR_G8_OFFV_12 "384"
#define RR_G8_OFFV_13 "416"
#define RR_G8_OFFV_14 "448"
#define RR_G8_OFFV_15 "480"
// one step of the group: A element from the row's slot, DMMA against the B fragment, the row's own t.t (sum(t) is the
// product with the last pin, which G8 plans keep at a column of ones: BatchPlanner::Chunk::reserve_pin_ones)
// (D0, D1) / S2: the accumulators of this step's chain (S1 is unused) - even and odd steps run two independent chains, a
// dependent DMMA every 32 cycles per warp would leave the pipe half idle
#define RR_G8_STEP_LO(s, A, D0, D1, S2, S1)                                                              \
    "ld.shared.f64 " A ", [wp+" RR_G8_OFF_LO(s) "];\n"                                                   \
    "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {" D0 ", " D1 "}, {" A "}, {" RR_PB(s) "}, {" D0 ", " D1 "};\n" \
    "fma.rn.f64 " S2 ", " A ", " A ", " S2 ";\n"
#define RR_G8_STEP_HI(s, k, A, D0, D1, S2, S1)                                                           \
    "ld.shared.f64 " A ", [wq+" RR_G8_OFF_LO(k) "];\n"                                                   \
    "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {" D0 ", " D1 "}, {" A "}, {" RR_PB(s) "}, {" D0 ", " D1 "};\n" \
    "fma.rn.f64 " S2 ", " A ", " A ", " S2 ";\n"
// B fragment of the lanes that own pin j (predicate pq) from a tile column (wp / wq = fragment base of both halves)
#define RR_G8_PB_LO(s) "@pq ld.shared.f64 " RR_PB(s) ", [wp+" RR_G8_OFF_LO(s) "];\n"
#define RR_G8_PB_HI(s, k) "@pq ld.shared.f64 " RR_PB(s) ", [wq+" RR_G8_OFF_LO(k) "];\n"
// ... from an engine column in global memory (ga / gb)
#define RR_G8_PBG_LO(s) "@pq ld.global.f64 " RR_PB(s) ", [ga+" RR_G8_OFF_LO(s) "];\n"
#define RR_G8_PBG_HI(s, k) "@pq ld.global.f64 " RR_PB(s) ", [gb+" RR_G8_OFF_LO(k) "];\n"

#define RR_G8_PINB(J)                                                                                    \
    "L_PINB" #J ":\n"                                                                                    \
    "mov.f64 " RR_P(J, 0) ", u0;\n mov.f64 " RR_P(J, 1) ", u1;\n mov.f64 " RR_P(J, 2) ", u2;\n"          \
    "mov.f64 " RR_P(J, 3) ", u3;\n bra.uni L_PINB_COMMON;\n"

// Two instantiations of the same asm body: full tiles (every sample valid: plain stores) and the partial tile at the
// end of the data (stores write zeros beyond n). No handler branches on the tile kind.
#define RR_G8_PARTIAL 0
#define RR_G8_FN rr_core_g8_full
#include "rr_sweep_core_g8_body.inc"
#undef RR_G8_PARTIAL
#undef RR_G8_FN
#define RR_G8_PARTIAL 1
#define RR_G8_FN rr_core_g8_partial
#include "rr_sweep_core_g8_body.inc"
#undef RR_G8_PARTIAL
#undef RR_G8_FN
