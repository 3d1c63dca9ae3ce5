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
#undef RR_MDCHK
#define RR_MDCHK "and.b32 x, w0, 0x8000;\n setp.ne.u32 p, x, 0;\n @p bra.uni L_STT;\n"
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
// one step of the group: A element from the row's slot, DMMA against the B fragment, the row's own t.t and sum(t)
// (D0, D1) / S2 / S1: the accumulators of this step's chain - even and odd steps run two independent chains, a
// dependent DMMA every 32 cycles per warp would leave the pipe half idle
#define RR_G8_STEP_LO(s, A, D0, D1, S2, S1)                                                              \
    "ld.shared.f64 " A ", [wp+" RR_G8_OFF_LO(s) "];\n"                                                   \
    "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {" D0 ", " D1 "}, {" A "}, {" RR_PB(s) "}, {" D0 ", " D1 "};\n" \
    "fma.rn.f64 " S2 ", " A ", " A ", " S2 ";\n add.rn.f64 " S1 ", " S1 ", " A ";\n"
#define RR_G8_STEP_HI(s, k, A, D0, D1, S2, S1)                                                           \
    "ld.shared.f64 " A ", [wq+" RR_G8_OFF_LO(k) "];\n"                                                   \
    "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {" D0 ", " D1 "}, {" A "}, {" RR_PB(s) "}, {" D0 ", " D1 "};\n" \
    "fma.rn.f64 " S2 ", " A ", " A ", " S2 ";\n add.rn.f64 " S1 ", " S1 ", " A ";\n"
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

// masked stores of a partial tile: samples beyond n are written as zeros (vbits = %45)
#define RR_G8_MASK4(D0, D1, D2, D3)                                                                      \
    "and.b32 slo, %45, 1;\n setp.ne.u32 q0, slo, 0;\n and.b32 slo, %45, 2;\n setp.ne.u32 q1, slo, 0;\n" \
    "and.b32 slo, %45, 4;\n setp.ne.u32 q2, slo, 0;\n and.b32 slo, %45, 8;\n setp.ne.u32 q3, slo, 0;\n" \
    "selp.f64 " D0 ", %0, 0d0000000000000000, q0;\n selp.f64 " D1 ", %1, 0d0000000000000000, q1;\n"      \
    "selp.f64 " D2 ", %2, 0d0000000000000000, q2;\n selp.f64 " D3 ", %3, 0d0000000000000000, q3;\n"
// store t to the tile column at byte address ADDR (a register). vbits = 15 in every thread of a full tile; a partial
// tile sets bit 4 in EVERY thread (the branch is block-uniform) and stores the masked copy
#define RR_G8_STORE(ADDR, LBL)                                                                           \
    "setp.ne.u32 p, %45, 15;\n"                                                                          \
    "@p bra.uni " LBL "_P;\n"                                                                            \
    "st.shared.v2.f64 [" ADDR "], {%0, %1};\n st.shared.v2.f64 [" ADDR "+2048], {%2, %3};\n"
// the masked variant, reached by the uniform branch above; LBL##_P ... ends in its own dispatch
#define RR_G8_STORE_PARTIAL(ADDR, LBL)                                                                   \
    LBL "_P:\n"                                                                                          \
    RR_G8_MASK4("f0", "f1", "f2", "f3")                                                                  \
    "st.shared.v2.f64 [" ADDR "], {f0, f1};\n st.shared.v2.f64 [" ADDR "+2048], {f2, f3};\n"

namespace rr {

__device__ __forceinline__ uint32_t rr_core_g8(double &t0, double &t1, double &t2, double &t3, double *B, double *PB,
                                               uint32_t &cnt, uint32_t vbits, uint32_t &ibp, uint32_t &ow0, uint32_t &ow1,
                                               double &oimm, uint32_t tile_sh, uint32_t frag_sh, double *acc_row,
                                               uint32_t gsel, uint32_t g, uint32_t q, uint32_t stage_w, const double *xg,
                                               int64_t ld_bytes, uint32_t stage_s, uint32_t comb_rd, uint32_t comb_word,
                                               uint32_t comb_bit, const double *xg_frag)
{
    uint32_t code;
    asm volatile(
        "{\n"
        ".reg .b32 w0, w1, n0, n1, nz, nw, op, col, x, idx, wp, wq, slo, shi, m0, m1, m2, mw;\n"
        ".reg .f32 fa, fb;\n"
        ".reg .f64 u0, u1, u2, u3, imm, v0, v1, v2, v3, v4, v5, v6, v7, f0, f1, f2, f3, a0, a1, a2, a3, a4, a5, a6, a7;\n"
        ".reg .f64 dr0, dr1, dr2, dr3, dn0, dn1, dn2, dn3, de0, de1, de2, de3, dq0, dq1, dq2, dq3;\n"
        ".reg .b32 ki0, ki1, ki2, ki3;\n"
        ".reg .f64 ta0, ta1, ta2, ta3, tr0, tr1, tr2, tr3, tz0, tz1, tz2, tz3, tm0, tm1, tm2, tm3;\n"
        ".reg .f64 tp0, tp1, tp2, tp3, tc0, tc1, tc2, tc3, tq0, tq1, tq2, tq3;\n"
        ".reg .pred p, pm, ps, po, q0, q1, q2, q3, pok, pq, pw;\n"
        ".reg .b64 ga, gb;\n"
        "TBL: .branchtargets L_END, L_WINEND, L_LOADC, L_ST, L_OTHER, L_LDG, L_NOP, L_NOP, "
        "L_ADDC, L_SUBC, L_RSUBC, L_MULC, L_DIVC, L_RDIVC, "
        "L_SIN, L_COS, L_LN, L_EXP, L_SQRT, L_SQR, L_OTHER, L_OTHER, L_GRAM8, L_PINBG, L_OTHER, L_OTHER, "
        "L_PIN0, L_PIN1, L_PIN2, L_PIN3, L_PIN4, L_PIN5, L_PIN6, L_PIN7, L_PIN8, L_PIN9, "
        "L_LDP0, L_LDP1, L_LDP2, L_LDP3, L_LDP4, L_LDP5, L_LDP6, L_LDP7, L_LDP8, L_LDP9, "
        "L_USEP0, L_USEP1, L_USEP2, L_USEP3, L_USEP4, L_USEP5, L_USEP6, L_USEP7, L_USEP8, L_USEP9, "
        "L_MULP0, L_MULP1, L_MULP2, L_MULP3, L_MULP4, L_MULP5, L_MULP6, L_MULP7, L_MULP8, L_MULP9, "
        "L_DIVP0, L_DIVP1, L_DIVP2, L_DIVP3, L_DIVP4, L_DIVP5, L_DIVP6, L_DIVP7, L_DIVP8, L_DIVP9, "
        "L_RDIVP0, L_RDIVP1, L_RDIVP2, L_RDIVP3, L_RDIVP4, L_RDIVP5, L_RDIVP6, L_RDIVP7, L_RDIVP8, L_RDIVP9, "
        "L_CMULP0, L_CMULP1, L_CMULP2, L_CMULP3, L_CMULP4, L_CMULP5, L_CMULP6, L_CMULP7, L_CMULP8, L_CMULP9, "
        "L_CDIVP0, L_CDIVP1, L_CDIVP2, L_CDIVP3, L_CDIVP4, L_CDIVP5, L_CDIVP6, L_CDIVP7, L_CDIVP8, L_CDIVP9, "
        "L_LOADM, L_ADDM, L_SUBM, L_RSUBM, L_MULM, L_DIVM, L_RDIVM, L_AXPY, L_OTHER, L_OTHER, "
        "L_CMULM, L_CDIVM, L_MULMM, L_MULMST, L_MULMMM, "
        "L_LDPMULM0, L_LDPMULM1, L_LDPMULM2, L_LDPMULM3, L_LDPMULM4, L_LDPMULM5, L_LDPMULM6, L_LDPMULM7, L_LDPMULM8, L_LDPMULM9, "
        "L_LDPDIVM0, L_LDPDIVM1, L_LDPDIVM2, L_LDPDIVM3, L_LDPDIVM4, L_LDPDIVM5, L_LDPDIVM6, L_LDPDIVM7, L_LDPDIVM8, L_LDPDIVM9, "
        "L_LDMDIVP0, L_LDMDIVP1, L_LDMDIVP2, L_LDMDIVP3, L_LDMDIVP4, L_LDMDIVP5, L_LDMDIVP6, L_LDMDIVP7, L_LDMDIVP8, L_LDMDIVP9, "
        "L_PINB0, L_PINB1, L_PINB2, L_PINB3, L_PINB4, L_PINB5, L_PINB6, L_PINB7;\n"
        "ld.shared.v4.b32 {n0, n1, nz, nw}, [%46];\n"
        RR_DISPATCH
        "L_NOP:\n"
        RR_DISPATCH
        "L_LOADC:\n"
        "mov.f64 %0, imm;\n mov.f64 %1, imm;\n mov.f64 %2, imm;\n mov.f64 %3, imm;\n"
        RR_DISPATCH
        "L_LOADM:\n"
        "mov.f64 %0, u0;\n mov.f64 %1, u1;\n mov.f64 %2, u2;\n mov.f64 %3, u3;\n"
        RR_DISPATCH
        "L_ST:\n"
        RR_G8_STORE("col", "ST_A")
        RR_DISPATCH
        RR_G8_STORE_PARTIAL("col", "ST_A")
        RR_DISPATCH
        /* tail of an instruction that carries RR_THEN_ST: store t to the tile column in bits 16-23 of w0 */
        "L_STT:\n"
        "shr.u32 x, w0, 16;\n and.b32 x, x, 255;\n mad.lo.u32 x, x, 4128, " G_TILE ";\n"
        RR_G8_STORE("x", "ST_B")
        RR_DISPATCH
        RR_G8_STORE_PARTIAL("x", "ST_B")
        RR_DISPATCH
        "L_LDG:\n"
        RR_RELOAD_W1
        "cvt.u64.u32 ga, w1;\n mul.lo.u64 ga, ga, " G_LD ";\n add.u64 ga, ga, " G_XG ";\n"
        "ld.global.v2.f64 {%0, %1}, [ga];\n ld.global.v2.f64 {%2, %3}, [ga+2048];\n"
        RR_DISPATCH
        RR_BIN_C("L_ADDC", "add.rn.f64")
        RR_BIN_C("L_SUBC", "sub.rn.f64")
        RR_RBIN_C("L_RSUBC", "sub.rn.f64")
        RR_BIN_C("L_MULC", "mul.rn.f64")
        RR_DIV4("L_DIVC", "%0", "%1", "%2", "%3", "imm", "imm", "imm", "imm")
        RR_DIV4("L_RDIVC", "imm", "imm", "imm", "imm", "%0", "%1", "%2", "%3")
        "L_SQRT:\n"
        "setp.eq.u32 pok, 0, 0;\n"
        RR_SQRT_FAST(0, "%0") RR_SQRT_FAST(1, "%1") RR_SQRT_FAST(2, "%2") RR_SQRT_FAST(3, "%3")
        "vote.sync.all.pred pok, pok, 0xffffffff;\n"
        "@!pok bra.uni L_SQRT_SLOW;\n"
        "mov.f64 %0, dq0;\n mov.f64 %1, dq1;\n mov.f64 %2, dq2;\n mov.f64 %3, dq3;\n"
        RR_DISPATCH
        "L_SQRT_SLOW:\n"
        "sqrt.rn.f64 %0, %0;\n sqrt.rn.f64 %1, %1;\n sqrt.rn.f64 %2, %2;\n sqrt.rn.f64 %3, %3;\n"
        RR_DISPATCH
        RR_FAST4("L_SIN", RR_TRIG_POLY(0, "%0", ""), RR_TRIG_POLY(1, "%1", ""), RR_TRIG_POLY(2, "%2", ""), RR_TRIG_POLY(3, "%3", ""))
        RR_FAST4("L_COS", RR_TRIG_POLY(0, "%0", "add.s32 ki0, ki0, 1;\n"), RR_TRIG_POLY(1, "%1", "add.s32 ki1, ki1, 1;\n"),
                 RR_TRIG_POLY(2, "%2", "add.s32 ki2, ki2, 1;\n"), RR_TRIG_POLY(3, "%3", "add.s32 ki3, ki3, 1;\n"))
        RR_FAST4("L_EXP", RR_EXP_FAST(0, "%0"), RR_EXP_FAST(1, "%1"), RR_EXP_FAST(2, "%2"), RR_EXP_FAST(3, "%3"))
        RR_FAST4("L_LN", RR_LOG_FAST(0, "%0"), RR_LOG_FAST(1, "%1"), RR_LOG_FAST(2, "%2"), RR_LOG_FAST(3, "%3"))
        "L_SQR:\n"
        "mul.rn.f64 %0, %0, %0;\n mul.rn.f64 %1, %1, %1;\n mul.rn.f64 %2, %2, %2;\n mul.rn.f64 %3, %3, %3;\n"
        RR_DISPATCH
        RR_BIN_M("L_ADDM", "add.rn.f64")
        RR_BIN_M("L_SUBM", "sub.rn.f64")
        RR_RBIN_M("L_RSUBM", "sub.rn.f64")
        "L_MULM:\n"
        "mul.rn.f64 %0, %0, u0;\n mul.rn.f64 %1, %1, u1;\n mul.rn.f64 %2, %2, u2;\n mul.rn.f64 %3, %3, u3;\n"
        RR_MDCHK RR_DISPATCH
        RR_DIV4("L_DIVM", "%0", "%1", "%2", "%3", "u0", "u1", "u2", "u3")
        RR_DIV4("L_RDIVM", "u0", "u1", "u2", "u3", "%0", "%1", "%2", "%3")
        "L_AXPY:\n"
        "mul.rn.f64 u0, imm, u0;\n mul.rn.f64 u1, imm, u1;\n mul.rn.f64 u2, imm, u2;\n mul.rn.f64 u3, imm, u3;\n"
        "add.rn.f64 %0, %0, u0;\n add.rn.f64 %1, %1, u1;\n add.rn.f64 %2, %2, u2;\n add.rn.f64 %3, %3, u3;\n"
        RR_DISPATCH
        RR_PIN_HANDLERS(0) RR_PIN_HANDLERS(1) RR_PIN_HANDLERS(2) RR_PIN_HANDLERS(3)
        RR_PIN_HANDLERS(4) RR_PIN_HANDLERS(5) RR_PIN_HANDLERS(6) RR_PIN_HANDLERS(7)
        RR_PIN_HANDLERS(8) RR_PIN_HANDLERS(9)
        RR_FUSED_HANDLERS(0) RR_FUSED_HANDLERS(1) RR_FUSED_HANDLERS(2) RR_FUSED_HANDLERS(3) RR_FUSED_HANDLERS(4)
        RR_FUSED_HANDLERS(5) RR_FUSED_HANDLERS(6) RR_FUSED_HANDLERS(7) RR_FUSED_HANDLERS(8) RR_FUSED_HANDLERS(9)
        "L_CMULM:\n" /* t = imm * tile[w1] */
        "mul.rn.f64 %0, imm, u0;\n mul.rn.f64 %1, imm, u1;\n mul.rn.f64 %2, imm, u2;\n mul.rn.f64 %3, imm, u3;\n"
        RR_MDCHK RR_DISPATCH
        "L_CDIVM:\n" /* t = imm / tile[w1] */
        RR_MOV4("%0", "%1", "%2", "%3", "u0", "u1", "u2", "u3")
        "bra.uni L_RDIVC;\n"
        "L_MULMM:\n" /* t = tile[w1] * tile[lo32(imm)] */
        "mov.b64 {slo, shi}, imm;\n mad.lo.u32 x, slo, 4128, " G_TILE ";\n"
        "ld.shared.v2.f64 {f0, f1}, [x];\n ld.shared.v2.f64 {f2, f3}, [x+2048];\n"
        "mul.rn.f64 %0, u0, f0;\n mul.rn.f64 %1, u1, f1;\n mul.rn.f64 %2, u2, f2;\n mul.rn.f64 %3, u3, f3;\n"
        RR_MDCHK RR_DISPATCH
        "L_MULMMM:\n" /* t = (tile[w1] * tile[lo32(imm)]) * tile[hi32(imm)] */
        "mov.b64 {slo, shi}, imm;\n mad.lo.u32 x, slo, 4128, " G_TILE ";\n mad.lo.u32 idx, shi, 4128, " G_TILE ";\n"
        "ld.shared.v2.f64 {f0, f1}, [x];\n ld.shared.v2.f64 {f2, f3}, [x+2048];\n"
        "ld.shared.v2.f64 {a0, a1}, [idx];\n ld.shared.v2.f64 {a2, a3}, [idx+2048];\n"
        "mul.rn.f64 %0, u0, f0;\n mul.rn.f64 %1, u1, f1;\n mul.rn.f64 %2, u2, f2;\n mul.rn.f64 %3, u3, f3;\n"
        "mul.rn.f64 %0, %0, a0;\n mul.rn.f64 %1, %1, a1;\n mul.rn.f64 %2, %2, a2;\n mul.rn.f64 %3, %3, a3;\n"
        RR_MDCHK RR_DISPATCH
        "L_MULMST:\n" /* t = t * tile[w1]; tile[lo32(imm)] = t */
        "mov.b64 {slo, shi}, imm;\n mad.lo.u32 idx, slo, 4128, " G_TILE ";\n"
        "mul.rn.f64 %0, %0, u0;\n mul.rn.f64 %1, %1, u1;\n mul.rn.f64 %2, %2, u2;\n mul.rn.f64 %3, %3, u3;\n"
        RR_G8_STORE("idx", "ST_C")
        RR_DISPATCH
        RR_G8_STORE_PARTIAL("idx", "ST_C")
        RR_DISPATCH
        /* ---- PINB j: pin j <- tile[w1] as operand register (u0..u3 hold the column) and as reduction partner ---- */
        RR_G8_PINB(0) RR_G8_PINB(1) RR_G8_PINB(2) RR_G8_PINB(3) RR_G8_PINB(4) RR_G8_PINB(5) RR_G8_PINB(6) RR_G8_PINB(7)
        "L_PINB_COMMON:\n"
        "bar.warp.sync 0xffffffff;\n" /* the column was stored by other lanes of this warp */
        RR_RELOAD_W1
        "sub.u32 x, op, " RR_STR(RR_G8_PINB0_VALUE) ";\n setp.eq.u32 pq, x, " G_G ";\n"
        "mad.lo.u32 wp, w1, 4128, " G_FRAG ";\n add.u32 wq, wp, 2048;\n"
        RR_G8_PB_LO(0) RR_G8_PB_LO(1) RR_G8_PB_LO(2) RR_G8_PB_LO(3) RR_G8_PB_LO(4) RR_G8_PB_LO(5) RR_G8_PB_LO(6) RR_G8_PB_LO(7)
        RR_G8_PB_LO(8) RR_G8_PB_LO(9) RR_G8_PB_LO(10) RR_G8_PB_LO(11) RR_G8_PB_LO(12) RR_G8_PB_LO(13) RR_G8_PB_LO(14) RR_G8_PB_LO(15)
        RR_G8_PB_HI(16, 0) RR_G8_PB_HI(17, 1) RR_G8_PB_HI(18, 2) RR_G8_PB_HI(19, 3) RR_G8_PB_HI(20, 4) RR_G8_PB_HI(21, 5)
        RR_G8_PB_HI(22, 6) RR_G8_PB_HI(23, 7) RR_G8_PB_HI(24, 8) RR_G8_PB_HI(25, 9) RR_G8_PB_HI(26, 10) RR_G8_PB_HI(27, 11)
        RR_G8_PB_HI(28, 12) RR_G8_PB_HI(29, 13) RR_G8_PB_HI(30, 14) RR_G8_PB_HI(31, 15)
        RR_DISPATCH
        /* ---- PINBG: pin (aux) <- engine column w1 from global memory, reduction partner only ---- */
        "L_PINBG:\n"
        RR_RELOAD_W1
        "shr.u32 x, w0, 8;\n and.b32 x, x, 255;\n setp.eq.u32 pq, x, " G_G ";\n"
        "cvt.u64.u32 ga, w1;\n mul.lo.u64 ga, ga, " G_LD ";\n add.u64 ga, ga, " G_XGF ";\n"
        "cvt.u64.u32 gb, 2048;\n add.u64 gb, gb, ga;\n"
        RR_G8_PBG_LO(0) RR_G8_PBG_LO(1) RR_G8_PBG_LO(2) RR_G8_PBG_LO(3) RR_G8_PBG_LO(4) RR_G8_PBG_LO(5) RR_G8_PBG_LO(6) RR_G8_PBG_LO(7)
        RR_G8_PBG_LO(8) RR_G8_PBG_LO(9) RR_G8_PBG_LO(10) RR_G8_PBG_LO(11) RR_G8_PBG_LO(12) RR_G8_PBG_LO(13) RR_G8_PBG_LO(14) RR_G8_PBG_LO(15)
        RR_G8_PBG_HI(16, 0) RR_G8_PBG_HI(17, 1) RR_G8_PBG_HI(18, 2) RR_G8_PBG_HI(19, 3) RR_G8_PBG_HI(20, 4) RR_G8_PBG_HI(21, 5)
        RR_G8_PBG_HI(22, 6) RR_G8_PBG_HI(23, 7) RR_G8_PBG_HI(24, 8) RR_G8_PBG_HI(25, 9) RR_G8_PBG_HI(26, 10) RR_G8_PBG_HI(27, 11)
        RR_G8_PBG_HI(28, 12) RR_G8_PBG_HI(29, 13) RR_G8_PBG_HI(30, 14) RR_G8_PBG_HI(31, 15)
        RR_DISPATCH
        /* ---- GRAM8: up to eight rows against the eight pins, themselves and ones ---- */
        "L_GRAM8:\n"
        /* the column slot behind this instruction sits in the prefetch registers: this lane's row is byte g & 3 of
           n1 (rows 0-3) or nz (rows 4-7); then the slot is skipped */
        "bar.warp.sync 0xffffffff;\n" /* the rows were stored by other lanes of this warp */
        "setp.lt.u32 p, " G_G ", 4;\n selp.b32 x, n1, nz, p;\n prmt.b32 x, x, 0, " G_GSEL ";\n"
        "mad.lo.u32 wp, x, 4128, " G_FRAG ";\n add.u32 wq, wp, 2048;\n"
        "ld.shared.b32 m0, [%46+-12];\n"
        "mov.b64 {m1, m2}, imm;\n"
        "add.u32 %46, %46, 16;\n"
        "ld.shared.v4.b32 {n0, n1, nz, nw}, [%46];\n"
        "mov.f64 v0, 0d0000000000000000;\n mov.f64 v1, 0d0000000000000000;\n"
        "mov.f64 v2, 0d0000000000000000;\n mov.f64 v3, 0d0000000000000000;\n"
        "mov.f64 v4, 0d0000000000000000;\n mov.f64 v5, 0d0000000000000000;\n"
        "mov.f64 v6, 0d0000000000000000;\n mov.f64 v7, 0d0000000000000000;\n"
        RR_G8_STEP_LO(0, "a0", "v0", "v1", "v2", "v3") RR_G8_STEP_LO(1, "a1", "v4", "v5", "v6", "v7") RR_G8_STEP_LO(2, "a2", "v0", "v1", "v2", "v3") RR_G8_STEP_LO(3, "a3", "v4", "v5", "v6", "v7")
        RR_G8_STEP_LO(4, "a4", "v0", "v1", "v2", "v3") RR_G8_STEP_LO(5, "a5", "v4", "v5", "v6", "v7") RR_G8_STEP_LO(6, "a6", "v0", "v1", "v2", "v3") RR_G8_STEP_LO(7, "a7", "v4", "v5", "v6", "v7")
        RR_G8_STEP_LO(8, "a0", "v0", "v1", "v2", "v3") RR_G8_STEP_LO(9, "a1", "v4", "v5", "v6", "v7") RR_G8_STEP_LO(10, "a2", "v0", "v1", "v2", "v3") RR_G8_STEP_LO(11, "a3", "v4", "v5", "v6", "v7")
        RR_G8_STEP_LO(12, "a4", "v0", "v1", "v2", "v3") RR_G8_STEP_LO(13, "a5", "v4", "v5", "v6", "v7") RR_G8_STEP_LO(14, "a6", "v0", "v1", "v2", "v3") RR_G8_STEP_LO(15, "a7", "v4", "v5", "v6", "v7")
        RR_G8_STEP_HI(16, 0, "a0", "v0", "v1", "v2", "v3") RR_G8_STEP_HI(17, 1, "a1", "v4", "v5", "v6", "v7") RR_G8_STEP_HI(18, 2, "a2", "v0", "v1", "v2", "v3") RR_G8_STEP_HI(19, 3, "a3", "v4", "v5", "v6", "v7")
        RR_G8_STEP_HI(20, 4, "a4", "v0", "v1", "v2", "v3") RR_G8_STEP_HI(21, 5, "a5", "v4", "v5", "v6", "v7") RR_G8_STEP_HI(22, 6, "a6", "v0", "v1", "v2", "v3") RR_G8_STEP_HI(23, 7, "a7", "v4", "v5", "v6", "v7")
        RR_G8_STEP_HI(24, 8, "a0", "v0", "v1", "v2", "v3") RR_G8_STEP_HI(25, 9, "a1", "v4", "v5", "v6", "v7") RR_G8_STEP_HI(26, 10, "a2", "v0", "v1", "v2", "v3") RR_G8_STEP_HI(27, 11, "a3", "v4", "v5", "v6", "v7")
        RR_G8_STEP_HI(28, 12, "a4", "v0", "v1", "v2", "v3") RR_G8_STEP_HI(29, 13, "a5", "v4", "v5", "v6", "v7") RR_G8_STEP_HI(30, 14, "a6", "v0", "v1", "v2", "v3") RR_G8_STEP_HI(31, 15, "a7", "v4", "v5", "v6", "v7")
        "add.rn.f64 v0, v0, v4;\n add.rn.f64 v1, v1, v5;\n add.rn.f64 v2, v2, v6;\n add.rn.f64 v3, v3, v7;\n"
        /* t.t and sum(t) of row g: the four lanes of the row join their quarters */
        "mov.b64 {slo, shi}, v2;\n"
        "shfl.sync.bfly.b32 slo, slo, 1, 31, 0xffffffff;\n shfl.sync.bfly.b32 shi, shi, 1, 31, 0xffffffff;\n"
        "mov.b64 f0, {slo, shi};\n add.rn.f64 v2, v2, f0;\n"
        "mov.b64 {slo, shi}, v3;\n"
        "shfl.sync.bfly.b32 slo, slo, 1, 31, 0xffffffff;\n shfl.sync.bfly.b32 shi, shi, 1, 31, 0xffffffff;\n"
        "mov.b64 f1, {slo, shi};\n add.rn.f64 v3, v3, f1;\n"
        "mov.b64 {slo, shi}, v2;\n"
        "shfl.sync.bfly.b32 slo, slo, 2, 31, 0xffffffff;\n shfl.sync.bfly.b32 shi, shi, 2, 31, 0xffffffff;\n"
        "mov.b64 f0, {slo, shi};\n add.rn.f64 v2, v2, f0;\n"
        "mov.b64 {slo, shi}, v3;\n"
        "shfl.sync.bfly.b32 slo, slo, 2, 31, 0xffffffff;\n shfl.sync.bfly.b32 shi, shi, 2, 31, 0xffffffff;\n"
        "mov.b64 f1, {slo, shi};\n add.rn.f64 v3, v3, f1;\n"
        /* warp totals -> the warp's staging row: outputs 10 g + 2 q, 10 g + 2 q + 1; lane q = 0 adds 10 g + 8, 10 g + 9 */
        "st.shared.v2.f64 [" G_STW "], {v0, v1};\n"
        "setp.eq.u32 p, " G_Q ", 0;\n"
        "@p st.shared.v2.f64 [" G_STS "], {v2, v3};\n"
        "bar.sync 1;\n"
        /* thread i < 80 owns output i: the four warps' totals in fixed order, one RED if the instruction wants it */
        "popc.b32 x, m0;\n popc.b32 idx, m1;\n"
        "setp.eq.u32 p, " G_CWORD ", 0;\n selp.b32 mw, m0, m1, p;\n setp.eq.u32 pw, " G_CWORD ", 2;\n selp.b32 mw, m2, mw, pw;\n"
        "setp.ge.u32 p, " G_CWORD ", 1;\n selp.b32 slo, x, 0, p;\n selp.b32 shi, idx, 0, pw;\n add.u32 slo, slo, shi;\n" /* wanted outputs in the words below mine */
        "add.u32 x, x, idx;\n popc.b32 idx, m2;\n add.u32 x, x, idx;\n"                                         /* x = all wanted outputs */
        "sub.u32 shi, " G_CBIT ", 1;\n and.b32 shi, shi, mw;\n popc.b32 shi, shi;\n add.u32 slo, slo, shi;\n"          /* + those below my bit */
        "add.u32 idx, %44, slo;\n add.u32 %44, %44, x;\n"
        "and.b32 mw, mw, " G_CBIT ";\n setp.ne.u32 p, mw, 0;\n setp.lt.and.u32 p, " G_CWORD ", 3, p;\n"
        "@p ld.shared.f64 f0, [" G_CRD "];\n @p ld.shared.f64 f1, [" G_CRD "+640];\n @p ld.shared.f64 f2, [" G_CRD "+1280];\n"
        "@p ld.shared.f64 f3, [" G_CRD "+1920];\n"
        "@p add.rn.f64 f0, f0, f1;\n @p add.rn.f64 f0, f0, f2;\n @p add.rn.f64 f0, f0, f3;\n"
        "mul.wide.u32 ga, idx, 8;\n add.u64 ga, ga, " G_ACC ";\n"
        "@p red.global.add.f64 [ga], f0;\n"
        "bar.sync 1;\n"
        RR_DISPATCH
        "L_OTHER:\n"
        RR_RELOAD_W1
        "mov.b32 %47, 2;\n mov.b32 %48, w0;\n mov.b32 %49, w1;\n mov.f64 %50, imm;\n"
        "bra.uni DONE;\n"
        "L_END:\n"
        "mov.b32 %47, 1;\n mov.b32 %48, 0;\n mov.b32 %49, 0;\n mov.f64 %50, imm;\n"
        "bra.uni DONE;\n"
        "L_WINEND:\n"
        "mov.b32 %47, 0;\n mov.b32 %48, 0;\n mov.b32 %49, 0;\n mov.f64 %50, 0d0000000000000000;\n"
        "DONE:\n"
        "}\n"
        : "+d"(t0), "+d"(t1), "+d"(t2), "+d"(t3),
          "+d"(B[0]), "+d"(B[1]), "+d"(B[2]), "+d"(B[3]), "+d"(B[4]), "+d"(B[5]), "+d"(B[6]), "+d"(B[7]),
          "+d"(B[8]), "+d"(B[9]), "+d"(B[10]), "+d"(B[11]), "+d"(B[12]), "+d"(B[13]), "+d"(B[14]), "+d"(B[15]),
          "+d"(B[16]), "+d"(B[17]), "+d"(B[18]), "+d"(B[19]), "+d"(B[20]), "+d"(B[21]), "+d"(B[22]), "+d"(B[23]),
          "+d"(B[24]), "+d"(B[25]), "+d"(B[26]), "+d"(B[27]), "+d"(B[28]), "+d"(B[29]), "+d"(B[30]), "+d"(B[31]),
          "+d"(B[32]), "+d"(B[33]), "+d"(B[34]), "+d"(B[35]), "+d"(B[36]), "+d"(B[37]), "+d"(B[38]), "+d"(B[39]),
          "+r"(cnt), "+r"(vbits), "+r"(ibp), "=r"(code), "=r"(ow0), "=r"(ow1), "=d"(oimm),
          "+d"(PB[0]), "+d"(PB[1]), "+d"(PB[2]), "+d"(PB[3]), "+d"(PB[4]), "+d"(PB[5]), "+d"(PB[6]), "+d"(PB[7]),
          "+d"(PB[8]), "+d"(PB[9]), "+d"(PB[10]), "+d"(PB[11]), "+d"(PB[12]), "+d"(PB[13]), "+d"(PB[14]), "+d"(PB[15]),
          "+d"(PB[16]), "+d"(PB[17]), "+d"(PB[18]), "+d"(PB[19]), "+d"(PB[20]), "+d"(PB[21]), "+d"(PB[22]), "+d"(PB[23]),
          "+d"(PB[24]), "+d"(PB[25]), "+d"(PB[26]), "+d"(PB[27]), "+d"(PB[28]), "+d"(PB[29]), "+d"(PB[30]), "+d"(PB[31])
        : "r"(tile_sh), "r"(frag_sh), "l"(acc_row), "r"(gsel), "r"(g), "r"(q), "r"(stage_w), "l"(xg), "l"(ld_bytes), "r"(stage_s),
          "r"(comb_rd), "r"(comb_word), "r"(comb_bit), "l"(xg_frag)
        : "memory");
    return code;
}

}  // namespace rr
