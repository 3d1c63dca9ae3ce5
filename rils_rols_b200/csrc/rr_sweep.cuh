// rr_sweep.cuh — the interpreter kernel (sm_100a).
//
// One launch sweeps a whole neighbourhood over all samples of this GPU:
//   grid  = (tile workers, program chunks); block = TH threads; S samples per thread.
//   A block loops over sample tiles of T = TH*S rows. Per tile it stages the feature columns its
//   chunk reads (plus y / centred y) into shared memory with TMA bulk copies (cp.async.bulk ->
//   UBLKCP) completing on an mbarrier, then every warp walks the chunk's instruction stream
//   (rr_isa.h). The stream itself is streamed through a double-buffered shared-memory window by
//   TMA as well, so instruction fetch is a broadcast LDS.128 instead of a trip to L2.
//   The per-sample state is the fp64 accumulator t[S] in registers; operands are tile columns
//   (conflict-free: lane <-> sample) or immediates, so X and y are read from HBM exactly once
//   per sweep.
//   A reduction (MDOT) forms the thread's partial sum over its S samples and feeds a
//   register-resident binary-counter butterfly: after 32 reductions each lane holds the warp
//   total of one of them (31 shuffle+add steps for 32 reductions instead of 160), which is added
//   with one coalesced fire-and-forget RED.ADD.F64 to the warp's PRIVATE accumulator row in global
//   memory (deterministic: one writer per address, fixed order). Rows are summed by
//   rr_reduce_rows afterwards.
//
// Semantics per opcode follow node::evaluate_inner, /root/reference/rils_rols_cpp/node.cpp:23-95
// (IEEE +,-,*,/ and sqrt are bit-identical to the CPU; sin/cos/log/exp/pow are CUDA libdevice,
// <= 1-2 ulp from glibc). The file is compiled with --fmad=false: products and sums are rounded
// separately like the reference's array-at-a-time evaluation; fused multiply-adds appear only
// where written explicitly (the reductions).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rr_isa.h"
#include "rr_sweep_core.cuh"

namespace rr {

constexpr int kInsWindow = 256;  // instructions per shared-memory window (4 KB), two windows
constexpr size_t kSweepStaticSmem = 2 * (kInsWindow + 1) * 16 + 2 * 2048 + 256;  // windows + combine buffers + mbarriers

struct SweepArgs {
    const double *X;        // engine matrix: columns (features, y, yc) of `ld` doubles
    int64_t ld;             // column stride, a multiple of 1024
    int64_t n;              // valid samples
    const RRIns *ins;       // all chunks; every chunk is padded to a multiple of kInsWindow
    const RRChunk *chunks;
    const int32_t *cols;
    double *acc;            // [gridDim.x * warps][acc_stride] per-warp accumulator rows
    int64_t acc_stride;
    double *stg;            // RI_STG target: column u at stg + u * ld_stg
    int64_t ld_stg;
    int32_t n_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// explicit shared-space accesses with 32-bit addresses: plain LDS/STS [R], no generic-window
// arithmetic in the interpreter loop. volatile keeps them ordered among themselves (a slot store
// followed by a load of the same slot), nothing else is constrained.
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// double-double helpers (RI_MDOTDD)
__device__ __forceinline__ void dd_add_prod(double &hi, double &lo, double a, double b)
{
    const double p = a * b;
    const double e = fma(a, b, -p);
    const double s = hi + p;
    const double bb = s - hi;
    const double err = (hi - (s - bb)) + (p - bb);
    hi = s;
    lo += err + e;
}
__device__ __forceinline__ void dd_add(double &hi, double &lo, double h2, double l2)
{
    const double s = hi + h2;
    const double bb = s - hi;
    const double err = (hi - (s - bb)) + (h2 - bb);
    const double t = lo + l2 + err;
    hi = s + t;
    lo = t - (hi - s);
}

// level-l combine of the butterfly: lanes with the mask bit clear keep the EARLIER reduction
__device__ __forceinline__ double bfly(double pending, double x, bool upper, int mask)
{
    const double send = upper ? pending : x;
    const double keep = upper ? x : pending;
    return keep + __shfl_xor_sync(0xffffffffu, send, mask);
}

// Transcendental and rarely generated operators run out of line on a per-thread scratch column of
// the tile: the caller parks t[] there, this function transforms it in place, the caller reloads.
// Keeping libdevice's branchy bodies out of the interpreter loop lets ptxas keep the accumulator
// in fixed registers for the cheap operators (the common case) instead of shuffling copies around
// every dispatch.
template <int S, int TH>
__device__ __noinline__ void rr_slow_op(uint32_t scratch, uint32_t w0, uint32_t operand, double imm)
{
    constexpr uint32_t SSTR = TH * 8u;
    const uint32_t op = w0 & 0xffu, aux = w0 >> 8;
#pragma unroll(S <= 2 ? S : 1)
    for (int s = 0; s < S; ++s) {
        double x = lds_f64(scratch + s * SSTR);
        double r;
        switch (op) {
        case RI_SIN: r = sin(x); break;
        case RI_COS: r = cos(x); break;
        case RI_LN: r = log(x); break;
        case RI_EXP: r = exp(x); break;
        default: {  // RI_RARE
            double u = (aux & RB_CONST) ? imm : lds_f64(operand + s * SSTR);
            if (aux & RB_SWAP) {
                const double tmp = x;
                x = u;
                u = tmp;
            }
            switch (aux & 0xfu) {
            case RR_POW: r = pow(x, u); break;
            case RR_LT: r = x < u ? 1.0 : 0.0; break;
            case RR_GT: r = x > u ? 1.0 : 0.0; break;
            case RR_EQ: r = x == u ? 1.0 : 0.0; break;
            case RR_NE: r = x != u ? 1.0 : 0.0; break;
            case RR_MIN: r = x < u ? x : u; break;  // a < b ? a : b, node.cpp:82
            default: r = x > u ? x : u; break;      // a > b ? a : b, node.cpp:88
            }
            break;
        }
        }
        sts_f64(scratch + s * SSTR, r);
    }
}

// SPECIAL = false: the production interpreter. SPECIAL = true additionally understands the
// double-double reductions (escalation plans) and the classifier-metric reduction; those plans are
// rare and run through a separate instantiation so their code does not burden the common one.
template <int S, int TH, bool SPECIAL>
__global__ void __launch_bounds__(TH, (S * TH >= 1024 ? 1 : 2)) rr_sweep_kernel(const SweepArgs a, const int scratch_col)
{
    constexpr int T = TH * S;
    constexpr int NW = TH / 32;
    constexpr int LOG2T = (T == 128 ? 7 : T == 256 ? 8 : T == 512 ? 9 : T == 1024 ? 10 : 11);
    static_assert((1 << LOG2T) == T, "tile height must be a power of two");
    extern __shared__ __align__(128) double rr_tile[];  // [columns][T]
    __shared__ __align__(16) uint4 ibuf[2][kInsWindow + 1];
    __shared__ __align__(16) double red[2][8][32];  // cross-warp combine of finished reductions
    __shared__ __align__(8) uint64_t mbar_tile;
    __shared__ __align__(8) uint64_t mbar_ins[2];

    const RRChunk ch = a.chunks[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *acc_row = a.acc + (size_t)blockIdx.x * (size_t)a.acc_stride + ch.dot_base;  // one row per block
    const uint4 *prog = reinterpret_cast<const uint4 *>(a.ins + ch.pc_begin);
    const int n_win = (ch.n_ins + kInsWindow - 1) / kInsWindow;
    const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4, up2 = lane & 2, up1 = lane & 1;
    const uint32_t out_slot = __brev((uint32_t)lane) >> 27;  // lane L ends up with reduction bitrev5(L)
    const uint32_t tile_sh = smem_u32(rr_tile) + (uint32_t)tid * 8u;  // this thread's row 0 of column 0
    const uint32_t red_sh = smem_u32(&red[0][warp][out_slot]);
    const uint32_t not_warp0 = warp != 0;
    constexpr uint32_t CSH = LOG2T + 3;                               // log2(bytes per tile column)
    constexpr uint32_t SSTR = TH * 8u;                                // byte stride between a thread's samples
    const uint32_t scratch = tile_sh + ((uint32_t)scratch_col << CSH);

    if (tid == 0) {
        mbar_init(&mbar_tile, 1);
        mbar_init(&mbar_ins[0], 1);
        mbar_init(&mbar_ins[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t tile_parity = 0, ins_parity0 = 0, ins_parity1 = 0;

    for (int tile_i = blockIdx.x; tile_i < a.n_tiles; tile_i += gridDim.x) {
        const int64_t base = (int64_t)tile_i * T;
        if (warp == 0) {
            // order this block's earlier generic-proxy accesses to the tile before the async writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) {
                mbar_expect_tx(&mbar_tile, (uint32_t)(ch.n_cols * T * 8));
                mbar_expect_tx(&mbar_ins[0], (uint32_t)(kInsWindow * 16));
                tma_load_1d(&ibuf[0][0], prog, (uint32_t)(kInsWindow * 16), &mbar_ins[0]);
            }
            __syncwarp();
            for (int c = lane; c < ch.n_cols; c += 32)
                tma_load_1d(rr_tile + ((size_t)c << LOG2T), a.X + (size_t)a.cols[ch.col_begin + c] * a.ld + base,
                            (uint32_t)(T * 8), &mbar_tile);
        }
        mbar_wait(&mbar_tile, tile_parity);
        tile_parity ^= 1u;

        const bool partial = base + T > a.n;
        bool valid[S];
#pragma unroll
        for (int s = 0; s < S; ++s) valid[s] = base + tid + s * TH < a.n;

        double t[S];
#pragma unroll
        for (int s = 0; s < S; ++s) t[s] = 0.0;
        double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0, l4 = 0.0;  // butterfly levels
        uint32_t cnt = 0;                                          // reductions emitted in this chunk
        uint32_t ddcnt = 0;

// feeds one reduction into the butterfly (named registers only: nothing is indexed dynamically)
#define RR_EMIT(val)                                                                         \
    do {                                                                                     \
        double x_ = (val);                                                                   \
        const uint32_t c_ = cnt++;                                                           \
        if (!(c_ & 1u)) { l0 = x_; break; }                                                  \
        x_ = bfly(l0, x_, up16, 16);                                                         \
        if (!(c_ & 2u)) { l1 = x_; break; }                                                  \
        x_ = bfly(l1, x_, up8, 8);                                                           \
        if (!(c_ & 4u)) { l2 = x_; break; }                                                  \
        x_ = bfly(l2, x_, up4, 4);                                                           \
        if (!(c_ & 8u)) { l3 = x_; break; }                                                  \
        x_ = bfly(l3, x_, up2, 2);                                                           \
        if (!(c_ & 16u)) { l4 = x_; break; }                                                 \
        x_ = bfly(l4, x_, up1, 1);                                                           \
        /* 32 reductions done: combine the block's warps in fixed order, one RED per reduction */ \
        const uint32_t buf_ = (c_ >> 5) & 1u;                                                \
        red[buf_][warp][out_slot] = x_;                                                      \
        __syncthreads();                                                                     \
        if (warp == 0) {                                                                     \
            double s_ = red[buf_][0][out_slot];                                              \
            for (int w_ = 1; w_ < NW; ++w_) s_ += red[buf_][w_][out_slot];                   \
            const uint32_t idx_ = (c_ & ~31u) + out_slot;                                    \
            if ((int32_t)idx_ < ch.n_dots) atomicAdd(acc_row + idx_, s_); /* RED.E.ADD.F64 */ \
        }                                                                                    \
    } while (0)

        bool running = true;
        for (int win = 0; running; ++win) {
            const int b = win & 1;
            // every warp has finished window win-1, so its buffer (the other one) may be refilled
            __syncthreads();
            if (tid == 0 && win + 1 < n_win) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&mbar_ins[b ^ 1], (uint32_t)(kInsWindow * 16));
                tma_load_1d(&ibuf[b ^ 1][0], prog + (size_t)(win + 1) * kInsWindow, (uint32_t)(kInsWindow * 16),
                            &mbar_ins[b ^ 1]);
            }
            if (b == 0) { mbar_wait(&mbar_ins[0], ins_parity0); ins_parity0 ^= 1u; }
            else { mbar_wait(&mbar_ins[1], ins_parity1); ins_parity1 ^= 1u; }
            const uint4 *ib = ibuf[b];
            if constexpr (!SPECIAL && S <= 4) {
                if (!partial) {
                    // full tile: the PTX core runs the window; it hands back what it does not implement
                    uint32_t ibp = smem_u32(ib);
                    const uint32_t ib_end = ibp + kInsWindow * 16u;
                    double t0 = t[0], t1 = t[S > 1 ? 1 : 0], t2 = t[S > 2 ? 2 : 0], t3 = t[S > 3 ? 3 : 0];
                    for (;;) {
                        uint32_t w0, w1;
                        double imm;
                        uint32_t code;
#define RR_CORE_ARGS t0, t1, t2, t3, l0, l1, l2, l3, l4, cnt, ibp, w0, w1, imm, ib_end, tile_sh, acc_row, ch.n_dots, \
                     out_slot, (uint32_t)lane, red_sh, not_warp0
                        if constexpr (S == 1 && NW == 4) code = rr_core_s1_w4<SSTR, CSH>(RR_CORE_ARGS);
                        else if constexpr (S == 2 && NW == 4) code = rr_core_s2_w4<SSTR, CSH>(RR_CORE_ARGS);
                        else if constexpr (S == 4 && NW == 4) code = rr_core_s4_w4<SSTR, CSH>(RR_CORE_ARGS);
                        else if constexpr (S == 1) code = rr_core_s1_w8<SSTR, CSH>(RR_CORE_ARGS);
                        else if constexpr (S == 2) code = rr_core_s2_w8<SSTR, CSH>(RR_CORE_ARGS);
                        else code = rr_core_s4_w8<SSTR, CSH>(RR_CORE_ARGS);
#undef RR_CORE_ARGS
                        if (code == 0) break;
                        if (code == 1) { running = false; break; }
                        const uint32_t col = tile_sh + (w1 << CSH);
                        if ((w0 & 0xffu) == RI_STG) {
                            double *p = a.stg + (size_t)w1 * a.ld_stg + base + tid;
                            p[0] = t0;
                            if (S > 1) p[TH] = t1;
                            if (S > 2) p[2 * TH] = t2;
                            if (S > 3) p[3 * TH] = t3;
                        } else {
                            sts_f64(scratch, t0);
                            if (S > 1) sts_f64(scratch + SSTR, t1);
                            if (S > 2) sts_f64(scratch + 2 * SSTR, t2);
                            if (S > 3) sts_f64(scratch + 3 * SSTR, t3);
                            rr_slow_op<S, TH>(scratch, w0, col, imm);
                            t0 = lds_f64(scratch);
                            if (S > 1) t1 = lds_f64(scratch + SSTR);
                            if (S > 2) t2 = lds_f64(scratch + 2 * SSTR);
                            if (S > 3) t3 = lds_f64(scratch + 3 * SSTR);
                        }
                    }
                    t[0] = t0;
                    if (S > 1) t[S > 1 ? 1 : 0] = t1;
                    if (S > 2) t[S > 2 ? 2 : 0] = t2;
                    if (S > 3) t[S > 3 ? 3 : 0] = t3;
                    continue;
                }
            }
            uint4 nx = ib[0];
#pragma unroll 1
            for (int pc = 0; pc < kInsWindow; ++pc) {
                const uint4 in = nx;
                nx = ib[pc + 1];  // one slot of padding follows each window
                const uint32_t w0 = in.x, w1 = in.y;
                const double imm = __hiloint2double((int)in.w, (int)in.z);
                const uint32_t col = tile_sh + (w1 << CSH);
#define COL(s) lds_f64(col + (s) * SSTR)
                switch (w0 & 0xffu) {
                case RI_END:
                    running = false;
                    pc = kInsWindow;
                    break;
                case RI_LOAD_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = imm;
                    break;
                case RI_LOAD_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = COL(s);
                    break;
                case RI_ST:
#pragma unroll
                    for (int s = 0; s < S; ++s) sts_f64(col + s * SSTR, t[s]);
                    break;
                case RI_STG: {
                    double *p = a.stg + (size_t)w1 * a.ld_stg + base + tid;
#pragma unroll
                    for (int s = 0; s < S; ++s)
                        if (valid[s]) p[s * TH] = t[s];
                    break;
                }
                case RI_ADD_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dadd_rn(t[s], imm);
                    break;
                case RI_ADD_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dadd_rn(t[s], COL(s));
                    break;
                case RI_SUB_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(t[s], imm);
                    break;
                case RI_SUB_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(t[s], COL(s));
                    break;
                case RI_RSUB_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(imm, t[s]);
                    break;
                case RI_RSUB_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(COL(s), t[s]);
                    break;
                case RI_MUL_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], imm);
                    break;
                case RI_MUL_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], COL(s));
                    break;
                case RI_DIV_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(t[s], imm);
                    break;
                case RI_DIV_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(t[s], COL(s));
                    break;
                case RI_RDIV_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(imm, t[s]);
                    break;
                case RI_RDIV_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(COL(s), t[s]);
                    break;
                case RI_AXPY:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dadd_rn(t[s], __dmul_rn(imm, COL(s)));
                    break;
                case RI_SIN:
                case RI_COS:
                case RI_LN:
                case RI_EXP:
                case RI_RARE:
#pragma unroll
                    for (int s = 0; s < S; ++s) sts_f64(scratch + s * SSTR, t[s]);
                    rr_slow_op<S, TH>(scratch, w0, col, imm);
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = lds_f64(scratch + s * SSTR);
                    break;
                case RI_SQRT:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = sqrt(t[s]);
                    break;
                case RI_SQR:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], t[s]);
                    break;
                case RI_MDOT: {
                    if (w0 & ((MD_ST | MD_LD) << 8)) {
                        const uint32_t fc = tile_sh + ((w0 >> 24) << CSH);
                        if (w0 & (MD_ST << 8)) {
#pragma unroll
                            for (int s = 0; s < S; ++s) sts_f64(fc + s * SSTR, t[s]);
                        } else {
#pragma unroll
                            for (int s = 0; s < S; ++s) t[s] = lds_f64(fc + s * SSTR);
                        }
                    }
                    if (w0 & (MD_SELF << 8)) {
                        double v = 0.0;
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            if (!partial || valid[s]) v = fma(t[s], t[s], v);
                        RR_EMIT(v);
                    }
                    if (w0 & (MD_ONE << 8)) {
                        double v = 0.0;
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            if (!partial || valid[s]) v += t[s];
                        RR_EMIT(v);
                    }
                    const uint32_t np = (w0 >> 16) & 0xffu;
                    uint32_t q0 = w1, q1 = in.z, q2 = in.w;  // six 16-bit partner columns
#pragma unroll 1
                    for (uint32_t j = 0; j < np; ++j) {
                        const uint32_t p = tile_sh + ((q0 & 0xffffu) << CSH);
                        q0 = __funnelshift_r(q0, q1, 16);
                        q1 = __funnelshift_r(q1, q2, 16);
                        q2 >>= 16;
                        double v = 0.0;
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const double x = lds_f64(p + s * SSTR);
                            if (!partial || valid[s]) v = fma(t[s], x, v);
                        }
                        RR_EMIT(v);
                    }
                    break;
                }
                case RI_MDOTDD: if constexpr (SPECIAL) {
                    // a double-double plan holds MDOTDD reductions only (rr_plan.cpp): they bypass the
                    // butterfly; output i occupies the (hi, lo) pair at 2i in the warp's private row
                    const uint32_t np = (w0 >> 16) & 0xffu;
                    const int has_self = (w0 >> 8) & 1, has_one = (w0 >> 9) & 1;
                    const int n_out = (int)np + has_self + has_one;
                    uint32_t q0 = w1, q1 = in.z, q2 = in.w;
#pragma unroll 1
                    for (int o = 0; o < n_out; ++o) {
                        int kind = 2;  // 0 self, 1 one, 2 column
                        if (has_self && o == 0) kind = 0;
                        else if (has_one && o == has_self) kind = 1;
                        const uint32_t p = tile_sh + ((q0 & 0xffffu) << CSH);
                        if (kind == 2) {
                            q0 = __funnelshift_r(q0, q1, 16);
                            q1 = __funnelshift_r(q1, q2, 16);
                            q2 >>= 16;
                        }
                        double hi = 0.0, lo = 0.0;
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            if (valid[s]) dd_add_prod(hi, lo, t[s], kind == 0 ? t[s] : (kind == 1 ? 1.0 : lds_f64(p + s * SSTR)));
#pragma unroll
                        for (int m = 16; m > 0; m >>= 1) {
                            const double h2 = __shfl_xor_sync(0xffffffffu, hi, m);
                            const double l2_ = __shfl_xor_sync(0xffffffffu, lo, m);
                            dd_add(hi, lo, h2, l2_);
                        }
                        // block combine in double-double (fixed warp order), one writer per (hi, lo) pair
                        if (lane == 0) {
                            red[0][warp][0] = hi;
                            red[0][warp][1] = lo;
                        }
                        __syncthreads();
                        if (tid == 0) {
                            double *q = acc_row + 2u * ddcnt;
                            double ah = q[0], al = q[1];
                            for (int w_ = 0; w_ < NW; ++w_) dd_add(ah, al, red[0][w_][0], red[0][w_][1]);
                            q[0] = ah;
                            q[1] = al;
                        }
                        __syncthreads();
                        ++ddcnt;
                    }
                    break;
                }
                case RI_CLSMET: if constexpr (SPECIAL) {
                    // rils_rols_cpp.cpp:51-86 on yhat = t, y = tile column w1
                    double acc = 0.0, ll = 0.0, al = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (!valid[s]) continue;
                        const double yp = t[s], yy = COL(s);
                        const double ypib = yp >= 0.5 ? 1.0 : 0.0;
                        const double yib = yy >= 0.5 ? 1.0 : 0.0;
                        if (ypib == yib) acc += 1.0;
                        const double prob = 1.0 / (1.0 + exp(-2.0 * (yp - 0.5)));
                        const double lli = (1.0 - yib) * log(1.0 - prob) + yib * log(prob);
                        ll -= lli;
                        al += fabs(yib - yp);
                    }
                    RR_EMIT(acc);
                    RR_EMIT(ll);
                    RR_EMIT(al);
                    break;
                }
                default:
                    break;
                }
#undef COL
            }
        }
        // drain the butterfly: pad the last group with zeros
        while (cnt & 31u) RR_EMIT(0.0);
#undef RR_EMIT
        __syncthreads();  // every warp is done with the tile before it is overwritten
    }
}

// out[i] = sum over rows of acc[row][i], fixed order (deterministic)
__global__ void rr_reduce_rows(const double *acc, int64_t stride, int32_t rows, int32_t n, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += acc[(size_t)r * stride + i];
    out[i] = s;
}

// same for double-double pairs laid out (hi, lo) at (2i, 2i+1): used when the plan is an MDOTDD plan
__global__ void rr_reduce_rows_dd(const double *acc, int64_t stride, int32_t rows, int32_t n_pairs, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    double hi = 0.0, lo = 0.0;
    for (int r = 0; r < rows; ++r) {
        const double *q = acc + (size_t)r * stride + 2 * (size_t)i;
        dd_add(hi, lo, q[0], q[1]);
    }
    out[2 * i] = hi;
    out[2 * i + 1] = lo;
}

}  // namespace rr
