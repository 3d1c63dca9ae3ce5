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
//   Reduction partners that many terms share (base-solution terms, centred target) live in 8 per-sample
//   pin registers; a reduction (RI_MDOT / RI_DOTM) forms the thread's partial over its S samples and
//   parks it in the warp's shared-memory ring; every 8 reductions the warp transposes the ring and stages
//   the 8 warp totals, and every 32 reductions the block's warps add up their staged totals in fixed order
//   and issue one RED.ADD.F64 per reduction to the BLOCK's accumulator row in global memory
//   (deterministic: one writer per address, fixed order; details in rr_sweep_core.cuh). Rows are summed
//   by rr_reduce_rows afterwards. Double-double plans keep one row per warp instead.
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

constexpr int kInsWindow = RR_INS_WINDOW;  // instructions per shared-memory window (1 KB), two windows
// Shared memory of a block besides its tile. Static: instruction windows + mbarriers, padded to 3072 bytes
// so that the dynamic part starts 4096-aligned in the shared WINDOW (1 KB of it is reserved by the
// system). Dynamic, in front of the tile: the reduction rings, 4 KB per warp (plus 256 bytes of staging
// row per warp behind them), aligned to 4096 bytes of the
// window at run time (the PTX core wraps its ring pointer with one LOP3, which needs that alignment).
// The host asks the kernel where its dynamic part starts (n_tiles < 0: probe) and adds the slack.
constexpr size_t kSweepStaticBytes = 3072;
constexpr size_t sweep_static_smem() { return kSweepStaticBytes; }
constexpr size_t sweep_ring_smem(int warps, int slack) { return (size_t)warps * (4096 + 256) + (size_t)slack; }

struct SweepArgs {
    const double *X;        // engine matrix: columns (features, y, yc) of `ld` doubles
    int64_t ld;             // column stride, a multiple of 1024
    int64_t n;              // valid samples
    const RRIns *ins;       // all chunks; every chunk is padded to a multiple of kInsWindow
    const RRChunk *chunks;
    const int32_t *cols;
    double *acc;            // [gridDim.x * acc_rows_per_block][acc_stride] accumulator rows
    int32_t acc_rows_per_block;  // 1, or the number of warps for double-double plans
    int64_t acc_stride;
    double *stg;            // RI_STG target: column u at stg + u * ld_stg
    int64_t ld_stg;
    int32_t dd_ring;        // double-double plans: 1 = warp reduction through the ring, 0 = shuffle tree
    int32_t tile_dbuf;      // tile buffers - 1 (0..3): buffers of tile_buf_doubles each; tiles are fetched that many iterations ahead
    int64_t tile_buf_doubles;
    int32_t n_tiles;        // < 0: probe, the kernel only reports the shared-window address of its dynamic part in acc[0]
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

// rarely generated binary operators (node.cpp:56-93), operand order already resolved
__device__ __forceinline__ double rr_rare(uint32_t which, double x, double u)
{
    switch (which) {
    case RR_POW: return pow(x, u);
    case RR_LT: return x < u ? 1.0 : 0.0;
    case RR_GT: return x > u ? 1.0 : 0.0;
    case RR_EQ: return x == u ? 1.0 : 0.0;
    case RR_NE: return x != u ? 1.0 : 0.0;
    case RR_MIN: return x < u ? x : u;  // a < b ? a : b, node.cpp:82
    default: return x > u ? x : u;      // a > b ? a : b, node.cpp:88
    }
}

// The warp's reduction ring and the block's staging rows (see rr_sweep_core.cuh). C++ twin of the PTX code,
// used by the generic interpreter and to drain what is pending at the end of a tile.
struct RingCtx {
    uint32_t ring_w;         // warp ring base | lane * 8 (shared-space byte address, base 4096-aligned)
    uint32_t ra[4];          // this lane's 4 read addresses in ring half 0 when the warp transposes
    uint32_t stage_sh;       // staging rows of the block: [warp][slot 0..3][8] doubles
    uint32_t stage_w;        // stage_sh + warp * 256 + (lane & 7) * 8
    double *acc_row;         // the block's accumulator row (+ chunk dot_base)
    uint32_t lane, warp, n_warps;
};
// slots [0, n_slots) of every warp's staging row -> the block's accumulator row (reductions base + 8 slot + r)
__device__ __forceinline__ void stage_combine(const RingCtx &rc, uint32_t cnt, uint32_t base, uint32_t n_slots)
{
    asm volatile("bar.sync 1;" ::: "memory");
    if (rc.warp < n_slots && rc.lane < 8u) {
        const uint32_t rd = rc.stage_sh + rc.warp * 64u + rc.lane * 8u;
        double s = lds_f64(rd);
        for (uint32_t w = 1; w < rc.n_warps; ++w) s += lds_f64(rd + w * 256u);
        const uint32_t idx = base + rc.warp * 8u + rc.lane;
        if (idx < cnt) atomicAdd(rc.acc_row + idx, s);  // RED.E.ADD.F64, one writer per address
    }
    asm volatile("bar.sync 1;" ::: "memory");
}
__device__ __forceinline__ void ring_flush(const RingCtx &rc, uint32_t cnt, uint32_t &fl)
{
    __syncwarp();
    const uint32_t h = (fl & 8u) << 8;
    double f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(f[2 * i]), "=d"(f[2 * i + 1]) : "r"(rc.ra[i] + h));
    double s = ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (rc.lane < 8u) sts_f64(rc.stage_w + ((fl & 24u) << 3), s);
    const bool full = (fl & 24u) == 24u;
    fl += 8;
    __syncwarp();
    if (full) stage_combine(rc, cnt, fl - 32u, 4u);  // warp-uniform AND block-uniform: every warp runs the same stream
}
// end of a tile: flush what is pending in the ring, then combine the partly filled group of the staging rows
__device__ __forceinline__ void ring_drain(const RingCtx &rc, uint32_t cnt, uint32_t &fl)
{
    while (fl < cnt) ring_flush(rc, cnt, fl);
    if (fl & 31u) stage_combine(rc, cnt, fl & ~31u, (fl >> 3) & 3u);
}
__device__ __forceinline__ void ring_emit(const RingCtx &rc, uint32_t &cnt, uint32_t &fl, double v)
{
    sts_f64(rc.ring_w + ((cnt & 15u) << 8), v);
    ++cnt;
    if (cnt - fl >= 8u) ring_flush(rc, cnt, fl);
}

// SPECIAL = false: the production interpreter. SPECIAL = true additionally understands the
// double-double reductions (escalation plans) and the classifier-metric reduction; those plans are
// rare and run through a separate instantiation so their code does not burden the common one.
//
// Sample ownership: with S == 4 a thread owns the pairs (2 tid, 2 tid + 1) of both halves of the tile
// (two LDS.128 per operand); otherwise samples tid + s * TH.
template <int S, int TH, bool SPECIAL>
__global__ void __launch_bounds__(TH, (S * TH >= 1024 ? 1 : 2)) rr_sweep_kernel(const SweepArgs a)
{
    constexpr int T = TH * S;
    constexpr int NW = TH / 32;
    constexpr bool PAIRS = (S == 4);
    constexpr uint32_t COLB = T * 8u;   // bytes per tile column
    constexpr uint32_t HALFB = T * 4u;  // byte offset of the second half of a column
    extern __shared__ __align__(128) unsigned char rr_dyn[];  // [slack][rings: NW x 16 rows x 32 lanes][tile: columns x T]
    __shared__ __align__(16) unsigned char rr_static[kSweepStaticBytes];
    static_assert(2 * (kInsWindow + 2) * 16 + 6 * 8 <= kSweepStaticBytes, "static shared memory layout");
    // each window is followed by its sentinel and one padding slot (the core prefetches one instruction ahead)
    uint4(*ibuf)[kInsWindow + 2] = reinterpret_cast<uint4(*)[kInsWindow + 2]>(rr_static);
    uint64_t *mbar_tile = reinterpret_cast<uint64_t *>(rr_static + 2 * (kInsWindow + 2) * 16);  // one per tile buffer
    uint64_t *mbar_ins = reinterpret_cast<uint64_t *>(rr_static + 2 * (kInsWindow + 2) * 16 + 32);

    if (a.n_tiles < 0) {
        if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) a.acc[0] = (double)smem_u32(rr_dyn);
        return;
    }
    const RRChunk ch = a.chunks[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4 *prog = reinterpret_cast<const uint4 *>(a.ins + ch.pc_begin);
    const int n_win = (ch.n_ins + kInsWindow - 1) / kInsWindow;
    // this thread's byte offset inside a column, and of its sample s relative to that
    const uint32_t tbase = PAIRS ? (uint32_t)tid * 16u : (uint32_t)tid * 8u;
    auto soff = [](int s) -> uint32_t {
        return PAIRS ? (uint32_t)(s >> 1) * HALFB + (uint32_t)(s & 1) * 8u : (uint32_t)s * (TH * 8u);
    };
    const uint32_t dyn_sh = smem_u32(rr_dyn);
    const uint32_t ring_sh = (dyn_sh + 4095u) & ~4095u;  // rings of warps 0..NW-1, each 4096-aligned
    const uint32_t stage_sh = ring_sh + NW * 4096u;          // staging rows: 256 bytes per warp
    double *const rr_tile0 = reinterpret_cast<double *>(rr_dyn + (ring_sh - dyn_sh) + NW * (4096u + 256u));
    const uint32_t tile_sh0 = stage_sh + NW * 256u + tbase;
    const int nbuf = 1 + a.tile_dbuf;  // 1..4 tile buffers
    const bool dbuf = nbuf > 1;
    // a program of one window stays in its buffer for the whole launch instead of being fetched again for every tile
    const bool resident_prog = n_win == 1;
    RingCtx rc;
    rc.lane = (uint32_t)lane;
    rc.warp = (uint32_t)warp;
    rc.n_warps = NW;
    rc.stage_sh = stage_sh;
    rc.stage_w = stage_sh + (uint32_t)warp * 256u + (uint32_t)(lane & 7) * 8u;
    rc.ring_w = ring_sh + (uint32_t)warp * 4096u + (uint32_t)lane * 8u;
    {
        const uint32_t r = lane & 7, q = lane >> 3;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i)
            rc.ra[i] = ring_sh + (uint32_t)warp * 4096u + r * 256u + (((4u * q + i + r) & 15u) << 4);
    }
    // one accumulator row per block; double-double plans (which bypass ring and staging) one per warp
    rc.acc_row = a.acc + ((size_t)blockIdx.x * a.acc_rows_per_block + (a.acc_rows_per_block > 1 ? warp : 0)) * (size_t)a.acc_stride +
                 ch.dot_base;

    if (tid == 0) {
        mbar_init(&mbar_tile[0], 1);
        mbar_init(&mbar_tile[1], 1);
        mbar_init(&mbar_tile[2], 1);
        mbar_init(&mbar_tile[3], 1);
        mbar_init(&mbar_ins[0], 1);
        mbar_init(&mbar_ins[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the sentinel behind each window: never overwritten by the window copies
        ibuf[0][kInsWindow] = make_uint4(RI_WINEND, 0, 0, 0);
        ibuf[1][kInsWindow] = make_uint4(RI_WINEND, 0, 0, 0);
        ibuf[0][kInsWindow + 1] = make_uint4(RI_END, 0, 0, 0);
        ibuf[1][kInsWindow + 1] = make_uint4(RI_END, 0, 0, 0);
    }
    __syncthreads();
    uint32_t ins_parity0 = 0, ins_parity1 = 0;
    // warp 0 fetches the staged columns of tile `ti` into tile buffer `bf` (TMA bulk copies on that buffer's mbarrier)
    auto fetch_tile = [&](int ti, int bf) {
        // order this block's earlier generic-proxy accesses to the buffer before the async writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) mbar_expect_tx(&mbar_tile[bf], (uint32_t)(ch.n_cols * T * 8));
        __syncwarp();
        double *dst = rr_tile0 + (size_t)bf * a.tile_buf_doubles;
        const int64_t b0 = (int64_t)ti * T;
        for (int c = lane; c < ch.n_cols; c += 32)
            tma_load_1d(dst + (size_t)c * T, a.X + (size_t)a.cols[ch.col_begin + c] * a.ld + b0, (uint32_t)(T * 8), &mbar_tile[bf]);
    };
    if (resident_prog) {
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&mbar_ins[0], (uint32_t)(kInsWindow * 16));
            tma_load_1d(&ibuf[0][0], prog, (uint32_t)(kInsWindow * 16), &mbar_ins[0]);
        }
        mbar_wait(&mbar_ins[0], 0u);
    }
    if (dbuf && warp == 0)
        for (int k = 0; k + 1 < nbuf; ++k)
            if ((int)blockIdx.x + k * (int)gridDim.x < a.n_tiles) fetch_tile((int)blockIdx.x + k * (int)gridDim.x, k);

    int it = 0;
    for (int tile_i = blockIdx.x; tile_i < a.n_tiles; tile_i += gridDim.x, ++it) {
        const int64_t base = (int64_t)tile_i * T;
        const int bf = it % nbuf;
        double *const rr_tile = rr_tile0 + (size_t)bf * a.tile_buf_doubles;
        (void)rr_tile;
        const uint32_t tile_sh = tile_sh0 + (uint32_t)bf * (uint32_t)(a.tile_buf_doubles * 8);
        if (warp == 0) {
            if (dbuf) {
                // the buffer of iteration it + nbuf - 1 was last read in the previous iteration, which ended in a block barrier
                const int ahead = tile_i + (nbuf - 1) * (int)gridDim.x;
                if (ahead < a.n_tiles) fetch_tile(ahead, (it + nbuf - 1) % nbuf);
            } else {
                fetch_tile(tile_i, 0);
            }
            if (!resident_prog && lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&mbar_ins[0], (uint32_t)(kInsWindow * 16));
                tma_load_1d(&ibuf[0][0], prog, (uint32_t)(kInsWindow * 16), &mbar_ins[0]);
            }
        }
        mbar_wait(&mbar_tile[bf], (uint32_t)(it / nbuf) & 1u);

        const bool partial = base + T > a.n;
        bool valid[S];
#pragma unroll
        for (int s = 0; s < S; ++s) valid[s] = base + (int64_t)((tbase + soff(s)) >> 3) < a.n;
        const double *xg = a.X + base + (tbase >> 3);  // this thread's first sample in engine column 0

        double t[S];
#pragma unroll
        for (int s = 0; s < S; ++s) t[s] = 0.0;
        uint32_t cnt = 0, fl = 0;  // reductions emitted / flushed in this chunk and tile
        uint32_t ddcnt = 0;
        // pins: registers of the PTX core (static indices only) or a local array of the generic path
        double pr[PAIRS ? RR_NREG * 4 : 1];
        double pl[RR_NREG][S];
        if constexpr (PAIRS) {
#pragma unroll
            for (int i = 0; i < RR_NREG * 4; ++i) pr[i] = 0.0;
        }
        int use_pin = -1;  // generic path: pin redirected into the next tile-column operand
        // RI_MDOT of the generic path (also the tail of an instruction carrying RR_THEN_MDOT)
        auto mdot = [&](const uint32_t w0) {
            if (w0 & (MD_SELF << 8)) {
                double v = 0.0;
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (!partial || valid[s]) v = fma(t[s], t[s], v);
                ring_emit(rc, cnt, fl, v);
            }
            if (w0 & (MD_ONE << 8)) {
                double v = 0.0;
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (!partial || valid[s]) v += t[s];
                ring_emit(rc, cnt, fl, v);
            }
            const uint32_t mask = (w0 >> 16) & 0xffu;
#pragma unroll 1
            for (int j = 0; j < RR_NPIN; ++j) {
                if (!((mask >> j) & 1u)) continue;
                double v = 0.0;
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (!partial || valid[s]) v = fma(t[s], pl[j][s], v);
                ring_emit(rc, cnt, fl, v);
            }
            if (w0 >> 24) {  // fused "then pin t"
                const int j = (int)(((w0 >> 24) - 1u) % RR_NREG);
#pragma unroll
                for (int s = 0; s < S; ++s) pl[j][s] = t[s];
            }
        };

        bool running = true;
        for (int win = 0; running; ++win) {
            const int b = win & 1;
            if (!resident_prog) {
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
            }
            const uint4 *ib = ibuf[b];
            if constexpr (!SPECIAL && PAIRS && NW == 4) {
                if (!partial) {
                    // full tile: the PTX core runs the window; it hands back what it does not implement
                    uint32_t ibp = smem_u32(ib);
                    double t0 = t[0], t1 = t[1], t2 = t[2], t3 = t[3];
                    for (;;) {
                        uint32_t w0, w1;
                        double imm;
                        const uint32_t code = rr_core_s4<COLB, HALFB>(t0, t1, t2, t3, pr, cnt, fl, ibp, w0, w1, imm, tile_sh,
                                                                      rc.ring_w, rc.acc_row, rc.lane, rc.ra[0], rc.ra[1],
                                                                      rc.ra[2], rc.ra[3], xg, a.ld * 8, rc.stage_w,
                                                                      stage_sh + (uint32_t)warp * 64u + (uint32_t)(lane & 7) * 8u,
                                                                      (uint32_t)warp * 8u + (uint32_t)(lane & 7),
                                                                      (warp < 4 && lane < 8) ? 1u : 0u);
                        if (code == 0) break;
                        if (code == 1) { running = false; break; }
                        switch (w0 & 0xffu) {
                        case RI_STG: {
                            double *p = a.stg + (size_t)w1 * a.ld_stg + base + (tbase >> 3);
                            *reinterpret_cast<double2 *>(p) = make_double2(t0, t1);
                            *reinterpret_cast<double2 *>(p + T / 2) = make_double2(t2, t3);
                            break;
                        }
                        // four independent evaluations: the compiler interleaves them
                        case RI_SIN: t0 = sin(t0); t1 = sin(t1); t2 = sin(t2); t3 = sin(t3); break;
                        case RI_COS: t0 = cos(t0); t1 = cos(t1); t2 = cos(t2); t3 = cos(t3); break;
                        case RI_LN: t0 = log(t0); t1 = log(t1); t2 = log(t2); t3 = log(t3); break;
                        case RI_EXP: t0 = exp(t0); t1 = exp(t1); t2 = exp(t2); t3 = exp(t3); break;
                        case RI_RARE: {
                            const uint32_t aux = w0 >> 8;
                            const uint32_t col = tile_sh + w1 * COLB;
                            double u0 = imm, u1 = imm, u2 = imm, u3 = imm;
                            if (!(aux & RB_CONST)) {
                                u0 = lds_f64(col); u1 = lds_f64(col + 8); u2 = lds_f64(col + HALFB); u3 = lds_f64(col + HALFB + 8);
                            }
                            const bool sw = aux & RB_SWAP;
                            t0 = sw ? rr_rare(aux & 0xfu, u0, t0) : rr_rare(aux & 0xfu, t0, u0);
                            t1 = sw ? rr_rare(aux & 0xfu, u1, t1) : rr_rare(aux & 0xfu, t1, u1);
                            t2 = sw ? rr_rare(aux & 0xfu, u2, t2) : rr_rare(aux & 0xfu, t2, u2);
                            t3 = sw ? rr_rare(aux & 0xfu, u3, t3) : rr_rare(aux & 0xfu, t3, u3);
                            break;
                        }
                        default: break;  // double-double / classifier reductions only exist in SPECIAL plans
                        }
                    }
                    t[0] = t0; t[S > 1 ? 1 : 0] = t1; t[S > 2 ? 2 : 0] = t2; t[S > 3 ? 3 : 0] = t3;
                    continue;
                }
            }
            uint4 nx = ib[0];
#pragma unroll 1
            for (int pc = 0; pc < kInsWindow; ++pc) {
                const uint4 in = nx;
                nx = ib[pc + 1];  // the sentinel slot follows each window
                const uint32_t w0 = in.x, w1 = in.y;
                const double imm = __hiloint2double((int)in.w, (int)in.z);
                const uint32_t col = tile_sh + w1 * COLB;
                const uint32_t op = w0 & 0xffu;
                // operand of the tile-column forms: the tile column or, after USEP, a pin
                double u[S];
                if (op >= RI_FIRST_M) {
#pragma unroll
                    for (int s = 0; s < S; ++s) u[s] = use_pin >= 0 ? pl[use_pin][s] : lds_f64(col + soff(s));
                    use_pin = -1;
                }
                if (op >= RI_PIN0 && op < RI_FIRST_M) {
                    // value-register forms, RR_NREG opcodes each: PIN, LDP, USEP and the fused MULP, DIVP, RDIVP, CMULP, CDIVP
                    const int j = (int)((op - RI_PIN0) % RR_NREG);
                    switch ((op - RI_PIN0) / RR_NREG) {
                    case 0:
#pragma unroll
                        for (int s = 0; s < S; ++s) pl[j][s] = t[s];
                        break;
                    case 1:
#pragma unroll
                        for (int s = 0; s < S; ++s) t[s] = pl[j][s];
                        break;
                    case 2: use_pin = j; break;
                    case 3:
#pragma unroll
                        for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], pl[j][s]);
                        break;
                    case 4:
#pragma unroll
                        for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(t[s], pl[j][s]);
                        break;
                    case 5:
#pragma unroll
                        for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(pl[j][s], t[s]);
                        break;
                    case 6:
#pragma unroll
                        for (int s = 0; s < S; ++s) t[s] = __dmul_rn(imm, pl[j][s]);
                        break;
                    default:
#pragma unroll
                        for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(imm, pl[j][s]);
                        break;
                    }
                    if ((w0 & RR_THEN_MDOT) && op >= RI_MULP0) mdot(w0);
                    continue;
                }
                if (op >= RI_LDPMUL_M0 && op < RI_PINB0) {
                    // fused register / tile-column forms: LDPMUL_M, LDPDIV_M, LDMDIVP
                    const int j = (int)((op - RI_LDPMUL_M0) % RR_NREG);
                    const uint32_t kind = (op - RI_LDPMUL_M0) / RR_NREG;
#pragma unroll
                    for (int s = 0; s < S; ++s)
                        t[s] = kind == 0 ? __dmul_rn(pl[j][s], u[s]) : (kind == 1 ? __ddiv_rn(pl[j][s], u[s]) : __ddiv_rn(u[s], pl[j][s]));
                    if (w0 & RR_THEN_MDOT) mdot(w0);
                    continue;
                }
                switch (op) {
                case RI_END:
                    running = false;
                    pc = kInsWindow;
                    break;
                case RI_NOP:
                case RI_COMBINE:  // this interpreter combines inside its flush (ring_flush)
                    break;
                case RI_LOAD_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = imm;
                    break;
                case RI_LOAD_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = u[s];
                    break;
                case RI_ST:
#pragma unroll
                    for (int s = 0; s < S; ++s) sts_f64(col + soff(s), t[s]);
                    break;
                case RI_STG: {
                    double *p = a.stg + (size_t)w1 * a.ld_stg + base + (tbase >> 3);
#pragma unroll
                    for (int s = 0; s < S; ++s)
                        if (valid[s]) p[soff(s) >> 3] = t[s];
                    break;
                }
                case RI_LDG: {
                    const double *p = xg + (size_t)w1 * a.ld;  // columns are zero padded to a multiple of 1024 rows
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = p[soff(s) >> 3];
                    break;
                }
                case RI_ADD_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dadd_rn(t[s], imm);
                    break;
                case RI_ADD_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dadd_rn(t[s], u[s]);
                    break;
                case RI_SUB_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(t[s], imm);
                    break;
                case RI_SUB_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(t[s], u[s]);
                    break;
                case RI_RSUB_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(imm, t[s]);
                    break;
                case RI_RSUB_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(u[s], t[s]);
                    break;
                case RI_MUL_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], imm);
                    break;
                case RI_MUL_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], u[s]);
                    break;
                case RI_CMUL_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(imm, u[s]);
                    break;
                case RI_CDIV_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(imm, u[s]);
                    break;
                case RI_MUL_MM: {
                    const uint32_t col2 = tile_sh + in.z * COLB;
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(u[s], lds_f64(col2 + soff(s)));
                    break;
                }
                case RI_MUL_M_ST: {
                    const uint32_t col2 = tile_sh + in.z * COLB;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        t[s] = __dmul_rn(t[s], u[s]);
                        sts_f64(col2 + soff(s), t[s]);
                    }
                    break;
                }
                case RI_DIV_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(t[s], imm);
                    break;
                case RI_DIV_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(t[s], u[s]);
                    break;
                case RI_RDIV_C:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(imm, t[s]);
                    break;
                case RI_RDIV_M:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(u[s], t[s]);
                    break;
                case RI_AXPY:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dadd_rn(t[s], __dmul_rn(imm, u[s]));
                    break;
                case RI_SIN:
#pragma unroll 1
                    for (int s = 0; s < S; ++s) t[s] = sin(t[s]);
                    break;
                case RI_COS:
#pragma unroll 1
                    for (int s = 0; s < S; ++s) t[s] = cos(t[s]);
                    break;
                case RI_LN:
#pragma unroll 1
                    for (int s = 0; s < S; ++s) t[s] = log(t[s]);
                    break;
                case RI_EXP:
#pragma unroll 1
                    for (int s = 0; s < S; ++s) t[s] = exp(t[s]);
                    break;
                case RI_RARE: {
                    const uint32_t aux = w0 >> 8;
#pragma unroll 1
                    for (int s = 0; s < S; ++s) {
                        const double v = (aux & RB_CONST) ? imm : lds_f64(col + soff(s));
                        t[s] = (aux & RB_SWAP) ? rr_rare(aux & 0xfu, v, t[s]) : rr_rare(aux & 0xfu, t[s], v);
                    }
                    break;
                }
                case RI_SQRT:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = sqrt(t[s]);
                    break;
                case RI_SQR:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], t[s]);
                    break;
                case RI_DOTM: {
                    double v = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s)
                        if (!partial || valid[s]) v = fma(t[s], u[s], v);
                    ring_emit(rc, cnt, fl, v);
                    break;
                }
                case RI_MDOT:
                    mdot(w0);
                    break;
                case RI_MDOTDD:
                case RI_DOTMDD: if constexpr (SPECIAL) {
                    // A double-double plan holds double-double reductions only (rr_plan.cpp): they bypass the
                    // ring; output i occupies the (hi, lo) pair at 2i in the warp's private row.
                    // All (up to 10) outputs of the instruction are formed together: each is a serial chain
                    // of ~100 dependent FP64 operations (error-free products, 5 double-double shuffle steps),
                    // and the unrolled code lets the scheduler interleave the chains. Outputs that the
                    // instruction does not ask for are computed on zeros and not stored.
                    const bool single = op == RI_DOTMDD;
                    const uint32_t mask = single ? 0u : (w0 >> 16) & 0xffu;
                    // wanted outputs in order: self, one, pins 0..7 (DOTMDD: the tile column takes slot 0)
                    const uint32_t want = single ? 1u : (((w0 >> 8) & 3u) | (mask << 2));
                    constexpr int NO = RR_NPIN + 2;
                    double hi[NO], lo[NO];
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        hi[o] = 0.0;
                        lo[o] = 0.0;
                    }
                    if (single) {
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            if (valid[s]) dd_add_prod(hi[0], lo[0], t[s], u[s]);
                    } else {
                        // all 10 potential outputs, wanted or not: ten independent chains per sample that the
                        // scheduler interleaves (a branch per output would serialise them); what is not wanted
                        // is not stored
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            if (!valid[s]) continue;
#pragma unroll
                            for (int o = 0; o < NO; ++o) {
                                const double v = o == 0 ? t[s] : (o == 1 ? 1.0 : pl[o >= 2 ? o - 2 : 0][s]);
                                dd_add_prod(hi[o], lo[o], t[s], v);
                            }
                        }
                    }
                    if (a.dd_ring) {
                        // Warp reduction through the warp's ring (unused by double-double plans otherwise): the
                        // k-th wanted output parks its 32 per-lane (hi, lo) pairs in row k (512 bytes); lane
                        // (q, r) adds up a quarter of row r in double-double (rotated order: conflict-free), two
                        // shuffle steps join the quarters, and lanes 0 .. n_out-1 each own one output's running
                        // pair in the warp's private accumulator row. ~5x fewer operations than a 5-level
                        // double-double shuffle tree per output, and the outputs' chains run in parallel lanes.
                        const int n_out = __popc(want);  // <= RR_MDOT_MAX_OUT = 8 rows = the 4096-byte ring
                        const uint32_t ring0 = rc.ring_w - rc.lane * 8u;
                        __syncwarp();
                        {
                            uint32_t wr = ring0 + rc.lane * 16u;
#pragma unroll
                            for (int o = 0; o < NO; ++o) {
                                if (!((want >> o) & 1u)) continue;
                                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(wr), "d"(hi[o]), "d"(lo[o]) : "memory");
                                wr += 512u;
                            }
                        }
                        __syncwarp();
                        const uint32_t r = rc.lane & 7u, q4 = rc.lane >> 3;
                        double sh = 0.0, sl = 0.0;
                        if ((int)r < n_out) {
                            const uint32_t rd = ring0 + r * 512u + q4 * 128u;
                            double ph[8], pl8[8];
#pragma unroll
                            for (uint32_t i = 0; i < 8; ++i)
                                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ph[i]), "=d"(pl8[i]) : "r"(rd + (((i + r) & 7u) << 4)));
                            // pairwise tree: 3 dependent double-double additions instead of 8
#pragma unroll
                            for (int w = 1; w < 8; w <<= 1)
#pragma unroll
                                for (int i = 0; i < 8; i += 2 * w) dd_add(ph[i], pl8[i], ph[i + w], pl8[i + w]);
                            sh = ph[0];
                            sl = pl8[0];
                        }
#pragma unroll
                        for (int m = 8; m <= 16; m <<= 1) {
                            const double h2 = __shfl_xor_sync(0xffffffffu, sh, m);
                            const double l2_ = __shfl_xor_sync(0xffffffffu, sl, m);
                            dd_add(sh, sl, h2, l2_);
                        }
                        if ((int)rc.lane < n_out) {
                            double2 *qp = reinterpret_cast<double2 *>(rc.acc_row + 2u * (ddcnt + rc.lane));
                            double2 cur = __ldcg(qp);
                            dd_add(cur.x, cur.y, sh, sl);
                            *qp = cur;
                        }
                        ddcnt += (uint32_t)n_out;
                        __syncwarp();
                    } else {
#pragma unroll
                    for (int m = 16; m > 0; m >>= 1) {
#pragma unroll
                        for (int o = 0; o < NO; ++o) {
                            const double h2 = __shfl_xor_sync(0xffffffffu, hi[o], m);
                            const double l2_ = __shfl_xor_sync(0xffffffffu, lo[o], m);
                            dd_add(hi[o], lo[o], h2, l2_);
                        }
                    }
                    // one writer per (hi, lo) pair of the warp's row
                    {
                        double *q = rc.acc_row + 2u * ddcnt;
                        const int n_out = __popc(want);
                        double2 cur[NO];
#pragma unroll
                        for (int o = 0; o < NO; ++o)  // all running pairs first: one L2 round trip, not ten
                            cur[o] = o < n_out ? __ldcg(reinterpret_cast<const double2 *>(q + 2 * o)) : make_double2(0.0, 0.0);
                        int k = 0;
#pragma unroll
                        for (int o = 0; o < NO; ++o) {
                            if (!((want >> o) & 1u)) continue;
                            // cur[] is indexed by the output's rank k among the wanted ones
                            double ah = 0.0, al = 0.0;
#pragma unroll
                            for (int r = 0; r < NO; ++r)
                                if (r == k) { ah = cur[r].x; al = cur[r].y; }
                            dd_add(ah, al, hi[o], lo[o]);
                            if (lane == 0) *reinterpret_cast<double2 *>(q + 2 * k) = make_double2(ah, al);
                            ++k;
                        }
                        ddcnt += (uint32_t)n_out;
                    }
                    }
                    if (!single && (w0 >> 24)) {  // fused "then pin t"
                        const int j = (int)(((w0 >> 24) - 1u) % RR_NREG);
#pragma unroll
                        for (int s = 0; s < S; ++s) pl[j][s] = t[s];
                    }
                    break;
                }
                case RI_CLSMET: if constexpr (SPECIAL) {
                    // rils_rols_cpp.cpp:51-86 on yhat = t, y = tile column w1
                    double acc = 0.0, ll = 0.0, al = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (!valid[s]) continue;
                        const double yp = t[s], yy = lds_f64(col + soff(s));
                        const double ypib = yp >= 0.5 ? 1.0 : 0.0;
                        const double yib = yy >= 0.5 ? 1.0 : 0.0;
                        if (ypib == yib) acc += 1.0;
                        const double prob = 1.0 / (1.0 + exp(-2.0 * (yp - 0.5)));
                        const double lli = (1.0 - yib) * log(1.0 - prob) + yib * log(prob);
                        ll -= lli;
                        al += fabs(yib - yp);
                    }
                    ring_emit(rc, cnt, fl, acc);
                    ring_emit(rc, cnt, fl, ll);
                    ring_emit(rc, cnt, fl, al);
                    break;
                }
                default:
                    break;
                }
                if ((w0 & RR_THEN_MDOT) && rr_md_fusable(op)) mdot(w0);
            }
        }
        // drain the ring and the staging rows
        ring_drain(rc, cnt, fl);
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
