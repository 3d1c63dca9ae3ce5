// rr_sweep.cuh — the interpreter kernel (sm_100a).
//
// One launch sweeps a whole neighbourhood over all samples of this GPU:
//   grid  = (tile workers, program chunks); block = 256 threads; S samples per thread.
//   A block loops over sample tiles of T = 256*S rows. Per tile it stages the feature
//   columns its chunk reads (plus y / centred y) into shared memory with TMA bulk copies
//   (cp.async.bulk -> UBLKCP) completing on an mbarrier, then every warp walks the chunk's
//   instruction stream (rr_isa.h). The per-sample state is the fp64 accumulator t[S] in
//   registers; operands are shared-memory tile columns (conflict-free: lane <-> sample) or
//   immediates, so X and y are read from HBM exactly once per sweep.
//   A DOT instruction forms the thread's partial sum over its S samples and feeds a
//   register-resident binary-counter butterfly: after 32 DOTs each lane holds the warp
//   total of one of them (31 shuffle+add steps for 32 reductions instead of 160), which is
//   added with one coalesced fire-and-forget RED.ADD.F64 to the warp's PRIVATE accumulator
//   row in global memory (deterministic: one writer per address, fixed order). Rows are
//   summed by rr_reduce_rows afterwards.
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

namespace rr {

constexpr int kSweepThreads = 256;
constexpr int kSweepWarps = kSweepThreads / 32;

struct SweepArgs {
    const double *X;        // engine matrix: columns (features, y, yc) of `ld` doubles
    int64_t ld;             // column stride, a multiple of 1024
    int64_t n;              // valid samples
    const RRIns *ins;
    const RRChunk *chunks;
    const int32_t *cols;
    double *acc;            // [gridDim.x * 8][acc_stride] per-warp accumulator rows
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

// One reduction enters the butterfly. lvl[l] holds a pending partial of level l (set when bit l
// of the running count is 1). Level-l combine: lanes with bit (4-l) clear keep the EARLIER dot,
// the others the later one; after 5 levels lane L holds dot number bitrev5(L) of the group.
struct DotState {
    double lvl[5];
    uint32_t cnt;  // dots emitted in this chunk so far
};

__device__ __forceinline__ void dot_flush(double x, uint32_t group, int lane, double *acc_row, int32_t n_dots)
{
    const uint32_t idx = group * 32u + (__brev((uint32_t)lane) >> 27);
    if ((int32_t)idx < n_dots) atomicAdd(acc_row + idx, x);  // result unused -> RED.E.ADD.F64
}

__device__ __forceinline__ void dot_emit(DotState &st, double x, int lane, double *acc_row, int32_t n_dots)
{
    const uint32_t c = st.cnt++;
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        if (((c >> l) & 1u) == 0u) {
            st.lvl[l] = x;
            return;
        }
        const int mask = 16 >> l;
        const bool upper = (lane & mask) != 0;
        const double p = st.lvl[l];
        const double send = upper ? p : x;
        const double keep = upper ? x : p;
        x = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
    dot_flush(x, c >> 5, lane, acc_row, n_dots);
}

// double-double helpers (RI_DOTDD)
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

template <int S>
__global__ void __launch_bounds__(kSweepThreads, (S >= 4 ? 1 : 2)) rr_sweep_kernel(const SweepArgs a)
{
    constexpr int TH = kSweepThreads;
    constexpr int T = TH * S;
    extern __shared__ __align__(128) unsigned char rr_smem[];
    double *tile = reinterpret_cast<double *>(rr_smem);
    __shared__ __align__(8) uint64_t mbar;

    const RRChunk ch = a.chunks[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *acc_row = a.acc + ((size_t)blockIdx.x * kSweepWarps + warp) * (size_t)a.acc_stride + ch.dot_base;
    const uint4 *prog = reinterpret_cast<const uint4 *>(a.ins + ch.pc_begin);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;

    for (int tile_i = blockIdx.x; tile_i < a.n_tiles; tile_i += gridDim.x) {
        const int64_t base = (int64_t)tile_i * T;
        if (warp == 0) {
            // order this block's earlier generic-proxy accesses to the tile before the async writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) mbar_expect_tx(&mbar, (uint32_t)(ch.n_cols * T * 8));
            __syncwarp();
            for (int c = lane; c < ch.n_cols; c += 32)
                tma_load_1d(tile + (size_t)c * T, a.X + (size_t)a.cols[ch.col_begin + c] * a.ld + base,
                            (uint32_t)(T * 8), &mbar);
        }
        mbar_wait(&mbar, parity);
        parity ^= 1u;

        const bool partial = base + T > a.n;
        bool valid[S];
#pragma unroll
        for (int s = 0; s < S; ++s) valid[s] = base + tid + s * TH < a.n;

        double t[S];
#pragma unroll
        for (int s = 0; s < S; ++s) t[s] = 0.0;
        DotState ds;
        uint32_t ddcnt = 0;
        ds.cnt = 0;
#pragma unroll
        for (int l = 0; l < 5; ++l) ds.lvl[l] = 0.0;

        uint4 nx = __ldg(prog);
        for (int pc = 0;; ++pc) {
            const uint4 in = nx;
            nx = __ldg(prog + pc + 1);  // the stream is padded: reading one past RI_END is safe
            const uint32_t w0 = in.x, w1 = in.y;
            const uint32_t op = w0 & 0xffu;
            const double imm = __hiloint2double((int)in.w, (int)in.z);
            if (op == RI_END) break;
            if (op >= RI_LOAD && op <= RI_AXPY && op != RI_ST && op != RI_STG) {
                // operand fetch
                double u[S];
                if (w0 & RF_CONST) {
#pragma unroll
                    for (int s = 0; s < S; ++s) u[s] = imm;
                } else {
                    const double *p = tile + (size_t)w1 * T + tid;
#pragma unroll
                    for (int s = 0; s < S; ++s) u[s] = p[s * TH];
                }
                if (op == RI_LOAD) {
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = u[s];
                    continue;
                }
                if (op == RI_AXPY) {
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dadd_rn(t[s], __dmul_rn(imm, u[s]));
                    continue;
                }
                if (w0 & RF_SWAP) {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const double x = t[s];
                        t[s] = u[s];
                        u[s] = x;
                    }
                }
                switch (op) {
                case RI_ADD:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dadd_rn(t[s], u[s]);
                    break;
                case RI_SUB:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dsub_rn(t[s], u[s]);
                    break;
                case RI_MUL:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], u[s]);
                    break;
                case RI_DIV:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = __ddiv_rn(t[s], u[s]);
                    break;
                case RI_POW:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = pow(t[s], u[s]);
                    break;
                case RI_LT:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = t[s] < u[s] ? 1.0 : 0.0;
                    break;
                case RI_GT:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = t[s] > u[s] ? 1.0 : 0.0;
                    break;
                case RI_EQ:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = t[s] == u[s] ? 1.0 : 0.0;
                    break;
                case RI_NE:
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = t[s] != u[s] ? 1.0 : 0.0;
                    break;
                case RI_MIN:  // a < b ? a : b, node.cpp:82
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = t[s] < u[s] ? t[s] : u[s];
                    break;
                case RI_MAX:  // a > b ? a : b, node.cpp:88
#pragma unroll
                    for (int s = 0; s < S; ++s) t[s] = t[s] > u[s] ? t[s] : u[s];
                    break;
                default:
                    break;
                }
                continue;
            }
            switch (op) {
            case RI_ST: {
                double *p = tile + (size_t)w1 * T + tid;
#pragma unroll
                for (int s = 0; s < S; ++s) p[s * TH] = t[s];
                break;
            }
            case RI_STG: {
                double *p = a.stg + (size_t)w1 * a.ld_stg + base + tid;
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (valid[s]) p[s * TH] = t[s];
                break;
            }
            case RI_SIN:
#pragma unroll
                for (int s = 0; s < S; ++s) t[s] = sin(t[s]);
                break;
            case RI_COS:
#pragma unroll
                for (int s = 0; s < S; ++s) t[s] = cos(t[s]);
                break;
            case RI_LN:
#pragma unroll
                for (int s = 0; s < S; ++s) t[s] = log(t[s]);
                break;
            case RI_EXP:
#pragma unroll
                for (int s = 0; s < S; ++s) t[s] = exp(t[s]);
                break;
            case RI_SQRT:
#pragma unroll
                for (int s = 0; s < S; ++s) t[s] = sqrt(t[s]);
                break;
            case RI_SQR:
#pragma unroll
                for (int s = 0; s < S; ++s) t[s] = __dmul_rn(t[s], t[s]);
                break;
            case RI_DOT:
            case RI_DOTDD: {
                const uint32_t ka = RR_DOT_KA(w0), kb = RR_DOT_KB(w0);
                double av[S], bv[S];
                if (ka == RD_TOS) {
#pragma unroll
                    for (int s = 0; s < S; ++s) av[s] = t[s];
                } else {
                    const double *p = tile + (size_t)(w1 & 0xffffu) * T + tid;
#pragma unroll
                    for (int s = 0; s < S; ++s) av[s] = p[s * TH];
                }
                if (kb == RD_TOS) {
#pragma unroll
                    for (int s = 0; s < S; ++s) bv[s] = t[s];
                } else if (kb == RD_ONE) {
#pragma unroll
                    for (int s = 0; s < S; ++s) bv[s] = 1.0;
                } else {
                    const double *p = tile + (size_t)(w1 >> 16) * T + tid;
#pragma unroll
                    for (int s = 0; s < S; ++s) bv[s] = p[s * TH];
                }
                if (op == RI_DOT) {
                    double v = 0.0;
                    if (!partial) {
#pragma unroll
                        for (int s = 0; s < S; ++s) v = fma(av[s], bv[s], v);
                    } else {
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            if (valid[s]) v = fma(av[s], bv[s], v);
                    }
                    dot_emit(ds, v, lane, acc_row, ch.n_dots);
                } else {
                    double hi = 0.0, lo = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s)
                        if (valid[s]) dd_add_prod(hi, lo, av[s], bv[s]);
#pragma unroll
                    for (int m = 16; m > 0; m >>= 1) {
                        const double h2 = __shfl_xor_sync(0xffffffffu, hi, m);
                        const double l2 = __shfl_xor_sync(0xffffffffu, lo, m);
                        dd_add(hi, lo, h2, l2);
                    }
                    // a double-double plan holds DOTDD reductions only (rr_plan.cpp): they bypass the
                    // butterfly; output i occupies the (hi, lo) pair at 2i in the warp's private row
                    if (lane == 0) {
                        double *q = acc_row + 2u * ddcnt;
                        double ah = q[0], al = q[1];
                        dd_add(ah, al, hi, lo);
                        q[0] = ah;
                        q[1] = al;
                    }
                    ++ddcnt;
                }
                break;
            }
            case RI_CLSMET: {
                // rils_rols_cpp.cpp:51-86 on yhat = t, y = tile column b
                const double *py = tile + (size_t)(w1 >> 16) * T + tid;
                double acc = 0.0, ll = 0.0, al = 0.0;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    if (!valid[s]) continue;
                    const double yp = t[s], yy = py[s * TH];
                    const double ypib = yp >= 0.5 ? 1.0 : 0.0;
                    const double yib = yy >= 0.5 ? 1.0 : 0.0;
                    if (ypib == yib) acc += 1.0;
                    const double prob = 1.0 / (1.0 + exp(-2.0 * (yp - 0.5)));
                    const double lli = (1.0 - yib) * log(1.0 - prob) + yib * log(prob);
                    ll -= lli;
                    al += fabs(yib - yp);
                }
                dot_emit(ds, acc, lane, acc_row, ch.n_dots);
                dot_emit(ds, ll, lane, acc_row, ch.n_dots);
                dot_emit(ds, al, lane, acc_row, ch.n_dots);
                break;
            }
            default:
                break;
            }
        }
        // drain the butterfly: pad the last group with zeros
        while (ds.cnt & 31u) dot_emit(ds, 0.0, lane, acc_row, ch.n_dots);
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

// same for double-double pairs laid out (hi, lo) at (i, i+1): used when the plan is a DOTDD plan
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
