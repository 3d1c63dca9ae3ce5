// rr_sweep_g8.cuh — the interpreter kernel for G8 plans (rr_isa.h RI_GRAM8; hot loop: rr_sweep_core_g8.cuh).
//
// Same launch shape and tile handling as rr_sweep_kernel<4, 128, false> (rr_sweep.cuh): grid = (tile workers, program
// chunks), 128 threads, 4 samples per thread, tiles of 512 rows staged by TMA bulk copies on an mbarrier, the
// instruction stream streamed through a double-buffered shared-memory window. What differs is the reduction path:
// no ring, no per-lane partials; the block's dynamic shared memory is [staging rows: 4 warps x 80 warp totals]
// [tile: columns of 512 doubles, 32 bytes of padding each]. Every tile - full or partial - runs in the PTX core.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rr_isa.h"
#include "rr_sweep.cuh"
#include "rr_sweep_core_g8.cuh"

namespace rr {

constexpr int kG8Threads = 128;
constexpr int kG8Tile = 512;                       // samples per tile
constexpr uint32_t kG8ColBytes = kG8Tile * 8 + 32;  // padded column stride
constexpr uint32_t kG8HalfBytes = kG8Tile * 4;
constexpr uint32_t kG8StageBytes = 4 * 80 * 8;      // 4 warps x 80 outputs
// instruction windows of 32 (the other kernels stream 64 at a time): the kilobyte that frees, with what was spare, is
// the tile's 27th column - a seventh row slot next to 20 staged features, and a group of seven rows costs the DMMA
// pipe what a group of six does
constexpr int kG8Window = RR_G8_INS_WINDOW;
constexpr size_t kG8StaticBytes = 2 * (kG8Window + 2) * 16 + 4 * 8;
static_assert(kG8ColBytes == 4128 && kG8HalfBytes == 2048, "rr_sweep_core_g8.cuh is written for this geometry");
constexpr size_t g8_dyn_smem(int cols) { return (size_t)kG8StageBytes + (size_t)cols * kG8ColBytes; }

__global__ void __launch_bounds__(kG8Threads, 2) rr_sweep_g8_kernel(const SweepArgs a)
{
    constexpr int T = kG8Tile;
    extern __shared__ __align__(128) unsigned char rr_dyn[];  // [staging][tile]
    __shared__ __align__(16) unsigned char rr_static[kG8StaticBytes];
    uint4(*ibuf)[kG8Window + 2] = reinterpret_cast<uint4(*)[kG8Window + 2]>(rr_static);
    uint64_t &mbar_tile = *reinterpret_cast<uint64_t *>(rr_static + 2 * (kG8Window + 2) * 16);
    uint64_t *mbar_ins = reinterpret_cast<uint64_t *>(rr_static + 2 * (kG8Window + 2) * 16 + 16);

    const RRChunk ch = a.chunks[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t g = (uint32_t)lane >> 2, q = (uint32_t)lane & 3u;
    const uint4 *prog = reinterpret_cast<const uint4 *>(a.ins + ch.pc_begin);
    const int n_win = (ch.n_ins + kG8Window - 1) / kG8Window;
    const uint32_t tbase = (uint32_t)tid * 16u;  // this thread's sample pair inside a column half
    const uint32_t dyn_sh = smem_u32(rr_dyn);
    const uint32_t stage_sh = dyn_sh;
    unsigned char *const tile_ptr = rr_dyn + kG8StageBytes;
    const uint32_t tile0_sh = dyn_sh + kG8StageBytes;
    const uint32_t tile_sh = tile0_sh + tbase;
    // fragment ownership: lane (g, q) of warp w holds samples 64 w + 4 step + q of each tile half
    const uint32_t frag_sh = tile0_sh + ((uint32_t)warp * 64u + q) * 8u;
    const uint32_t stage_w = stage_sh + (uint32_t)warp * 640u + (g * 10u + 2u * q) * 8u;
    const uint32_t stage_s = stage_sh + (uint32_t)warp * 640u + (g * 10u + 8u) * 8u;
    const uint32_t comb_rd = stage_sh + (uint32_t)(tid < 80 ? tid : 0) * 8u;
    const uint32_t comb_word = tid < 80 ? (uint32_t)tid >> 5 : 3u;
    const uint32_t comb_bit = 1u << (tid & 31);
    const uint32_t gsel = 0x4440u | (g & 3u);
    double *const acc_row = a.acc + (size_t)blockIdx.x * (size_t)a.acc_stride + ch.dot_base;

    if (tid == 0) {
        mbar_init(&mbar_tile, 1);
        mbar_init(&mbar_ins[0], 1);
        mbar_init(&mbar_ins[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ibuf[0][kG8Window] = make_uint4(RI_WINEND, 0, 0, 0);
        ibuf[1][kG8Window] = make_uint4(RI_WINEND, 0, 0, 0);
        ibuf[0][kG8Window + 1] = make_uint4(RI_END, 0, 0, 0);
        ibuf[1][kG8Window + 1] = make_uint4(RI_END, 0, 0, 0);
    }
    __syncthreads();
    uint32_t tile_parity = 0, ins_parity0 = 0, ins_parity1 = 0;
    // value registers and the B fragment live across tiles only as registers: a chunk's stream sets every pin before it
    // reads it, and what a group multiplies with an unset pin is never stored
    double pr[RR_NREG * 4], pb[32];
#pragma unroll
    for (int i = 0; i < RR_NREG * 4; ++i) pr[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) pb[i] = 0.0;

    for (int tile_i = blockIdx.x; tile_i < a.n_tiles; tile_i += gridDim.x) {
        const int64_t base = (int64_t)tile_i * T;
        if (warp == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) {
                mbar_expect_tx(&mbar_tile, (uint32_t)(ch.n_cols * T * 8));
                mbar_expect_tx(&mbar_ins[0], (uint32_t)(kG8Window * 16));
                tma_load_1d(&ibuf[0][0], prog, (uint32_t)(kG8Window * 16), &mbar_ins[0]);
            }
            __syncwarp();
            for (int c = lane; c < ch.n_cols; c += 32)
                tma_load_1d(tile_ptr + (size_t)c * kG8ColBytes, a.X + (size_t)a.cols[ch.col_begin + c] * a.ld + base, (uint32_t)(T * 8),
                            &mbar_tile);
        }
        mbar_wait(&mbar_tile, tile_parity);
        tile_parity ^= 1u;

        // this thread's samples: pairs (2 tid, 2 tid + 1) of both tile halves
        uint32_t vbits = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (base + (int64_t)(s >> 1) * (T / 2) + 2 * tid + (s & 1) < a.n) vbits |= 1u << s;
        const bool partial = base + T > a.n;  // block-uniform
        const double *xg = a.X + base + 2 * tid;                       // this thread's first sample in engine column 0
        const double *xg_frag = a.X + base + (int64_t)warp * 64 + q;   // this lane's first fragment sample

        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        uint32_t cnt = 0;
        bool running = true;
        for (int win = 0; running; ++win) {
            const int b = win & 1;
            __syncthreads();  // every warp has finished window win-1, so its buffer (the other one) may be refilled
            if (tid == 0 && win + 1 < n_win) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&mbar_ins[b ^ 1], (uint32_t)(kG8Window * 16));
                tma_load_1d(&ibuf[b ^ 1][0], prog + (size_t)(win + 1) * kG8Window, (uint32_t)(kG8Window * 16), &mbar_ins[b ^ 1]);
            }
            if (b == 0) { mbar_wait(&mbar_ins[0], ins_parity0); ins_parity0 ^= 1u; }
            else { mbar_wait(&mbar_ins[1], ins_parity1); ins_parity1 ^= 1u; }
            uint32_t ibp = smem_u32(ibuf[b]);
            for (;;) {
                uint32_t w0, w1;
                double imm;
                const uint32_t code =
                    partial ? rr_core_g8_partial(t0, t1, t2, t3, pr, pb, cnt, vbits, ibp, w0, w1, imm, tile_sh, frag_sh, acc_row, gsel, g, q,
                                                 stage_w, xg, a.ld * 8, stage_s, comb_rd, comb_word, comb_bit, xg_frag)
                            : rr_core_g8_full(t0, t1, t2, t3, pr, pb, cnt, vbits, ibp, w0, w1, imm, tile_sh, frag_sh, acc_row, gsel, g, q,
                                              stage_w, xg, a.ld * 8, stage_s, comb_rd, comb_word, comb_bit, xg_frag);
                if (code == 0) break;
                if (code == 1) { running = false; break; }
                // what the core does not implement: libdevice transcendentals outside the fast ranges, rare operators
                switch (w0 & 0xffu) {
                case RI_SIN: t0 = sin(t0); t1 = sin(t1); t2 = sin(t2); t3 = sin(t3); break;
                case RI_COS: t0 = cos(t0); t1 = cos(t1); t2 = cos(t2); t3 = cos(t3); break;
                case RI_LN: t0 = log(t0); t1 = log(t1); t2 = log(t2); t3 = log(t3); break;
                case RI_EXP: t0 = exp(t0); t1 = exp(t1); t2 = exp(t2); t3 = exp(t3); break;
                case RI_RARE: {
                    const uint32_t aux = w0 >> 8;
                    const uint32_t col = tile_sh + w1 * kG8ColBytes;
                    double u0 = imm, u1 = imm, u2 = imm, u3 = imm;
                    if (!(aux & RB_CONST)) {
                        u0 = lds_f64(col); u1 = lds_f64(col + 8); u2 = lds_f64(col + kG8HalfBytes); u3 = lds_f64(col + kG8HalfBytes + 8);
                    }
                    const bool sw = aux & RB_SWAP;
                    t0 = sw ? rr_rare(aux & 0xfu, u0, t0) : rr_rare(aux & 0xfu, t0, u0);
                    t1 = sw ? rr_rare(aux & 0xfu, u1, t1) : rr_rare(aux & 0xfu, t1, u1);
                    t2 = sw ? rr_rare(aux & 0xfu, u2, t2) : rr_rare(aux & 0xfu, t2, u2);
                    t3 = sw ? rr_rare(aux & 0xfu, u3, t3) : rr_rare(aux & 0xfu, t3, u3);
                    break;
                }
                default: break;  // G8 plans hold no other instruction
                }
            }
        }
        __syncthreads();  // every warp is done with the tile before it is overwritten
    }
}

}  // namespace rr
