// rr_engine.cu — the C ABI of include/rr_b200.h on top of the planner and the sm_100a kernels.
//
// Data layout in HBM (per engine = per GPU shard): one column-major matrix of d + 2 columns with
// column stride ld = round_up(n, 1024) doubles, zero padded: features 0..d-1 (the vector<ArrayXd>
// layout of /root/reference/rils_rols_cpp/rils_rols_cpp.cpp:675-698), y, and y - mean(y).
// Everything else (instruction streams, per-block accumulator rows, reduced dots, per-candidate
// solve workspaces, the materialised term matrix of the exact path) lives in grow-only device
// buffers owned by the engine. There is no CPU fallback: without a CUDA device of compute
// capability 10.x every entry point fails with RR_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rr_b200.h"
#include "rr_exact.cuh"
#include "rr_isa.h"
#include "rr_plan.h"
#include "rr_solve.cuh"
#include "rr_sweep.cuh"

namespace {

thread_local std::string g_thread_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {  // pinned
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

double env_double(const char *name, double dflt)
{
    const char *s = std::getenv(name);
    return s && *s ? std::atof(s) : dflt;
}
int env_int(const char *name, int dflt)
{
    const char *s = std::getenv(name);
    return s && *s ? std::atoi(s) : dflt;
}

constexpr int64_t kLdAlign = 1024;
// shared memory per SM usable by blocks: 227 KB per block opt-in limit; the kernel's static part
// (instruction windows + mbarriers) and the 1 KB the driver reserves per block come off the top
constexpr size_t kSmemPerBlockMax = 232448;
constexpr size_t kSmemPerSM = 233472;
constexpr size_t kSmemReserved = 1024;

__global__ void k_transpose_rowmajor(const double *__restrict__ Xr, int64_t n, int32_t d, double *__restrict__ Xc,
                                     int64_t ld)
{
    // 32 x 32 tiles through shared memory: coalesced on both sides
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t i = i0 + r;
        const int j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && j < d) ? Xr[i * d + j] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r;
        const int64_t i = i0 + threadIdx.x;
        if (j < d && i < n) Xc[(int64_t)j * ld + i] = tile[threadIdx.x][r];
    }
}

// deterministic two-stage sum of f(y): mode 0 -> sum y ; mode 1 -> writes yc = y - mean and sums yc, yc^2
__global__ void k_y_stats(const double *y, double *yc, int64_t n, double mean, int mode, double *partial)
{
    __shared__ double s0[256], s1[256];
    double a = 0.0, b = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (mode == 0) {
            a += y[i];
        } else {
            const double v = y[i] - mean;
            yc[i] = v;
            a += v;
            b = fma(v, v, b);
        }
    }
    s0[threadIdx.x] = a;
    s1[threadIdx.x] = b;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) {
            s0[threadIdx.x] += s0[threadIdx.x + st];
            s1[threadIdx.x] += s1[threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s0[0];
        partial[2 * blockIdx.x + 1] = s1[0];
    }
}

__global__ void k_take_ssr(const int32_t *list, int32_t n_list, const int32_t *rbegin, const int32_t *rcand_dot,
                           const double *rdots, double *ssr, uint32_t *flags)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_list) return;
    const int c = list[li];
    const double v = rdots[rcand_dot[rbegin[li]]];
    ssr[c] = v;
    if (!isfinite(v)) flags[c] |= RR_RES_NONFINITE;
}

// DFMA-only microkernel: 8 independent chains per thread
__global__ void k_fp64_peak(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace

struct rr_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t flags = 0;
    int64_t n = 0, ld = 0, n_total = 0;
    int32_t d = 0;
    int exact_max_n = 4096;
    int s_pref = 0, th_pref = 0, occ_pref = 0, slots_pref = 0;  // 0 = auto (env RR_B200_S / _TH / _OCC / _SLOTS)
    DevBuf X;        // (d + 2) * ld doubles
    double y_mean = 0, sst = 0, sum_yc = 0;
    rr::SolveConsts sc{};
    rr_allreduce_fn allreduce = nullptr;
    void *allreduce_user = nullptr;
    int rank = 0, world = 1;
    // grow-only work buffers
    DevBuf d_ins2, d_chunks2, d_cols2, d_acc2;  // second sweep in flight (run_gram plans the batch in two halves)
    DevBuf d_ins, d_chunks, d_cols, d_acc, d_dots, d_rdots, d_tab, d_rtab, d_ws, d_wsoff, d_list, d_coef, d_cs,
        d_nzp, d_ssr, d_flags, d_status, d_delta, d_V, d_A, d_rhs, d_aux, d_perm, d_ctb, d_tid, d_misc, d_gather,
        d_t0, d_t1, d_t2, d_t3, d_t4;  // small per-pass tables of the Gram path
    HostBuf h_stage, h_out;
    rr_stats stats{};
    std::string error;
    float sweep_ms_accum = 0.f;

    int fail(int code, const std::string &msg)
    {
        error = msg;
        return code;
    }
    int cuda_fail(cudaError_t e, const char *what)
    {
        error = std::string(what) + ": " + cudaGetErrorString(e);
        return RR_ERR_CUDA;
    }
    void free_all()
    {
        for (DevBuf *b : {&X, &d_ins2, &d_chunks2, &d_cols2, &d_acc2, &d_ins, &d_chunks, &d_cols, &d_acc, &d_dots, &d_rdots, &d_tab, &d_rtab, &d_ws, &d_wsoff,
                          &d_list, &d_coef, &d_cs, &d_nzp, &d_ssr, &d_flags, &d_status, &d_delta, &d_V, &d_A, &d_rhs,
                          &d_aux, &d_perm, &d_ctb, &d_tid, &d_misc, &d_gather, &d_t0, &d_t1, &d_t2, &d_t3, &d_t4})
            b->release();
        h_stage.release();
        h_out.release();
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

#define CU(call)                                              \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return e->cuda_fail(_e, #call); \
    } while (0)

namespace {

// launch shape of the interpreter: S samples per thread, TH threads per block, `occ` blocks per SM
// planned for (bounds the shared-memory tile and therefore the number of value slots)
struct SweepCfg {
    int S, TH, occ;
    int slack = 4096;  // bytes between the start of the kernel's dynamic shared memory and the 4096-aligned rings
    int T() const { return S * TH; }
    size_t dyn_smem_budget() const  // for the tile
    {
        const size_t per_block = std::min(kSmemPerBlockMax, kSmemPerSM / (size_t)occ - kSmemReserved);
        return per_block - rr::sweep_static_smem() - rr::sweep_ring_smem(TH / 32, slack);
    }
    // columns the planner may use
    int tile_cols() const { return (int)std::min<size_t>(dyn_smem_budget() / ((size_t)T() * 8), 0x3000); }
};

using SweepKernel = void (*)(const rr::SweepArgs);
template <bool SP> SweepKernel sweep_kernel_sel(const SweepCfg &c)
{
    if (c.TH == 128) {
        switch (c.S) {
        case 1: return rr::rr_sweep_kernel<1, 128, SP>;
        case 2: return rr::rr_sweep_kernel<2, 128, SP>;
        default: return rr::rr_sweep_kernel<4, 128, SP>;
        }
    }
    switch (c.S) {
    case 1: return rr::rr_sweep_kernel<1, 256, SP>;
    case 2: return rr::rr_sweep_kernel<2, 256, SP>;
    default: return rr::rr_sweep_kernel<4, 256, SP>;
    }
}
SweepKernel sweep_kernel_for(const SweepCfg &c, bool special)
{
    return special ? sweep_kernel_sel<true>(c) : sweep_kernel_sel<false>(c);
}

// Where does this kernel's dynamic shared memory start in the shared window? (asked once per kernel
// variant: the kernel reports it in probe mode.) Returns the bytes up to the next multiple of 4096.
int sweep_slack(rr_engine *e, SweepKernel kern, int TH)
{
    static std::mutex mu;
    static std::map<const void *, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find((const void *)kern);
    if (it != cache.end()) return it->second;
    int slack = 4096;
    if (e->d_misc.ensure(64) == cudaSuccess) {
        rr::SweepArgs a;
        std::memset(&a, 0, sizeof(a));
        a.acc = e->d_misc.as<double>();
        a.n_tiles = -1;
        kern<<<1, TH, 0, e->stream>>>(a);
        double v = -1.0;
        if (cudaMemcpyAsync(&v, e->d_misc.p, 8, cudaMemcpyDeviceToHost, e->stream) == cudaSuccess &&
            cudaStreamSynchronize(e->stream) == cudaSuccess && v >= 0.0) {
            slack = (int)((4096u - ((uint32_t)v & 4095u)) & 4095u);
            cache[(const void *)kern] = slack;
        }
    }
    return slack;
}

SweepCfg choose_cfg(rr_engine *e)
{
    SweepCfg c;
    // 4 samples per thread (the PTX core of rr_sweep_core.cuh) in 128-thread blocks, 2 blocks per SM:
    // per-dispatch overhead wants more samples per thread, latency hiding wants more warps, and the
    // shared-memory tile caps the samples in flight per SM (profiles/r1_config_sweep.txt)
    const bool big = e->n >= (1 << 15);
    c.S = e->s_pref ? e->s_pref : (big ? 4 : 1);
    c.TH = e->th_pref ? e->th_pref : 128;
    c.occ = e->occ_pref ? e->occ_pref : (c.T() >= 1024 ? 1 : (c.T() >= 512 ? 2 : (c.T() >= 256 ? 3 : 4)));
    c.slack = std::max(sweep_slack(e, sweep_kernel_for(c, false), c.TH), sweep_slack(e, sweep_kernel_for(c, true), c.TH));
    // the tile must stage the feature columns a chunk can touch plus a few value slots (cached terms
    // live in pins first)
    const int want = std::min(e->d, 24) + 3;
    while (c.tile_cols() < want && c.occ > 1) --c.occ;
    while (c.tile_cols() < want && c.S > 1) c.S /= 2;
    return c;
}

template <typename T> int upload(rr_engine *e, DevBuf &buf, const T *src, size_t count)
{
    const size_t bytes = count * sizeof(T);
    CU(buf.ensure(std::max<size_t>(bytes, 16)));
    if (bytes) CU(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, e->stream));
    e->stats.h2d_bytes += bytes;
    return RR_OK;
}

// grid.x of a sweep for a plan with n_chunks chunks
int sweep_gx(rr_engine *e, const SweepCfg &c, bool special, size_t smem, int n_chunks, int n_tiles)
{
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_kernel_for(c, special), c.TH, smem);
    occ = std::max(1, occ);
    const int slots = e->sm_count * occ;
    int gx = std::max(1, slots / std::max(1, n_chunks));
    return std::min(gx, n_tiles);
}

// Runs one plan: zero accumulators, launch the interpreter, reduce rows into `dots` (device),
// all-reduce across ranks when sharded. dd: the plan holds DOTDD reductions only.
// set / dots_off / sync: run_gram plans a large batch in two halves and launches the first while it plans the
// second, so two sweeps can be in flight: `set` picks the device buffers and the event pair, the reduced dots
// land at dots + dots_off (the caller has sized `dots` for both: growing it here would drop the first half),
// and with sync = false the call returns right after the launches (finish_sweep reads the time later).
int run_sweep(rr_engine *e, const rr::SweepPlan &P, const SweepCfg &cfg, DevBuf &dots, bool dd, double *stg, int64_t ld_stg,
              int set = 0, size_t dots_off = 0, bool sync = true)
{
    DevBuf &d_ins = set ? e->d_ins2 : e->d_ins, &d_chunks = set ? e->d_chunks2 : e->d_chunks;
    DevBuf &d_cols = set ? e->d_cols2 : e->d_cols, &d_acc = set ? e->d_acc2 : e->d_acc;
    cudaEvent_t ev0 = e->ev[set ? 4 : 2], ev1 = e->ev[set ? 5 : 3];
    if (P.chunks.empty()) return RR_OK;
    const int T = cfg.T();
    const int NW = cfg.TH / 32;
    const int n_tiles = (int)((e->n + T - 1) / T);
    const size_t smem = (size_t)std::max(P.max_tile_cols, 1) * T * 8 + rr::sweep_ring_smem(NW, cfg.slack);
    if (smem > cfg.dyn_smem_budget() + rr::sweep_ring_smem(NW, cfg.slack)) return e->fail(RR_ERR_INVALID, "internal: plan exceeds the shared-memory tile");
    bool special = dd;
    for (const RRIns &x : P.ins)
        if (RR_OP(x.w0) == RI_CLSMET) { special = true; break; }
    SweepKernel kern = sweep_kernel_for(cfg, special);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemPerBlockMax - rr::sweep_static_smem())));
    const int n_chunks = (int)P.chunks.size();
    int gx = sweep_gx(e, cfg, special, smem, n_chunks, std::max(1, n_tiles));
    const int64_t stride = round_up(std::max(P.n_dots, 1), 32) + 32;
    // keep the accumulator rows within a sane budget
    const size_t row_budget = (size_t)env_double("RR_B200_ACC_BYTES", 6e9);
    const int rpb = dd ? NW : 1;  // accumulator rows per block: double-double plans keep one per warp
    while (gx > 1 && (size_t)gx * rpb * stride * 8 > row_budget) gx = (gx + 1) / 2;
    const int rows = gx * rpb;

    // the kernel streams whole windows of kInsWindow instructions: pad the tail with ENDs
    std::vector<RRIns> ins(P.ins);
    RRIns endi;
    std::memset(&endi, 0, sizeof(endi));
    ins.resize(P.ins.size() + rr::kInsWindow, endi);
    int rc;
    if ((rc = upload(e, d_ins, ins.data(), ins.size()))) return rc;
    if ((rc = upload(e, d_chunks, P.chunks.data(), P.chunks.size()))) return rc;
    if ((rc = upload(e, d_cols, P.cols.data(), P.cols.size()))) return rc;
    if (P.n_dots > 0) {
        CU(d_acc.ensure((size_t)rows * stride * 8));
        CU(cudaMemsetAsync(d_acc.p, 0, (size_t)rows * stride * 8, e->stream));
        CU(dots.ensure((dots_off + (size_t)stride) * 8));
    } else {
        CU(d_acc.ensure(64));
    }
    rr::SweepArgs a;
    a.X = e->X.as<double>();
    a.ld = e->ld;
    a.n = e->n;
    a.ins = d_ins.as<RRIns>();
    a.chunks = d_chunks.as<RRChunk>();
    a.cols = d_cols.as<int32_t>();
    a.acc = d_acc.as<double>();
    a.acc_stride = stride;
    a.acc_rows_per_block = rpb;
    a.stg = stg;
    a.ld_stg = ld_stg;
    a.n_tiles = n_tiles;
    a.dd_ring = env_int("RR_B200_DD_RING", 1);
    CU(cudaEventRecord(ev0, e->stream));
    kern<<<dim3(gx, n_chunks), cfg.TH, smem, e->stream>>>(a);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ev1, e->stream));
    e->stats.sweep_launches++;
    e->stats.kernel_launches++;
    if (P.n_dots > 0) {
        if (!dd) {
            rr::rr_reduce_rows<<<(P.n_dots + 255) / 256, 256, 0, e->stream>>>(d_acc.as<double>(), stride, rows,
                                                                            P.n_dots, dots.as<double>() + dots_off);
        } else {
            rr::rr_reduce_rows_dd<<<(P.n_dots / 2 + 255) / 256, 256, 0, e->stream>>>(
                d_acc.as<double>(), stride, rows, P.n_dots / 2, dots.as<double>() + dots_off);
        }
        CU(cudaGetLastError());
        e->stats.kernel_launches++;
        if (e->allreduce && e->world > 1) {
            if (!dd) {
                if (e->allreduce(dots.as<double>() + dots_off, (size_t)P.n_dots, e->stream, e->allreduce_user))
                    return e->fail(RR_ERR_COLLECTIVE, "all-reduce hook failed");
            } else {
                // double-double pairs must not be summed in fp64: gather every rank's pairs with a
                // sum-of-disjoint-segments all-reduce (exact), then add them in double-double here
                const size_t len = (size_t)P.n_dots;
                CU(e->d_gather.ensure(len * e->world * 8));
                CU(cudaMemsetAsync(e->d_gather.p, 0, len * e->world * 8, e->stream));
                CU(cudaMemcpyAsync(e->d_gather.as<double>() + len * e->rank, dots.as<double>() + dots_off, len * 8, cudaMemcpyDeviceToDevice,
                                   e->stream));
                if (e->allreduce(e->d_gather.p, len * e->world, e->stream, e->allreduce_user))
                    return e->fail(RR_ERR_COLLECTIVE, "all-reduce hook failed");
                rr::rr_reduce_rows_dd<<<(P.n_dots / 2 + 255) / 256, 256, 0, e->stream>>>(
                    e->d_gather.as<double>(), (int64_t)len, e->world, P.n_dots / 2, dots.as<double>() + dots_off);
                CU(cudaGetLastError());
                e->stats.kernel_launches++;
            }
        }
    }
    e->stats.distinct_dots += P.n_dot_ins;
    e->stats.w_shared += P.w_issued;
    if (!sync) return RR_OK;
    // sweep time is read after the synchronisation
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    e->sweep_ms_accum += ms;
    return RR_OK;
}

// after the stream has been synchronised: account the time of a sweep that was launched with sync = false
void finish_sweep(rr_engine *e, int set)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->ev[set ? 4 : 2], e->ev[set ? 5 : 3]) == cudaSuccess) e->sweep_ms_accum += ms;
}

rr::PlanLimits limits_for(rr_engine *e, const SweepCfg &cfg, int n_cand)
{
    rr::PlanLimits lim;
    lim.tile_cols = cfg.tile_cols();
    if (e->slots_pref > 0) lim.max_slots = e->slots_pref;
    lim.no_cse = (e->flags & RR_FLAG_NO_CSE) != 0;
    lim.fuse = env_int("RR_B200_FUSE", 1) != 0;
    lim.mdot_rows = cfg.S == 4 && cfg.TH == 128;  // what rr_core_s4 expects behind every RI_MDOT carrier
    const int T = cfg.T();
    const int n_tiles = (int)std::max<int64_t>(1, (e->n + T - 1) / T);
    // enough independent program chunks to occupy the GPU when there are few sample tiles
    const int want_blocks = e->sm_count * std::max(2, cfg.occ);
    lim.target_chunks = n_tiles >= want_blocks ? 1 : std::min(std::max(1, n_cand), (want_blocks + n_tiles - 1) / n_tiles);
    return lim;
}

int compute_y_stats(rr_engine *e)
{
    // mean and centred sums, deterministic; all-reduced when sharded
    const int blocks = 296;
    CU(e->d_misc.ensure(blocks * 2 * 8 + 64));
    std::vector<double> part(blocks * 2);
    double *y = e->X.as<double>() + (size_t)e->d * e->ld;
    double *yc = e->X.as<double>() + (size_t)(e->d + 1) * e->ld;
    k_y_stats<<<blocks, 256, 0, e->stream>>>(y, yc, e->n, 0.0, 0, e->d_misc.as<double>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(part.data(), e->d_misc.p, blocks * 2 * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    double sum_y = 0.0;
    for (int i = 0; i < blocks; ++i) sum_y += part[2 * i];
    double tot[2] = {sum_y, (double)e->n};
    if (e->allreduce && e->world > 1) {
        CU(cudaMemcpyAsync(e->d_misc.p, tot, 16, cudaMemcpyHostToDevice, e->stream));
        if (e->allreduce(e->d_misc.p, 2, e->stream, e->allreduce_user)) return e->fail(RR_ERR_COLLECTIVE, "all-reduce hook failed");
        CU(cudaMemcpyAsync(tot, e->d_misc.p, 16, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    e->n_total = (int64_t)llround(tot[1]);
    e->y_mean = tot[0] / tot[1];
    k_y_stats<<<blocks, 256, 0, e->stream>>>(y, yc, e->n, e->y_mean, 1, e->d_misc.as<double>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(part.data(), e->d_misc.p, blocks * 2 * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < blocks; ++i) { s1 += part[2 * i]; s2 += part[2 * i + 1]; }
    double tot2[2] = {s1, s2};
    if (e->allreduce && e->world > 1) {
        CU(cudaMemcpyAsync(e->d_misc.p, tot2, 16, cudaMemcpyHostToDevice, e->stream));
        if (e->allreduce(e->d_misc.p, 2, e->stream, e->allreduce_user)) return e->fail(RR_ERR_COLLECTIVE, "all-reduce hook failed");
        CU(cudaMemcpyAsync(tot2, e->d_misc.p, 16, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    e->sum_yc = tot2[0];
    e->sst = tot2[1];
    e->sc.n_total = (double)e->n_total;
    e->sc.y_mean = e->y_mean;
    e->sc.sum_yc = e->sum_yc;
    e->sc.sst = e->sst;
    e->sc.rho_accurate = env_double("RR_B200_RHO_ACCURATE", 1e-5);
    e->sc.rho_escalate = env_double("RR_B200_RHO_ESCALATE", 1e-11);
    e->sc.ssr_rel_tol = env_double("RR_B200_SSR_TOL", 1e-11);
    return RR_OK;
}

int create_common(const double *Xsrc, const double *y, int64_t n, int32_t d, int32_t device, uint32_t flags,
                  bool rowmajor, rr_engine **out)
{
    if (!out) { g_thread_error = "out is null"; return RR_ERR_INVALID; }
    *out = nullptr;
    if (!Xsrc || !y || n <= 0 || d <= 0) { g_thread_error = "rr_engine_create: bad arguments"; return RR_ERR_INVALID; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        g_thread_error = "no CUDA device: this engine has no CPU fallback";
        return RR_ERR_NO_DEVICE;
    }
    if (device < 0) cudaGetDevice(&device);
    if (device >= count) { g_thread_error = "device ordinal out of range"; return RR_ERR_INVALID; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        g_thread_error = std::string("device ") + prop.name + " is not compute capability 10.x (sm_100a kernels only)";
        return RR_ERR_NO_DEVICE;
    }
    rr_engine *e = new rr_engine();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->flags = flags;
    e->n = n;
    e->d = d;
    e->n_total = n;
    e->ld = round_up(n, kLdAlign);
    e->exact_max_n = env_int("RR_B200_EXACT_MAX_N", 4096);
    e->s_pref = env_int("RR_B200_S", 0);
    e->th_pref = env_int("RR_B200_TH", 0);
    e->occ_pref = env_int("RR_B200_OCC", 0);
    e->slots_pref = env_int("RR_B200_SLOTS", 0);
    if (e->th_pref != 128 && e->th_pref != 256) e->th_pref = 0;
    if (e->s_pref != 1 && e->s_pref != 2 && e->s_pref != 4 && !(e->s_pref == 8 && e->th_pref == 128)) e->s_pref = 0;
    auto bail = [&](int code) {
        g_thread_error = e->error;
        e->free_all();
        delete e;
        return code;
    };
#define CUC(call)                                                     \
    do {                                                              \
        cudaError_t _e = (call);                                      \
        if (_e != cudaSuccess) { e->cuda_fail(_e, #call); return bail(RR_ERR_CUDA); } \
    } while (0)
    CUC(cudaSetDevice(device));
    CUC(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (auto &ev : e->ev) CUC(cudaEventCreate(&ev));
    const size_t cols = (size_t)d + 2;
    CUC(e->X.ensure(cols * e->ld * 8));
    CUC(cudaMemsetAsync(e->X.p, 0, cols * e->ld * 8, e->stream));
    const bool on_dev = (flags & RR_FLAG_X_DEVICE) != 0;
    const cudaMemcpyKind kind = on_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!rowmajor) {
        CUC(cudaMemcpy2DAsync(e->X.p, e->ld * 8, Xsrc, (size_t)n * 8, (size_t)n * 8, d, kind, e->stream));
    } else {
        const double *src = Xsrc;
        if (!on_dev) {
            CUC(e->d_V.ensure((size_t)n * d * 8));
            CUC(cudaMemcpyAsync(e->d_V.p, Xsrc, (size_t)n * d * 8, cudaMemcpyHostToDevice, e->stream));
            src = e->d_V.as<double>();
        }
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((d + 31) / 32));
        k_transpose_rowmajor<<<grid, dim3(32, 8), 0, e->stream>>>(src, n, d, e->X.as<double>(), e->ld);
        CUC(cudaGetLastError());
    }
    CUC(cudaMemcpyAsync(e->X.as<double>() + (size_t)d * e->ld, y, (size_t)n * 8, kind, e->stream));
    CUC(cudaStreamSynchronize(e->stream));
    const int rc = compute_y_stats(e);
    if (rc) return bail(rc);
#undef CUC
    *out = e;
    return RR_OK;
}

// ---- result staging -------------------------------------------------------------------------
int ensure_result_buffers(rr_engine *e, int n_cand, int n_coef)
{
    CU(e->d_coef.ensure((size_t)std::max(n_coef, 1) * 8));
    CU(e->d_cs.ensure((size_t)std::max(n_coef, 1) * 8));
    CU(e->d_nzp.ensure((size_t)n_cand * 4));
    CU(e->d_ssr.ensure((size_t)n_cand * 8));
    CU(e->d_flags.ensure((size_t)n_cand * 4));
    CU(e->d_status.ensure((size_t)n_cand * 4));
    CU(e->d_delta.ensure((size_t)n_cand * 8));
    CU(cudaMemsetAsync(e->d_flags.p, 0, (size_t)n_cand * 4, e->stream));
    CU(cudaMemsetAsync(e->d_nzp.p, 0, (size_t)n_cand * 4, e->stream));
    return RR_OK;
}

int download_results(rr_engine *e, const rr_batch *b, rr_result *res, bool with_coef)
{
    const int nc = b->n_cand;
    const int n_coef = b->cand_term_begin[nc] + nc;
    if (res->ssr) CU(cudaMemcpyAsync(res->ssr, e->d_ssr.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, e->stream));
    if (res->flags) CU(cudaMemcpyAsync(res->flags, e->d_flags.p, (size_t)nc * 4, cudaMemcpyDeviceToHost, e->stream));
    if (res->nonzero_pivots)
        CU(cudaMemcpyAsync(res->nonzero_pivots, e->d_nzp.p, (size_t)nc * 4, cudaMemcpyDeviceToHost, e->stream));
    if (with_coef && res->coef)
        CU(cudaMemcpyAsync(res->coef, e->d_coef.p, (size_t)n_coef * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += (size_t)nc * 16 + (with_coef ? (size_t)n_coef * 8 : 0);
    return RR_OK;
}

// ---- EVAL_ONLY ------------------------------------------------------------------------------
int run_eval(rr_engine *e, const rr_batch *b, rr::BatchPlanner &bp, rr_result *res)
{
    const SweepCfg S = choose_cfg(e);
    rr::PlanLimits lim = limits_for(e, S, b->n_cand);
    rr::ColIds cols{e->d, e->d + 1};
    rr::SweepPlan P;
    std::vector<int32_t> cand_dot;
    std::string err = bp.plan_eval(lim, cols, false, P, cand_dot);
    if (!err.empty()) return e->fail(RR_ERR_INVALID, err);
    int rc = run_sweep(e, P, S, e->d_dots, false, nullptr, 0);
    if (rc) return rc;
    std::vector<double> dots(std::max(P.n_dots, 1));
    CU(cudaMemcpyAsync(dots.data(), e->d_dots.p, (size_t)P.n_dots * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += (size_t)P.n_dots * 8;
    for (int c = 0; c < b->n_cand; ++c) {
        const double v = dots[cand_dot[c]];
        res->ssr[c] = v;
        if (res->flags) res->flags[c] = std::isfinite(v) ? 0u : (uint32_t)RR_RES_NONFINITE;
        if (res->nonzero_pivots) res->nonzero_pivots[c] = 0;
        if (!std::isfinite(v)) e->stats.nonfinite++;
    }
    return RR_OK;
}

// ---- OLS_FIT, exact path ----------------------------------------------------------------------
int run_exact(rr_engine *e, const rr_batch *b, rr::BatchPlanner &bp, rr_result *res)
{
    const SweepCfg S = choose_cfg(e);
    rr::PlanLimits lim = limits_for(e, S, bp.n_terms_distinct());
    rr::ColIds cols{e->d, e->d + 1};
    rr::SweepPlan P;
    std::string err = bp.plan_materialise(lim, cols, P);
    if (!err.empty()) return e->fail(RR_ERR_INVALID, err);
    const int nc = b->n_cand;
    const int n_terms = b->cand_term_begin[nc];
    const int n_coef = n_terms + nc;
    int rc = ensure_result_buffers(e, nc, n_coef);
    if (rc) return rc;
    CU(e->d_V.ensure((size_t)std::max(bp.n_terms_distinct(), 1) * e->ld * 8));
    rc = run_sweep(e, P, S, e->d_dots, false, e->d_V.as<double>(), e->ld);
    if (rc) return rc;
    if ((rc = upload(e, e->d_ctb, b->cand_term_begin, (size_t)nc + 1))) return rc;
    if ((rc = upload(e, e->d_tid, bp.term_ids().data(), bp.term_ids().size()))) return rc;
    const int kmax = bp.max_k();
    const int n = (int)e->n;
    // candidates per launch bounded by the workspace budget
    const size_t budget = (size_t)env_double("RR_B200_EXACT_WS_BYTES", 4e9);
    const size_t per_cand = ((size_t)n * (kmax + 1)) * 8;
    int Q = (int)std::min<size_t>((size_t)nc, std::max<size_t>(32, budget / per_cand));
    Q = (int)round_up(Q, 32);
    CU(e->d_A.ensure((size_t)n * kmax * Q * 8));
    CU(e->d_rhs.ensure((size_t)n * Q * 8));
    CU(e->d_aux.ensure((size_t)Q * 5 * kmax * 8));
    CU(e->d_perm.ensure((size_t)Q * kmax * 4));
    const double *ycol = e->X.as<double>() + (size_t)e->d * e->ld;
    for (int lo = 0; lo < nc; lo += Q) {
        rr::ExactArgs a;
        a.V = e->d_V.as<double>();
        a.ldv = e->ld;
        a.y = ycol;
        a.n = n;
        a.term_ids = e->d_tid.as<int32_t>();
        a.cand_term_begin = e->d_ctb.as<int32_t>();
        a.cand_lo = lo;
        a.cand_hi = std::min(nc, lo + Q);
        a.Q = Q;
        a.kmax = kmax;
        a.A = e->d_A.as<double>();
        a.rhs = e->d_rhs.as<double>();
        a.aux = e->d_aux.as<double>();
        a.perm = e->d_perm.as<int32_t>();
        a.coef = e->d_coef.as<double>();
        a.nzp = e->d_nzp.as<int32_t>();
        a.flags = e->d_flags.as<uint32_t>();
        const int threads = 64;
        rr::rr_exact_qr<<<(a.cand_hi - lo + threads - 1) / threads, threads, 0, e->stream>>>(a);
        CU(cudaGetLastError());
        e->stats.kernel_launches++;
    }
    rr::ResidColsArgs r;
    r.V = e->d_V.as<double>();
    r.ldv = e->ld;
    r.y = ycol;
    r.n = n;
    r.term_ids = e->d_tid.as<int32_t>();
    r.cand_term_begin = e->d_ctb.as<int32_t>();
    r.n_cand = nc;
    r.coef = e->d_coef.as<double>();
    r.ssr = e->d_ssr.as<double>();
    r.flags = e->d_flags.as<uint32_t>();
    rr::rr_resid_cols<<<(nc * 32 + 255) / 256, 256, 0, e->stream>>>(r);
    CU(cudaGetLastError());
    e->stats.kernel_launches++;
    e->stats.exact += nc;
    return download_results(e, b, res, true);
}

// ---- OLS_FIT, Gram path -----------------------------------------------------------------------
int run_gram(rr_engine *e, const rr_batch *b, rr::BatchPlanner &bp, rr_result *res)
{
    const bool verbose = env_int("RR_B200_VERBOSE", 0) != 0;
    auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tlast = tnow();
    auto phase = [&](const char *name) {
        if (!verbose) return;
        cudaStreamSynchronize(e->stream);
        const double t = tnow();
        std::fprintf(stderr, "[rr_b200] %-28s %8.2f ms\n", name, t - tlast);
        tlast = t;
    };

    const SweepCfg S = choose_cfg(e);
    const int nc = b->n_cand;
    const int n_terms = b->cand_term_begin[nc];
    const int n_coef = n_terms + nc;
    rr::PlanLimits lim = limits_for(e, S, nc);
    rr::ColIds cols{e->d, e->d + 1};
    int rc = ensure_result_buffers(e, nc, n_coef);
    if (rc) return rc;

    // pass 1: Gram / A^T yc / column sums, shared across candidates.
    // A large neighbourhood on a large data set is planned in two halves, the second on a helper thread: its
    // planning overlaps the first half's planning, launch and sweep (planning is host work, ~1 us per
    // candidate; with the rows sharded over several GPUs it is otherwise a visible part of the step). The halves share nothing but the
    // base solution's terms, which each half evaluates and reduces once (a few dozen instructions).
    rr::SweepPlan P1;
    std::vector<int32_t> tab, tab_begin;
    std::string err;
    const bool pipelined = env_int("RR_B200_PIPELINE", 1) != 0 && nc >= 1024 && S.S == 4 && lim.target_chunks == 1;
    if (!pipelined) {
        err = bp.plan_gram(lim, cols, nullptr, false, P1, tab, tab_begin);
        if (!err.empty()) return e->fail(RR_ERR_INVALID, err);
        phase("plan gram");
        rc = run_sweep(e, P1, S, e->d_dots, false, nullptr, 0);
        if (rc) return rc;
        phase("sweep gram");
    } else {
        // the reduced dots of both halves live in one vector: size it before the first launch
        size_t tab_total = 0;
        for (int c = 0; c < nc; ++c) {
            const size_t m = (size_t)bp.k_of(c) - 1;
            tab_total += m * (m + 1) / 2 + 2 * m;
        }
        CU(e->d_dots.ensure((tab_total + 256) * 8));
        const int half = nc / 2;
        std::vector<int32_t> la(half), lb(nc - half);
        for (int c = 0; c < half; ++c) la[c] = c;
        for (int c = half; c < nc; ++c) lb[c - half] = c;
        // the second half is planned on a helper thread (plan_gram only reads the analysed batch): it overlaps
        // the first half's planning and launch, and - when an all-reduce hook blocks this thread until the
        // first sweep is done - the first sweep as well
        rr::SweepPlan P2;
        std::vector<int32_t> tab2, tab2_begin;
        std::string err2;
        auto plan_second = [&]() { err2 = bp.plan_gram(lim, cols, &lb, false, P2, tab2, tab2_begin); };
        std::thread planner2;
        struct Joiner {  // whatever path leaves this scope, the helper is joined first (it writes to the locals above)
            std::thread &t;
            ~Joiner()
            {
                if (t.joinable()) t.join();
            }
        } joiner{planner2};
        bool threaded = true;
        try {
            planner2 = std::thread(plan_second);
        } catch (...) {
            threaded = false;  // no thread to be had: plan the second half here, after the first launch
        }
        err = bp.plan_gram(lim, cols, &la, false, P1, tab, tab_begin);
        if (!err.empty()) {
            if (threaded) planner2.join();
            return e->fail(RR_ERR_INVALID, err);
        }
        rc = run_sweep(e, P1, S, e->d_dots, false, nullptr, 0, 0, 0, false);
        if (threaded) planner2.join();
        else if (!rc) plan_second();
        if (rc) { cudaStreamSynchronize(e->stream); return rc; }
        if (!err2.empty()) { cudaStreamSynchronize(e->stream); return e->fail(RR_ERR_INVALID, err2); }
        const size_t off2 = (size_t)round_up(std::max(P1.n_dots, 1), 32);
        if (off2 + (size_t)P2.n_dots > tab_total + 256) { cudaStreamSynchronize(e->stream); return e->fail(RR_ERR_INVALID, "internal: dot vector too small"); }
        rc = run_sweep(e, P2, S, e->d_dots, false, nullptr, 0, 1, off2, false);
        if (rc) { cudaStreamSynchronize(e->stream); return rc; }
        const int32_t t0 = (int32_t)tab.size();
        for (int32_t id : tab2) tab.push_back(id + (int32_t)off2);
        for (size_t i = 1; i < tab2_begin.size(); ++i) tab_begin.push_back(tab2_begin[i] + t0);
        CU(cudaStreamSynchronize(e->stream));
        finish_sweep(e, 0);
        finish_sweep(e, 1);
        phase("plan + sweep gram (two halves)");
    }

    // per-candidate solve
    std::vector<int64_t> wsoff(nc + 1, 0);
    for (int c = 0; c < nc; ++c) {
        const int64_t kk = bp.k_of(c);
        wsoff[c + 1] = wsoff[c] + 2 * kk * kk + 10 * kk;
    }
    CU(e->d_ws.ensure((size_t)wsoff[nc] * 8));
    if ((rc = upload(e, e->d_wsoff, wsoff.data(), wsoff.size()))) return rc;
    if ((rc = upload(e, e->d_tab, tab.data(), tab.size()))) return rc;
    // tab_begin is uploaded behind the table in the same buffer family
    DevBuf &d_tabb = e->d_rtab;  // reused later for the residual tables; order of use is sequential
    if ((rc = upload(e, d_tabb, tab_begin.data(), tab_begin.size()))) return rc;
    if ((rc = upload(e, e->d_ctb, b->cand_term_begin, (size_t)nc + 1))) return rc;
    rr::GramArgs g;
    g.dots = e->d_dots.as<double>();
    g.cand_dot = e->d_tab.as<int32_t>();
    g.cand_dot_begin = d_tabb.as<int32_t>();
    g.list = nullptr;
    g.cand_term_begin = e->d_ctb.as<int32_t>();
    g.n_list = nc;
    g.ws = e->d_ws.as<double>();
    g.ws_begin = e->d_wsoff.as<int64_t>();
    g.k = e->sc;
    g.coef = e->d_coef.as<double>();
    g.coef_snapped = e->d_cs.as<double>();
    g.nzp = e->d_nzp.as<int32_t>();
    g.ssr = e->d_ssr.as<double>();
    g.flags = e->d_flags.as<uint32_t>();
    g.status = e->d_status.as<uint32_t>();
    rr::rr_gram_solve<<<(nc + 63) / 64, 64, 0, e->stream>>>(g);
    CU(cudaGetLastError());
    e->stats.kernel_launches++;

    phase("gram solve launch");
    std::vector<uint32_t> status(nc);
    CU(cudaMemcpyAsync(status.data(), e->d_status.p, (size_t)nc * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += (size_t)nc * 4;
    std::vector<int32_t> refine, escalate;
    for (int c = 0; c < nc; ++c) {
        if (status[c] & rr::ST_NEED_ESCALATE) escalate.push_back(c);
        else if (status[c] & rr::ST_NEED_REFINE) refine.push_back(c);
    }

    // Helper: upload the subset tables (Gram table offsets of the listed candidates, workspace offsets)
    auto make_sub = [&](const std::vector<int32_t> &list, DevBuf &d_list, DevBuf &d_begin, DevBuf &d_wso,
                        bool dd_ws) -> int {
        std::vector<int32_t> begin(list.size() + 1, 0);
        std::vector<int64_t> wso(list.size() + 1, 0);
        for (size_t i = 0; i < list.size(); ++i) {
            begin[i] = tab_begin[list[i]];
            const int64_t kk = bp.k_of(list[i]);
            wso[i + 1] = wso[i] + 2 * kk * kk + 10 * kk;
        }
        (void)dd_ws;
        int r;
        if ((r = upload(e, d_list, list.data(), list.size()))) return r;
        if ((r = upload(e, d_begin, begin.data(), begin.size()))) return r;
        if ((r = upload(e, d_wso, wso.data(), wso.size()))) return r;
        CU(e->d_ws.ensure((size_t)wso[list.size()] * 8));
        return RR_OK;
    };

    // pass 2 (rare): double-double Gram for numerically singular candidates
    if (!escalate.empty()) {
        rr::SweepPlan Pd;
        std::vector<int32_t> dtab, dtab_begin;
        err = bp.plan_gram(lim, cols, &escalate, true, Pd, dtab, dtab_begin);
        if (!err.empty()) return e->fail(RR_ERR_INVALID, err);
        rc = run_sweep(e, Pd, S, e->d_rdots, true, nullptr, 0);
        if (rc) return rc;
        DevBuf &d_dt = e->d_t0, &d_dtb = e->d_t1, &d_l = e->d_t2, &d_wo = e->d_t3;
        auto cleanup = [&]() {};
        std::vector<int64_t> wso(escalate.size() + 1, 0);
        for (size_t i = 0; i < escalate.size(); ++i) {
            const int64_t kk = bp.k_of(escalate[i]);
            wso[i + 1] = wso[i] + 2 * kk * kk + 10 * kk;
        }
        if ((rc = upload(e, d_dt, dtab.data(), dtab.size())) || (rc = upload(e, d_dtb, dtab_begin.data(), dtab_begin.size())) ||
            (rc = upload(e, d_l, escalate.data(), escalate.size())) || (rc = upload(e, d_wo, wso.data(), wso.size()))) {
            cleanup();
            return rc;
        }
        cudaError_t ce = e->d_ws.ensure((size_t)wso[escalate.size()] * 8);
        if (ce != cudaSuccess) { cleanup(); return e->cuda_fail(ce, "ws"); }
        rr::GramArgs gd = g;
        gd.dots = e->d_rdots.as<double>();
        gd.cand_dot = d_dt.as<int32_t>();
        gd.cand_dot_begin = d_dtb.as<int32_t>();
        gd.list = d_l.as<int32_t>();
        gd.n_list = (int)escalate.size();
        gd.ws = e->d_ws.as<double>();
        gd.ws_begin = d_wo.as<int64_t>();
        rr::rr_gram_solve_dd<<<((int)escalate.size() + 63) / 64, 64, 0, e->stream>>>(gd);
        ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        cleanup();
        if (ce != cudaSuccess) return e->cuda_fail(ce, "rr_gram_solve_dd");
        e->stats.kernel_launches++;
        e->stats.dd += escalate.size();
    }

    phase("solve + dd pass");
    // pass 3: explicit residuals for the refine set (and the SSR of the escalated set)
    std::vector<int32_t> pending(refine);
    bool first_round = true;
    for (int round = 0; round < 3 && (!pending.empty() || (first_round && !escalate.empty())); ++round) {
        std::vector<int32_t> list(pending);
        if (first_round) list.insert(list.end(), escalate.begin(), escalate.end());
        std::vector<double> cs(n_coef);
        CU(cudaMemcpyAsync(cs.data(), e->d_cs.p, (size_t)n_coef * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        e->stats.d2h_bytes += (size_t)n_coef * 8;
        // candidates whose coefficients are not finite cannot be refined
        std::vector<int32_t> ok;
        for (int32_t c : list) {
            bool fin = true;
            for (int i = b->cand_term_begin[c] + c; i < b->cand_term_begin[c + 1] + c + 1; ++i) fin = fin && std::isfinite(cs[i]);
            if (fin) ok.push_back(c);
        }
        const size_t n_pending_ok = std::count_if(ok.begin(), ok.end(), [&](int32_t c) {
            return std::find(pending.begin(), pending.end(), c) != pending.end();
        });
        (void)n_pending_ok;
        rr::SweepPlan Pr;
        std::vector<int32_t> rtab, rtab_begin;
        err = bp.plan_residual(lim, cols, ok, cs.data(), Pr, rtab, rtab_begin);
        if (!err.empty()) return e->fail(RR_ERR_INVALID, err);
        rc = run_sweep(e, Pr, S, e->d_rdots, false, nullptr, 0);
        if (rc) return rc;
        // split `ok` back into refine (pending) and escalated members, keeping residual-table offsets
        std::vector<int32_t> l_ref, l_esc, rb_ref, rb_esc;
        {
            std::vector<char> is_pending(nc, 0);
            for (int32_t c : pending) is_pending[c] = 1;
            for (size_t i = 0; i < ok.size(); ++i) {
                if (is_pending[ok[i]]) { l_ref.push_back(ok[i]); rb_ref.push_back(rtab_begin[i]); }
                else { l_esc.push_back(ok[i]); rb_esc.push_back(rtab_begin[i]); }
            }
        }
        DevBuf &d_rt = e->d_t0, &d_l = e->d_t1, &d_b = e->d_t2, &d_wo = e->d_t3, &d_rb = e->d_t4;
        auto cleanup = [&]() {};
        if ((rc = upload(e, d_rt, rtab.data(), rtab.size()))) { cleanup(); return rc; }
        if (!l_esc.empty()) {
            if ((rc = upload(e, d_l, l_esc.data(), l_esc.size())) || (rc = upload(e, d_rb, rb_esc.data(), rb_esc.size()))) { cleanup(); return rc; }
            k_take_ssr<<<((int)l_esc.size() + 127) / 128, 128, 0, e->stream>>>(d_l.as<int32_t>(), (int)l_esc.size(), d_rb.as<int32_t>(),
                                                                             d_rt.as<int32_t>(), e->d_rdots.as<double>(),
                                                                             e->d_ssr.as<double>(), e->d_flags.as<uint32_t>());
            e->stats.kernel_launches++;
            cudaError_t ce = cudaStreamSynchronize(e->stream);
            if (ce != cudaSuccess) { cleanup(); return e->cuda_fail(ce, "k_take_ssr"); }
        }
        std::vector<int32_t> next;
        if (!l_ref.empty()) {
            if ((rc = make_sub(l_ref, d_l, d_b, d_wo, false)) || (rc = upload(e, d_rb, rb_ref.data(), rb_ref.size()))) { cleanup(); return rc; }
            rr::RefineArgs ra;
            ra.g = g;
            ra.g.list = d_l.as<int32_t>();
            ra.g.n_list = (int)l_ref.size();
            ra.g.cand_dot_begin = d_b.as<int32_t>();
            ra.g.ws = e->d_ws.as<double>();
            ra.g.ws_begin = d_wo.as<int64_t>();
            ra.rdots = e->d_rdots.as<double>();
            ra.rcand_dot = d_rt.as<int32_t>();
            ra.rcand_dot_begin = d_rb.as<int32_t>();
            ra.delta_rel = e->d_delta.as<double>();
            rr::rr_refine_update<<<((int)l_ref.size() + 63) / 64, 64, 0, e->stream>>>(ra);
            e->stats.kernel_launches++;
            std::vector<double> delta(nc);
            cudaError_t ce = cudaGetLastError();
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(delta.data(), e->d_delta.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, e->stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
            if (ce != cudaSuccess) { cleanup(); return e->cuda_fail(ce, "rr_refine_update"); }
            const double again = env_double("RR_B200_REFINE_AGAIN", 1e-6);
            for (int32_t c : l_ref)
                if (delta[c] > again) next.push_back(c);
            if (round == 0) e->stats.refined += l_ref.size();
        }
        cleanup();
        pending.swap(next);
        first_round = false;
    }
    phase("residual pass");
    rc = download_results(e, b, res, true);
    phase("download");
    return rc;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int rr_abi_version(void) { return RR_ABI_VERSION; }

const char *rr_last_error(const rr_engine *e) { return e ? e->error.c_str() : g_thread_error.c_str(); }

int rr_engine_create(const double *X, const double *y, int64_t n, int32_t d, int32_t device, uint32_t flags,
                     rr_engine **out)
{
    return create_common(X, y, n, d, device, flags, false, out);
}

int rr_engine_create_rowmajor(const double *X, const double *y, int64_t n, int32_t d, int32_t device, uint32_t flags,
                              rr_engine **out)
{
    return create_common(X, y, n, d, device, flags, true, out);
}

void rr_engine_destroy(rr_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    e->free_all();
    delete e;
}

int rr_engine_set_allreduce(rr_engine *e, rr_allreduce_fn fn, void *user, int32_t rank, int32_t world)
{
    if (!e) return RR_ERR_INVALID;
    if (world < 1 || rank < 0 || rank >= world) return e->fail(RR_ERR_INVALID, "bad rank/world");
    CU(cudaSetDevice(e->device));
    e->allreduce = fn;
    e->allreduce_user = user;
    e->rank = rank;
    e->world = fn ? world : 1;
    return compute_y_stats(e);
}

int rr_engine_get_info(const rr_engine *e, rr_engine_info *info)
{
    if (!e || !info) return RR_ERR_INVALID;
    info->n = e->n;
    info->n_total = e->n_total;
    info->d = e->d;
    info->device = e->device;
    info->y_mean = e->y_mean;
    info->sst = e->sst;
    info->sm_count = e->sm_count;
    info->exact_max_n = e->exact_max_n;
    return RR_OK;
}

int rr_get_stats(const rr_engine *e, rr_stats *stats)
{
    if (!e || !stats) return RR_ERR_INVALID;
    *stats = e->stats;
    return RR_OK;
}

int rr_score_batch(rr_engine *e, const rr_batch *b, rr_result *res)
{
    if (!e) { g_thread_error = "null engine"; return RR_ERR_INVALID; }
    if (!b || !res || !res->ssr) return e->fail(RR_ERR_INVALID, "null batch/result (result.ssr is required)");
    if (b->mode != RR_MODE_EVAL_ONLY && b->mode != RR_MODE_OLS_FIT) return e->fail(RR_ERR_INVALID, "bad mode");
    if (b->n_cand == 0) return RR_OK;
    CU(cudaSetDevice(e->device));
    rr::BatchPlanner bp(b, e->d);
    std::string err = bp.analyse((e->flags & RR_FLAG_NO_CSE) != 0);
    if (!err.empty()) return e->fail(RR_ERR_INVALID, "malformed batch: " + err);
    e->stats.h2d_bytes = e->stats.d2h_bytes = 0;
    e->stats.w_shared = 0.0;
    e->sweep_ms_accum = 0.f;
    const uint64_t dd0 = e->stats.distinct_dots;
    (void)dd0;
    CU(cudaEventRecord(e->ev[0], e->stream));
    int rc;
    if (b->mode == RR_MODE_EVAL_ONLY) {
        rc = run_eval(e, b, bp, res);
    } else {
        const bool sharded = e->allreduce && e->world > 1;
        bool exact = !sharded && e->n_total <= e->exact_max_n;
        if (e->flags & RR_FLAG_FORCE_GRAM) exact = false;
        if ((e->flags & RR_FLAG_FORCE_EXACT) && !sharded) exact = true;
        rc = exact ? run_exact(e, b, bp, res) : run_gram(e, b, bp, res);
    }
    if (rc) return rc;
    CU(cudaEventRecord(e->ev[1], e->stream));
    CU(cudaEventSynchronize(e->ev[1]));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
    e->stats.last_batch_ms = ms;
    e->stats.last_sweep_ms = e->sweep_ms_accum;
    e->stats.batches++;
    e->stats.candidates += b->n_cand;
    e->stats.term_instances += bp.n_term_instances();
    e->stats.distinct_terms += bp.n_terms_distinct();
    e->stats.w_contract = bp.w_contract();
    if (b->mode == RR_MODE_OLS_FIT) {
        for (int c = 0; c < b->n_cand; ++c) {
            const int64_t k = bp.k_of(c);
            e->stats.dot_instances += k * (k + 1) / 2 + k;
        }
        if (res->flags)
            for (int c = 0; c < b->n_cand; ++c)
                if (res->flags[c] & RR_RES_NONFINITE) e->stats.nonfinite++;
    } else {
        e->stats.dot_instances += b->n_cand;
    }
    return RR_OK;
}

int rr_classifier_metrics(rr_engine *e, const rr_batch *b, double *accuracy, double *log_loss, double *abs_loss)
{
    if (!e) { g_thread_error = "null engine"; return RR_ERR_INVALID; }
    if (!b || b->mode != RR_MODE_EVAL_ONLY) return e->fail(RR_ERR_INVALID, "classifier metrics need an EVAL_ONLY batch");
    if (b->n_cand == 0) return RR_OK;
    CU(cudaSetDevice(e->device));
    rr::BatchPlanner bp(b, e->d);
    std::string err = bp.analyse((e->flags & RR_FLAG_NO_CSE) != 0);
    if (!err.empty()) return e->fail(RR_ERR_INVALID, "malformed batch: " + err);
    const SweepCfg S = choose_cfg(e);
    rr::PlanLimits lim = limits_for(e, S, b->n_cand);
    rr::ColIds cols{e->d, e->d + 1};
    rr::SweepPlan P;
    std::vector<int32_t> cand_dot;
    err = bp.plan_eval(lim, cols, true, P, cand_dot);
    if (!err.empty()) return e->fail(RR_ERR_INVALID, err);
    int rc = run_sweep(e, P, S, e->d_dots, false, nullptr, 0);
    if (rc) return rc;
    std::vector<double> dots(std::max(P.n_dots, 1));
    CU(cudaMemcpyAsync(dots.data(), e->d_dots.p, (size_t)P.n_dots * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    const double n = (double)e->n_total;
    for (int c = 0; c < b->n_cand; ++c) {
        const int id = cand_dot[c];
        if (accuracy) accuracy[c] = dots[id] / n;
        if (log_loss) log_loss[c] = dots[id + 1] / n;
        if (abs_loss) abs_loss[c] = dots[id + 2] / n;
    }
    return RR_OK;
}

static int predict_common(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts, int32_t n_consts,
                          const double *X, int64_t n, int32_t d, bool rowmajor, double *out)
{
    if (!e) { g_thread_error = "null engine"; return RR_ERR_INVALID; }
    if (!code || code_len <= 0 || !X || !out || n <= 0 || d <= 0) return e->fail(RR_ERR_INVALID, "rr_predict: bad arguments");
    // a temporary engine over the caller's matrix (y is a dummy column) shares the kernels
    std::vector<double> ydummy((size_t)n, 0.0);
    rr_engine *t = nullptr;
    int rc = create_common(X, ydummy.data(), n, d, e->device, 0, rowmajor, &t);
    if (rc) return e->fail(rc, g_thread_error);
    int32_t ctb[2] = {0, 1}, tcb[2] = {0, code_len};
    rr_batch b;
    std::memset(&b, 0, sizeof(b));
    b.mode = RR_MODE_EVAL_ONLY;
    b.n_cand = 1;
    b.cand_term_begin = ctb;
    b.term_code_begin = tcb;
    b.code = code;
    b.consts = consts;
    b.n_consts = n_consts;
    rr::BatchPlanner bp(&b, d);
    std::string err = bp.analyse(true);
    if (err.empty()) {
        const SweepCfg S = choose_cfg(t);
        rr::PlanLimits lim = limits_for(t, S, 1);
        lim.target_chunks = 1;
        rr::SweepPlan P;
        err = bp.plan_materialise(lim, rr::ColIds{d, d + 1}, P);
        if (err.empty()) {
            cudaError_t ce = t->d_V.ensure((size_t)t->ld * 8);
            if (ce != cudaSuccess) err = cudaGetErrorString(ce);
            if (err.empty()) {
                rc = run_sweep(t, P, S, t->d_dots, false, t->d_V.as<double>(), t->ld);
                if (rc) err = t->error;
            }
            if (err.empty()) {
                ce = cudaMemcpyAsync(out, t->d_V.p, (size_t)n * 8, cudaMemcpyDeviceToHost, t->stream);
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(t->stream);
                if (ce != cudaSuccess) err = cudaGetErrorString(ce);
            }
        }
    }
    e->stats.kernel_launches += t->stats.kernel_launches;
    e->stats.sweep_launches += t->stats.sweep_launches;
    rr_engine_destroy(t);
    if (!err.empty()) return e->fail(RR_ERR_INVALID, "rr_predict: " + err);
    return RR_OK;
}

int rr_predict(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts, int32_t n_consts,
               const double *X, int64_t n, int32_t d, double *out)
{
    return predict_common(e, code, code_len, consts, n_consts, X, n, d, false, out);
}

int rr_predict_rowmajor(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts, int32_t n_consts,
                        const double *X, int64_t n, int32_t d, double *out)
{
    return predict_common(e, code, code_len, consts, n_consts, X, n, d, true, out);
}

int rr_measure_fp64_peak(rr_engine *e, double *dfma_per_second)
{
    if (!e || !dfma_per_second) return RR_ERR_INVALID;
    CU(cudaSetDevice(e->device));
    const int blocks = e->sm_count * 8, threads = 256, iters = 4096;
    CU(e->d_misc.ensure((size_t)blocks * threads * 8));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e->ev[0], e->stream));
        k_fp64_peak<<<blocks, threads, 0, e->stream>>>(e->d_misc.as<double>(), iters, 1.0000001, 1e-9);
        CU(cudaGetLastError());
        CU(cudaEventRecord(e->ev[1], e->stream));
        CU(cudaEventSynchronize(e->ev[1]));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
        const double rate = (double)blocks * threads * iters * 16.0 * 8.0 / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    e->stats.kernel_launches += 5;
    *dfma_per_second = best;
    return RR_OK;
}

}  // extern "C"
