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
#include <dlfcn.h>
#include <nccl.h>  // types and enumerators only: the library is loaded at run time (see NcclApi)

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rr_b200.h"
#include "rr_exact.cuh"
#include "rr_isa.h"
#include "rr_plan.h"
#include "rr_solve.cuh"
#include "rr_sweep.cuh"
#include "rr_sweep_g8.cuh"
#include "rr_sweep_r8.cuh"

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

// NCCL is bound at run time: a process that never shards rows needs no NCCL at all, and under torchrun the copy
// torch has already loaded (same SONAME) is the one that is picked up - two NCCL copies in one process do not mix.
struct NcclApi {
    void *h = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;

    template <typename F> bool sym(F &f, const char *name)
    {
        f = reinterpret_cast<F>(dlsym(h, name));
        if (!f) err = std::string("NCCL symbol missing: ") + name;
        return f != nullptr;
    }
    bool load()
    {
        if (h) return true;
        const char *env = std::getenv("RR_B200_NCCL_LIB");
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy this process already has (torch's)
        if (!h && env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            const char *de = dlerror();
            err = std::string("cannot load libnccl.so.2: ") + (de ? de : "?");
            return false;
        }
        const bool ok = sym(GetUniqueId, "ncclGetUniqueId") && sym(CommInitRank, "ncclCommInitRank") &&
                        sym(CommInitAll, "ncclCommInitAll") && sym(CommDestroy, "ncclCommDestroy") &&
                        sym(AllReduce, "ncclAllReduce") && sym(Reduce, "ncclReduce") && sym(AllGather, "ncclAllGather") &&
                        sym(GroupStart, "ncclGroupStart") && sym(GroupEnd, "ncclGroupEnd") &&
                        sym(GetErrorString, "ncclGetErrorString");
        if (!ok) h = nullptr;
        return ok;
    }
};
NcclApi &nccl()
{
    static NcclApi api;
    return api;
}
std::mutex g_nccl_mu;

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

// Rows [0, rows) of a row-major chunk Xr (rows x d) -> feature-major engine matrix: row r goes to position
// pos[r] (pos == nullptr: dst0 + r; pos[r] < 0: not selected). 32 x 32 tiles through shared memory: the reads are
// coalesced; with an index the writes are scattered 8-byte stores (the shuffle of rils_rols_cpp.cpp:777-795).
__global__ void k_ingest_rows(const double *__restrict__ Xr, const double *__restrict__ yr, int64_t rows, int32_t d,
                              const int32_t *__restrict__ pos, int64_t dst0, int64_t dst_lo, int64_t dst_hi,
                              double *__restrict__ Xc, int64_t ld, double *__restrict__ yc)
{
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t i = i0 + r;
        const int j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < rows && j < d) ? Xr[i * d + j] : 0.0;
    }
    __syncthreads();
    const int64_t i = i0 + threadIdx.x;
    int64_t dst = -1;
    if (i < rows) {
        dst = pos ? (int64_t)pos[i] : dst0 + i;
        if (dst < dst_lo || dst >= dst_hi) dst = -1;
    }
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r;
        if (j < d && dst >= 0) Xc[(int64_t)j * ld + (dst - dst_lo)] = tile[threadIdx.x][r];
    }
    if (blockIdx.y == 0 && threadIdx.y == 0 && dst >= 0 && yr) yc[dst - dst_lo] = yr[i];
}

// relevant_features (rils_rols_cpp.cpp:753-770), per feature column: pass 0 -> sum x; pass 1 -> sum (x - mean)^2 and
// sum (x - y)^2. Deterministic: fixed grid, block partials summed on the host in order.
__global__ void k_feature_stats(const double *__restrict__ X, int64_t ld, const double *__restrict__ y, int64_t n,
                                const double *__restrict__ mean, int pass, double *__restrict__ partial)
{
    __shared__ double s0[256], s1[256];
    const int j = blockIdx.y;
    const double *x = X + (size_t)j * ld;
    const double m = pass ? mean[j] : 0.0;
    double a = 0.0, b = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        if (pass == 0) {
            a += v;
        } else {
            a = fma(v - m, v - m, a);
            b = fma(v - y[i], v - y[i], b);
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
        partial[2 * ((size_t)j * gridDim.x + blockIdx.x)] = s0[0];
        partial[2 * ((size_t)j * gridDim.x + blockIdx.x) + 1] = s1[0];
    }
}

// predict_proba epilogue: (1 - p, p) with p = 1 / (1 + exp(-2 (yhat - 0.5))), rils_rols_cpp.cpp:69
__global__ void k_proba(const double *__restrict__ yhat, int64_t n, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double p = 1.0 / (1.0 + exp(-2.0 * (yhat[i] - 0.5)));
    out[2 * i] = 1.0 - p;
    out[2 * i + 1] = p;
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
    // sample sharding, NCCL inside the engine. Single process: this object is shard 0 (the leader) and owns
    // `peers`, one engine per further device, each with a communicator of one ncclCommInitAll clique.
    // One process per GPU: `comm` is this rank's communicator (rr_engine_comm_init).
    std::vector<rr_engine *> peers;
    ncclComm_t comm = nullptr;
    bool comm_ranks = false;  // comm spans processes (world / rank above); false with peers: comm spans this object's shards
    int64_t n_object = 0;     // rows held by this object over all its shards
    DevBuf d_px, d_pr, d_pout;  // predict(): feature-major chunk, row-major chunk, output chunk
    // grow-only work buffers
    DevBuf d_ins2, d_chunks2, d_cols2, d_acc2;  // second sweep in flight (run_gram plans the batch in two halves)
    DevBuf d_ins, d_chunks, d_cols, d_acc, d_dots, d_rdots, d_tab, d_rtab, d_ws, d_wsoff, d_list, d_coef, d_cs,
        d_nzp, d_ssr, d_flags, d_status, d_delta, d_V, d_A, d_rhs, d_aux, d_perm, d_ctb, d_cmask, d_tid, d_misc, d_gather,
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
    int nccl_fail(ncclResult_t r, const char *what)
    {
        error = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
        return RR_ERR_COLLECTIVE;
    }
    bool grouped() const { return !peers.empty(); }  // rows spread over devices of this object
    bool sharded() const { return peers.empty() && world > 1 && (allreduce || comm); }  // ... over other processes
    bool rows_split() const { return grouped() || sharded(); }
    void free_all()
    {
        for (rr_engine *p : peers) {
            cudaSetDevice(p->device);
            p->free_all();
            delete p;
        }
        peers.clear();
        cudaSetDevice(device);
        if (comm && nccl().CommDestroy) nccl().CommDestroy(comm);
        comm = nullptr;
        for (DevBuf *b : {&d_px, &d_pr, &d_pout}) b->release();
        for (DevBuf *b : {&X, &d_ins2, &d_chunks2, &d_cols2, &d_acc2, &d_ins, &d_chunks, &d_cols, &d_acc, &d_dots, &d_rdots, &d_tab, &d_rtab, &d_ws, &d_wsoff,
                          &d_list, &d_coef, &d_cs, &d_nzp, &d_ssr, &d_flags, &d_status, &d_delta, &d_V, &d_A, &d_rhs,
                          &d_aux, &d_perm, &d_ctb, &d_cmask, &d_tid, &d_misc, &d_gather, &d_t0, &d_t1, &d_t2, &d_t3, &d_t4})
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
#define NC(call)                                               \
    do {                                                       \
        ncclResult_t _r = (call);                              \
        if (_r != ncclSuccess) return e->nccl_fail(_r, #call); \
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

// which of an engine's reduced-dot buffers a sweep writes (the same member on every shard)
using DotsMember = DevBuf rr_engine::*;
// the matrix a launch sweeps: the engine's resident one, or a chunk of predict()'s input
struct XView {
    const double *X;
    int64_t ld, n;
};
std::vector<rr_engine *> shards_of(rr_engine *e)
{
    std::vector<rr_engine *> v{e};
    v.insert(v.end(), e->peers.begin(), e->peers.end());
    return v;
}

// One shard, one plan: upload the plan, zero the accumulators, launch the interpreter on s->stream, reduce the block
// rows into (s->*dots) + dots_off. No synchronisation. `e` (the leader) only receives the error text.
// tile columns a G8 plan may use (rr_sweep_g8.cuh: padded columns behind the staging rows, two blocks per SM)
int g8_tile_cols()
{
    const size_t per_block = kSmemPerSM / 2 - kSmemReserved - rr::kG8StaticBytes;
    return (int)((per_block - rr::kG8StageBytes) / rr::kG8ColBytes);
}

// tile columns an R8 plan may use (rr_sweep_r8.cuh: 256-sample tiles, two blocks per SM)
int r8_tile_cols()
{
    const size_t per_block = kSmemPerSM / rr::kR8BlocksPerSM - kSmemReserved - rr::kR8StaticBytes;
    return (int)((per_block - rr::kR8StageBytes) / rr::kR8ColBytes);
}

int launch_shard(rr_engine *e, rr_engine *s, const rr::SweepPlan &P, const std::vector<RRIns> &ins_padded, const SweepCfg &cfg,
                 const XView &view, DotsMember dots, bool dd, double *stg, int64_t ld_stg, int set, size_t dots_off, bool g8 = false,
                 bool mark_begin = true)
{
    DevBuf &d_ins = set ? s->d_ins2 : s->d_ins, &d_chunks = set ? s->d_chunks2 : s->d_chunks;
    DevBuf &d_cols = set ? s->d_cols2 : s->d_cols, &d_acc = set ? s->d_acc2 : s->d_acc;
    cudaEvent_t ev0 = s->ev[set ? 4 : 2], ev1 = s->ev[set ? 5 : 3];
    CU(cudaSetDevice(s->device));
    const bool r8 = P.r8;
    const int T = r8 ? rr::kR8Tile : cfg.T();
    const int NW = cfg.TH / 32;
    const int n_tiles = (int)((view.n + T - 1) / T);
    const size_t tile_bytes = (size_t)std::max(P.max_tile_cols, 1) * T * 8;
    if (r8 ? P.max_tile_cols > r8_tile_cols() : (g8 ? P.max_tile_cols > g8_tile_cols() : tile_bytes > cfg.dyn_smem_budget()))
        return e->fail(RR_ERR_INVALID, "internal: plan exceeds the shared-memory tile");
    // two tile buffers when they fit: the next tile's columns are fetched while this one is interpreted (what an
    // HBM-bound sweep - one small program over many rows - needs; the big neighbourhoods fill the tile and do not care)
    // up to four tile buffers when they fit (an HBM-bound sweep wants several tiles in flight per block)
    const int tile_dbuf = (g8 || r8) ? 0 : (int)std::min<size_t>((size_t)std::max(0, env_int("RR_B200_TILE_BUFS", 4) - 1), cfg.dyn_smem_budget() / tile_bytes - 1);
    const size_t smem = r8 ? rr::r8_dyn_smem(std::max(P.max_tile_cols, 1)) : g8 ? rr::g8_dyn_smem(std::max(P.max_tile_cols, 1)) : tile_bytes * (size_t)(tile_dbuf + 1) + rr::sweep_ring_smem(NW, cfg.slack);
    bool special = dd;
    for (const RRIns &x : P.ins)
        if (RR_OP(x.w0) == RI_CLSMET) { special = true; break; }
    SweepKernel kern = r8 ? (SweepKernel)rr::rr_sweep_r8_kernel : g8 ? (SweepKernel)rr::rr_sweep_g8_kernel : sweep_kernel_for(cfg, special);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            r8 ? (int)rr::r8_dyn_smem(r8_tile_cols())
                               : g8 ? (int)rr::g8_dyn_smem(g8_tile_cols()) : (int)(kSmemPerBlockMax - rr::sweep_static_smem())));
    const int n_chunks = (int)P.chunks.size();
    int gx;
    if (g8 || r8) {
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, cfg.TH, smem);
        gx = std::min(std::max(1, s->sm_count * std::max(1, occ) / std::max(1, n_chunks)), std::max(1, n_tiles));
    } else {
        gx = sweep_gx(s, cfg, special, smem, n_chunks, std::max(1, n_tiles));
    }
    const int64_t stride = round_up(std::max(P.n_dots, 1), 32) + 32;
    // keep the accumulator rows within a sane budget
    const size_t row_budget = (size_t)env_double("RR_B200_ACC_BYTES", 6e9);
    const int rpb = dd ? NW : 1;  // accumulator rows per block: double-double plans keep one per warp
    while (gx > 1 && (size_t)gx * rpb * stride * 8 > row_budget) gx = (gx + 1) / 2;
    const int rows = gx * rpb;
    {
        int rc;
        if ((rc = upload(s, d_ins, ins_padded.data(), ins_padded.size())) || (rc = upload(s, d_chunks, P.chunks.data(), P.chunks.size())) ||
            (rc = upload(s, d_cols, P.cols.data(), P.cols.size()))) {
            if (s != e) e->error = s->error;
            return rc;
        }
        if (s != e) e->stats.h2d_bytes += (ins_padded.size() * sizeof(RRIns) + P.chunks.size() * sizeof(RRChunk) + P.cols.size() * 4);
    }
    DevBuf &dbuf = s->*dots;
    if (P.n_dots > 0) {
        CU(d_acc.ensure((size_t)rows * stride * 8));
        CU(cudaMemsetAsync(d_acc.p, 0, (size_t)rows * stride * 8, s->stream));
        CU(dbuf.ensure((dots_off + (size_t)stride) * 8));
    } else {
        CU(d_acc.ensure(64));
    }
    rr::SweepArgs a;
    a.X = view.X;
    a.ld = view.ld;
    a.n = view.n;
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
    a.tile_dbuf = tile_dbuf;
    a.tile_buf_doubles = (int64_t)(tile_bytes / 8);
    if (mark_begin) CU(cudaEventRecord(ev0, s->stream));
    kern<<<dim3(gx, n_chunks), r8 ? rr::kR8Threads : cfg.TH, smem, s->stream>>>(a);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ev1, s->stream));
    e->stats.sweep_launches++;
    e->stats.kernel_launches++;
    if (P.n_dots > 0) {
        if (!dd) {
            rr::rr_reduce_rows<<<(P.n_dots + 255) / 256, 256, 0, s->stream>>>(d_acc.as<double>(), stride, rows, P.n_dots,
                                                                            dbuf.as<double>() + dots_off);
        } else {
            rr::rr_reduce_rows_dd<<<(P.n_dots / 2 + 255) / 256, 256, 0, s->stream>>>(d_acc.as<double>(), stride, rows, P.n_dots / 2,
                                                                                   dbuf.as<double>() + dots_off);
        }
        CU(cudaGetLastError());
        e->stats.kernel_launches++;
    }
    return RR_OK;
}

// Sum `count` doubles at (shard->*dots) + off over all shards / ranks, stream-ordered, no host synchronisation:
// single-process engines reduce to the leader (only it solves), communicator and hook engines all-reduce.
// dd: the buffer holds (hi, lo) pairs that must not be summed in fp64: every rank's pairs are gathered and
// added in double-double by the caller's rr_reduce_rows_dd.
int reduce_over_ranks(rr_engine *e, DotsMember dots, size_t off, size_t count, bool dd)
{
    if (!count) return RR_OK;
    NcclApi &N = nccl();
    if (e->grouped()) {
        std::vector<rr_engine *> sh = shards_of(e);
        const size_t G = sh.size();
        if (dd)
            for (rr_engine *s : sh) {
                CU(cudaSetDevice(s->device));
                CU(s->d_gather.ensure(count * G * 8));
            }
        NC(N.GroupStart());
        for (rr_engine *s : sh) {
            double *buf = (s->*dots).as<double>() + off;
            if (!dd) NC(N.Reduce(buf, buf, count, ncclDouble, ncclSum, 0, s->comm, s->stream));
            else NC(N.AllGather(buf, s->d_gather.p, count, ncclDouble, s->comm, s->stream));
        }
        NC(N.GroupEnd());
        e->stats.collectives++;
        CU(cudaSetDevice(e->device));
        if (dd) {
            rr::rr_reduce_rows_dd<<<(int)((count / 2 + 255) / 256), 256, 0, e->stream>>>(e->d_gather.as<double>(), (int64_t)count, (int)G,
                                                                                       (int)(count / 2), (e->*dots).as<double>() + off);
            CU(cudaGetLastError());
            e->stats.kernel_launches++;
        }
        return RR_OK;
    }
    if (!e->sharded()) return RR_OK;
    double *buf = (e->*dots).as<double>() + off;
    if (!dd) {
        if (e->comm) {
            NC(N.AllReduce(buf, buf, count, ncclDouble, ncclSum, e->comm, e->stream));
            e->stats.collectives++;
        } else if (e->allreduce(buf, count, e->stream, e->allreduce_user)) {
            return e->fail(RR_ERR_COLLECTIVE, "all-reduce hook failed");
        }
        return RR_OK;
    }
    CU(e->d_gather.ensure(count * e->world * 8));
    if (e->comm) {
        NC(N.AllGather(buf, e->d_gather.p, count, ncclDouble, e->comm, e->stream));
        e->stats.collectives++;
    } else {
        // sum-of-disjoint-segments all-reduce (exact)
        CU(cudaMemsetAsync(e->d_gather.p, 0, count * e->world * 8, e->stream));
        CU(cudaMemcpyAsync(e->d_gather.as<double>() + count * e->rank, buf, count * 8, cudaMemcpyDeviceToDevice, e->stream));
        if (e->allreduce(e->d_gather.p, count * e->world, e->stream, e->allreduce_user))
            return e->fail(RR_ERR_COLLECTIVE, "all-reduce hook failed");
    }
    rr::rr_reduce_rows_dd<<<(int)((count / 2 + 255) / 256), 256, 0, e->stream>>>(e->d_gather.as<double>(), (int64_t)count, e->world,
                                                                               (int)(count / 2), buf);
    CU(cudaGetLastError());
    e->stats.kernel_launches++;
    return RR_OK;
}

// Runs one plan over every shard of the engine: launch the interpreter, reduce rows into the shard's `dots`, sum over
// shards / ranks. dd: the plan holds DOTDD reductions only.
// set / dots_off / sync: run_gram plans a large batch in pieces and launches a piece while it plans the next, so two
// sweeps can be in flight: `set` picks the device buffers and the event pair, the reduced dots land at
// dots + dots_off (the caller has sized `dots` for all pieces: growing it here would drop the earlier ones),
// and with sync = false the call returns right after the launches (finish_sweep reads the time later).
int run_sweep(rr_engine *e, const rr::SweepPlan &P, const SweepCfg &cfg, DotsMember dots, bool dd, double *stg, int64_t ld_stg,
              int set = 0, size_t dots_off = 0, bool sync = true, bool g8 = false, bool reduce = true, bool mark_begin = true)
{
    if (P.chunks.empty()) return RR_OK;
    // the kernel streams whole windows of kInsWindow instructions: pad the tail with ENDs
    std::vector<RRIns> ins(P.ins);
    RRIns endi;
    std::memset(&endi, 0, sizeof(endi));
    ins.resize(P.ins.size() + rr::kInsWindow, endi);
    int rc = RR_OK;
    for (rr_engine *s : shards_of(e)) {
        const XView view{s->X.as<double>(), s->ld, s->n};
        if ((rc = launch_shard(e, s, P, ins, cfg, view, dots, dd, stg, ld_stg, set, dots_off, g8, mark_begin))) break;
    }
    if (e->grouped()) cudaSetDevice(e->device);
    if (rc) return rc;
    if (reduce && P.n_dots > 0 && (rc = reduce_over_ranks(e, dots, dots_off, (size_t)P.n_dots, dd))) return rc;
    e->stats.distinct_dots += P.n_dot_ins;
    if (P.r8) {
        e->stats.row_groups += P.n_gram_groups;
        e->stats.row_group_rows += P.n_gram_rows;
    }
    e->stats.w_shared += P.w_issued;
    if (!sync) return RR_OK;
    // sweep time is read after the synchronisation
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev[set ? 4 : 2], e->ev[set ? 5 : 3]);
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

// deterministic partial sums over one shard's rows: mode 0 -> {sum y, 0}; mode 1 -> writes yc = y - mean, {sum yc, sum yc^2}
int y_partial(rr_engine *e, rr_engine *s, int mode, double mean, double out[2])
{
    const int blocks = 296;
    CU(cudaSetDevice(s->device));
    CU(s->d_misc.ensure(blocks * 2 * 8 + 64));
    std::vector<double> part(blocks * 2);
    double *y = s->X.as<double>() + (size_t)s->d * s->ld;
    double *yc = s->X.as<double>() + (size_t)(s->d + 1) * s->ld;
    k_y_stats<<<blocks, 256, 0, s->stream>>>(y, yc, s->n, mean, mode, s->d_misc.as<double>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(part.data(), s->d_misc.p, blocks * 2 * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    out[0] = out[1] = 0.0;
    for (int i = 0; i < blocks; ++i) { out[0] += part[2 * i]; out[1] += part[2 * i + 1]; }
    return RR_OK;
}

// a few doubles summed over the ranks of a communicator / hook (engine creation only)
int allreduce_small(rr_engine *e, double *vals, int count)
{
    if (!e->sharded()) return RR_OK;
    CU(cudaSetDevice(e->device));
    CU(e->d_misc.ensure(64 + (size_t)count * 8));
    CU(cudaMemcpyAsync(e->d_misc.p, vals, (size_t)count * 8, cudaMemcpyHostToDevice, e->stream));
    if (e->comm) {
        NC(nccl().AllReduce(e->d_misc.p, e->d_misc.p, (size_t)count, ncclDouble, ncclSum, e->comm, e->stream));
    } else if (e->allreduce(e->d_misc.p, (size_t)count, e->stream, e->allreduce_user)) {
        return e->fail(RR_ERR_COLLECTIVE, "all-reduce hook failed");
    }
    CU(cudaMemcpyAsync(vals, e->d_misc.p, (size_t)count * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return RR_OK;
}

int compute_y_stats(rr_engine *e)
{
    // mean and centred sums, deterministic: block partials in fixed order, shards in order on the host,
    // ranks by the collective
    std::vector<rr_engine *> sh = shards_of(e);
    double tot[2] = {0.0, 0.0};
    int rc;
    for (rr_engine *s : sh) {
        double p[2];
        if ((rc = y_partial(e, s, 0, 0.0, p))) return rc;
        tot[0] += p[0];
        tot[1] += (double)s->n;
    }
    if ((rc = allreduce_small(e, tot, 2))) return rc;
    e->n_total = (int64_t)llround(tot[1]);
    e->y_mean = tot[0] / tot[1];
    double tot2[2] = {0.0, 0.0};
    for (rr_engine *s : sh) {
        double p[2];
        if ((rc = y_partial(e, s, 1, e->y_mean, p))) return rc;
        tot2[0] += p[0];
        tot2[1] += p[1];
    }
    if ((rc = allreduce_small(e, tot2, 2))) return rc;
    CU(cudaSetDevice(e->device));
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

// what a shard ingests: rows of the caller's matrix (host, or device with RR_FLAG_X_DEVICE)
struct Ingest {
    const double *X = nullptr, *y = nullptr;
    int64_t n_src = 0;        // rows of the caller's matrix
    bool rowmajor = false;
    const int32_t *pos = nullptr;  // host, n_src entries: engine row of each source row (-1: not selected); nullptr: identity
    int64_t lo = 0, hi = 0;   // engine rows [lo, hi) belong to this shard
};

// One engine = one device = one block of rows. Nothing is synchronised against other shards here.
int create_shard(const Ingest &in, int32_t d, int32_t device, uint32_t flags, rr_engine **out)
{
    *out = nullptr;
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
    const int64_t n = in.hi - in.lo;
    rr_engine *e = new rr_engine();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->flags = flags;
    e->n = n;
    e->d = d;
    e->n_total = n;
    e->n_object = n;
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
    double *ycol = e->X.as<double>() + (size_t)d * e->ld;
    if (!in.rowmajor) {
        // feature-major source: column j of the shard = rows [lo, hi) of column j (no index support: see the caller)
        CUC(cudaMemcpy2DAsync(e->X.p, e->ld * 8, in.X + in.lo, (size_t)in.n_src * 8, (size_t)n * 8, d, kind, e->stream));
        CUC(cudaMemcpyAsync(ycol, in.y + in.lo, (size_t)n * 8, kind, e->stream));
    } else {
        // row-major source in row chunks: upload, then transpose (and scatter through the index) on the device.
        // Without an index only this shard's rows are uploaded; with one every source row may land here.
        const int64_t src0 = in.pos ? 0 : in.lo, src1 = in.pos ? in.n_src : in.hi;
        const int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(src1 - src0, (int64_t)(env_double("RR_B200_INGEST_CHUNK_BYTES", 256e6) / (8.0 * d))));
        if (!on_dev) {
            CUC(e->d_V.ensure((size_t)chunk * d * 8));
            CUC(e->d_rhs.ensure((size_t)chunk * 8));
        }
        if (in.pos) CUC(e->d_perm.ensure((size_t)chunk * 4));
        for (int64_t r0 = src0; r0 < src1; r0 += chunk) {
            const int64_t rows = std::min(chunk, src1 - r0);
            const double *xsrc = in.X + (size_t)r0 * d, *ysrc = in.y + r0;
            if (!on_dev) {
                CUC(cudaMemcpyAsync(e->d_V.p, xsrc, (size_t)rows * d * 8, cudaMemcpyHostToDevice, e->stream));
                CUC(cudaMemcpyAsync(e->d_rhs.p, ysrc, (size_t)rows * 8, cudaMemcpyHostToDevice, e->stream));
                xsrc = e->d_V.as<double>();
                ysrc = e->d_rhs.as<double>();
            }
            const int32_t *pos = nullptr;
            if (in.pos) {
                CUC(cudaMemcpyAsync(e->d_perm.p, in.pos + r0, (size_t)rows * 4, cudaMemcpyHostToDevice, e->stream));
                pos = e->d_perm.as<int32_t>();
            }
            dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((d + 31) / 32));
            k_ingest_rows<<<grid, dim3(32, 8), 0, e->stream>>>(xsrc, ysrc, rows, d, pos, r0, in.lo, in.hi, e->X.as<double>(), e->ld, ycol);
            CUC(cudaGetLastError());
            e->stats.kernel_launches++;
        }
    }
    CUC(cudaStreamSynchronize(e->stream));
#undef CUC
    *out = e;
    return RR_OK;
}

int create_common(const double *Xsrc, const double *y, int64_t n, int32_t d, int32_t device, uint32_t flags,
                  bool rowmajor, rr_engine **out)
{
    if (!out) { g_thread_error = "out is null"; return RR_ERR_INVALID; }
    *out = nullptr;
    if (!Xsrc || !y || n <= 0 || d <= 0) { g_thread_error = "rr_engine_create: bad arguments"; return RR_ERR_INVALID; }
    const auto t0 = std::chrono::steady_clock::now();
    Ingest in;
    in.X = Xsrc;
    in.y = y;
    in.n_src = n;
    in.rowmajor = rowmajor;
    in.lo = 0;
    in.hi = n;
    rr_engine *e = nullptr;
    int rc = create_shard(in, d, device, flags, &e);
    if (rc) return rc;
    if ((rc = compute_y_stats(e))) {
        g_thread_error = e->error;
        e->free_all();
        delete e;
        return rc;
    }
    e->stats.ingest_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = e;
    return RR_OK;
}

// rr_engine_create_sharded: n_gpus shards in this process, optional row selection
int create_sharded(const double *Xsrc, const double *y, int64_t n_src, int32_t d, const int32_t *row_index, int64_t n_rows,
                   int32_t n_gpus, uint32_t flags, rr_engine **out)
{
    if (!out) { g_thread_error = "out is null"; return RR_ERR_INVALID; }
    *out = nullptr;
    if (!Xsrc || !y || n_src <= 0 || d <= 0 || n_rows <= 0 || (!row_index && n_rows != n_src)) {
        g_thread_error = "rr_engine_create_sharded: bad arguments";
        return RR_ERR_INVALID;
    }
    const bool rowmajor = (flags & RR_FLAG_X_ROWMAJOR) != 0;
    if (row_index && (!rowmajor || (flags & RR_FLAG_X_DEVICE))) {
        g_thread_error = "rr_engine_create_sharded: a row index needs a row-major host matrix";
        return RR_ERR_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        g_thread_error = "no CUDA device: this engine has no CPU fallback";
        return RR_ERR_NO_DEVICE;
    }
    int G = n_gpus <= 0 ? count : std::min<int>(n_gpus, count);
    G = (int)std::max<int64_t>(1, std::min<int64_t>(G, n_rows / 1024));  // no shard below one tile's worth of rows
    const auto t0 = std::chrono::steady_clock::now();
    // engine row of every source row
    std::vector<int32_t> pos;
    if (row_index) {
        pos.assign((size_t)n_src, -1);
        for (int64_t i = 0; i < n_rows; ++i) {
            const int32_t r = row_index[i];
            if (r < 0 || r >= n_src) { g_thread_error = "rr_engine_create_sharded: row index out of range"; return RR_ERR_INVALID; }
            pos[(size_t)r] = (int32_t)i;  // a row listed twice keeps its last position (the reference's index is a permutation)
        }
    }
    const bool verbose = env_int("RR_B200_VERBOSE", 0) != 0;
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    if (verbose) std::fprintf(stderr, "[rr_b200] ingest: index inverted at        %8.2f ms\n", since());
    std::vector<rr_engine *> sh;
    auto bail = [&](int code) {
        for (rr_engine *s : sh) {
            cudaSetDevice(s->device);
            s->free_all();
            delete s;
        }
        return code;
    };
    int device0 = 0;
    if (G == 1) cudaGetDevice(&device0);
    for (int g = 0; g < G; ++g) {
        Ingest in;
        in.X = Xsrc;
        in.y = y;
        in.n_src = n_src;
        in.rowmajor = rowmajor;
        in.pos = row_index ? pos.data() : nullptr;
        in.lo = n_rows * g / G;
        in.hi = n_rows * (g + 1) / G;
        rr_engine *s = nullptr;
        const int rc = create_shard(in, d, G == 1 ? device0 : g, flags & ~(uint32_t)RR_FLAG_X_ROWMAJOR, &s);
        if (rc) return bail(rc);
        sh.push_back(s);
        if (verbose) std::fprintf(stderr, "[rr_b200] ingest: shard %d resident at         %8.2f ms\n", g, since());
    }
    rr_engine *e = sh[0];
    e->n_object = n_rows;
    if (G > 1) {
        std::lock_guard<std::mutex> lock(g_nccl_mu);
        NcclApi &N = nccl();
        if (!N.load()) { g_thread_error = N.err; return bail(RR_ERR_COLLECTIVE); }
        std::vector<ncclComm_t> comms(G);
        std::vector<int> devs(G);
        for (int g = 0; g < G; ++g) devs[g] = sh[g]->device;
        const ncclResult_t r = N.CommInitAll(comms.data(), G, devs.data());
        if (r != ncclSuccess) { g_thread_error = std::string("ncclCommInitAll: ") + N.GetErrorString(r); return bail(RR_ERR_COLLECTIVE); }
        for (int g = 0; g < G; ++g) {
            sh[g]->comm = comms[g];
            sh[g]->rank = g;
            sh[g]->world = G;
        }
        e->peers.assign(sh.begin() + 1, sh.end());
        cudaSetDevice(e->device);
    }
    const int rc = compute_y_stats(e);
    if (rc) {
        g_thread_error = e->error;
        cudaSetDevice(e->device);
        e->free_all();  // releases the peers too
        delete e;
        return rc;
    }
    e->n_total = n_rows;
    e->stats.ingest_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (verbose) std::fprintf(stderr, "[rr_b200] ingest: communicator + target statistics at %8.2f ms\n", since());
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
    int rc = run_sweep(e, P, S, &rr_engine::d_dots, false, nullptr, 0);
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
    rc = run_sweep(e, P, S, &rr_engine::d_dots, false, e->d_V.as<double>(), e->ld);
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
    const double *ycol = e->X.as<double>() + (size_t)e->d * e->ld;
    // One warp per candidate with the matrix in shared memory (rr_exact_qr_warp) whenever a candidate's matrix fits:
    // candidates are bucketed by column count so that one outlier (expand() can blow a candidate up to dozens of
    // terms) does not cost every warp its shared memory. Larger n: one thread per candidate in global memory.
    const size_t smem_cap = 200 * 1024;
    auto per_warp_bytes = [&](int kcap) { return ((size_t)n * (kcap + 1) + 5 * (size_t)kcap + (size_t)(kcap + 1) / 2 + 1) * 8; };
    const int kcap_a = std::min(kmax, 12);
    const int wpb_a = (int)std::min<size_t>(8, smem_cap / per_warp_bytes(kcap_a));
    const int wpb_b = (int)std::min<size_t>(8, smem_cap / per_warp_bytes(kmax));
    const bool warp_path = wpb_a >= 1 && wpb_b >= 1 && env_int("RR_B200_EXACT_WARP", 1) != 0;
    if (warp_path) {
        std::vector<int32_t> la, lb;
        for (int c = 0; c < nc; ++c) (bp.k_of(c) <= kcap_a ? la : lb).push_back(c);
        std::vector<int32_t> both(la);
        both.insert(both.end(), lb.begin(), lb.end());
        if ((rc = upload(e, e->d_list, both.data(), both.size()))) return rc;
        rr::ExactArgs a;
        std::memset(&a, 0, sizeof(a));
        a.V = e->d_V.as<double>();
        a.ldv = e->ld;
        a.y = ycol;
        a.n = n;
        a.term_ids = e->d_tid.as<int32_t>();
        a.cand_term_begin = e->d_ctb.as<int32_t>();
        a.kmax = kmax;
        a.coef = e->d_coef.as<double>();
        a.nzp = e->d_nzp.as<int32_t>();
        a.flags = e->d_flags.as<uint32_t>();
        CU(cudaFuncSetAttribute(rr::rr_exact_qr_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        struct Bucket { size_t begin, count; int kcap, wpb; };
        for (const Bucket &bk : {Bucket{0, la.size(), kcap_a, wpb_a}, Bucket{la.size(), lb.size(), kmax, wpb_b}}) {
            if (!bk.count) continue;
            rr::ExactWarpArgs w;
            w.a = a;
            w.list = e->d_list.as<int32_t>() + bk.begin;
            w.n_list = (int32_t)bk.count;
            w.kcap = bk.kcap;
            const size_t smem = per_warp_bytes(bk.kcap) * bk.wpb;
            rr::rr_exact_qr_warp<<<(unsigned)((bk.count + bk.wpb - 1) / bk.wpb), bk.wpb * 32, smem, e->stream>>>(w);
            CU(cudaGetLastError());
            e->stats.kernel_launches++;
        }
    } else {
    CU(e->d_A.ensure((size_t)n * kmax * Q * 8));
    CU(e->d_rhs.ensure((size_t)n * Q * 8));
    CU(e->d_aux.ensure((size_t)Q * 5 * kmax * 8));
    CU(e->d_perm.ensure((size_t)Q * kmax * 4));
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
// A contiguous run of a batch's candidates as a batch of its own (no copy: the candidate offsets are rebased, the term
// offsets stay absolute into the caller's code array), with its own planner: the pieces of a step are ANALYSED and
// planned independently, each on its own thread, so that what precedes the first kernel launch is the analysis and
// the plan of the first, small piece only.
struct RangeBatch {
    int32_t c0 = 0, c1 = 0;
    std::vector<int32_t> ctb;
    rr_batch view{};
    std::unique_ptr<rr::BatchPlanner> bp;
    void init(const rr_batch *b, int32_t lo, int32_t hi, int32_t d)
    {
        c0 = lo;
        c1 = hi;
        const int32_t t0 = b->cand_term_begin[lo];
        ctb.resize((size_t)(hi - lo) + 1);
        for (int32_t c = lo; c <= hi; ++c) ctb[(size_t)(c - lo)] = b->cand_term_begin[c] - t0;
        view = *b;
        view.n_cand = hi - lo;
        view.cand_term_begin = ctb.data();
        view.term_code_begin = b->term_code_begin + t0;
        bp.reset(new rr::BatchPlanner(&view, d));
    }
};
// An arbitrary list of a batch's candidates gathered into a batch of its own (the escalation and refinement passes
// touch a few dozen candidates: they are analysed on their own instead of dragging the whole neighbourhood's analysis)
struct ListBatch {
    std::vector<int32_t> cand, ctb, tcb;
    std::vector<uint32_t> code;
    rr_batch view{};
    std::unique_ptr<rr::BatchPlanner> bp;
    std::string init(const rr_batch *b, const std::vector<int32_t> &list, int32_t d, bool no_cse)
    {
        cand = list;
        ctb.assign(1, 0);
        tcb.assign(1, 0);
        code.clear();
        for (int32_t c : list) {
            for (int32_t t = b->cand_term_begin[c]; t < b->cand_term_begin[c + 1]; ++t) {
                code.insert(code.end(), b->code + b->term_code_begin[t], b->code + b->term_code_begin[t + 1]);
                tcb.push_back((int32_t)code.size());
            }
            ctb.push_back((int32_t)tcb.size() - 1);
        }
        view = *b;
        view.n_cand = (int32_t)list.size();
        view.cand_term_begin = ctb.data();
        view.term_code_begin = tcb.data();
        view.code = code.data();
        view.n_code = (int32_t)code.size();
        bp.reset(new rr::BatchPlanner(&view, d));
        return bp->analyse(no_cse);
    }
};

// structural validation of an OLS_FIT batch that is going to be analysed piecewise (what BatchPlanner::analyse checks
// about the offset arrays, without touching the code)
std::string validate_offsets(const rr_batch *b)
{
    if (!b->cand_term_begin || !b->term_code_begin) return "null batch arrays";
    if (b->cand_term_begin[0] != 0 || b->term_code_begin[0] != 0) return "offset arrays must start at 0";
    for (int32_t c = 0; c < b->n_cand; ++c)
        if (b->cand_term_begin[c + 1] < b->cand_term_begin[c]) return "candidate offsets must not decrease";
    const int32_t n_terms = b->cand_term_begin[b->n_cand];
    if (n_terms > 0 && !b->code) return "null batch arrays";
    for (int32_t t = 0; t < n_terms; ++t)
        if (b->term_code_begin[t + 1] <= b->term_code_begin[t]) return "empty term program";
    if (b->n_consts < 0 || (b->n_consts > 0 && !b->consts)) return "bad constant pool";
    if (b->n_code < 0) return "negative code length";
    if (b->n_code > 0 && n_terms > 0 && b->term_code_begin[n_terms] > b->n_code) return "term offsets exceed the code length";
    return "";
}

int run_gram(rr_engine *e, const rr_batch *b, rr_result *res)
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
    {
        const std::string verr = validate_offsets(b);
        if (!verr.empty()) return e->fail(RR_ERR_INVALID, "malformed batch: " + verr);
    }
    const int n_terms = b->cand_term_begin[nc];
    const int n_coef = n_terms + nc;
    const bool no_cse = (e->flags & RR_FLAG_NO_CSE) != 0;
    auto k_of = [&](int c) { return b->cand_term_begin[c + 1] - b->cand_term_begin[c] + 1; };
    rr::PlanLimits lim = limits_for(e, S, nc);
    rr::ColIds cols{e->d, e->d + 1};
    int rc = ensure_result_buffers(e, nc, n_coef);
    if (rc) return rc;

    // pass 1: Gram / A^T yc / column sums, shared across candidates.
    // The neighbourhood is planned in PIECES - contiguous runs of candidates - on helper threads, and every piece is
    // launched as soon as its plan exists: planning (host work, ~1 us per candidate) hides behind the sweeps of the
    // pieces before it instead of preceding the whole step, which is what bounds a step once the rows are sharded
    // over several GPUs. Pieces share nothing but the base solution's terms, which each piece evaluates and pins
    // once (a few dozen instructions). On the 4-samples-per-thread shape the pieces are G8 plans (reductions by
    // DMMA, rr_sweep_g8.cuh); candidates with more terms than there are pins form one classic piece of their own.
    // All pieces write into one dot vector, which is summed over shards / ranks ONCE behind the last piece.
    struct Piece {
        RangeBatch rb;                       // its candidates, analysed on their own
        std::vector<int32_t> narrow, wide;   // local candidate indices by plan kind
        rr::SweepPlan P, Pw;                 // the G8 (or classic) plan of `narrow`, the classic plan of `wide`
        std::vector<int32_t> tab, tab_begin, tabw, tabw_begin;
        std::string err;
        size_t off = 0, offw = 0;
        bool g8 = false;
    };
    const bool big_shape = S.S == 4 && S.TH == 128 && lim.target_chunks == 1;
    const bool use_g8 = big_shape && env_int("RR_B200_G8", 1) != 0 && e->d + 3 <= g8_tile_cols();
    std::vector<Piece> pieces;
    {
        // piece sizes: a small first piece (its analysis + plan is all that precedes the first launch), the rest in equal
        // parts; small batches are one piece
        const bool pipelined = env_int("RR_B200_PIPELINE", 1) != 0 && nc >= 1024 && big_shape;
        // Pieces are planned on their own threads from the start, so piece k's plan exists after ~1.6 us per candidate
        // of it and is needed when the sweeps of the pieces before it end (~2.9e-6 us per candidate and row of this
        // shard). Fewer, larger pieces share more (2^24 rows on one GPU: two pieces 199 ms per step, four 203.5, six
        // 210.5; one piece: 194 ms of sweeps behind 7 ms of planning), so every piece is made as large as the sweeps in
        // front of it can hide: 128 + the rest at 2^24 rows, 128 + ~600 + ~2900 + the rest at 2^21 rows per GPU.
        // RR_B200_PIECES = k forces the old shape (a first piece, then k - 1 equal parts).
        const int forced = env_int("RR_B200_PIECES", 0);
        std::vector<int32_t> cut{0};
        if (pipelined && forced != 1) {
            const int32_t first = std::min<int32_t>(nc, std::max(64, env_int("RR_B200_FIRST_PIECE", 128)));
            cut.push_back(first);
            if (forced > 1) {
                for (int i = 1; i < forced; ++i) cut.push_back(first + (int32_t)((int64_t)(nc - first) * i / (forced - 1)));
            } else {
                const double t_sweep = env_double("RR_B200_SWEEP_US_PER_CAND_ROW", 2.9e-6) * (double)e->n;
                const double t_plan = env_double("RR_B200_PLAN_US_PER_CAND", 1.6);
                int32_t cum = first;
                while (cum < nc) {
                    int32_t next = (int32_t)std::min<double>((double)(nc - cum), std::max(256.0, (200.0 + t_sweep * cum) / t_plan));
                    if (nc - cum - next < 128) next = nc - cum;
                    cum += next;
                    cut.push_back(cum);
                }
            }
        } else {
            cut.push_back(nc);
        }
        cut.erase(std::unique(cut.begin(), cut.end()), cut.end());
        pieces.resize(cut.size() - 1);
        for (size_t i = 0; i + 1 < cut.size(); ++i) {
            Piece &pc = pieces[i];
            pc.rb.init(b, cut[i], cut[i + 1], e->d);
            pc.g8 = use_g8;
            for (int32_t c = cut[i]; c < cut[i + 1]; ++c) (use_g8 && k_of(c) - 1 > RR_NPIN - 1 ? pc.wide : pc.narrow).push_back(c - cut[i]);
        }
    }
    rr::PlanLimits lim_g8 = lim;
    lim_g8.tile_cols = g8_tile_cols();
    lim_g8.g8 = true;
    lim_g8.ins_window = rr::kG8Window;
    lim_g8.mdot_rows = false;
    // R8 plans (the row machine, rr_sweep_r8.cuh) on request (RR_B200_R8=1): measured slower than G8 plans on the headline
    // neighbourhood ("the row machine"), so they are not the default. A piece whose rows do not fill their
    // groups (or that the row machine cannot plan) falls back to its G8 plan.
    const bool use_r8 = use_g8 && env_int("RR_B200_R8", 0) != 0 && e->d + 4 <= r8_tile_cols();
    const double r8_min_fill = env_double("RR_B200_R8_MIN_FILL", 3.0);
    rr::PlanLimits lim_r8 = lim;
    lim_r8.tile_cols = r8_tile_cols();
    std::string err;
    auto plan_piece = [&](Piece &pc) {
        pc.err = pc.rb.bp->analyse(no_cse);
        if (!pc.err.empty()) { pc.err = "malformed batch: " + pc.err; return; }
        if (!pc.narrow.empty() && use_r8) {
            const std::string e8 = pc.rb.bp->plan_gram_r8(lim_r8, cols, &pc.narrow, pc.P, pc.tab, pc.tab_begin);
            const double work = (double)pc.P.n_gram_groups + (double)pc.P.n_stored_evals;
            if (e8.empty() && (double)pc.P.n_gram_rows >= r8_min_fill * std::max(work, 1.0)) 
            {
                if (!pc.wide.empty()) pc.err = pc.rb.bp->plan_gram(lim, cols, &pc.wide, false, pc.Pw, pc.tabw, pc.tabw_begin);
                return;
            }
            pc.P = rr::SweepPlan();
            pc.tab.clear();
            pc.tab_begin.clear();
        }
        if (!pc.narrow.empty())
            pc.err = pc.g8 ? pc.rb.bp->plan_gram_g8(lim_g8, cols, &pc.narrow, pc.P, pc.tab, pc.tab_begin)
                           : pc.rb.bp->plan_gram(lim, cols, &pc.narrow, false, pc.P, pc.tab, pc.tab_begin);
        if (pc.err.empty() && !pc.wide.empty()) pc.err = pc.rb.bp->plan_gram(lim, cols, &pc.wide, false, pc.Pw, pc.tabw, pc.tabw_begin);
    };
    std::vector<int32_t> tab, tab_begin;
    std::vector<uint32_t> cmask((size_t)nc, 0u);  // per candidate: terms that are constant by construction (rr_plan.h)
    {
        // the reduced dots of all pieces live in one vector: size it before the first launch
        size_t tab_total = 0;
        for (int c = 0; c < nc; ++c) {
            const size_t m = (size_t)k_of(c) - 1;
            tab_total += m * (m + 1) / 2 + 2 * m;
        }
        const size_t dots_cap = tab_total + 128 * pieces.size() + 256;
        for (rr_engine *s : shards_of(e)) {
            CU(cudaSetDevice(s->device));
            CU(s->d_dots.ensure(dots_cap * 8));
        }
        CU(cudaSetDevice(e->device));
        std::vector<std::thread> helpers(pieces.size());
        struct Joiner {  // whatever path leaves this scope, the helpers are joined first (they write into `pieces`)
            std::vector<std::thread> &t;
            ~Joiner()
            {
                for (std::thread &x : t)
                    if (x.joinable()) x.join();
            }
        } joiner{helpers};
        std::vector<char> threaded(pieces.size(), 0);
        for (size_t i = 1; i < pieces.size(); ++i) {
            try {
                helpers[i] = std::thread(plan_piece, std::ref(pieces[i]));
                threaded[i] = 1;
            } catch (...) {  // no thread to be had: that piece is planned here, when its turn comes
            }
        }
        size_t off = 0;
        int launches = 0;
        for (size_t i = 0; i < pieces.size(); ++i) {
            Piece &pc = pieces[i];
            const double tp0 = tnow();
            if (threaded[i]) helpers[i].join();
            else plan_piece(pc);
            if (verbose) std::fprintf(stderr, "[rr_b200]   piece %zu (%s, %zu + %zu cand) plan wait %6.2f ms, %zu ins, %d dots\n", i, pc.P.r8 ? "r8" : pc.g8 ? "g8" : "classic",
                                      pc.narrow.size(), pc.wide.size(), tnow() - tp0, pc.P.ins.size(), pc.P.n_dots);
            if (!pc.err.empty()) { cudaStreamSynchronize(e->stream); return e->fail(RR_ERR_INVALID, pc.err); }
            for (int w = 0; w < 2; ++w) {
                const rr::SweepPlan &P = w ? pc.Pw : pc.P;
                if (P.chunks.empty()) continue;
                (w ? pc.offw : pc.off) = off;
                if (off + (size_t)P.n_dots > dots_cap) { cudaStreamSynchronize(e->stream); return e->fail(RR_ERR_INVALID, "internal: dot vector too small"); }
                rc = run_sweep(e, P, S, &rr_engine::d_dots, false, nullptr, 0, launches & 1, off, false, pc.g8 && w == 0 && !P.r8, false, launches == 0);
                if (rc) { cudaStreamSynchronize(e->stream); return rc; }
                ++launches;
                off += (size_t)round_up(std::max(P.n_dots, 1), 32);
            }
            e->stats.term_instances += pc.rb.bp->n_term_instances();
            e->stats.distinct_terms += pc.rb.bp->n_terms_distinct();
            e->stats.w_contract += pc.rb.bp->w_contract();
            for (int32_t lc = 0; lc < pc.rb.c1 - pc.rb.c0; ++lc) cmask[(size_t)(pc.rb.c0 + lc)] = pc.rb.bp->cand_const_mask(lc);
        }
        if ((rc = reduce_over_ranks(e, &rr_engine::d_dots, 0, off, false))) { cudaStreamSynchronize(e->stream); return rc; }
        // per-candidate tables in candidate order
        tab_begin.assign(1, 0);
        for (size_t i = 0; i < pieces.size(); ++i) {
            const Piece &pc = pieces[i];
            std::vector<std::pair<int32_t, int32_t>> where((size_t)(pc.rb.c1 - pc.rb.c0));  // local candidate -> (plan, index in its list)
            for (size_t j = 0; j < pc.narrow.size(); ++j) where[(size_t)pc.narrow[j]] = {0, (int32_t)j};
            for (size_t j = 0; j < pc.wide.size(); ++j) where[(size_t)pc.wide[j]] = {1, (int32_t)j};
            for (size_t lc = 0; lc < where.size(); ++lc) {
                const bool w = where[lc].first != 0;
                const std::vector<int32_t> &tb = w ? pc.tabw_begin : pc.tab_begin, &tt = w ? pc.tabw : pc.tab;
                const size_t base = w ? pc.offw : pc.off;
                for (int32_t k = tb[(size_t)where[lc].second]; k < tb[(size_t)where[lc].second + 1]; ++k) tab.push_back(tt[(size_t)k] + (int32_t)base);
                tab_begin.push_back((int32_t)tab.size());
            }
        }
        CU(cudaStreamSynchronize(e->stream));
        // the pieces' sweeps as one span: from the first launch (event 2) to the end of the last
        if (launches > 0) {
            float ms = 0.f;
            const int last = (launches - 1) & 1;
            if (cudaEventElapsedTime(&ms, e->ev[2], e->ev[last ? 5 : 3]) == cudaSuccess) e->sweep_ms_accum += ms;
        }
        phase("plan + sweep gram (pieces)");
    }

    // per-candidate solve
    std::vector<int64_t> wsoff(nc + 1, 0);
    for (int c = 0; c < nc; ++c) {
        const int64_t kk = k_of(c);
        wsoff[c + 1] = wsoff[c] + 2 * kk * kk + 10 * kk;
    }
    CU(e->d_ws.ensure((size_t)wsoff[nc] * 8));
    if ((rc = upload(e, e->d_wsoff, wsoff.data(), wsoff.size()))) return rc;
    if ((rc = upload(e, e->d_tab, tab.data(), tab.size()))) return rc;
    // tab_begin is uploaded behind the table in the same buffer family
    DevBuf &d_tabb = e->d_rtab;  // reused later for the residual tables; order of use is sequential
    if ((rc = upload(e, d_tabb, tab_begin.data(), tab_begin.size()))) return rc;
    if ((rc = upload(e, e->d_ctb, b->cand_term_begin, (size_t)nc + 1))) return rc;
    if ((rc = upload(e, e->d_cmask, cmask.data(), cmask.size()))) return rc;
    rr::GramArgs g;
    g.dots = e->d_dots.as<double>();
    g.cand_const = e->d_cmask.as<uint32_t>();
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
            const int64_t kk = k_of(list[i]);
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
        ListBatch lbd;
        err = lbd.init(b, escalate, e->d, no_cse);
        if (err.empty()) err = lbd.bp->plan_gram(lim, cols, nullptr, true, Pd, dtab, dtab_begin);
        if (!err.empty()) return e->fail(RR_ERR_INVALID, err);
        rc = run_sweep(e, Pd, S, &rr_engine::d_rdots, true, nullptr, 0);
        if (rc) return rc;
        DevBuf &d_dt = e->d_t0, &d_dtb = e->d_t1, &d_l = e->d_t2, &d_wo = e->d_t3;
        auto cleanup = [&]() {};
        std::vector<int64_t> wso(escalate.size() + 1, 0);
        for (size_t i = 0; i < escalate.size(); ++i) {
            const int64_t kk = k_of(escalate[i]);
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
        {
            // the listed candidates as a batch of their own, their snapped coefficients in that batch's layout
            ListBatch lbr;
            err = lbr.init(b, ok, e->d, no_cse);
            std::vector<double> cs_local((size_t)lbr.ctb.back() + ok.size());
            std::vector<int32_t> all(ok.size());
            for (size_t i = 0; i < ok.size(); ++i) {
                all[i] = (int32_t)i;
                const int32_t c = ok[i], kk = k_of(c);
                std::copy(cs.begin() + (b->cand_term_begin[c] + c), cs.begin() + (b->cand_term_begin[c] + c + kk), cs_local.begin() + (lbr.ctb[i] + (int32_t)i));
            }
            if (err.empty()) err = lbr.bp->plan_residual(lim, cols, all, cs_local.data(), Pr, rtab, rtab_begin);
        }
        if (!err.empty()) return e->fail(RR_ERR_INVALID, err);
        rc = run_sweep(e, Pr, S, &rr_engine::d_rdots, false, nullptr, 0);
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
namespace {

// predict(): the program is planned once (materialising plan: interpreter + RI_STG epilogue) and the caller's
// matrix is streamed through the engine in row chunks - upload, transpose on the device when it is row-major, sweep,
// download - on the engine's own stream and grow-only buffers; the chunks of a multi-GPU engine go round the shards.
// No engine is created, no target statistics are computed, nothing of the resident data set is touched.
int predict_impl(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts, int32_t n_consts,
                 const double *X, int64_t n, int32_t d, bool rowmajor, bool proba, double *out)
{
    if (!e) { g_thread_error = "null engine"; return RR_ERR_INVALID; }
    if (!code || code_len <= 0 || !X || !out || n <= 0 || d <= 0) return e->fail(RR_ERR_INVALID, "rr_predict: bad arguments");
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
    b.n_code = code_len;
    rr::BatchPlanner bp(&b, d);
    std::string err = bp.analyse(true);
    if (!err.empty()) return e->fail(RR_ERR_INVALID, "rr_predict: " + err);
    const int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(4096, (int64_t)(env_double("RR_B200_PREDICT_CHUNK_BYTES", 128e6) / (8.0 * d))));
    // launch shape for the chunk height (the engine's own shape follows its resident n)
    SweepCfg S;
    {
        rr_engine probe;  // only the fields choose_cfg reads
        probe.n = chunk;
        probe.d = d;
        probe.s_pref = e->s_pref;
        probe.th_pref = e->th_pref;
        probe.occ_pref = e->occ_pref;
        probe.d_misc = e->d_misc;
        probe.stream = e->stream;
        CU(cudaSetDevice(e->device));
        S = choose_cfg(&probe);
        e->d_misc = probe.d_misc;  // the probe may have grown the shared scratch buffer
        probe.d_misc = DevBuf();
        probe.stream = nullptr;
    }
    rr::PlanLimits lim;
    lim.tile_cols = S.tile_cols();
    lim.no_cse = true;
    lim.fuse = env_int("RR_B200_FUSE", 1) != 0;
    lim.mdot_rows = S.S == 4 && S.TH == 128;
    lim.target_chunks = 1;
    rr::SweepPlan P;
    err = bp.plan_materialise(lim, rr::ColIds{d, d + 1}, P);
    if (!err.empty()) return e->fail(RR_ERR_INVALID, "rr_predict: " + err);
    std::vector<RRIns> ins(P.ins);
    RRIns endi;
    std::memset(&endi, 0, sizeof(endi));
    ins.resize(P.ins.size() + rr::kInsWindow, endi);
    std::vector<rr_engine *> sh = shards_of(e);
    const int64_t ldc = round_up(chunk, kLdAlign);
    int rc = RR_OK;
    int64_t ci = 0;
    for (int64_t r0 = 0; r0 < n && !rc; r0 += chunk, ++ci) {
        rr_engine *s = sh[(size_t)(ci % (int64_t)sh.size())];
        const int64_t rows = std::min(chunk, n - r0);
        CU(cudaSetDevice(s->device));
        CU(s->d_px.ensure((size_t)d * ldc * 8));
        CU(s->d_pout.ensure((size_t)ldc * 8 * (proba ? 3 : 1)));
        // columns are read in whole tiles: the tail beyond `rows` must hold finite numbers
        CU(cudaMemsetAsync(s->d_px.p, 0, (size_t)d * ldc * 8, s->stream));
        if (rowmajor) {
            CU(s->d_pr.ensure((size_t)chunk * d * 8));
            CU(cudaMemcpyAsync(s->d_pr.p, X + (size_t)r0 * d, (size_t)rows * d * 8, cudaMemcpyHostToDevice, s->stream));
            dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((d + 31) / 32));
            k_ingest_rows<<<grid, dim3(32, 8), 0, s->stream>>>(s->d_pr.as<double>(), nullptr, rows, d, nullptr, 0, 0, rows,
                                                              s->d_px.as<double>(), ldc, nullptr);
            CU(cudaGetLastError());
            e->stats.kernel_launches++;
        } else {
            CU(cudaMemcpy2DAsync(s->d_px.p, (size_t)ldc * 8, X + r0, (size_t)n * 8, (size_t)rows * 8, d, cudaMemcpyHostToDevice, s->stream));
        }
        const XView view{s->d_px.as<double>(), ldc, rows};
        rc = launch_shard(e, s, P, ins, S, view, &rr_engine::d_dots, false, s->d_pout.as<double>(), ldc, 0, 0);
        if (rc) break;
        if (proba) {
            k_proba<<<(unsigned)((rows + 255) / 256), 256, 0, s->stream>>>(s->d_pout.as<double>(), rows, s->d_pout.as<double>() + ldc);
            CU(cudaGetLastError());
            e->stats.kernel_launches++;
            CU(cudaMemcpyAsync(out + 2 * r0, s->d_pout.as<double>() + ldc, (size_t)rows * 16, cudaMemcpyDeviceToHost, s->stream));
        } else {
            CU(cudaMemcpyAsync(out + r0, s->d_pout.p, (size_t)rows * 8, cudaMemcpyDeviceToHost, s->stream));
        }
    }
    for (rr_engine *s : sh) {
        cudaSetDevice(s->device);
        const cudaError_t ce = cudaStreamSynchronize(s->stream);
        if (ce != cudaSuccess && !rc) rc = e->cuda_fail(ce, "rr_predict");
    }
    cudaSetDevice(e->device);
    return rc;
}

int score_batch_impl(rr_engine *e, const rr_batch *b, rr_result *res)
{
    if (!e) { g_thread_error = "null engine"; return RR_ERR_INVALID; }
    if (!b || !res || !res->ssr) return e->fail(RR_ERR_INVALID, "null batch/result (result.ssr is required)");
    if (b->mode != RR_MODE_EVAL_ONLY && b->mode != RR_MODE_OLS_FIT) return e->fail(RR_ERR_INVALID, "bad mode");
    if (b->n_cand < 0) return e->fail(RR_ERR_INVALID, "malformed batch: negative candidate count");
    if (b->n_cand == 0) return RR_OK;
    const auto t_host0 = std::chrono::steady_clock::now();
    CU(cudaSetDevice(e->device));
    // the device clock of the batch starts before the host-side analysis: `last_batch_ms` is what a caller waits for
    CU(cudaEventRecord(e->ev[0], e->stream));
    e->stats.h2d_bytes = e->stats.d2h_bytes = 0;
    e->stats.w_shared = 0.0;
    e->sweep_ms_accum = 0.f;
    bool gram = false;
    if (b->mode == RR_MODE_OLS_FIT) {
        const bool split = e->rows_split();
        bool exact = !split && e->n_total <= e->exact_max_n;
        if (e->flags & RR_FLAG_FORCE_GRAM) exact = false;
        if ((e->flags & RR_FLAG_FORCE_EXACT) && !split) exact = true;
        gram = !exact;
    }
    int rc;
    const uint64_t ti0 = e->stats.term_instances, dt0 = e->stats.distinct_terms;
    if (gram) {
        // the Gram path analyses the batch piece by piece, on the threads that plan the pieces
        e->stats.w_contract = 0.0;
        rc = run_gram(e, b, res);
        if (rc) { e->stats.term_instances = ti0; e->stats.distinct_terms = dt0; return rc; }
    } else {
        rr::BatchPlanner bp(b, e->d);
        std::string err = bp.analyse((e->flags & RR_FLAG_NO_CSE) != 0);
        if (!err.empty()) return e->fail(RR_ERR_INVALID, "malformed batch: " + err);
        rc = b->mode == RR_MODE_EVAL_ONLY ? run_eval(e, b, bp, res) : run_exact(e, b, bp, res);
        if (rc) return rc;
        e->stats.term_instances += bp.n_term_instances();
        e->stats.distinct_terms += bp.n_terms_distinct();
        e->stats.w_contract = bp.w_contract();
    }
    CU(cudaEventRecord(e->ev[1], e->stream));
    CU(cudaEventSynchronize(e->ev[1]));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
    e->stats.last_batch_ms = ms;
    e->stats.last_sweep_ms = e->sweep_ms_accum;
    e->stats.last_host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    e->stats.batches++;
    e->stats.candidates += b->n_cand;
    if (b->mode == RR_MODE_OLS_FIT) {
        for (int c = 0; c < b->n_cand; ++c) {
            const int64_t k = b->cand_term_begin[c + 1] - b->cand_term_begin[c] + 1;
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

int classifier_metrics_impl(rr_engine *e, const rr_batch *b, double *accuracy, double *log_loss, double *abs_loss)
{
    if (!e) { g_thread_error = "null engine"; return RR_ERR_INVALID; }
    if (!b || b->mode != RR_MODE_EVAL_ONLY) return e->fail(RR_ERR_INVALID, "classifier metrics need an EVAL_ONLY batch");
    if (b->n_cand < 0) return e->fail(RR_ERR_INVALID, "malformed batch: negative candidate count");
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
    int rc = run_sweep(e, P, S, &rr_engine::d_dots, false, nullptr, 0);
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

int feature_r2_impl(rr_engine *e, double *r2)
{
    if (!e || !r2) { g_thread_error = "rr_feature_r2: null argument"; return RR_ERR_INVALID; }
    if (e->sharded()) return e->fail(RR_ERR_INVALID, "rr_feature_r2: not available on a rank of a multi-process engine");
    const int blocks = 64, d = e->d;
    std::vector<rr_engine *> sh = shards_of(e);
    std::vector<double> part((size_t)d * blocks * 2), sum(d, 0.0), mean(d), sst(d, 0.0), ssr(d, 0.0);
    for (int pass = 0; pass < 2; ++pass) {
        for (rr_engine *s : sh) {
            CU(cudaSetDevice(s->device));
            CU(s->d_aux.ensure((size_t)d * blocks * 2 * 8 + (size_t)d * 8));
            double *d_part = s->d_aux.as<double>(), *d_mean = d_part + (size_t)d * blocks * 2;
            if (pass) CU(cudaMemcpyAsync(d_mean, mean.data(), (size_t)d * 8, cudaMemcpyHostToDevice, s->stream));
            k_feature_stats<<<dim3(blocks, d), 256, 0, s->stream>>>(s->X.as<double>(), s->ld, s->X.as<double>() + (size_t)d * s->ld, s->n,
                                                                   d_mean, pass, d_part);
            CU(cudaGetLastError());
            e->stats.kernel_launches++;
            CU(cudaMemcpyAsync(part.data(), d_part, part.size() * 8, cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            for (int j = 0; j < d; ++j)
                for (int bi = 0; bi < blocks; ++bi) {
                    const double a = part[2 * ((size_t)j * blocks + bi)], b2 = part[2 * ((size_t)j * blocks + bi) + 1];
                    if (pass == 0) sum[j] += a;
                    else { sst[j] += a; ssr[j] += b2; }
                }
        }
        if (pass == 0)
            for (int j = 0; j < d; ++j) mean[j] = sum[j] / (double)e->n_object;
    }
    CU(cudaSetDevice(e->device));
    for (int j = 0; j < d; ++j) r2[j] = 1 - ssr[j] / sst[j];  // R2(X[j], y), rils_rols_cpp.cpp:40-45 with :763's argument order
    return RR_OK;
}

int read_rows_impl(rr_engine *e, int64_t row0, int64_t rows, double *Xout, double *yout)
{
    if (!e) { g_thread_error = "null engine"; return RR_ERR_INVALID; }
    if (row0 < 0 || rows <= 0 || row0 + rows > e->n_object) return e->fail(RR_ERR_INVALID, "rr_engine_read_rows: range");
    int64_t base = 0;
    for (rr_engine *s : shards_of(e)) {
        const int64_t lo = std::max(row0, base), hi = std::min(row0 + rows, base + s->n);
        if (lo < hi) {
            CU(cudaSetDevice(s->device));
            if (Xout)
                CU(cudaMemcpy2DAsync(Xout + (lo - row0), (size_t)rows * 8, s->X.as<double>() + (lo - base), (size_t)s->ld * 8, (size_t)(hi - lo) * 8,
                                     s->d, cudaMemcpyDeviceToHost, s->stream));
            if (yout)
                CU(cudaMemcpyAsync(yout + (lo - row0), s->X.as<double>() + (size_t)s->d * s->ld + (lo - base), (size_t)(hi - lo) * 8,
                                   cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
        }
        base += s->n;
    }
    CU(cudaSetDevice(e->device));
    return RR_OK;
}

// No C++ exception crosses the C boundary: allocation failures and anything else thrown by the planner or the
// standard library become return codes.
template <typename F> int guarded(rr_engine *e, F &&f)
{
    try {
        return f();
    } catch (const std::bad_alloc &) {
        if (e) e->error = "out of host memory";
        else g_thread_error = "out of host memory";
        return RR_ERR_NOMEM;
    } catch (const std::exception &ex) {
        if (e) e->error = std::string("internal error: ") + ex.what();
        else g_thread_error = std::string("internal error: ") + ex.what();
        return RR_ERR_INVALID;
    } catch (...) {
        if (e) e->error = "internal error";
        else g_thread_error = "internal error";
        return RR_ERR_INVALID;
    }
}

}  // namespace

extern "C" {

int rr_abi_version(void) { return RR_ABI_VERSION; }

const char *rr_last_error(const rr_engine *e) { return e ? e->error.c_str() : g_thread_error.c_str(); }

int rr_engine_create(const double *X, const double *y, int64_t n, int32_t d, int32_t device, uint32_t flags,
                     rr_engine **out)
{
    return guarded(nullptr, [&] { return create_common(X, y, n, d, device, flags, false, out); });
}

int rr_engine_create_rowmajor(const double *X, const double *y, int64_t n, int32_t d, int32_t device, uint32_t flags,
                              rr_engine **out)
{
    return guarded(nullptr, [&] { return create_common(X, y, n, d, device, flags, true, out); });
}

int rr_engine_create_sharded(const double *X, const double *y, int64_t n, int32_t d, const int32_t *row_index, int64_t n_rows,
                             int32_t n_gpus, uint32_t flags, rr_engine **out)
{
    return guarded(nullptr, [&] { return create_sharded(X, y, n, d, row_index, n_rows, n_gpus, flags, out); });
}

void rr_engine_destroy(rr_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    e->free_all();
    delete e;
}

int rr_comm_unique_id(void *id128)
{
    if (!id128) { g_thread_error = "rr_comm_unique_id: null"; return RR_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    NcclApi &N = nccl();
    if (!N.load()) { g_thread_error = N.err; return RR_ERR_COLLECTIVE; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    const ncclResult_t r = N.GetUniqueId(&id);
    if (r != ncclSuccess) { g_thread_error = std::string("ncclGetUniqueId: ") + N.GetErrorString(r); return RR_ERR_COLLECTIVE; }
    std::memcpy(id128, &id, 128);
    return RR_OK;
}

int rr_engine_comm_init(rr_engine *e, const void *id128, int32_t rank, int32_t world)
{
    if (!e) { g_thread_error = "null engine"; return RR_ERR_INVALID; }
    if (!id128 || world < 1 || rank < 0 || rank >= world) return e->fail(RR_ERR_INVALID, "bad rank/world");
    if (e->grouped()) return e->fail(RR_ERR_INVALID, "this engine already shards its rows over its own devices");
    return guarded(e, [&]() -> int {
        NcclApi &N = nccl();
        {
            std::lock_guard<std::mutex> lock(g_nccl_mu);
            if (!N.load()) return e->fail(RR_ERR_COLLECTIVE, N.err);
        }
        CU(cudaSetDevice(e->device));
        if (e->comm) { N.CommDestroy(e->comm); e->comm = nullptr; }
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        NC(N.CommInitRank(&e->comm, world, id, rank));
        e->allreduce = nullptr;
        e->comm_ranks = true;
        e->rank = rank;
        e->world = world;
        return compute_y_stats(e);
    });
}

int rr_engine_set_allreduce(rr_engine *e, rr_allreduce_fn fn, void *user, int32_t rank, int32_t world)
{
    if (!e) return RR_ERR_INVALID;
    if (world < 1 || rank < 0 || rank >= world) return e->fail(RR_ERR_INVALID, "bad rank/world");
    if (e->grouped()) return e->fail(RR_ERR_INVALID, "this engine already shards its rows over its own devices");
    return guarded(e, [&]() -> int {
        CU(cudaSetDevice(e->device));
        if (e->comm) { nccl().CommDestroy(e->comm); e->comm = nullptr; e->comm_ranks = false; }
        e->allreduce = fn;
        e->allreduce_user = user;
        e->rank = rank;
        e->world = fn ? world : 1;
        return compute_y_stats(e);
    });
}

int rr_engine_get_info(const rr_engine *e, rr_engine_info *info)
{
    if (!e || !info) return RR_ERR_INVALID;
    info->n = e->n_object;
    info->n_total = e->n_total;
    info->d = e->d;
    info->device = e->device;
    info->y_mean = e->y_mean;
    info->sst = e->sst;
    info->sm_count = e->sm_count;
    info->exact_max_n = e->exact_max_n;
    info->n_gpus = 1 + (int32_t)e->peers.size();
    info->world = e->world;
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
    return guarded(e, [&] { return score_batch_impl(e, b, res); });
}

int rr_classifier_metrics(rr_engine *e, const rr_batch *b, double *accuracy, double *log_loss, double *abs_loss)
{
    return guarded(e, [&] { return classifier_metrics_impl(e, b, accuracy, log_loss, abs_loss); });
}

int rr_predict(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts, int32_t n_consts,
               const double *X, int64_t n, int32_t d, double *out)
{
    return guarded(e, [&] { return predict_impl(e, code, code_len, consts, n_consts, X, n, d, false, false, out); });
}

int rr_predict_rowmajor(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts, int32_t n_consts,
                        const double *X, int64_t n, int32_t d, double *out)
{
    return guarded(e, [&] { return predict_impl(e, code, code_len, consts, n_consts, X, n, d, true, false, out); });
}

int rr_predict_proba_rowmajor(rr_engine *e, const uint32_t *code, int32_t code_len, const double *consts, int32_t n_consts,
                              const double *X, int64_t n, int32_t d, double *out)
{
    return guarded(e, [&] { return predict_impl(e, code, code_len, consts, n_consts, X, n, d, true, true, out); });
}

int rr_feature_r2(rr_engine *e, double *r2)
{
    return guarded(e, [&] { return feature_r2_impl(e, r2); });
}

int rr_engine_read_rows(rr_engine *e, int64_t row0, int64_t rows, double *Xout, double *yout)
{
    return guarded(e, [&] { return read_rows_impl(e, row0, rows, Xout, yout); });
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
