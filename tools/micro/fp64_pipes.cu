// Microbenchmark: FP64 throughput of DFMA vs DMMA (mma.sync f64) shapes on sm_100a, and the latency of
// the indirect branch (brx.idx) that the interpreter's dispatcher pays. Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_dfma(double *out, int iters, double a, double b)
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

// m8n8k4: A 1 reg, B 1 reg, C 2 regs per thread; NCH independent accumulator chains
template <int NCH> __global__ void k_dmma884(double *out, int iters, double a, double b)
{
    double c[NCH][2];
#pragma unroll
    for (int j = 0; j < NCH; ++j) c[j][0] = c[j][1] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// m16n8k4: A 2 regs, B 1, C 4
template <int NCH> __global__ void k_dmma1684(double *out, int iters, double a, double b)
{
    double c[NCH][4];
#pragma unroll
    for (int j = 0; j < NCH; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// m16n8k16: A 8 regs, B 4, C 4
template <int NCH> __global__ void k_dmma16816(double *out, int iters, double a, double b)
{
    double c[NCH][4];
#pragma unroll
    for (int j = 0; j < NCH; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K> double time_kernel(K launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best * 1e-3;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sm = p.multiProcessorCount;
    printf("device %s, %d SMs\n", p.name, sm);
    double *out; CK(cudaMalloc(&out, (size_t)sm * 16 * 1024 * 8));
    const int iters = 2048;
    for (int wps : {4, 8, 16, 32}) {  // warps per SM
        const int blocks = sm, threads = wps * 32;
        double t = time_kernel([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fma = (double)blocks * threads * iters * 128.0;
        printf("warps/SM %2d  DFMA            %7.2f TFMA/s (%6.2f TFLOP/s)\n", wps, fma / t * 1e-12, 2 * fma / t * 1e-12);
#define RUN(KERN, NCH, FMA_PER_MMA, U, NAME)                                                                     \
        {                                                                                                        \
            double tt = time_kernel([&] { KERN<NCH><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });       \
            double f = (double)blocks * wps * iters * (double)(U) * NCH * (FMA_PER_MMA);                         \
            printf("warps/SM %2d  %-12s x%d %7.2f TFMA/s (%6.2f TFLOP/s)\n", wps, NAME, NCH, f / tt * 1e-12, 2 * f / tt * 1e-12); \
        }
        RUN(k_dmma884, 1, 256.0, 8, "m8n8k4")
        RUN(k_dmma884, 4, 256.0, 8, "m8n8k4")
        RUN(k_dmma884, 8, 256.0, 8, "m8n8k4")
        RUN(k_dmma1684, 4, 512.0, 8, "m16n8k4")
        RUN(k_dmma16816, 2, 2048.0, 4, "m16n8k16")
        RUN(k_dmma16816, 4, 2048.0, 4, "m16n8k16")
    }
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}
