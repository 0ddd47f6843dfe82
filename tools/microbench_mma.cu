// Throughput probes for the ct-pt GEMM design decision (north_star (3)): legacy tensor-core paths that
// compile for sm_100a — mma.sync int8 (m16n8k32, u8 x u8 -> s32) and FP64 (m8n8k4) — next to DFMA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_mma.bin tools/microbench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_imma(int iters, int *out)
{
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    int c[8][4] = {};
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    int s = 0;
    for (int u = 0; u < 8; u++)
        s += c[u][0] + c[u][1] + c[u][2] + c[u][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(int iters, double *out)
{
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    double c[8][2] = {};
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[u][0]), "+d"(c[u][1])
                         : "d"(a), "d"(b));
        }
    }
    double s = 0;
    for (int u = 0; u < 8; u++)
        s += c[u][0] + c[u][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(int iters, double *out)
{
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    double c[8] = { 1, 2, 3, 4, 5, 6, 7, 8 };
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int u = 0; u < 8; u++)
            c[u] = fma(c[u], b, a);
    }
    double s = 0;
    for (int u = 0; u < 8; u++)
        s += c[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    const int ctas = 148 * 8, thr = 256, iters = 4096;
    void *buf;
    cudaMalloc(&buf, (size_t)ctas * thr * 8);
    float ms = timeit([&] { k_imma<<<ctas, thr>>>(iters, (int *)buf); });
    double macs = (double)ctas * (thr / 32) * iters * 8 * 16 * 8 * 32;
    printf("{\"op\": \"mma.sync m16n8k32 u8\", \"ms\": %.3f, \"int8_TMAC_per_s\": %.1f, \"TOPS\": %.1f}\n", ms, macs / ms / 1e9, 2 * macs / ms / 1e9);
    ms = timeit([&] { k_dmma<<<ctas, thr>>>(iters, (double *)buf); });
    macs = (double)ctas * (thr / 32) * iters * 8 * 8 * 8 * 4;
    printf("{\"op\": \"mma.sync m8n8k4 f64\", \"ms\": %.3f, \"fp64_TFMA_per_s\": %.2f, \"TFLOPS\": %.1f}\n", ms, macs / ms / 1e9, 2 * macs / ms / 1e9);
    ms = timeit([&] { k_dfma<<<ctas, thr>>>(iters, (double *)buf); });
    macs = (double)ctas * thr * iters * 8;
    printf("{\"op\": \"DFMA\", \"ms\": %.3f, \"fp64_TFMA_per_s\": %.2f, \"TFLOPS\": %.1f}\n", ms, macs / ms / 1e9, 2 * macs / ms / 1e9);
    return 0;
}
