// Throughput microbenchmarks that decide the NTT arithmetic: 32x32->64 integer multiply-add
// (IMAD.WIDE), 64-bit mul.hi, FP64 FMA, and the two candidate modular multiplications.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

template <int MODE>
__global__ void k(u64 *out, int iters, u64 seed, double p, double pinv, u64 q, u64 wq)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    u64 a[8];
    double d[8];
    for (int i = 0; i < 8; i++)
    {
        a[i] = seed * (tid + 1) + i * 0x9E3779B97F4A7C15ull;
        d[i] = (double)((a[i] >> 14) % (u64)p) - p / 2;
    }
    const unsigned w0 = (unsigned)seed | 1, w1 = (unsigned)(seed >> 32) | 1;
    const double w = (double)(seed % (u64)p) - p / 2;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (MODE == 0)
            { // IMAD.WIDE.U32 with 64-bit accumulate
                a[i] += (u64)(unsigned)a[i] * w0;
                a[i] += (u64)(unsigned)(a[i] >> 32) * w1;
            }
            else if (MODE == 1)
            { // mul.hi.u64
                a[i] = __umul64hi(a[i], wq) + 1;
            }
            else if (MODE == 2)
            { // DFMA
                d[i] = fma(d[i], 1.0000001, 0.5);
                d[i] = fma(d[i], 0.9999999, -0.5);
            }
            else if (MODE == 3)
            { // Shoup lazy modmul (integer)
                u64 hi = __umul64hi(a[i], wq);
                a[i] = seed * a[i] - hi * q;
            }
            else if (MODE == 4)
            { // FP64 modmul, one reduction step (6 ops)
                double h = __dmul_rn(d[i], w);
                double l = __fma_rn(d[i], w, -h);
                double t = __dadd_rn(__dadd_rn(__dmul_rn(h, pinv), 6755399441055744.0), -6755399441055744.0);
                d[i] = __dadd_rn(__fma_rn(-t, p, h), l);
            }
        }
    }
    u64 s = 0;
    for (int i = 0; i < 8; i++)
    {
        s += a[i] + (u64)d[i];
    }
    out[tid] = s;
}

template <int MODE>
void run(const char *name, double ops_per_iter)
{
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    u64 *out;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    const double p = 2251799780917249.0;
    k<MODE><<<blocks, threads>>>(out, 16, 12345, p, 1.0 / p, 2251799780917249ull, 0x123456789abcdefull);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 12345, p, 1.0 / p, 2251799780917249ull, 0x123456789abcdefull);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double total = (double)blocks * threads * iters * 8 * ops_per_iter;
    printf("%-34s %8.3f ms  %8.2f Gop/s (thread-level)  %6.2f ops/clk/SM @1.965GHz\n", name, ms, total / ms / 1e6,
           total / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main()
{
    run<0>("IMAD.WIDE.U32 (64-bit acc)", 2);
    run<1>("mul.hi.u64", 1);
    run<2>("DFMA", 2);
    run<3>("Shoup lazy modmul (int64)", 1);
    run<4>("FP64 modmul (6 DP ops)", 1);
    return 0;
}
