// FP64 issue-rate micro-benchmark for sm_100a: cycles per DMMA (mma.sync.m8n8k4.f64) and per DFMA warp instruction on one
// SM, as a function of the number of resident warps.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(double *out, long long *ticks, int iters)
{
    double acc[NACC][2];
    for (int i = 0; i < NACC; i++) acc[i][0] = acc[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(acc[i], a, b);
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) ticks[0] = t1 - t0;
}
template <int NACC>
__global__ void k_dfma(double *out, long long *ticks, int iters)
{
    double acc[NACC];
    for (int i = 0; i < NACC; i++) acc[i] = i;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) ticks[0] = t1 - t0;
}
int main()
{
    double *out;
    long long *tk, h;
    cudaMalloc(&out, 8 * 1024);
    cudaMalloc(&tk, 8);
    const int iters = 2000;
    for (int nt : {32, 128, 256, 512, 1024}) {
        k_dmma<8><<<1, nt>>>(out, tk, iters);
        cudaMemcpy(&h, tk, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / (iters * 8.0);  // cycles per DMMA of ONE warp
        printf("DMMA  threads %4d: %.2f clk per warp-DMMA in a warp's stream; SM rate = %.1f FMA/clk\n", nt, per, (nt / 32) * 256.0 / per);
        k_dfma<8><<<1, nt>>>(out, tk, iters);
        cudaMemcpy(&h, tk, 8, cudaMemcpyDeviceToHost);
        const double per2 = (double)h / (iters * 8.0);
        printf("DFMA  threads %4d: %.2f clk per warp-DFMA in a warp's stream; SM rate = %.1f FMA/clk\n", nt, per2, (nt / 32) * 32.0 / per2);
    }
    // latency: one dependent chain
    k_dmma<1><<<1, 32>>>(out, tk, iters);
    cudaMemcpy(&h, tk, 8, cudaMemcpyDeviceToHost);
    printf("DMMA dependent-chain latency: %.1f clk\n", (double)h / iters);
    k_dfma<1><<<1, 32>>>(out, tk, iters);
    cudaMemcpy(&h, tk, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent-chain latency: %.1f clk\n", (double)h / iters);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
