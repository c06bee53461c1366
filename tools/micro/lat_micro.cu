// Dependent-chain latencies on sm_100a: DFMA, DMUL, rsqrt(double), 1.0/x, sqrt, shfl of a double, LDS, clock overhead.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1;} } while (0)
__global__ void lat(double *out, long long *cyc, double a, double b)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = a + lane;
    sm[lane + 32] = b;
    __syncthreads();
    double x = a;
    long long t0, t1;
    const int N = 256;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, b, a);
    t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0) / N;
    out[0] = x;
    // DMUL chain
    x = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * b;
    t1 = clock64();
    if (lane == 0) cyc[1] = (t1 - t0) / N;
    out[1] = x;
    // rsqrt chain
    x = a + 3.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; i++) x = rsqrt(x) + 2.0;
    t1 = clock64();
    if (lane == 0) cyc[2] = (t1 - t0) / 64;
    out[2] = x;
    // 1/x chain
    x = a + 3.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; i++) x = 1.0 / x + 2.0;
    t1 = clock64();
    if (lane == 0) cyc[3] = (t1 - t0) / 64;
    out[3] = x;
    // sqrt chain
    x = a + 3.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; i++) x = sqrt(x) + 2.0;
    t1 = clock64();
    if (lane == 0) cyc[4] = (t1 - t0) / 64;
    out[4] = x;
    // shfl chain (double = two 32-bit shuffles)
    x = a + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64();
    if (lane == 0) cyc[5] = (t1 - t0) / N;
    out[5] = x;
    // dependent LDS chain (pointer chasing through shared memory)
    __shared__ int nxt[32];
    nxt[lane] = (lane * 7 + 3) & 31;
    __syncthreads();
    int p = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = nxt[p];
    t1 = clock64();
    if (lane == 0) cyc[6] = (t1 - t0) / N;
    out[6] = p;
    // float rsqrt + one Newton step in double
    x = a + 3.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; i++) {
        double y = (double)rsqrtf((float)x);
        y = y * fma(-0.5 * x, y * y, 1.5);
        x = y + 2.0;
    }
    t1 = clock64();
    if (lane == 0) cyc[7] = (t1 - t0) / 64;
    out[7] = x;
    // __syncthreads cost with 512 threads is measured in the second kernel
}
__global__ void synclat(long long *cyc)
{
    long long t0 = clock64();
    for (int i = 0; i < 64; i++) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = (t1 - t0) / 64;
}
// global round trip: dependent ld.cg chain through an L2-resident array
__global__ void glat(const int *nxt, long long *cyc, int *out)
{
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < 64; i++) p = __ldcg(nxt + p);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[9] = (t1 - t0) / 64;
    out[threadIdx.x] = p;
}
int main()
{
    double *out; long long *cyc; int *nxt, *o2;
    CK(cudaMalloc(&out, 64 * 8)); CK(cudaMalloc(&cyc, 16 * 8)); CK(cudaMalloc(&nxt, 4096 * 4)); CK(cudaMalloc(&o2, 4096 * 4));
    int h[4096]; for (int i = 0; i < 4096; i++) h[i] = (i * 577 + 1234) & 4095;
    CK(cudaMemcpy(nxt, h, sizeof(h), cudaMemcpyHostToDevice));
    for (int rep = 0; rep < 2; rep++) {
        lat<<<1, 32>>>(out, cyc, 1.000001, 0.999999);
        synclat<<<1, 512>>>(cyc);
        glat<<<1, 32>>>(nxt, cyc, o2);
        CK(cudaDeviceSynchronize());
    }
    long long hc[16];
    CK(cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
    const char *nm[] = {"DFMA", "DMUL", "rsqrt(double)", "1.0/x", "sqrt(double)", "shfl(double)", "LDS chase", "rsqrtf+Newton", "__syncthreads(512)", "ld.cg chase (L2)"};
    for (int i = 0; i < 10; i++) printf("%-20s %lld cycles\n", nm[i], hc[i]);
    return 0;
}
