// Micro-benchmark of the resident path's sweep phase: d[j][f] = sum_i X[i][j] R[i][f] for an L2-resident row-major X
// (n x p) and FT chain slots, with several thread mappings.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
// Usage: sweep_micro [n p]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int NT = 512;
__device__ __forceinline__ double2 ldg2(const double *p) { double2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v; }

__device__ __forceinline__ double ldg1(const double *p) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
// V0: thread = (column pair, row phase), 2 columns x FT chains, R rows broadcast from smem with LDS.128
// MODE: 0 full, 1 no LDS (R constant), 2 no LDG (x constant)
template <int FT, int MODE, int U>
__global__ void __launch_bounds__(NT, 1) v0(const double *X, long long ldx, int n, int p, const double *R, double *out, int nsweep, int iters)
{
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x;
    const int P2 = (p + 1) / 2, WpA = (P2 + nsweep - 1) / nsweep;
    const int q0 = min(P2, (int)blockIdx.x * WpA), Wp = min(P2, q0 + WpA) - q0;
    if (Wp <= 0) return;
    const int RP = min(NT / Wp, 32);
    const int cp = tid % Wp, rp = tid / Wp;
    for (int e = tid; e < n * FT; e += NT) sm[e] = R[e];
    __syncthreads();
    double a0[FT], a1[FT];
#pragma unroll
    for (int f = 0; f < FT; f++) a0[f] = a1[f] = 0.0;
    for (int itr = 0; itr < iters; itr++) {
    asm volatile("" ::: "memory");
    __syncthreads();
    if (rp < RP) {
        const double *xp = X + 2 * (size_t)(q0 + cp);
        for (int i0 = rp; i0 < n; i0 += RP * U) {
            double2 xv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * RP;
                if (MODE == 2) xv[u] = make_double2(1.0 + i, 2.0);
                else xv[u] = i < n ? ldg2(xp + (size_t)i * ldx) : make_double2(0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * RP;
                if (i < n) {
                    const double *r = sm + (size_t)i * FT;
#pragma unroll
                    for (int f = 0; f < FT; f += 2) {
                        double2 g;
                        if (MODE == 1) g = make_double2(1.0 + f, 0.5);
                        else g = *reinterpret_cast<const double2 *>(r + f);
                        a0[f] = fma(xv[u].x, g.x, a0[f]);
                        a1[f] = fma(xv[u].y, g.x, a1[f]);
                        a0[f + 1] = fma(xv[u].x, g.y, a0[f + 1]);
                        a1[f + 1] = fma(xv[u].y, g.y, a1[f + 1]);
                    }
                }
            }
        }
    }
    }
    double s = 0;
#pragma unroll
    for (int f = 0; f < FT; f++) s += a0[f] + a1[f];
    if (s == 123.456) out[tid] = s;
}

// V3: thread = (group of 4 columns, chain half, row phase): 4 columns x FT/2 chains; the two chain halves of a column
// group sit in the same warp (lanes l and l+16) so their x loads coalesce.
template <int FT, int U>
__global__ void __launch_bounds__(NT, 1) v3(const double *X, long long ldx, int n, int p, const double *R, double *out, int nsweep, int iters)
{
    constexpr int FH = FT / 2;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int P4 = (p + 3) / 4, WgA = (P4 + nsweep - 1) / nsweep;
    const int g0 = min(P4, (int)blockIdx.x * WgA), Wg = min(P4, g0 + WgA) - g0;  // column groups of this CTA
    if (Wg <= 0) return;
    for (int e = tid; e < n * FT; e += NT) sm[e] = R[e];
    __syncthreads();
    // items = (group, row phase); 16 items per warp, 256 items per CTA
    const int item = wid * 16 + (lane & 15), half = lane >> 4;
    const int RP = min(256 / Wg, 32);
    const int cg = item % Wg, rp = item / Wg;
    double acc[4][FH];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int f = 0; f < FH; f++) acc[c][f] = 0.0;
    for (int itr = 0; itr < iters; itr++) {
    asm volatile("" ::: "memory");
    __syncthreads();
    if (rp < RP) {
        const double *xp = X + 4 * (size_t)(g0 + cg);
        for (int i0 = rp; i0 < n; i0 += RP * U) {
            double2 xa[U], xb[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * RP;
                xa[u] = i < n ? ldg2(xp + (size_t)i * ldx) : make_double2(0, 0);
                xb[u] = i < n ? ldg2(xp + (size_t)i * ldx + 2) : make_double2(0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * RP;
                if (i < n) {
                    const double *r = sm + (size_t)i * FT + half * FH;
                    double g[FH];
#pragma unroll
                    for (int f = 0; f < FH; f += 2) {
                        if (f + 1 < FH && (FH % 2 == 0)) {
                            const double2 t = *reinterpret_cast<const double2 *>(r + f);
                            g[f] = t.x;
                            g[f + 1] = t.y;
                        } else {
                            g[f] = r[f];
                            if (f + 1 < FH) g[f + 1] = r[f + 1];
                        }
                    }
#pragma unroll
                    for (int f = 0; f < FH; f++) {
                        acc[0][f] = fma(xa[u].x, g[f], acc[0][f]);
                        acc[1][f] = fma(xa[u].y, g[f], acc[1][f]);
                        acc[2][f] = fma(xb[u].x, g[f], acc[2][f]);
                        acc[3][f] = fma(xb[u].y, g[f], acc[3][f]);
                    }
                }
            }
        }
    }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int f = 0; f < FH; f++) s += acc[c][f];
    if (s == 123.456) out[tid] = s;
}

// V4: FP64 tensor cores.  warp tile = 8 columns x (8*NTL chains), k = 4 rows per mma; lanes: A[m = lane/4][k = lane%4],
// B[k = lane%4][n = lane/4].  Each warp walks a row phase.
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int NTL, int U>
__global__ void __launch_bounds__(NT, 1) v4(const double *X, long long ldx, int n, int p, const double *R, int FS, double *out, int nsweep, int iters)
{
    extern __shared__ __align__(16) double sm[];  // R padded to [n4][8*NTL]
    constexpr int FP = 8 * NTL;
    constexpr int FPS = NTL == 2 ? 20 : 12;  // row stride == 4 (mod 16) doubles resp. 12: fragment loads are bank-conflict free
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int P8 = (p + 7) / 8, WgA = (P8 + nsweep - 1) / nsweep;
    const int g0 = min(P8, (int)blockIdx.x * WgA), Wg = min(P8, g0 + WgA) - g0;  // 8-column groups
    if (Wg <= 0) return;
    const int n4 = (n + 3) & ~3;
    for (int e = tid; e < n4 * FP; e += NT) {
        const int i = e / FP, f = e % FP;
        sm[(size_t)i * FPS + f] = (i < n && f < FS) ? R[(size_t)i * FS + f] : 0.0;
    }
    __syncthreads();
    // 16 warps: warp = (column group cg, row phase rp)
    const int RP = 16 / Wg > 0 ? 16 / Wg : 1;
    const int cg = wid % Wg, rp = wid / Wg;
    double acc[NTL][2];
#pragma unroll
    for (int t = 0; t < NTL; t++) acc[t][0] = acc[t][1] = 0.0;
    for (int itr = 0; itr < iters; itr++) {
    asm volatile("" ::: "memory");
    __syncthreads();
    if (rp < RP && wid < Wg * RP) {
        const int m = lane >> 2, k = lane & 3;
        const double *xp = X + 8 * (size_t)(g0 + cg) + m;
        const int nk = n4 / 4;  // k-steps
        for (int s0 = rp; s0 < nk; s0 += RP * U) {
            double a[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = (s0 + u * RP) * 4 + k;
                a[u] = i < n ? ldg1(xp + (size_t)i * ldx) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int s = s0 + u * RP;
                if (s < nk) {
                    const double *r = sm + (size_t)(s * 4 + k) * FPS + m;
#pragma unroll
                    for (int t = 0; t < NTL; t++) dmma(acc[t], a[u], r[8 * t]);
                }
            }
        }
    }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < NTL; t++) s += acc[t][0] + acc[t][1];
    if (s == 123.456) out[tid] = s;
}

template <class F>
static float timeit(F launch, int reps)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 5; i++) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; i++) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms * 1000.f / reps;
}

int main(int argc, char **argv)
{
    const int n = argc > 2 ? atoi(argv[1]) : 1000, p = argc > 2 ? atoi(argv[2]) : 5000;
    const long long ldx = (p + 1) & ~1;
    double *X, *R, *out;
    CK(cudaMalloc(&X, (size_t)n * ldx * 8 + 64));
    CK(cudaMalloc(&R, (size_t)(n + 4) * 16 * 8));
    CK(cudaMalloc(&out, 4096 * 8));
    std::vector<double> h((size_t)n * ldx, 0.5);
    CK(cudaMemcpy(X, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    std::vector<double> r((size_t)(n + 4) * 16, 0.25);
    CK(cudaMemcpy(R, r.data(), r.size() * 8, cudaMemcpyHostToDevice));
    const int reps = 20;
    const int sweepers[] = {137, 142, 148};
    for (int ns : sweepers) {
        printf("---- n=%d p=%d sweepers=%d (X = %.1f MB)\n", n, p, ns, n * ldx * 8 / 1e6);
#define RUN0(FT, MODE, U, name)                                                                                   \
    {                                                                                                             \
        const size_t smem = (size_t)n * FT * 8;                                                                   \
        CK(cudaFuncSetAttribute(v0<FT, MODE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        float u1 = timeit([&] { v0<FT, MODE, U><<<ns, NT, smem>>>(X, ldx, n, p, R, out, ns, 10); }, reps);        \
        float u2 = timeit([&] { v0<FT, MODE, U><<<ns, NT, smem>>>(X, ldx, n, p, R, out, ns, 60); }, reps);        \
        float us = (u2 - u1) / 50.f;                                                                              \
        printf("%-44s %7.2f us  (%.2f TB/s of X)\n", name, us, n * ldx * 8 / us / 1e6);                           \
    }
        RUN0(12, 0, 8, "v0 FT=12 2col x 12ch, U=8");
        RUN0(12, 0, 4, "v0 FT=12 U=4");
        RUN0(12, 1, 8, "v0 FT=12 no LDS");
        RUN0(12, 2, 8, "v0 FT=12 no LDG");
        RUN0(6, 0, 8, "v0 FT=6 U=8");
        RUN0(6, 1, 8, "v0 FT=6 no LDS");
        RUN0(6, 2, 8, "v0 FT=6 no LDG");
        RUN0(2, 0, 8, "v0 FT=2 U=8");
#define RUN3(FT, U, name)                                                                                         \
    {                                                                                                             \
        const size_t smem = (size_t)n * FT * 8;                                                                   \
        CK(cudaFuncSetAttribute(v3<FT, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        float u1 = timeit([&] { v3<FT, U><<<ns, NT, smem>>>(X, ldx, n, p, R, out, ns, 10); }, reps);              \
        float u2 = timeit([&] { v3<FT, U><<<ns, NT, smem>>>(X, ldx, n, p, R, out, ns, 60); }, reps);              \
        float us = (u2 - u1) / 50.f;                                                                              \
        printf("%-44s %7.2f us  (%.2f TB/s of X)\n", name, us, n * ldx * 8 / us / 1e6);                           \
    }
        RUN3(12, 4, "v3 FT=12 4col x 6ch (half-warp split), U=4");
        RUN3(12, 8, "v3 FT=12 U=8");
        RUN3(6, 4, "v3 FT=6 4col x 3ch, U=4");
        RUN3(16, 4, "v3 FT=16 4col x 8ch, U=4");
#define RUN4(NTL, U, FS, name)                                                                                    \
    {                                                                                                             \
        const size_t smem = (size_t)((n + 3) & ~3) * 20 * 8;                                                    \
        CK(cudaFuncSetAttribute(v4<NTL, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        float u1 = timeit([&] { v4<NTL, U><<<ns, NT, smem>>>(X, ldx, n, p, R, FS, out, ns, 10); }, reps);         \
        float u2 = timeit([&] { v4<NTL, U><<<ns, NT, smem>>>(X, ldx, n, p, R, FS, out, ns, 60); }, reps);         \
        float us = (u2 - u1) / 50.f;                                                                              \
        printf("%-44s %7.2f us  (%.2f TB/s of X)\n", name, us, n * ldx * 8 / us / 1e6);                           \
    }
        RUN4(2, 8, 12, "v4 DMMA 16 chain slots, U=8");
        RUN4(2, 4, 12, "v4 DMMA 16 chain slots, U=4");
        RUN4(1, 8, 6, "v4 DMMA 8 chain slots, U=8");
    }
    return 0;
}
