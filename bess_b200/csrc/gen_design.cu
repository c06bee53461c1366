// Device-side design generator: the x of the reference's gen.data (R/R/gen.data.R:110-118, cortype 1; cortype 2 and 3 at
// the end of the file) drawn directly
// in HBM, row-major n x p fp64 -- the layout pywrap_bess takes -- so a benchmark design never crosses PCIe.
//   x_i ~ MVN(0, Sigma),  Sigma_jk = rho^|j-k|          (rho = 0, the default of gen.data: iid N(0,1))
// A stationary AR(1) process along the columns has exactly that covariance:
//   x_ij = rho * x_i,j-1 + sqrt(1 - rho^2) * z_ij,   z iid N(0,1).
// z_ij is a pure function of (seed, i, j): Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
// SC'11) on the counter (j as a signed 64-bit integer, i, 0) with the seed as key, two 53-bit uniforms, Box-Muller
// (cosine branch).  R's Mersenne-Twister stream cannot be reproduced without R; what is kept is the
// distribution; the parity tests check the stream against an independent numpy restatement (tests/test_gen_design.py).
//
// One warp walks one (row, column segment): 32 columns per step, the AR(1) recurrence inside the step is a 5-stage
// decayed inclusive scan over the lanes (v_j += rho^d * v_{j-d}), the carry from the previous step enters as
// carry * rho^(lane+1).  A segment starts `kwarm` columns early with a zero carry: after kwarm columns the missing
// tail of the stationary sum is below rho^kwarm <= 2^-60, so segments are independent of each other and of where the
// design is cut.  Writes are coalesced (a warp writes 256 contiguous bytes per step).
#include <cmath>
#include <cstdint>

#include "device_utils.cuh"
#include "kernels.cuh"

namespace bess {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// z_ij: standard normal, a pure function of (seed, i, j)
__device__ __forceinline__ double normal_at(unsigned long long seed, int i, long long j)
{
    uint32_t r[4];
    philox4x32_10((uint32_t)((unsigned long long)j & 0xffffffffull), (uint32_t)((unsigned long long)j >> 32), (uint32_t)i, 0u,
                  (uint32_t)(seed & 0xffffffffull), (uint32_t)(seed >> 32), r);
    const unsigned long long a = ((unsigned long long)r[0] << 32) | r[1];
    const unsigned long long b = ((unsigned long long)r[2] << 32) | r[3];
    const double u1 = ((double)(a >> 11) + 0.5) * 0x1.0p-53;  // (0, 1)
    const double u2 = ((double)(b >> 11) + 0.5) * 0x1.0p-53;
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__global__ void __launch_bounds__(128) gen_design_kernel(double *X, long long ld, int n, long long p, double rho,
                                                         unsigned long long seed, long long seg, int kwarm, long long nseg)
{
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= (long long)n * nseg) return;
    const int i = (int)(w / nseg);
    const long long j_begin = (w % nseg) * seg;
    const long long j_end = min(p, j_begin + seg);
    const double s = sqrt(1.0 - rho * rho);
    const double r1 = rho, r2 = r1 * r1, r4 = r2 * r2, r8 = r4 * r4, r16 = r8 * r8;
    double rl = rho;  // rho^(lane + 1)
    for (int q = 0; q < lane; q++) rl *= rho;
    double carry = 0.0;
    double *row = X + (size_t)i * ld;
    for (long long jb = j_begin - kwarm; jb < j_end; jb += 32) {
        const long long j = jb + lane;
        double v = s * normal_at(seed, i, j);
        if (rho != 0.0) {
            double t;
            t = __shfl_up_sync(0xffffffffu, v, 1);  if (lane >= 1) v = fma(r1, t, v);
            t = __shfl_up_sync(0xffffffffu, v, 2);  if (lane >= 2) v = fma(r2, t, v);
            t = __shfl_up_sync(0xffffffffu, v, 4);  if (lane >= 4) v = fma(r4, t, v);
            t = __shfl_up_sync(0xffffffffu, v, 8);  if (lane >= 8) v = fma(r8, t, v);
            t = __shfl_up_sync(0xffffffffu, v, 16); if (lane >= 16) v = fma(r16, t, v);
            v = fma(carry, rl, v);
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
        if (j >= j_begin && j < j_end) row[j] = v;
    }
}

void launch_gen_design(double *X, long long ld, int n, long long p, double rho, unsigned long long seed, cudaStream_t st)
{
    if (!(rho > -1.0 && rho < 1.0)) throw EngineError{"gen_design: rho must be in (-1, 1)"};
    int kwarm = 0;
    if (rho != 0.0) {
        const double k = std::ceil(60.0 * std::log(2.0) / -std::log(std::fabs(rho)));
        if (k > 1.0e6) throw EngineError{"gen_design: |rho| too close to 1"};
        kwarm = ((int)k + 31) & ~31;
    }
    // segments long enough to amortise the warm-up, short enough to fill the machine
    long long seg = std::max<long long>(4096, 16LL * kwarm);
    seg = (seg + 31) & ~31LL;
    const long long nseg = (p + seg - 1) / seg;
    const long long warps = (long long)n * nseg;
    gen_design_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(X, ld, n, p, rho, seed, seg, kwarm, nseg);
    CUDA_CHECK(cudaGetLastError());
}

// ---- cortype 2 (R/R/gen.data.R:114-116): Sigma = rho + (1 - rho) I, i.e. one common factor per row:
//   x_ij = sqrt(rho) * f_i + sqrt(1 - rho) * z_ij,  f_i = z_i,-1 (a column index no design column uses)
__global__ void gen_design_exch_kernel(double *X, long long ld, int n, long long p, double rho, unsigned long long seed)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= p) return;
    const double f = normal_at(seed, i, -1);
    X[(size_t)i * ld + j] = sqrt(rho) * f + sqrt(1.0 - rho) * normal_at(seed, i, j);
}

// ---- cortype 3 (R/R/gen.data.R:167-181; the design of python/bess/gen_data.py:25-30): X iid N(0,1), columns centred and
// scaled to norm sqrt(n), then x_j = X_j + rho (X_{j-1} + X_{j+1}) for 1 <= j <= p-2 and x_j = X_j at both ends.
// Pass 1 (column statistics; thread per column, the rows of a warp's 32 columns are coalesced): mean, then the norm of
// the centred column.  Pass 2 (one CTA per row, tiles left to right, IN PLACE): the left neighbour of a tile's first
// column has already been overwritten by this CTA, so its normalised value is carried over; the right halo has not.
__global__ void band_stats_kernel(const double *X, long long ld, int n, long long p, double *mean, double *scale)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    double s = 0.0;
    for (int i = 0; i < n; i++) s += X[(size_t)i * ld + j];
    const double mu = s / (double)n;
    double q = 0.0;
    for (int i = 0; i < n; i++) {
        const double t = X[(size_t)i * ld + j] - mu;
        q = fma(t, t, q);
    }
    mean[j] = mu;
    scale[j] = sqrt((double)n) / sqrt(q);
}
constexpr int BAND_TB = 1024;
__global__ void __launch_bounds__(BAND_TB) band_mix_kernel(double *X, long long ld, long long p, double rho, const double *mean,
                                                           const double *scale)
{
    __shared__ double t[BAND_TB + 2];
    double *row = X + (size_t)blockIdx.x * ld;
    const int tid = threadIdx.x;
    double carry = 0.0;  // normalised value of the column left of the tile
    for (long long j0 = 0; j0 < p; j0 += BAND_TB) {
        const long long j = j0 + tid;
        const double xn = j < p ? (row[j] - mean[j]) * scale[j] : 0.0;
        t[tid + 1] = xn;
        if (tid == 0) {
            t[0] = carry;
            const long long jr = j0 + BAND_TB;
            t[BAND_TB + 1] = jr < p ? (row[jr] - mean[jr]) * scale[jr] : 0.0;
        }
        __syncthreads();
        carry = t[BAND_TB];  // last column of this tile, before it is overwritten
        if (j < p) row[j] = (j >= 1 && j <= p - 2) ? xn + rho * (t[tid] + t[tid + 2]) : xn;
        __syncthreads();
    }
}

// cortype 1: AR(1) covariance rho^|j-k| (launch_gen_design); 2: exchangeable; 3: banded.  scratch: 2 * p doubles (cortype 3).
void launch_gen_design_cortype(double *X, long long ld, int n, long long p, double rho, unsigned long long seed, int cortype,
                               double *scratch, cudaStream_t st)
{
    if (cortype == 1) {
        launch_gen_design(X, ld, n, p, rho, seed, st);
    } else if (cortype == 2) {
        if (!(rho >= 0.0 && rho < 1.0)) throw EngineError{"gen_design: cortype 2 needs rho in [0, 1)"};
        dim3 grid((unsigned)((p + 255) / 256), (unsigned)n);
        gen_design_exch_kernel<<<grid, 256, 0, st>>>(X, ld, n, p, rho, seed);
        CUDA_CHECK(cudaGetLastError());
    } else if (cortype == 3) {
        if (p < 3) throw EngineError{"gen_design: cortype 3 needs p >= 3"};
        launch_gen_design(X, ld, n, p, 0.0, seed, st);
        band_stats_kernel<<<(unsigned)((p + 127) / 128), 128, 0, st>>>(X, ld, n, p, scratch, scratch + p);
        CUDA_CHECK(cudaGetLastError());
        band_mix_kernel<<<(unsigned)n, BAND_TB, 0, st>>>(X, ld, p, rho, scratch, scratch + p);
        CUDA_CHECK(cudaGetLastError());
    } else {
        throw EngineError{"gen_design: cortype must be 1, 2 or 3"};
    }
}

}  // namespace bess
