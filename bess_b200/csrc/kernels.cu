// Hand-written sm_100a kernels for the BeSS primal-dual active-set hot path.
//
//   dual_sweep_kernel   d = X^T g, h = (X.X)^T w (+ Cox risk-set suffix term) for every chain of a batch in ONE
//                       pass over X.  X is row-major n x p (the layout pywrap_bess hands over,
//                       /root/reference/src/utilities.cpp:13-25), so a thread owns 1-2 adjacent columns, walks the
//                       rows with coalesced 16-byte loads and needs no cross-lane reduction; the per-row gradient
//                       vectors of all chains are staged into shared memory with 1-D bulk TMA copies
//                       (cp.async.bulk + mbarrier), double buffered.
//                       Replaces the GEMVs + per-column 1x1 sqrt()/ldlt() objects of
//                       Algorithm.h:1097-1129 (Lm), :1206-1263 (Logistic), :1324-1367 (Poisson), :1569-1640 (Cox).
//   finish_kernel       reduces the row-split partials and applies the splicing sacrifice (same citations).
//   topk_slices_kernel  exact top-k by MSB-first radix select on the fp64 bit patterns held in shared memory,
//                       ordered compaction => ascending indices (utilities.cpp:179-199 max_k + slice_assignment).
//   chain_begin/chain_fit  one CTA per chain: gather X_A (utilities.cpp:132-140), active-set fit
//                       (Algorithm.h:1131-1135 / 1148-1204 / 1273-1322 / 1377-1490), cycle test (Algorithm.h:164-170),
//                       and the gradient vectors of the next sweep.
//   loss_kernel         Metric.h train_loss / fold test losses.
#include "kernels.cuh"

#include <cfloat>
#include <cmath>
#include <cstdio>

namespace bess {

#define CUDA_CHECK(x)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (x);                                                                           \
        if (e_ != cudaSuccess) {                                                                        \
            throw EngineError{std::string(#x) + ": " + cudaGetErrorString(e_)};                         \
        }                                                                                               \
    } while (0)

// =====================================================================================================
// small device helpers
// =====================================================================================================
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum; every thread gets the result.  sh: >= 33 doubles.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sh)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < NT / 32 ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

// block exclusive scan of one value per thread (thread order); returns exclusive prefix, *total = block total.
// The exclusive value is obtained by SHIFTING the inclusive scan, never by subtracting the thread's own value:
// "inclusive - own" cancels catastrophically when one term dwarfs the prefix (Cox risk sets span e^+-30).
template <int NT>
__device__ __forceinline__ double block_excl_scan(double v, double *sh, double *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    double excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 0.0;
    __syncthreads();
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        double w = lane < NT / 32 ? sh[lane] : 0.0;
        double winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        double wex = __shfl_up_sync(0xffffffffu, winc, 1);
        if (lane == 0) wex = 0.0;
        if (lane < NT / 32) sh[lane] = wex;  // exclusive warp offsets
        if (lane == 31) sh[32] = winc;
    }
    __syncthreads();
    *total = sh[32];
    return sh[wid] + excl;
}

// In-place inclusive scans over v[0..nr): each thread owns a contiguous chunk.
template <int NT>
__device__ void block_prefix_scan(double *v, int nr, double *sh)
{
    const int per = (nr + NT - 1) / NT;
    const int b = min(nr, (int)threadIdx.x * per), e = min(nr, b + per);
    double s = 0.0;
    for (int i = b; i < e; i++) s += v[i];
    double tot;
    double run = block_excl_scan<NT>(s, sh, &tot);
    for (int i = b; i < e; i++) {
        run += v[i];
        v[i] = run;
    }
    __syncthreads();
}
// suffix: v[i] <- sum_{k >= i} v[k]
template <int NT>
__device__ void block_suffix_scan(double *v, int nr, double *sh)
{
    const int per = (nr + NT - 1) / NT;
    // thread t owns the chunk counted from the END so that thread order == scan order
    const int e = max(0, nr - (int)threadIdx.x * per), b = max(0, e - per);
    double s = 0.0;
    for (int i = e - 1; i >= b; i--) s += v[i];
    double tot;
    double run = block_excl_scan<NT>(s, sh, &tot);
    for (int i = e - 1; i >= b; i--) {
        run += v[i];
        v[i] = run;
    }
    __syncthreads();
}

__device__ __forceinline__ double clampd(double v, double c) { return v > c ? c : (v < -c ? -c : v); }

// =====================================================================================================
// mbarrier + 1-D bulk TMA (cp.async.bulk) helpers
// =====================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// =====================================================================================================
// dual sweep
// =====================================================================================================
template <int MODE>
struct SweepTraits {
    static constexpr int NV = MODE == MODE_D ? 1 : (MODE == MODE_DH ? 2 : 4);  // staged vectors
    static constexpr int NQ = MODE == MODE_D ? 1 : (MODE == MODE_DH ? 2 : 5);  // partial outputs per column
};

template <int FT, int MODE, int CPT>
__global__ void __launch_bounds__(SWEEP_NT) dual_sweep_kernel(const Dev d)
{
    constexpr int NV = SweepTraits<MODE>::NV;
    constexpr int NQ = SweepTraits<MODE>::NQ;
    constexpr bool REV = (MODE == MODE_COX);  // risk-set suffix sums: walk rows from the last to the first
    constexpr int RC = SWEEP_RC;
    constexpr int U = 8;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tile = reinterpret_cast<double *>(smem_raw);  // [2][NV][RC][FT]
    __shared__ __align__(8) uint64_t bar[2];
    if (d.gate && *d.gate == 0) return;  // every chain of the batch already stopped (speculative launch)

    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const int r0 = s * d.rows_per_split;
    const int r1 = min(d.n, r0 + d.rows_per_split);
    const int nrows = r1 - r0;
    const int nchunks = (nrows + RC - 1) / RC;
    const long long j0 = ((long long)blockIdx.x * SWEEP_NT + tid) * CPT;
    const bool active = j0 < d.p;

    double accd[CPT][FT];
    double acch[MODE >= MODE_DH ? CPT : 1][MODE >= MODE_DH ? FT : 1];
    double s1[MODE == MODE_COX ? CPT : 1][MODE == MODE_COX ? FT : 1];
    double accA[MODE == MODE_COX ? CPT : 1][MODE == MODE_COX ? FT : 1];
    double accB[MODE == MODE_COX ? CPT : 1][MODE == MODE_COX ? FT : 1];
#pragma unroll
    for (int c = 0; c < CPT; c++)
#pragma unroll
        for (int f = 0; f < FT; f++) {
            accd[c][f] = 0.0;
            if constexpr (MODE >= MODE_DH) acch[c][f] = 0.0;
            if constexpr (MODE == MODE_COX) { s1[c][f] = 0.0; accA[c][f] = 0.0; accB[c][f] = 0.0; }
        }
    double c2acc = 0.0;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const double *vecs[4] = {d.G, d.W, d.TH, d.C2};
    auto issue = [&](int q, int stage) {
        const int c = REV ? (nchunks - 1 - q) : q;
        const int cr0 = r0 + c * RC;
        const int crows = min(RC, r1 - cr0);
        const uint32_t bytes = (uint32_t)(((crows + 1) & ~1) * FT * 8);  // vectors are padded to an even row count
        mbar_expect_tx(&bar[stage], bytes * NV);
#pragma unroll
        for (int v = 0; v < NV; v++)
            bulk_g2s(tile + (size_t)(stage * NV + v) * RC * FT, vecs[v] + (size_t)cr0 * FT, bytes, &bar[stage]);
    };
    if (tid == 0 && nchunks > 0) {
        issue(0, 0);
        if (nchunks > 1) issue(1, 1);
    }

    for (int q = 0; q < nchunks; q++) {
        const int stage = q & 1;
        mbar_wait(&bar[stage], (q >> 1) & 1);
        const int c = REV ? (nchunks - 1 - q) : q;
        const int cr0 = r0 + c * RC;
        const int crows = min(RC, r1 - cr0);
        const double *tg = tile + (size_t)(stage * NV + 0) * RC * FT;
        const double *tw = tile + (size_t)(stage * NV + (NV > 1 ? 1 : 0)) * RC * FT;
        const double *tt = tile + (size_t)(stage * NV + (NV > 2 ? 2 : 0)) * RC * FT;
        const double *tc = tile + (size_t)(stage * NV + (NV > 3 ? 3 : 0)) * RC * FT;
        if (active) {
            for (int rr = 0; rr < crows; rr += U) {
                double xv[U][CPT];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int lr = REV ? (crows - 1 - (rr + u)) : (rr + u);
                    if (rr + u < crows) {
                        const double *xp = d.X + (size_t)(cr0 + lr) * d.ldx + j0;
                        if constexpr (CPT == 2) {
                            const double2 t = __ldg(reinterpret_cast<const double2 *>(xp));
                            xv[u][0] = t.x;
                            xv[u][CPT - 1] = t.y;
                        } else {
                            xv[u][0] = __ldg(xp);
                        }
                    } else {
#pragma unroll
                        for (int cc = 0; cc < CPT; cc++) xv[u][cc] = 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (rr + u < crows) {
                        const int lr = REV ? (crows - 1 - (rr + u)) : (rr + u);
                        double xx[CPT];
#pragma unroll
                        for (int cc = 0; cc < CPT; cc++) xx[cc] = xv[u][cc] * xv[u][cc];
#pragma unroll
                        for (int f = 0; f < FT; f++) {
                            const double g = tg[lr * FT + f];
#pragma unroll
                            for (int cc = 0; cc < CPT; cc++) accd[cc][f] = fma(xv[u][cc], g, accd[cc][f]);
                            if constexpr (MODE >= MODE_DH) {
                                const double w = tw[lr * FT + f];
#pragma unroll
                                for (int cc = 0; cc < CPT; cc++) acch[cc][f] = fma(xx[cc], w, acch[cc][f]);
                            }
                            if constexpr (MODE == MODE_COX) {
                                const double th = tt[lr * FT + f];
                                const double c2 = tc[lr * FT + f];
#pragma unroll
                                for (int cc = 0; cc < CPT; cc++) {
                                    s1[cc][f] = fma(xv[u][cc], th, s1[cc][f]);
                                    const double t = c2 * s1[cc][f];
                                    accA[cc][f] = fma(t, s1[cc][f], accA[cc][f]);
                                    accB[cc][f] += t;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (MODE == MODE_COX && blockIdx.x == 0 && tid < FT) {
            for (int lr = 0; lr < crows; lr++) c2acc += tc[lr * FT + tid];
        }
        __syncthreads();
        if (tid == 0 && q + 2 < nchunks) issue(q + 2, stage);
    }

    if (MODE == MODE_COX && blockIdx.x == 0 && tid < FT) d.c2sum[s * FT + tid] = c2acc;
    if (!active) return;
    // partials: part[((s*NQ + q)*FT + f) * pstride + j]
#pragma unroll
    for (int f = 0; f < FT; f++) {
        auto put = [&](int q, double v0, double v1) {
            double *dst = d.part + ((size_t)(s * NQ + q) * FT + f) * d.pstride + j0;
            if (CPT == 2)
                *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
            else
                *dst = v0;
        };
        put(0, accd[0][f], accd[CPT - 1][f]);
        if constexpr (MODE >= MODE_DH) put(1, acch[0][f], acch[CPT - 1][f]);
        if constexpr (MODE == MODE_COX) {
            put(2, s1[0][f], s1[CPT - 1][f]);
            put(3, accA[0][f], accA[CPT - 1][f]);
            put(4, accB[0][f], accB[CPT - 1][f]);
        }
    }
}

template <int FT, int MODE, int CPT>
static void launch_sweep_t(const Dev &d, cudaStream_t st)
{
    constexpr int NV = SweepTraits<MODE>::NV;
    const size_t smem = (size_t)2 * NV * SWEEP_RC * FT * sizeof(double);
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(dual_sweep_kernel<FT, MODE, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        configured = true;
    }
    const long long cols_per_cta = (long long)SWEEP_NT * CPT;
    dim3 grid((unsigned)((d.p + cols_per_cta - 1) / cols_per_cta), (unsigned)d.S);
    dual_sweep_kernel<FT, MODE, CPT><<<grid, SWEEP_NT, smem, st>>>(d);
    CUDA_CHECK(cudaGetLastError());
}

template <int MODE>
static void launch_sweep_m(const Dev &d, cudaStream_t st)
{
    // columns per thread: 2 (16-byte loads) unless the accumulator set would not fit in registers
    switch (d.FS) {
        case 1: launch_sweep_t<1, MODE, 2>(d, st); break;
        case 2: launch_sweep_t<2, MODE, 2>(d, st); break;
        case 4: launch_sweep_t<4, MODE, MODE == MODE_COX ? 1 : 2>(d, st); break;
        case 6: launch_sweep_t<6, MODE, MODE == MODE_COX ? 1 : 2>(d, st); break;
        case 8: launch_sweep_t<8, MODE, MODE == MODE_D ? 2 : 1>(d, st); break;
        case 12: launch_sweep_t<12, MODE, MODE == MODE_D ? 2 : 1>(d, st); break;
        case 16: launch_sweep_t<16, MODE, 1>(d, st); break;
        default: throw EngineError{"dual sweep: unsupported chain tile FS=" + std::to_string(d.FS)};
    }
}

void launch_dual_sweep(const Dev &d, int mode, cudaStream_t st)
{
    if (mode == MODE_D) launch_sweep_m<MODE_D>(d, st);
    else if (mode == MODE_DH) launch_sweep_m<MODE_DH>(d, st);
    else launch_sweep_m<MODE_COX>(d, st);
}

// =====================================================================================================
// finish: reduce row-split partials, apply the sacrifice
// =====================================================================================================
template <int EPI>
__global__ void __launch_bounds__(256) finish_kernel(const Dev d, int mode, const BatchDesc b, double *raw_out)
{
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= d.p) return;
    if (d.gate && *d.gate == 0) return;
    const int c = b.chain[blockIdx.y];  // chain id == slot in the sweep vectors
    const int NQ = mode == MODE_D ? 1 : (mode == MODE_DH ? 2 : 5);
    const int FT = d.FS;
    double dsum = 0.0, hsum = 0.0, R = 0.0;
    if (mode != MODE_COX) {
        for (int s = 0; s < d.S; s++) {
            dsum += d.part[((size_t)(s * NQ + 0) * FT + c) * d.pstride + j];
            if (mode == MODE_DH) hsum += d.part[((size_t)(s * NQ + 1) * FT + c) * d.pstride + j];
        }
    } else {
        double carry = 0.0;
        for (int s = d.S - 1; s >= 0; s--) {
            const size_t base = ((size_t)(s * NQ) * FT + c) * d.pstride + j;
            const size_t qs = (size_t)FT * d.pstride;
            dsum += d.part[base];
            hsum += d.part[base + qs];
            const double Ts = d.part[base + 2 * qs], As = d.part[base + 3 * qs], Bs = d.part[base + 4 * qs];
            R += As + 2.0 * carry * Bs + carry * carry * d.c2sum[s * FT + c];
            carry += Ts;
        }
    }
    if (EPI == EPI_RAW) {
        raw_out[(size_t)(0 * FT + c) * d.pstride + j] = dsum;
        if (mode >= MODE_DH) raw_out[(size_t)(1 * FT + c) * d.pstride + j] = hsum;
        return;
    }
    double out;
    if (EPI == EPI_SCREEN_LM) {
        const double bq = dsum / hsum;  // one-column least squares (screening.cpp:46)
        out = bq * bq;
    } else {
        const double beta = d.betaD[(size_t)c * d.pstride + j];
        if (EPI == EPI_SACR_LM) {
            // Phi = sqrt(x_j.x_j / n) (utilities.cpp:142-151), invPhi = 1/Phi (:167-177); Algorithm.h:1116-1122
            const double phi = sqrt(d.xtx[(size_t)c * d.pstride + j] / (double)d.ntrain[c]);
            const double t = phi * beta + (1.0 / phi) * dsum;
            out = t * t;
        } else if (EPI == EPI_SACR_GLM) {
            const double phi = sqrt(hsum);  // Algorithm.h:1238-1257 / 1342-1361
            const double t = phi * beta + (1.0 / phi) * dsum;
            out = t * t;
        } else {
            // Algorithm.h:1626-1634: l1 = -dsum, l2 = hsum - R, d = -l1/l2, bd = |beta + d| * sqrt(l2)
            const double l2 = hsum - R;
            out = fabs(beta + dsum / l2) * sqrt(l2);
        }
    }
    d.bd[(size_t)c * d.pstride + j] = out;
}

void launch_finish(const Dev &d, int mode, int epi, const BatchDesc &b, double *raw_out, cudaStream_t st)
{
    dim3 grid((unsigned)((d.p + 255) / 256), (unsigned)b.nch);
    switch (epi) {
        case EPI_RAW: finish_kernel<EPI_RAW><<<grid, 256, 0, st>>>(d, mode, b, raw_out); break;
        case EPI_SACR_LM: finish_kernel<EPI_SACR_LM><<<grid, 256, 0, st>>>(d, mode, b, raw_out); break;
        case EPI_SACR_GLM: finish_kernel<EPI_SACR_GLM><<<grid, 256, 0, st>>>(d, mode, b, raw_out); break;
        case EPI_SACR_COX: finish_kernel<EPI_SACR_COX><<<grid, 256, 0, st>>>(d, mode, b, raw_out); break;
        default: finish_kernel<EPI_SCREEN_LM><<<grid, 256, 0, st>>>(d, mode, b, raw_out); break;
    }
    CUDA_CHECK(cudaGetLastError());
}

// always_select -> DBL_MAX (utilities.cpp:190-199)
__global__ void pin_kernel(double *vals, long long stride, int nch, const int *idx, int nidx, const int *gate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nidx * nch) return;
    if (gate && *gate == 0) return;
    vals[(size_t)(i / nidx) * stride + idx[i % nidx]] = DBL_MAX;
}
void launch_pin(const Dev &d, double *vals, long long stride, int nch, const int *idx, int nidx, cudaStream_t st)
{
    if (nidx <= 0) return;
    const int tot = nidx * nch;
    pin_kernel<<<(tot + 255) / 256, 256, 0, st>>>(vals, stride, nch, idx, nidx, d.gate);
    CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================================
// exact top-k: radix select in shared memory + ordered compaction
// =====================================================================================================
// One CTA selects the top min(k, len) keys of its slice and writes them IN INPUT ORDER (so candidate lists stay
// index-ascending through every stage and the final list needs no sort).  Total order: larger key first, then
// lower index first.  grid = (nslices, nchains).
__global__ void __launch_bounds__(TOPK_NT) topk_slices_kernel(const double *__restrict__ keys_in,
                                                              const int *__restrict__ idx_in, long long in_stride,
                                                              int n_in, int k, int slice_len, double *keys_out,
                                                              int *idx_out, long long out_stride, int *final_out,
                                                              int final_ld, int *tie, const int *gate)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    if (gate && *gate == 0) return;
    __shared__ int hist[256];
    __shared__ unsigned long long sh_prefix;
    __shared__ int sh_krem, sh_neq;
    __shared__ unsigned long long scan_sh[34];

    const int tid = threadIdx.x;
    const int s = blockIdx.x, f = blockIdx.y;
    const int b0 = s * slice_len;
    const int len = min(slice_len, n_in - b0);
    if (len <= 0) return;
    const int kk = min(k, len);
    const double *kin = keys_in + (size_t)f * in_stride + b0;
    const int *iin = idx_in ? idx_in + (size_t)f * in_stride + b0 : nullptr;
    const int out_per_slice = min(k, slice_len);

    for (int i = tid; i < len; i += TOPK_NT) {
        const double v = kin[i];
        unsigned long long u = (unsigned long long)__double_as_longlong(v);
        if (!(v == v) || v < 0.0) u = 0ull;  // NaN / negative never happen for a sacrifice; rank them last
        keys[i] = u;
    }
    unsigned long long thr = 0ull;
    int krem = kk, neq = len;
    if (kk < len) {
        unsigned long long prefix = 0ull, mask = 0ull;
        for (int pass = 7; pass >= 0; pass--) {
            const int shift = pass * 8;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            for (int i = tid; i < len; i += TOPK_NT) {
                const unsigned long long u = keys[i];
                if ((u & mask) == prefix) atomicAdd(&hist[(int)((u >> shift) & 255ull)], 1);
            }
            __syncthreads();
            if (tid < 32) {
                // lane owns bins [8*lane, 8*lane+8); find the digit where the count from the top reaches krem
                int loc[8], tot = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) { loc[q] = hist[tid * 8 + q]; tot += loc[q]; }
                // suffix sums over lanes (higher lanes = larger digits)
                int suf = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_down_sync(0xffffffffu, suf, o);
                    if (tid + o < 32) suf += t;
                }
                const int above = suf - tot;  // keys with a digit in a higher lane
                if (above < krem && suf >= krem) {
                    int acc = above;
                    for (int q = 7; q >= 0; q--) {
                        if (acc + loc[q] >= krem) {
                            sh_prefix = prefix | ((unsigned long long)(tid * 8 + q) << shift);
                            sh_krem = krem - acc;
                            sh_neq = loc[q];
                            break;
                        }
                        acc += loc[q];
                    }
                }
            }
            __syncthreads();
            prefix = sh_prefix;
            krem = sh_krem;
            neq = sh_neq;
            mask |= 255ull << shift;
            __syncthreads();
        }
        thr = prefix;
    } else {
        __syncthreads();
    }
    // ordered compaction.  thread owns a contiguous chunk.
    const int per = (len + TOPK_NT - 1) / TOPK_NT;
    const int cb = min(len, tid * per), ce = min(len, cb + per);
    unsigned int cgt = 0, ceq = 0;
    if (kk < len) {
        for (int i = cb; i < ce; i++) {
            const unsigned long long u = keys[i];
            cgt += (u > thr);
            ceq += (u == thr);
        }
    } else {
        cgt = ce - cb;
    }
    // exclusive scan of (gt, eq) packed in 64 bits
    unsigned long long v = ((unsigned long long)cgt << 32) | ceq, inc = v;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scan_sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned long long w = scan_sh[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        scan_sh[lane] = winc - w;
    }
    __syncthreads();
    const unsigned long long ex = scan_sh[wid] + (inc - v);
    int gt_before = (int)(ex >> 32), eq_before = (int)(ex & 0xffffffffull);
    for (int i = cb; i < ce; i++) {
        const unsigned long long u = keys[i];
        bool sel;
        if (kk >= len) sel = true;
        else if (u > thr) sel = true;
        else if (u == thr) sel = eq_before < krem;
        else sel = false;
        if (sel) {
            const int pos = gt_before + min(eq_before, kk < len ? krem : 0);
            const int id = iin ? iin[i] : (b0 + i);
            if (final_out) {
                final_out[(size_t)f * final_ld + pos] = id;
            } else {
                keys_out[(size_t)f * out_stride + (size_t)s * out_per_slice + pos] = __longlong_as_double((long long)u);
                idx_out[(size_t)f * out_stride + (size_t)s * out_per_slice + pos] = id;
            }
        }
        if (kk < len) {
            gt_before += (u > thr);
            eq_before += (u == thr);
        } else {
            gt_before++;
        }
    }
    if (final_out && tie && tid == 0) tie[f] = (kk < len && neq > krem) ? 1 : 0;
}

void launch_topk(const double *vals, long long stride, int n_in, int k, int nch, int *out_idx, int out_ld, int *tie,
                 double *ck0, int *ci0, double *ck1, int *ci1, long long cstride, cudaStream_t st, const int *gate)
{
    if (k > n_in) throw EngineError{"top-k: k > number of candidates"};
    const double *kin = vals;
    const int *iin = nullptr;
    long long in_stride = stride;
    int cur_n = n_in;
    int pp = 0;
    for (int guard = 0; guard < 16; guard++) {
        if (cur_n <= TOPK_LMAX) {
            const size_t smem = (size_t)cur_n * 8;
            dim3 grid(1, nch);
            topk_slices_kernel<<<grid, TOPK_NT, smem, st>>>(kin, iin, in_stride, cur_n, k, cur_n, nullptr, nullptr, 0,
                                                            out_idx, out_ld, tie, gate);
            CUDA_CHECK(cudaGetLastError());
            return;
        }
        // slice length: keep >= ~600 CTAs busy but never so short that nothing is filtered
        int slice = 8192;
        while (slice < TOPK_LMAX && k * 2 > slice) slice *= 2;
        if (k * 2 > slice)
            throw EngineError{"top-k: k=" + std::to_string(k) + " too large for the shared-memory radix select (max " +
                              std::to_string(TOPK_LMAX / 2) + " when p > " + std::to_string(TOPK_LMAX) + ")"};
        const int nsl = (cur_n + slice - 1) / slice;
        const int per = k < slice ? k : slice;
        // the last slice may be shorter than `per`: its candidates are packed at s*per; we compact counts on the host
        // side by making the stage output dense: slice s writes min(k, len_s) entries; only the LAST slice can be
        // short, so the dense length is (nsl-1)*per + min(k, len_last).
        const int len_last = cur_n - (nsl - 1) * slice;
        const int out_n = (nsl - 1) * per + (k < len_last ? k : len_last);
        double *ko = pp ? ck1 : ck0;
        int *io = pp ? ci1 : ci0;
        if ((long long)out_n > cstride) throw EngineError{"top-k: candidate scratch too small"};
        dim3 grid(nsl, nch);
        topk_slices_kernel<<<grid, TOPK_NT, (size_t)slice * 8, st>>>(kin, iin, in_stride, cur_n, k, slice, ko, io,
                                                                    cstride, nullptr, 0, nullptr, gate);
        CUDA_CHECK(cudaGetLastError());
        kin = ko;
        iin = io;
        in_stride = cstride;
        cur_n = out_n;
        pp ^= 1;
    }
    throw EngineError{"top-k: did not converge"};
}

// =====================================================================================================
// chain kernels: gather, active-set fits, cycle test, gradient vectors
// =====================================================================================================
constexpr int FIT_TILE_DOUBLES = 8192;  // 64 KB row tile for the Gram
constexpr int FIT_SMEM_MS = 64;         // Gram matrices up to 64 x 64 live in shared memory

struct FitSmem {
    double *tile;     // FIT_TILE_DOUBLES
    double *scratch;  // FIT_NT * 16
    double *Ssm;      // FIT_SMEM_MS^2
    double *b0, *b1, *rhs, *dg;  // ldA each
    double *red;      // 40
};
__device__ __forceinline__ FitSmem carve_fit_smem(unsigned char *raw, int ldA)
{
    FitSmem s;
    double *p = reinterpret_cast<double *>(raw);
    s.tile = p; p += FIT_TILE_DOUBLES;
    s.scratch = p; p += FIT_NT * 16;
    s.Ssm = p; p += FIT_SMEM_MS * FIT_SMEM_MS;
    s.b0 = p; p += ldA;
    s.b1 = p; p += ldA;
    s.rhs = p; p += ldA;
    s.dg = p; p += ldA;
    s.red = p; p += 40;
    return s;
}
size_t fit_smem_bytes(const Dev &d)
{
    return sizeof(double) * ((size_t)FIT_TILE_DOUBLES + FIT_NT * 16 + FIT_SMEM_MS * FIT_SMEM_MS + 4 * (size_t)d.ldA + 40);
}

// S (mm x mm, both triangles) = sum_r wt[r] * V[r][a] * V[r][b]; V row-major [nr][ldv] in global memory.
__device__ void block_syrk(const double *__restrict__ V, int ldv, int nr, int mm, const double *__restrict__ wt,
                           double *S, int lds, const FitSmem &sm)
{
    const int tid = threadIdx.x;
    const int mb = (mm + 3) >> 2, mp = mb * 4;
    const int nblk = mb * (mb + 1) / 2;
    int R = FIT_TILE_DOUBLES / (mp + 1);
    if (R > 512) R = 512;
    double *tile = sm.tile;
    double *tw = sm.tile + (size_t)R * mp;
    const int nsl = nblk >= FIT_NT ? 1 : FIT_NT / nblk;
    const int nbatch = nsl > 1 ? 1 : (nblk + FIT_NT - 1) / FIT_NT;
    const int lane = tid & 31, wid = tid >> 5;
    for (int batch = 0; batch < nbatch; batch++) {
        int blk, sl;
        bool valid;
        if (nsl > 1) {
            blk = tid % nblk;
            sl = tid / nblk;
            valid = sl < nsl;
        } else {
            blk = batch * FIT_NT + tid;
            sl = 0;
            valid = blk < nblk;
        }
        int bi = 0, bj = 0;
        if (valid) {
            bi = (int)((sqrt(8.0 * (double)blk + 1.0) - 1.0) * 0.5);
            while ((bi + 1) * (bi + 2) / 2 <= blk) bi++;
            while (bi * (bi + 1) / 2 > blk) bi--;
            bj = blk - bi * (bi + 1) / 2;
        }
        double acc[16];
#pragma unroll
        for (int e = 0; e < 16; e++) acc[e] = 0.0;
        for (int rb = 0; rb < nr; rb += R) {
            const int rc = min(R, nr - rb);
            __syncthreads();
            for (int r = wid; r < rc; r += FIT_NT / 32) {
                const double *src = V + (size_t)(rb + r) * ldv;
                for (int cidx = lane; cidx < mp; cidx += 32) tile[r * mp + cidx] = cidx < mm ? src[cidx] : 0.0;
                if (lane == 0) tw[r] = wt ? wt[rb + r] : 1.0;
            }
            __syncthreads();
            if (valid) {
                for (int r = sl; r < rc; r += nsl) {
                    const double w = tw[r];
                    const double *ta = tile + r * mp + 4 * bi;
                    const double *tb = tile + r * mp + 4 * bj;
                    const double2 a01 = *reinterpret_cast<const double2 *>(ta);
                    const double2 a23 = *reinterpret_cast<const double2 *>(ta + 2);
                    const double2 b01 = *reinterpret_cast<const double2 *>(tb);
                    const double2 b23 = *reinterpret_cast<const double2 *>(tb + 2);
                    const double a[4] = {a01.x * w, a01.y * w, a23.x * w, a23.y * w};
                    const double bb[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int qa = 0; qa < 4; qa++)
#pragma unroll
                        for (int qb = 0; qb < 4; qb++) acc[qa * 4 + qb] = fma(a[qa], bb[qb], acc[qa * 4 + qb]);
                }
            }
        }
        if (nsl > 1) {
            __syncthreads();
            if (valid) {
#pragma unroll
                for (int e = 0; e < 16; e++) sm.scratch[(size_t)(sl * nblk + blk) * 16 + e] = acc[e];
            }
            __syncthreads();
            for (int idx = tid; idx < nblk * 16; idx += FIT_NT) {
                const int bk = idx >> 4, e = idx & 15;
                double v = 0.0;
                for (int q = 0; q < nsl; q++) v += sm.scratch[(size_t)(q * nblk + bk) * 16 + e];
                int ci = (int)((sqrt(8.0 * (double)bk + 1.0) - 1.0) * 0.5);
                while ((ci + 1) * (ci + 2) / 2 <= bk) ci++;
                while (ci * (ci + 1) / 2 > bk) ci--;
                const int cj = bk - ci * (ci + 1) / 2;
                const int a = 4 * ci + (e >> 2), bcol = 4 * cj + (e & 3);
                if (a < mm && bcol < mm && a >= bcol) {
                    S[(size_t)a * lds + bcol] = v;
                    S[(size_t)bcol * lds + a] = v;
                }
            }
        } else if (valid) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int a = 4 * bi + (e >> 2), bcol = 4 * bj + (e & 3);
                if (a < mm && bcol < mm && a >= bcol) {
                    S[(size_t)a * lds + bcol] = acc[e];
                    S[(size_t)bcol * lds + a] = acc[e];
                }
            }
        }
    }
    __syncthreads();
}

// Cholesky solve of the leading mm x mm block of S (lower triangle used, destroyed); x <- S^{-1} x.
// (The reference uses Eigen's pivoted ldlt()/colPivHouseholderQr(); for the SPD, well-conditioned active-set
//  systems of this path the solutions agree to ~1e-13 relative.)
__device__ void block_chol_solve(double *S, int lds, int mm, double *x, double *dg)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int j = 0; j < mm; j++) {
        __syncthreads();
        const double djj = sqrt(S[(size_t)j * lds + j]);
        if (tid == 0) dg[j] = djj;
        const double inv = 1.0 / djj;
        for (int i = j + 1 + tid; i < mm; i += FIT_NT) S[(size_t)i * lds + j] *= inv;
        __syncthreads();
        for (int i = j + 1 + wid; i < mm; i += FIT_NT / 32) {
            const double lij = S[(size_t)i * lds + j];
            for (int c = j + 1 + lane; c <= i; c += 32) S[(size_t)i * lds + c] -= lij * S[(size_t)c * lds + j];
        }
    }
    __syncthreads();
    if (wid == 0) {
        for (int j = 0; j < mm; j++) {  // L z = x
            double part = 0.0;
            for (int c = lane; c < j; c += 32) part += S[(size_t)j * lds + c] * x[c];
            part = warp_sum(part);
            if (lane == 0) x[j] = (x[j] - part) / dg[j];
            __syncwarp();
        }
        for (int j = mm - 1; j >= 0; j--) {  // L^T x = z
            double part = 0.0;
            for (int i = j + 1 + lane; i < mm; i += 32) part += S[(size_t)i * lds + j] * x[i];
            part = warp_sum(part);
            if (lane == 0) x[j] = (x[j] - part) / dg[j];
            __syncwarp();
        }
    }
    __syncthreads();
}

struct ChainCtx {
    int c, nt, T, off, m;  // m = number of columns of the design incl. intercept
    double *XA;
    const double *y, *w;
    double *v[NVEC];
    double *S;
    int lds;
};

__device__ __forceinline__ double row_dot(const double *row, const double *b, int m)
{
    double s = 0.0;
    for (int a = 0; a < m; a++) s = fma(row[a], b[a], s);
    return s;
}

// ---- gaussian: Algorithm.h:1131-1135
__device__ void fit_lm(const ChainCtx &cx, int ldA, const FitSmem &sm, double *beta_out)
{
    const int T = cx.T;
    for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) cx.XA[(size_t)r * ldA + T] = cx.y[r];
    __syncthreads();
    block_syrk(cx.XA, ldA, cx.nt, T + 1, nullptr, cx.S, cx.lds, sm);
    for (int a = threadIdx.x; a < T; a += FIT_NT) sm.rhs[a] = cx.S[(size_t)T * cx.lds + a];
    __syncthreads();
    block_chol_solve(cx.S, cx.lds, T, sm.rhs, sm.dg);
    for (int a = threadIdx.x; a < T; a += FIT_NT) beta_out[a] = sm.rhs[a];
    __syncthreads();
}

// ---- binomial: Algorithm.h:1148-1204.  Design columns: [1 | X_A | z]
__device__ double logit_eval(const ChainCtx &cx, int ldA, const double *beta, const FitSmem &sm)
{
    double ll = 0.0;
    for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) {
        const double eu = row_dot(cx.XA + (size_t)r * ldA, beta, cx.m);
        const double e = exp(clampd(eu, 30.0));
        const double pi = e / (1.0 + e);
        cx.v[0][r] = eu;
        cx.v[1][r] = pi;
        ll += (cx.y[r] * log(pi) + (1.0 - cx.y[r]) * log(1.0 - pi)) * cx.w[r];
    }
    return block_sum<FIT_NT>(ll, sm.red);
}
__device__ void logit_wz(const ChainCtx &cx, int ldA, bool floor_w)
{
    for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) {
        const double pi = cx.v[1][r];
        double W = pi * (1.0 - pi);
        if (floor_w && W < 0.001) W = 0.001;
        cx.XA[(size_t)r * ldA + cx.m] = cx.v[0][r] + (cx.y[r] - pi) / W;
        cx.v[2][r] = W * cx.w[r];
    }
    __syncthreads();
}
__device__ void irls_solve(const ChainCtx &cx, int ldA, const FitSmem &sm, double *beta_out)
{
    block_syrk(cx.XA, ldA, cx.nt, cx.m + 1, cx.v[2], cx.S, cx.lds, sm);
    for (int a = threadIdx.x; a < cx.m; a += FIT_NT) sm.rhs[a] = cx.S[(size_t)cx.m * cx.lds + a];
    __syncthreads();
    block_chol_solve(cx.S, cx.lds, cx.m, sm.rhs, sm.dg);
    for (int a = threadIdx.x; a < cx.m; a += FIT_NT) beta_out[a] = sm.rhs[a];
    __syncthreads();
}
__device__ void fit_logistic(const ChainCtx &cx, int ldA, const FitSmem &sm)
{
    double *b0 = sm.b0, *b1 = sm.b1;
    for (int a = threadIdx.x; a < cx.m; a += FIT_NT) b0[a] = 0.0;
    __syncthreads();
    double ll0 = logit_eval(cx, ldA, b0, sm);
    logit_wz(cx, ldA, false);
    irls_solve(cx, ldA, sm, b1);
    for (int j = 0; j < 30; j++) {
        const double ll1 = logit_eval(cx, ldA, b1, sm);
        if (fabs(ll0 - ll1) / (0.1 + fabs(ll1)) < 1e-6) break;
        for (int a = threadIdx.x; a < cx.m; a += FIT_NT) b0[a] = b1[a];
        ll0 = ll1;
        __syncthreads();
        logit_wz(cx, ldA, true);
        irls_solve(cx, ldA, sm, b1);
    }
    // result: b0 (the iterate before the last solve)
}

// ---- poisson: Algorithm.h:1273-1322
__device__ void fit_poisson(const ChainCtx &cx, int ldA, double coef0_in, const FitSmem &sm)
{
    double *b0 = sm.b0;
    for (int a = threadIdx.x; a < cx.m; a += FIT_NT) b0[a] = a == 0 ? coef0_in : 0.0;
    __syncthreads();
    for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) {
        const double eta = row_dot(cx.XA + (size_t)r * ldA, b0, cx.m);
        cx.v[0][r] = eta;
        cx.v[1][r] = exp(eta);
    }
    __syncthreads();
    double ll0 = 1e5;
    for (int j = 0; j < 50; j++) {
        for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) {
            const double e = cx.v[1][r];
            cx.v[2][r] = e * cx.w[r];
            cx.XA[(size_t)r * ldA + cx.m] = cx.v[0][r] + (cx.y[r] - e) / e;
        }
        __syncthreads();
        irls_solve(cx, ldA, sm, b0);
        double ll = 0.0;
        for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) {
            const double eta = clampd(row_dot(cx.XA + (size_t)r * ldA, b0, cx.m), 30.0);
            double e = exp(eta);
            if (e < 0.001) e = 0.001;
            cx.v[0][r] = eta;
            cx.v[1][r] = e;
            ll += (cx.y[r] * eta - e) * cx.w[r];
        }
        const double ll1 = block_sum<FIT_NT>(ll, sm.red);
        if (fabs(ll0 - ll1) / fabs(0.1 + ll0) < 1e-6) break;
        ll0 = ll1;
    }
}

// ---- cox: Algorithm.h:1377-1490 (+ loglik_cox, coxph.cpp:16-40)
// loglik at beta: theta = exp(clip(X_A beta)), S0 = suffix(theta); sum status*w*log(theta/S0)
__device__ double cox_loglik(const ChainCtx &cx, int ldA, const double *beta, double *th, double *s0, const FitSmem &sm)
{
    for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) {
        const double t = exp(clampd(row_dot(cx.XA + (size_t)r * ldA, beta, cx.m), 30.0));
        th[r] = t;
        s0[r] = t;
    }
    __syncthreads();
    block_suffix_scan<FIT_NT>(s0, cx.nt, sm.red);
    double ll = 0.0;
    for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) ll += log(th[r] / s0[r]) * cx.y[r] * cx.w[r];
    return block_sum<FIT_NT>(ll, sm.red);
}
// XB[r][a] = suffix_r(theta * XA[.][a]) / S0[r]   (risk-set means), chunked two-pass scan over rows
__device__ void cox_riskset_means(const ChainCtx &cx, int ldA, double *XB, const double *th, const double *s0,
                                  const FitSmem &sm)
{
    const int m = cx.m, nt = cx.nt;
    // rows per chunk: chunk sums [nch][m] must fit the scratch region (FIT_NT*16 doubles)
    const int nch_max = max(1, (FIT_NT * 16) / m);
    const int CH = max(32, (nt + nch_max - 1) / nch_max);
    const int nch = (nt + CH - 1) / CH;
    double *csum = sm.scratch;  // [nch][m]   (needs nch*m <= FIT_NT*16)
    for (int it = threadIdx.x; it < nch * m; it += FIT_NT) {
        const int ch = it / m, a = it % m;
        const int rb = ch * CH, re = min(nt, rb + CH);
        double s = 0.0;
        for (int r = re - 1; r >= rb; r--) s += th[r] * cx.XA[(size_t)r * ldA + a];
        csum[it] = s;
    }
    __syncthreads();
    // exclusive suffix over chunks per column (sequential over nch, parallel over columns)
    for (int a = threadIdx.x; a < m; a += FIT_NT) {
        double run = 0.0;
        for (int ch = nch - 1; ch >= 0; ch--) {
            const double t = csum[ch * m + a];
            csum[ch * m + a] = run;
            run += t;
        }
    }
    __syncthreads();
    for (int it = threadIdx.x; it < nch * m; it += FIT_NT) {
        const int ch = it / m, a = it % m;
        const int rb = ch * CH, re = min(nt, rb + CH);
        double s = csum[it];
        for (int r = re - 1; r >= rb; r--) {
            s += th[r] * cx.XA[(size_t)r * ldA + a];
            XB[(size_t)r * ldA + a] = s / s0[r];
        }
    }
    __syncthreads();
}
__device__ int g_dbg_cox_iters = 30;  // debug knob (bess_b200_debug_set key 1); 30 = reference behaviour
__device__ void fit_cox(const ChainCtx &cx, int ldA, double *XB, double *S2, int lds2, const FitSmem &sm)
{
    const int max_newton = g_dbg_cox_iters;
    const int m = cx.m, nt = cx.nt;
    double *b0 = sm.b0, *b1 = sm.b1;
    double *th = cx.v[0], *s0 = cx.v[1], *ev = cx.v[2], *om = cx.v[3], *gv = cx.v[4];
    for (int a = threadIdx.x; a < m; a += FIT_NT) b0[a] = 0.0;
    __syncthreads();
    double ll0 = 1e5;
    for (int l = 1; l <= max_newton; l++) {
        // theta (no weights here, Algorithm.h:1423), S0
        for (int r = threadIdx.x; r < nt; r += FIT_NT) {
            const double t = exp(clampd(row_dot(cx.XA + (size_t)r * ldA, b0, m), 30.0));
            th[r] = t;
            s0[r] = t;
        }
        __syncthreads();
        block_suffix_scan<FIT_NT>(s0, nt, sm.red);
        // e = w*status; C = prefix(e/S0); omega = theta*C; gvec = e - omega
        for (int r = threadIdx.x; r < nt; r += FIT_NT) {
            ev[r] = cx.w[r] * cx.y[r];
            om[r] = ev[r] / s0[r];
        }
        __syncthreads();
        block_prefix_scan<FIT_NT>(om, nt, sm.red);
        for (int r = threadIdx.x; r < nt; r += FIT_NT) {
            om[r] *= th[r];
            gv[r] = ev[r] - om[r];
        }
        __syncthreads();
        cox_riskset_means(cx, ldA, XB, th, s0, sm);
        // g = X_A^T gv : use the augmented column trick with unit weights: column m of XA <- gv
        for (int r = threadIdx.x; r < nt; r += FIT_NT) cx.XA[(size_t)r * ldA + m] = gv[r];
        __syncthreads();
        // P1 = X_A^T diag(omega) X_A  (last row unused); g from an unweighted product of [X_A | gv]
        block_syrk(cx.XA, ldA, nt, m, om, cx.S, cx.lds, sm);
        block_syrk(XB, ldA, nt, m, ev, S2, lds2, sm);
        // g_a = sum_r XA[r][a]*gv[r]
        {
            // (a, slice) decomposition, deterministic reduction through scratch
            const int nsl = max(1, FIT_NT / m);
            for (int it = threadIdx.x; it < m * nsl; it += FIT_NT) {
                const int a = it % m, sl = it / m;
                double s = 0.0;
                for (int r = sl; r < nt; r += nsl) s = fma(cx.XA[(size_t)r * ldA + a], gv[r], s);
                sm.scratch[it] = s;
            }
            __syncthreads();
            for (int a = threadIdx.x; a < m; a += FIT_NT) {
                double s = 0.0;
                for (int sl = 0; sl < nsl; sl++)
                    if (a + sl * m < m * nsl) s += sm.scratch[sl * m + a];
                sm.rhs[a] = s;
            }
            __syncthreads();
        }
        // P = P1 - P2 = -h ;  h d = g  =>  d = -P^{-1} g
        for (int it = threadIdx.x; it < m * m; it += FIT_NT) {
            const int a = it / m, bcol = it % m;
            cx.S[(size_t)a * cx.lds + bcol] -= S2[(size_t)a * lds2 + bcol];
        }
        __syncthreads();
        block_chol_solve(cx.S, cx.lds, m, sm.rhs, sm.dg);  // rhs = P^{-1} g = -d
        // line search (Algorithm.h:1474-1481): beta1 = beta0 - 0.5^mm * d = beta0 + 0.5^mm * rhs
        int mm = 1;
        double step = 0.5;
        for (int a = threadIdx.x; a < m; a += FIT_NT) b1[a] = b0[a] + step * sm.rhs[a];
        __syncthreads();
        double ll1 = cox_loglik(cx, ldA, b1, cx.v[5], cx.v[6], sm);
        while (ll0 > ll1 && mm < 5) {
            mm++;
            step *= 0.5;
            __syncthreads();
            for (int a = threadIdx.x; a < m; a += FIT_NT) b1[a] = b0[a] + step * sm.rhs[a];
            __syncthreads();
            ll1 = cox_loglik(cx, ldA, b1, cx.v[5], cx.v[6], sm);
        }
        if (fabs(ll0 - ll1) / fabs(0.1 + ll0) < 1e-5) break;
        __syncthreads();
        for (int a = threadIdx.x; a < m; a += FIT_NT) b0[a] = b1[a];
        ll0 = ll1;
        __syncthreads();
    }
}

// Gradient vectors of the next dual sweep from the chain's current (A, beta_A, coef0); X_A is in cx.XA.
__device__ void chain_gradient(const Dev &d, const ChainCtx &cx, const double *bsl /*smem slopes*/, int ks, double coef0,
                               const FitSmem &sm)
{
    const int c = cx.c, FS = d.FS, nt = cx.nt;
    const int *rows = d.rows + (size_t)c * d.n;
    const int fam = d.family;
    if (fam != FAM_COX) {
        for (int r = threadIdx.x; r < nt; r += FIT_NT) {
            double eta = coef0;
            const double *row = cx.XA + (size_t)r * d.ldA + cx.off;
            for (int a = 0; a < ks; a++) eta = fma(row[a], bsl[a], eta);
            const size_t o = (size_t)rows[r] * FS + c;
            if (fam == FAM_LM) {
                d.G[o] = (cx.y[r] - eta) / (double)nt;  // Algorithm.h:1109 (coef0 == 0 for gaussian)
            } else if (fam == FAM_LOGIT) {
                const double e = exp(clampd(eta, 30.0));  // Algorithm.h:1223-1236
                const double pr = e / (e + 1.0);
                d.G[o] = cx.w[r] * (cx.y[r] - pr);
                d.W[o] = cx.w[r] * pr * (1.0 - pr);
            } else {
                const double e = exp(eta);  // Algorithm.h:1338-1341 (not clamped)
                d.G[o] = (cx.y[r] - e) * cx.w[r];
                d.W[o] = e * cx.w[r];
            }
        }
        __syncthreads();
        return;
    }
    // cox, Algorithm.h:1579-1630 restated with prefix/suffix sums
    double *th = cx.v[0], *s0 = cx.v[1], *cc = cx.v[2];
    for (int r = threadIdx.x; r < nt; r += FIT_NT) {
        double eta = 0.0;
        const double *row = cx.XA + (size_t)r * d.ldA + cx.off;
        for (int a = 0; a < ks; a++) eta = fma(row[a], bsl[a], eta);
        const double t = cx.w[r] * exp(clampd(eta, 30.0));
        th[r] = t;
        s0[r] = t;
    }
    __syncthreads();
    block_suffix_scan<FIT_NT>(s0, nt, sm.red);
    for (int r = threadIdx.x; r < nt; r += FIT_NT) cc[r] = (cx.y[r] != 0.0 ? cx.w[r] : 0.0) / s0[r];
    __syncthreads();
    block_prefix_scan<FIT_NT>(cc, nt, sm.red);
    for (int r = threadIdx.x; r < nt; r += FIT_NT) {
        const double e = cx.y[r] != 0.0 ? cx.w[r] : 0.0;
        const double om = th[r] * cc[r];
        const size_t o = (size_t)rows[r] * FS + c;
        d.G[o] = e - om;
        d.W[o] = om;
        d.TH[o] = th[r];
        d.C2[o] = e / (s0[r] * s0[r]);
    }
    __syncthreads();
}

__device__ __forceinline__ ChainCtx make_ctx(const Dev &d, int c, int T, const FitSmem &sm)
{
    ChainCtx cx;
    cx.c = c;
    cx.nt = d.ntrain[c];
    cx.T = T;
    cx.off = (d.family == FAM_LOGIT || d.family == FAM_POISSON) ? 1 : 0;
    cx.m = T + cx.off;
    cx.XA = d.XA + (size_t)c * d.n * d.ldA;
    cx.y = d.ytr + (size_t)c * d.n;
    cx.w = d.wtr + (size_t)c * d.n;
    for (int q = 0; q < NVEC; q++) cx.v[q] = d.vec + ((size_t)c * NVEC + q) * d.n;
    const int mmax = cx.m + 1;
    if (mmax <= FIT_SMEM_MS) {
        cx.S = sm.Ssm;
        cx.lds = FIT_SMEM_MS;
    } else {
        cx.S = d.Smat + (size_t)c * 2 * d.ldA * d.ldA;
        cx.lds = d.ldA;
    }
    return cx;
}

// Start of a batch (Algorithm::fit prologue, Algorithm.h:141-148): coef0 <- coef0_init, l <- 0, A_list.col(0) <- 0,
// gradient vectors from beta_init.
__global__ void __launch_bounds__(FIT_NT, 1) chain_begin_kernel(const Dev d, const BatchDesc b)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FitSmem sm = carve_fit_smem(smem_raw, d.ldA);
    const int c = b.chain[blockIdx.x];
    // Algorithm::coef0_init is only refreshed when the path starts a new step (path.cpp:57); CV folds of the same
    // step inherit it.
    double level;
    if (!d.warm) level = 0.0;
    else if (b.new_path_step) level = d.coef0[0];
    else level = *d.coef0_level;
    __syncthreads();
    int ks = d.ks[c];
    if (!d.warm) {
        // cold start: beta_init = 0
        for (int a = threadIdx.x; a < ks; a += FIT_NT) d.betaD[(size_t)c * d.pstride + d.A[(size_t)c * d.kcap + a]] = 0.0;
        ks = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0 && b.new_path_step) *d.coef0_level = level;
        if (blockIdx.x == 0) *d.n_active = b.nch;
        d.tie_acc[c] = 0;
        if (!(c == 0 && b.new_path_step && d.warm)) d.coef0[c] = level;
        d.ks[c] = ks;
        d.l[c] = 0;
        d.done[c] = 0;
    }
    int *h0 = d.hist + (size_t)c * MAX_HIST * d.kcap;
    for (int a = threadIdx.x; a < b.T; a += FIT_NT) h0[a] = 0;
    for (int a = threadIdx.x; a < ks; a += FIT_NT) sm.b0[a] = d.bA[(size_t)c * d.kcap + a];
    __syncthreads();
    ChainCtx cx = make_ctx(d, c, ks, sm);
    chain_gradient(d, cx, sm.b0, ks, level, sm);
}

// One PDAS iteration after the top-k (Algorithm.h:154-170) + gradient vectors for the next one.
__global__ void __launch_bounds__(FIT_NT, 1) chain_fit_kernel(const Dev d, const BatchDesc b)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FitSmem sm = carve_fit_smem(smem_raw, d.ldA);
    const int c = b.chain[blockIdx.x];
    if (d.done[c]) return;
    const int T = b.T;
    ChainCtx cx = make_ctx(d, c, T, sm);
    const int ldA = d.ldA;
    const int *Anew = d.Anew + (size_t)c * d.kcap;
    const int *rows = d.rows + (size_t)c * d.n;
    int *Acur = d.A + (size_t)c * d.kcap;
    const int ks_old = d.ks[c];
    const double coef0_in = d.coef0[c];

    // clear the dense beta on the old support (Algorithm.h:159)
    for (int a = threadIdx.x; a < ks_old; a += FIT_NT) d.betaD[(size_t)c * d.pstride + Acur[a]] = 0.0;
    // gather X_A (utilities.cpp:132-140): XA[r][off + a] = X[rows[r]][A[a]]
    for (int it0 = threadIdx.x; it0 < cx.nt * T; it0 += 4 * FIT_NT) {
        double val[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int it = it0 + q * FIT_NT;
            if (it < cx.nt * T) val[q] = __ldg(d.X + (size_t)rows[it / T] * d.ldx + Anew[it % T]);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int it = it0 + q * FIT_NT;
            if (it < cx.nt * T) cx.XA[(size_t)(it / T) * ldA + cx.off + (it % T)] = val[q];
        }
    }
    if (cx.off)
        for (int r = threadIdx.x; r < cx.nt; r += FIT_NT) cx.XA[(size_t)r * ldA] = 1.0;
    __syncthreads();

    double coef0 = coef0_in;
    const double *slopes;
    if (d.family == FAM_LM) {
        fit_lm(cx, ldA, sm, sm.b0);
        slopes = sm.b0;
    } else if (d.family == FAM_LOGIT) {
        fit_logistic(cx, ldA, sm);
        coef0 = sm.b0[0];
        slopes = sm.b0 + 1;
    } else if (d.family == FAM_POISSON) {
        fit_poisson(cx, ldA, coef0_in, sm);
        coef0 = sm.b0[0];
        slopes = sm.b0 + 1;
    } else {
        fit_cox(cx, ldA, d.XB + (size_t)c * d.n * ldA, d.Smat + ((size_t)c * 2 + 1) * ldA * ldA, ldA, sm);
        slopes = sm.b0;
    }
    __syncthreads();
    // scatter (Algorithm.h:159-163), record A, cycle test (Algorithm.h:164-170)
    const int l = d.l[c] + 1;
    int *hl = d.hist + ((size_t)c * MAX_HIST + l) * d.kcap;
    for (int a = threadIdx.x; a < T; a += FIT_NT) {
        const int j = Anew[a];
        Acur[a] = j;
        hl[a] = j;
        d.bA[(size_t)c * d.kcap + a] = slopes[a];
        d.betaD[(size_t)c * d.pstride + j] = slopes[a];
    }
    __syncthreads();
    int seen = 0;
    for (int ll = 0; ll < l && !seen; ll++) {
        const int *hp = d.hist + ((size_t)c * MAX_HIST + ll) * d.kcap;
        int same = 1;
        for (int a = threadIdx.x; a < T; a += FIT_NT) same &= (hp[a] == Anew[a]);
        seen = __syncthreads_and(same);
    }
    const int finished = seen || l >= d.max_iter;
    if (threadIdx.x == 0) {
        d.l[c] = seen ? l : (l >= d.max_iter ? d.max_iter + 1 : l);
        d.ks[c] = T;
        d.coef0[c] = coef0;
        d.done[c] = finished;
        d.tie_acc[c] += d.tie[c];
        if (finished) atomicSub(d.n_active, 1);
    }
    if (finished) return;
    chain_gradient(d, cx, slopes, T, coef0, sm);
}

void debug_set(int key, int val)
{
    if (key == 1) CUDA_CHECK(cudaMemcpyToSymbol(g_dbg_cox_iters, &val, sizeof(int)));
}

void configure_kernels()
{
    CUDA_CHECK(cudaFuncSetAttribute(topk_slices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_LMAX * 8));
}

void launch_chain_begin(const Dev &d, const BatchDesc &b, cudaStream_t st)
{
    const size_t smem = fit_smem_bytes(d);
    CUDA_CHECK(cudaFuncSetAttribute(chain_begin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chain_begin_kernel<<<b.nch, FIT_NT, smem, st>>>(d, b);
    CUDA_CHECK(cudaGetLastError());
}
void launch_chain_fit(const Dev &d, const BatchDesc &b, cudaStream_t st)
{
    const size_t smem = fit_smem_bytes(d);
    CUDA_CHECK(cudaFuncSetAttribute(chain_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chain_fit_kernel<<<b.nch, FIT_NT, smem, st>>>(d, b);
    CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================================
// losses: Metric.h:145-148/190 (lm), :266-290/336-351 (logistic), :426-440/489 (poisson), :565-568/609 (cox)
// =====================================================================================================
__global__ void __launch_bounds__(FIT_NT) loss_kernel(const Dev d, const LossDesc jobs, const int *testrows,
                                                      const int *ntest, const double *y, const double *w,
                                                      const double *lfact, double *scratch, double *out)
{
    __shared__ double red[40];
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *bsl = reinterpret_cast<double *>(smem_raw);           // [kcap]
    int *As = reinterpret_cast<int *>(bsl + d.kcap);              // [kcap]
    const int job = blockIdx.x;
    const int c = jobs.chain[job], kind = jobs.kind[job], fold = jobs.fold[job];
    const int ks = d.ks[c];
    const double coef0 = d.coef0[c];
    for (int a = threadIdx.x; a < ks; a += FIT_NT) {
        bsl[a] = d.bA[(size_t)c * d.kcap + a];
        As[a] = d.A[(size_t)c * d.kcap + a];
    }
    __syncthreads();
    const int nr = kind == 0 ? d.n : ntest[fold];
    const int *rl = kind == 0 ? nullptr : testrows + (size_t)fold * d.n;
    double *e_s = scratch + (size_t)job * 2 * d.n;
    double *c_s = e_s + d.n;
    double acc = 0.0;
    for (int r = threadIdx.x; r < nr; r += FIT_NT) {
        const int i = rl ? rl[r] : r;
        double eta = d.family == FAM_LM || d.family == FAM_COX ? 0.0 : coef0;
        const double *row = d.X + (size_t)i * d.ldx;
        for (int a = 0; a < ks; a++) eta = fma(row[As[a]], bsl[a], eta);
        if (d.family == FAM_LM) {
            const double t = y[i] - eta;
            acc += t * t;
        } else if (d.family == FAM_LOGIT) {
            const double e = exp(clampd(eta, kind == 0 ? 30.0 : 25.0));
            const double pr = e / (e + 1.0);
            acc += w[i] * (y[i] * log(pr) + (1.0 - y[i]) * log(1.0 - pr));
        } else if (d.family == FAM_POISSON) {
            const double ec = clampd(eta, 30.0);
            acc += (y[i] * ec - exp(ec) - lfact[i]) * w[i];
        } else {
            const double ec = clampd(eta, 30.0);
            e_s[r] = exp(ec);
            c_s[r] = e_s[r];
        }
    }
    double res;
    if (d.family == FAM_COX) {
        __syncthreads();
        block_suffix_scan<FIT_NT>(c_s, nr, red);
        for (int r = threadIdx.x; r < nr; r += FIT_NT) {
            const int i = rl ? rl[r] : r;
            acc += log(e_s[r] / c_s[r]) * y[i] * w[i];
        }
        res = -2.0 * block_sum<FIT_NT>(acc, red);
    } else {
        const double s = block_sum<FIT_NT>(acc, red);
        if (d.family == FAM_LM) res = kind == 0 ? s / (double)d.n : s / (double)(2 * nr);
        else if (d.family == FAM_LOGIT) res = -2.0 * s;
        else res = kind == 0 ? -2.0 * s : -s;
    }
    if (threadIdx.x == 0) out[job] = res;
}

void launch_losses(const Dev &d, const LossDesc &jobs, const int *testrows, const int *ntest, const double *y,
                   const double *w, const double *lfact, double *scratch, double *out, cudaStream_t st)
{
    const size_t smem = (size_t)d.kcap * (sizeof(double) + sizeof(int));
    loss_kernel<<<jobs.njobs, FIT_NT, smem, st>>>(d, jobs, testrows, ntest, y, w, lfact, scratch, out);
    CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================================
// normalisation / screening helpers
// =====================================================================================================
// X[i][j] = (X[i][j] - sub[j]) * mul[j] * rowmul[i]     (normalize.cpp:30-45 + Data.h:70-77)
__global__ void center_scale_kernel(double *X, long long ldx, int n, int p, const double *sub, const double *mul,
                                    const double *rowmul)
{
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (j >= p) return;
    const bool two = j + 1 < p;
    const double s0 = sub ? sub[j] : 0.0, s1 = (sub && two) ? sub[j + 1] : 0.0;
    const double m0 = mul ? mul[j] : 1.0, m1 = (mul && two) ? mul[j + 1] : 1.0;
    const int r0 = blockIdx.y * 64, r1 = min(n, r0 + 64);
    for (int i = r0; i < r1; i++) {
        double2 *ptr = reinterpret_cast<double2 *>(X + (size_t)i * ldx + j);
        double2 v = *ptr;
        const double rm = rowmul ? rowmul[i] : 1.0;
        v.x = (v.x - s0) * m0 * rm;
        v.y = two ? (v.y - s1) * m1 * rm : 0.0;
        *ptr = v;
    }
}
void launch_center_scale(double *X, long long ldx, int n, int p, const double *sub, const double *mul,
                         const double *rowmul, cudaStream_t st)
{
    dim3 grid((unsigned)(((long long)(p + 1) / 2 + 127) / 128), (unsigned)((n + 63) / 64));
    center_scale_kernel<<<grid, 128, 0, st>>>(X, ldx, n, p, sub, mul, rowmul);
    CUDA_CHECK(cudaGetLastError());
}

// Xn[i][jn] = X[i][cols[jn]]   (screening.cpp:83-88)
__global__ void gather_cols_kernel(const double *X, long long ldx, int n, const int *cols, int pnew, double *Xn,
                                   long long ldn)
{
    const int jn = blockIdx.x * blockDim.x + threadIdx.x;
    if (jn >= pnew) return;
    const int src = cols[jn];
    const int r0 = blockIdx.y * 32, r1 = min(n, r0 + 32);
    for (int i = r0; i < r1; i++) Xn[(size_t)i * ldn + jn] = X[(size_t)i * ldx + src];
}
void launch_gather_cols(const double *X, long long ldx, int n, const int *cols, int pnew, double *Xn, long long ldn,
                        cudaStream_t st)
{
    dim3 grid((pnew + 127) / 128, (n + 31) / 32);
    gather_cols_kernel<<<grid, 128, 0, st>>>(X, ldx, n, cols, pnew, Xn, ldn);
    CUDA_CHECK(cudaGetLastError());
}

__global__ void gather_cols_pos_kernel(const double *X, long long ldx, int n, const int *cols, const int *pos, int m,
                                       double *dst, long long ld)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const int src = cols[q], dp = pos[q];
    const int r0 = blockIdx.y * 32, r1 = min(n, r0 + 32);
    for (int i = r0; i < r1; i++) dst[(size_t)i * ld + dp] = X[(size_t)i * ldx + src];
}
void launch_gather_cols_pos(const double *X, long long ldx, int n, const int *cols, const int *pos, int m, double *dst,
                            long long ld, cudaStream_t st)
{
    dim3 grid((m + 127) / 128, (n + 31) / 32);
    gather_cols_pos_kernel<<<grid, 128, 0, st>>>(X, ldx, n, cols, pos, m, dst, ld);
    CUDA_CHECK(cudaGetLastError());
}

// Marginal GLM utilities for screening (screening.cpp:48-61): one thread per column, the column is re-read from
// L2/HBM each Newton/IRLS step (coalesced across the warp because X is row-major).
//   binomial: logit_fit, logistic.cpp:61-157 (2-parameter IRLS, no W floor, returns the previous iterate)
//   poisson : poisson_fit, poisson.cpp:84-137, restated LITERALLY incl. the vector*vector product that evaluates to
//             X.col(i)*expeta_w(0) with asserts off (:117) and the wrong-sign step (:122,:128)
//   cox     : cox_fit, coxph.cpp:42-109 (1-parameter damped Newton, clamp +-50)
__device__ __forceinline__ void solve2(double a, double b, double c, double r0, double r1, double &x0, double &x1)
{
    // [[a b][b c]] x = r  (2x2 symmetric)
    const double det = a * c - b * b;
    x0 = (c * r0 - b * r1) / det;
    x1 = (a * r1 - b * r0) / det;
}

__global__ void __launch_bounds__(128) screen_glm_kernel(const double *__restrict__ X, long long ldx, int n, int p,
                                                         const double *__restrict__ y, const double *__restrict__ w,
                                                         int family, double *util)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ys = reinterpret_cast<double *>(smem_raw);
    double *ws = ys + n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        ys[i] = y[i];
        ws[i] = w[i];
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const double *xc = X + j;
    if (family == FAM_LOGIT) {
        double b00 = 0.0, b01 = 0.0;  // beta0
        double b10, b11;              // beta1
        // step 0 at beta0 = 0: Pi = 0.5
        double ll0 = 0.0, s00 = 0, s01 = 0, s11 = 0, r0 = 0, r1 = 0;
        for (int i = 0; i < n; i++) {
            const double x = __ldg(xc + (size_t)i * ldx);
            const double pi = 0.5;
            ll0 += (ys[i] * log(pi) + (1.0 - ys[i]) * log(1.0 - pi)) * ws[i];
            const double W = pi * (1.0 - pi);
            const double Z = (ys[i] - pi) / W;
            const double ww = W * ws[i];
            s00 += ww; s01 += ww * x; s11 += ww * x * x; r0 += ww * Z; r1 += ww * x * Z;
        }
        solve2(s00, s01, s11, r0, r1, b10, b11);
        for (int it = 0; it < 30; it++) {
            double ll1 = 0.0;
            s00 = s01 = s11 = r0 = r1 = 0.0;
            for (int i = 0; i < n; i++) {
                const double x = __ldg(xc + (size_t)i * ldx);
                const double eu = b10 + b11 * x;
                const double e = exp(clampd(eu, 30.0));
                const double pi = e / (1.0 + e);
                ll1 += (ys[i] * log(pi) + (1.0 - ys[i]) * log(1.0 - pi)) * ws[i];
                const double W = pi * (1.0 - pi);
                const double Z = eu + (ys[i] - pi) / W;
                const double ww = W * ws[i];
                s00 += ww; s01 += ww * x; s11 += ww * x * x; r0 += ww * Z; r1 += ww * x * Z;
            }
            if (fabs(ll0 - ll1) / (0.1 + fabs(ll1)) < 1e-6) break;
            b00 = b10; b01 = b11; ll0 = ll1;
            solve2(s00, s01, s11, r0, r1, b10, b11);
        }
        (void)b00;
        util[j] = b01 * b01;
    } else if (family == FAM_POISSON) {
        double b0 = 0.0, b1 = 0.0;
        for (int it = 0; it < 100; it++) {
            double g0 = 0, g1 = 0, h00 = 0, h01 = 0, h11 = 0, ll0 = 0, ew0 = 0;
            for (int i = 0; i < n; i++) {
                const double x = __ldg(xc + (size_t)i * ldx);
                const double eta = clampd(b0 + b1 * x, 30.0);
                const double e = exp(eta);
                if (i == 0) ew0 = e * ws[0];
                const double r = (ys[i] - e) * ws[i];
                g0 += r; g1 += x * r;
                h00 += 1.0; h01 += x; h11 += x * x;  // scaled by ew0 below (poisson.cpp:117 as evaluated)
                ll0 += (ys[i] * eta - e) * ws[i];
            }
            h00 *= ew0; h01 *= ew0; h11 *= ew0;
            double d0, d1;
            solve2(h00, h01, h11, g0, g1, d0, d1);
            int m = 0;
            double step = 1.0;
            double n0 = b0 - d0, n1 = b1 - d1, ll1 = 0.0;
            for (;;) {
                ll1 = 0.0;
                for (int i = 0; i < n; i++) {
                    const double x = __ldg(xc + (size_t)i * ldx);
                    const double eta = clampd(n0 + n1 * x, 30.0);
                    ll1 += (ys[i] * eta - exp(eta)) * ws[i];
                }
                if (!(ll0 >= ll1 && m < 10)) break;
                m++;
                step *= 0.2;
                n0 = b0 - step * d0;
                n1 = b1 - step * d1;
            }
            b0 = n0; b1 = n1;
            if (fabs(ll0 - ll1) / fabs(ll0) < 1e-8) break;
        }
        util[j] = b1 * b1;
    } else {  // cox, one parameter
        double b0 = 0.0, ll0 = 1e5;
        for (int l = 1; l <= 30; l++) {
            // one backward pass: S0, S1, S2 running suffix sums
            double S0 = 0, S1 = 0, S2 = 0, g = 0, hneg = 0;
            for (int i = n - 1; i >= 0; i--) {
                const double x = __ldg(xc + (size_t)i * ldx);
                const double th = exp(clampd(b0 * x, 50.0));
                S0 += th; S1 += th * x; S2 += th * x * x;
                const double e = ws[i] * ys[i];
                const double xb = S1 / S0;
                g += (x - xb) * e;
                hneg += (S2 / S0 - xb * xb) * e;  // = -h
            }
            const double d = -g / hneg;  // h d = g
            int m = 1;
            double step = 0.5;
            double b1 = b0 - step * d, ll1;
            for (;;) {
                double S = 0, acc = 0;
                for (int i = n - 1; i >= 0; i--) {
                    const double x = __ldg(xc + (size_t)i * ldx);
                    const double th = exp(clampd(b1 * x, 30.0));
                    S += th;
                    acc += log(th / S) * ys[i] * ws[i];
                }
                ll1 = acc;
                if (!(ll0 > ll1 && m < 5)) break;
                m++;
                step *= 0.5;
                b1 = b0 - step * d;
            }
            if (fabs(ll0 - ll1) / fabs(0.1 + ll0) < 1e-5) break;
            b0 = b1;
            ll0 = ll1;
        }
        util[j] = b0 * b0;
    }
}
void launch_screen_glm(const double *X, long long ldx, int n, int p, const double *y, const double *w, int family,
                       double *util, cudaStream_t st)
{
    const size_t smem = (size_t)2 * n * sizeof(double);
    CUDA_CHECK(cudaFuncSetAttribute(screen_glm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    screen_glm_kernel<<<(p + 127) / 128, 128, smem, st>>>(X, ldx, n, p, y, w, family, util);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace bess
