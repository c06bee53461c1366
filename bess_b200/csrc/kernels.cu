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
#include "device_utils.cuh"

#include <cfloat>
#include <cmath>
#include <cstdio>

namespace bess {

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
        case 24: launch_sweep_t<24, MODE, 1>(d, st); break;
        case 32: launch_sweep_t<32, MODE, 1>(d, st); break;
        default: throw EngineError{"dual sweep: unsupported chain tile FS=" + std::to_string(d.FS)};
    }
}

void launch_dual_sweep(const Dev &d, int mode, cudaStream_t st)
{
    if (sweep_uses_tma(d)) {
        launch_dual_sweep_tma(d, mode, st);
        return;
    }
    if (mode == MODE_D) launch_sweep_m<MODE_D>(d, st);
    else if (mode == MODE_DH) launch_sweep_m<MODE_DH>(d, st);
    else launch_sweep_m<MODE_COX>(d, st);
}

// =====================================================================================================
// finish: reduce row-split partials, apply the sacrifice
// =====================================================================================================
// sacrifice / screening utility from the reduced sums (the tail of the finish epilogue)
template <int EPI>
__device__ __forceinline__ double finish_epilogue(const Dev &d, int c, long long j, double dsum, double hsum, double R)
{
    double out;
    if (EPI == EPI_SCREEN_LM) {
        const double bq = dsum / hsum;  // one-column least squares (screening.cpp:46)
        out = bq * bq;
    } else {
        const double beta = d.betaD[(size_t)c * d.pstride + j];
        const double lam2 = 2.0 * d.lambda;  // L0L2 ridge term (Algorithm.h:1109, 1236/1246, 1341/1350, 1629-1630)
        if (EPI == EPI_SACR_LM) {
            // Phi = sqrt(2*lambda + x_j.x_j / n) (utilities.cpp:142-151), invPhi = 1/Phi (:167-177); Algorithm.h:1116-1122
            const double phi = sqrt(lam2 + d.xtx[(size_t)c * d.pstride + j] / (double)d.ntrain[c]);
            const double t = phi * beta + (1.0 / phi) * (dsum - lam2 * beta);
            out = t * t;
        } else if (EPI == EPI_SACR_GLM) {
            const double phi = sqrt(hsum + lam2);  // Algorithm.h:1238-1257 / 1342-1361
            const double t = phi * beta + (1.0 / phi) * (dsum - lam2 * beta);
            out = t * t;
        } else {
            // Algorithm.h:1626-1634: l1 = -dsum + 2*lambda*beta, l2 = hsum - R + 2*lambda, d = -l1/l2, bd = |beta + d| * sqrt(l2)
            const double l2 = hsum - R + lam2;
            out = fabs(beta + (dsum - lam2 * beta) / l2) * sqrt(l2);
        }
    }
    return out;
}

// value of the finish epilogue EPI for chain c, column j: reduces the row-split partials in a fixed order and applies the
// sacrifice / screening formula.  EPI_RAW writes the reduced sums to raw_out and returns 0.
template <int EPI>
__device__ __forceinline__ double finish_value(const Dev &d, int mode, int c, long long j, double *raw_out)
{
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
        return 0.0;
    }
    return finish_epilogue<EPI>(d, c, j, dsum, hsum, R);
}

// The same for NB columns j0, j0 + jstep, ... of one thread at once (gaussian / GLM sweeps): the partials of all NB columns
// are in flight together instead of one dependent chain of S loads per column.  Per column the partials are still added
// in the order s = 0 .. S-1, so the result is bit-identical to finish_value.
template <int EPI, int NB>
__device__ __forceinline__ void finish_batch(const Dev &d, int mode, int c, long long j0, long long jstep, int nb, double (&out)[NB])
{
    const int NQ = mode == MODE_D ? 1 : 2;
    const int FT = d.FS;
    double ds[NB], hs[NB];
#pragma unroll
    for (int q = 0; q < NB; q++) ds[q] = 0.0, hs[q] = 0.0;
#pragma unroll 2
    for (int s = 0; s < d.S; s++) {
        const double *pd = d.part + ((size_t)(s * NQ + 0) * FT + c) * d.pstride + j0;
        const double *ph = d.part + ((size_t)(s * NQ + (NQ - 1)) * FT + c) * d.pstride + j0;
        double vd[NB], vh[NB];
#pragma unroll
        for (int q = 0; q < NB; q++) {
            vd[q] = q < nb ? pd[q * jstep] : 0.0;
            vh[q] = (q < nb && mode == MODE_DH) ? ph[q * jstep] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < NB; q++) {
            ds[q] += vd[q];
            hs[q] += vh[q];
        }
    }
#pragma unroll
    for (int q = 0; q < NB; q++)
        if (q < nb) out[q] = finish_epilogue<EPI>(d, c, j0 + q * jstep, ds[q], hs[q], 0.0);
}


template <int EPI>
__global__ void __launch_bounds__(256) finish_kernel(const Dev d, int mode, const BatchDesc b, double *raw_out)
{
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= d.p) return;
    if (d.gate && *d.gate == 0) return;
    const int c = b.chain[blockIdx.y];  // chain id == slot in the sweep vectors
    const double out = finish_value<EPI>(d, mode, c, j, raw_out);
    if (EPI != EPI_RAW) d.bd[(size_t)c * d.pstride + j] = out;
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
__device__ __forceinline__ unsigned long long topk_key(double v)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    if (!(v == v) || v < 0.0) u = 0ull;  // NaN / negative never happen for a sacrifice; rank them last
    return u;
}

// Select + ordered compaction over `len` keys already in shared memory (block-synchronised by the caller's first barrier
// inside).  Outputs as in topk_slices_kernel.
__device__ __forceinline__ void topk_body(unsigned long long *keys, int len, int k, int slice_len, const int *iin, int b0, int s,
                                          int f, double *keys_out, int *idx_out, long long out_stride, int *final_out,
                                          int final_ld, int *tie)
{
    __shared__ int hist[256];
    __shared__ unsigned long long sh_prefix;
    __shared__ int sh_krem, sh_neq, sh_cnt;
    __shared__ unsigned long long scan_sh[34];
    __shared__ unsigned long long small_bin[32];
    const int tid = threadIdx.x;
    const int kk = min(k, len);
    const int out_per_slice = min(k, slice_len);
    unsigned long long thr = 0ull;
    int krem = kk, neq = len;
    if (kk < len) {
        unsigned long long prefix = 0ull, mask = 0ull;
        bool early = false;
        for (int pass = 7; pass >= 0; pass--) {
            const int shift = pass * 8;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            // (a warp-aggregated variant -- __match_any_sync on the digit, one atomicAdd per distinct digit -- was measured
            // and lost: 0.135 -> 0.247 ms for the four stages of the screening's top-5000; match.any costs more than the
            // same-address shared atomics it saves; so did a single-ballot variant that counts the lanes sharing the digit
            // of the warp's first candidate with one add: 0.134 -> 0.160 ms -- the hardware already merges same-address
            // shared atomics of a warp)
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
            // every key of the boundary bin is wanted: the select is decided, the remaining digits cannot change it
            if (neq == krem && pass > 0 && prefix > 0ull) {
                early = true;
                break;
            }
            // few keys left in the boundary bin (the usual case after two digits: the largest sacrifices are spread over
            // orders of magnitude): rank them in one warp instead of walking the remaining digits
            if (neq <= 32 && pass > 0) {
                if (tid == 0) sh_cnt = 0;
                __syncthreads();
                for (int i = tid; i < len; i += TOPK_NT) {
                    const unsigned long long u = keys[i];
                    if ((u & mask) == prefix) small_bin[atomicAdd(&sh_cnt, 1)] = u;
                }
                __syncthreads();
                if (tid < 32 && tid < neq) {
                    const unsigned long long u = small_bin[tid];
                    int gt = 0, ge = 0;
                    for (int j = 0; j < neq; j++) {
                        const unsigned long long v = small_bin[j];
                        gt += (v > u);
                        ge += (v >= u);
                    }
                    if (gt < krem && krem <= ge) {  // u is the krem-th largest of the bin (all its duplicates agree)
                        sh_prefix = u;
                        sh_krem = krem - gt;
                        sh_neq = ge - gt;
                    }
                }
                __syncthreads();
                prefix = sh_prefix;
                krem = sh_krem;
                neq = sh_neq;
                __syncthreads();
                break;
            }
        }
        if (early) {
            thr = prefix - 1ull;  // u > thr  <=>  u >= prefix (lower digits zero): all of the boundary bin and above
            krem = 0;             // nothing is taken from keys equal to thr (they sit in a lower bin)
            neq = 0;
        } else {
            thr = prefix;
        }
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
    const int tid = threadIdx.x;
    const int s = blockIdx.x, f = blockIdx.y;
    const int b0 = s * slice_len;
    const int len = min(slice_len, n_in - b0);
    if (len <= 0) return;
    const double *kin = keys_in + (size_t)f * in_stride + b0;
    const int *iin = idx_in ? idx_in + (size_t)f * in_stride + b0 : nullptr;
    for (int i = tid; i < len; i += TOPK_NT) keys[i] = topk_key(kin[i]);
    topk_body(keys, len, k, slice_len, iin, b0, s, f, keys_out, idx_out, out_stride, final_out, final_ld, tie);
}

// Fused finish + pin + top-k for designs of at most TOPK_LMAX columns: the CTA of chain c reduces the sweep partials and
// applies the sacrifice itself (finish_value), pins the always-include columns to DBL_MAX (utilities.cpp:190-199) and
// selects -- the sacrifice vector never goes to memory.  grid = (1, chains cmin..cmax); chains that already met the
// stopping rule are skipped.
template <int EPI>
__global__ void __launch_bounds__(TOPK_NT) topk_fused_kernel(const Dev d, int mode, int cmin, int k, const int *always,
                                                             int n_always)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    if (d.gate && *d.gate == 0) return;
    const int tid = threadIdx.x;
    const int c = cmin + blockIdx.y;
    if (d.done[c]) return;
    const int len = d.p;
    if (mode != MODE_COX) {
        constexpr int NB = 8;
        for (int i0 = tid; i0 < len; i0 += NB * TOPK_NT) {
            const int nb = min(NB, (len - i0 + TOPK_NT - 1) / TOPK_NT);
            double v[NB];
            finish_batch<EPI, NB>(d, mode, c, i0, TOPK_NT, nb, v);
#pragma unroll
            for (int q = 0; q < NB; q++)
                if (q < nb) keys[i0 + q * TOPK_NT] = topk_key(v[q]);
        }
    } else {
        for (int i = tid; i < len; i += TOPK_NT) keys[i] = topk_key(finish_value<EPI>(d, mode, c, i, nullptr));
    }
    if (n_always > 0) {
        __syncthreads();
        for (int q = tid; q < n_always; q += TOPK_NT) keys[always[q]] = (unsigned long long)__double_as_longlong(DBL_MAX);
    }
    topk_body(keys, len, k, len, nullptr, 0, 0, 0, nullptr, nullptr, 0, d.Anew + (size_t)c * d.kcap, d.kcap, d.tie + c);
}
void launch_topk_fused(const Dev &d, int mode, int epi, int cmin, int nspan, int k, const int *always, int n_always,
                       cudaStream_t st)
{
    if (k > d.p) throw EngineError{"top-k: k > number of candidates"};
    if (d.p > TOPK_LMAX) throw EngineError{"fused top-k: too many columns"};
    const size_t smem = (size_t)d.p * 8;
    dim3 grid(1, nspan);
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(topk_fused_kernel<EPI_SACR_LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_LMAX * 8));
        CUDA_CHECK(cudaFuncSetAttribute(topk_fused_kernel<EPI_SACR_GLM>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_LMAX * 8));
        CUDA_CHECK(cudaFuncSetAttribute(topk_fused_kernel<EPI_SACR_COX>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_LMAX * 8));
        configured = true;
    }
    if (epi == EPI_SACR_LM) topk_fused_kernel<EPI_SACR_LM><<<grid, TOPK_NT, smem, st>>>(d, mode, cmin, k, always, n_always);
    else if (epi == EPI_SACR_GLM) topk_fused_kernel<EPI_SACR_GLM><<<grid, TOPK_NT, smem, st>>>(d, mode, cmin, k, always, n_always);
    else topk_fused_kernel<EPI_SACR_COX><<<grid, TOPK_NT, smem, st>>>(d, mode, cmin, k, always, n_always);
    CUDA_CHECK(cudaGetLastError());
}

void launch_topk(const double *vals, long long stride, int n_in, int k, int nch, int *out_idx, int out_ld, int *tie,
                 double *ck0, int *ci0, double *ck1, int *ci1, long long cstride, cudaStream_t st, const int *gate,
                 const int *idx0)
{
    if (k > n_in) throw EngineError{"top-k: k > number of candidates"};
    const double *kin = vals;
    const int *iin = idx0;
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
// column-sharded mode: candidate exchange and active-column exchange (the collectives themselves are NCCL calls in
// engine.cu; these kernels pack / unpack their payloads)
// =====================================================================================================
__global__ void pack_candidates_kernel(const double *vals, long long stride, const int *sel, int sel_ld, int kloc, int kpad,
                                       long long offset, Cand *out, int out_ld, const int *gate)
{
    if (gate && *gate == 0) return;
    const int f = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= kpad) return;
    Cand cnd;
    if (a < kloc) {
        const int j = sel[(size_t)f * sel_ld + a];
        cnd.v = vals[(size_t)f * stride + j];
        cnd.idx = (long long)j + offset;
    } else {
        cnd.v = -1.0;
        cnd.idx = 2147483647LL;
    }
    out[(size_t)f * out_ld + a] = cnd;
}
void launch_pack_candidates(const double *vals, long long stride, const int *sel, int sel_ld, int kloc, int kpad,
                            long long offset, int nch, Cand *out, int out_ld, const int *gate, cudaStream_t st)
{
    dim3 grid((kpad + 127) / 128, nch);
    pack_candidates_kernel<<<grid, 128, 0, st>>>(vals, stride, sel, sel_ld, kloc, kpad, offset, out, out_ld, gate);
    CUDA_CHECK(cudaGetLastError());
}
__global__ void unpack_candidates_kernel(const Cand *in, int world, int nch, int in_ld, int k, double *mv, int *mi,
                                         long long mstride, const int *gate)
{
    if (gate && *gate == 0) return;
    const int f = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= world * k) return;
    const int q = e / k, a = e - q * k;
    const Cand cnd = in[((size_t)q * nch + f) * in_ld + a];
    mv[(size_t)f * mstride + e] = cnd.v;
    mi[(size_t)f * mstride + e] = (int)cnd.idx;
}
void launch_unpack_candidates(const Cand *in, int world, int nch, int in_ld, int k, double *mv, int *mi, long long mstride,
                              const int *gate, cudaStream_t st)
{
    dim3 grid((world * k + 127) / 128, nch);
    unpack_candidates_kernel<<<grid, 128, 0, st>>>(in, world, nch, in_ld, k, mv, mi, mstride, gate);
    CUDA_CHECK(cudaGetLastError());
}
__global__ void gather_active_kernel(const Dev d, const BatchDesc b, double *AXs)
{
    if (d.gate && *d.gate == 0) return;
    const int c = b.chain[blockIdx.y];
    if (d.done[c]) return;
    const int T = b.T;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= d.n * T) return;
    const int i = e / T, a = e - i * T;
    const int j = d.Anew[(size_t)c * d.kcap + a] - d.col_lo;
    AXs[((size_t)c * d.n + i) * T + a] = (j >= 0 && j < d.p) ? __ldg(d.X + (size_t)i * d.ldx + j) : 0.0;
}
void launch_gather_active(const Dev &d, const BatchDesc &b, double *AXs, cudaStream_t st)
{
    dim3 grid((d.n * b.T + 255) / 256, b.nch);
    gather_active_kernel<<<grid, 256, 0, st>>>(d, b, AXs);
    CUDA_CHECK(cudaGetLastError());
}
__global__ void gather_owned_cols_kernel(const double *X, long long ldx, int n, int p_local, long long col_lo, const int *sel,
                                         int m, double *Xn, long long ldn)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const long long j = (long long)sel[q] - col_lo;
    const bool own = j >= 0 && j < p_local;
    const int r0 = blockIdx.y * 32, r1 = min(n, r0 + 32);
    for (int i = r0; i < r1; i++) Xn[(size_t)i * ldn + q] = own ? X[(size_t)i * ldx + j] : 0.0;
}
void launch_gather_owned_cols(const double *X, long long ldx, int n, int p_local, long long col_lo, const int *sel, int m,
                              double *Xn, long long ldn, cudaStream_t st)
{
    dim3 grid((m + 127) / 128, (n + 31) / 32);
    gather_owned_cols_kernel<<<grid, 128, 0, st>>>(X, ldx, n, p_local, col_lo, sel, m, Xn, ldn);
    CUDA_CHECK(cudaGetLastError());
}

void configure_kernels()
{
    CUDA_CHECK(cudaFuncSetAttribute(topk_slices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_LMAX * 8));
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
        if (d.sharded) {
            const double *row = d.AXk + ((size_t)c * d.n + i) * d.ldXk;
            for (int a = 0; a < ks; a++) eta = fma(row[a], bsl[a], eta);
        } else {
            const double *row = d.X + (size_t)i * d.ldx;
            for (int a = 0; a < ks; a++) eta = fma(row[As[a]], bsl[a], eta);
        }
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
    // + 16: nvcc reads the index list two entries at a time (LDS.64) in the loop remainder, i.e. up to one int past the end
    const size_t smem = (size_t)d.kcap * (sizeof(double) + sizeof(int)) + 16;
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
    // eight rows in flight per thread: the pass is a pure stream (40 MB in, 40 MB out on a screened design), a dependent
    // load -> store per row left the memory system idle most of the time (39 us measured for 80 MB)
    for (int i0 = r0; i0 < r1; i0 += 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (i0 + u < r1) v[u] = *reinterpret_cast<const double2 *>(X + (size_t)(i0 + u) * ldx + j);
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (i0 + u < r1) {
                const double rm = rowmul ? rowmul[i0 + u] : 1.0;
                double2 o;
                o.x = (v[u].x - s0) * m0 * rm;
                o.y = two ? (v[u].y - s1) * m1 * rm : 0.0;
                *reinterpret_cast<double2 *>(X + (size_t)(i0 + u) * ldx + j) = o;
            }
    }
}
// normx_j = sqrt(h_j), mul_j = sqrt(n) / normx_j  (normalize.cpp:36-45; IEEE sqrt and division: the same bits as the host)
__global__ void norm_factors_kernel(const double *h, int p, double sn, double *norm_out, double *mul_out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const double nx = sqrt(h[j]);
    norm_out[j] = nx;
    mul_out[j] = sn / nx;
}
void launch_norm_factors(const double *h, int p, double sn, double *norm_out, double *mul_out, cudaStream_t st)
{
    norm_factors_kernel<<<(p + 255) / 256, 256, 0, st>>>(h, p, sn, norm_out, mul_out);
    CUDA_CHECK(cudaGetLastError());
}

// Normalisation of a design that fits L2 (the screened design of config 5: 40 MB) in ONE launch: a CTA owns 8 adjacent
// columns for all rows and makes the three passes of normalize.cpp:20-86 over its own slab (64 KB at n = 1000, re-read
// from L2 / L1) -- weighted column means, norms of the CENTRED columns, scaling (+ the sqrt(w) row factors of add_weight,
// Data.h:70-77).  Per element the arithmetic is that of center_scale_kernel ((x - mean) rounded once, then * mul * rm);
// only the summation order of the two column sums differs from the sweep kernels (fixed: rows warp-strided, warps added
// in order).  Replaces 7 launches and 6 passes over X by 1 launch and 4.
constexpr int NR_NT = 256;
constexpr int NR_TC = 8;  // columns per CTA: a warp load covers 4 rows x 64 bytes; p / 8 CTAs keep every SM streaming
__global__ void __launch_bounds__(NR_NT) normalize_resident_kernel(double *X, long long ldx, int n, int p, const double *gmean,
                                                                   const double *wnorm, const double *rowmul, double sn,
                                                                   double *mean_out, double *norm_out)
{
    __shared__ double red[NR_NT / 32][NR_TC];
    __shared__ double cmean[NR_TC], cmul[NR_TC];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = NR_NT / 32, RS = NW * 4;  // rows per CTA step
    const int c = lane & (NR_TC - 1), rg = lane >> 3;
    const long long j = (long long)blockIdx.x * NR_TC + c;
    const bool live = j < p;
    double *col = X + (live ? j : 0);
    const int ifirst = wid * 4 + rg;
    // column sum of f(x_i, i) over this thread's rows, 8 loads in flight; then lanes of the same column, then warps
    auto column_sum = [&](auto f) {
        double s = 0.0;
        for (int i0 = ifirst; i0 < n; i0 += 8 * RS) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = (live && i0 + u * RS < n) ? col[(size_t)(i0 + u * RS) * ldx] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (i0 + u * RS < n) s = f(v[u], i0 + u * RS, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        __syncthreads();
        if (lane < NR_TC) red[wid][lane] = s;
        __syncthreads();
        double t = 0.0;
        if (threadIdx.x < NR_TC) {
#pragma unroll
            for (int q = 0; q < NW; q++) t += red[q][threadIdx.x];
        }
        return t;  // valid on threads < NR_TC (thread = column of the tile)
    };
    // pass 1: mean_j = sum_i gmean_i x_ij  (gmean = w / n; nullptr: no centring, cox)
    double mean = 0.0;
    if (gmean) {
        const double t = column_sum([&](double v, int i, double s) { return fma(v, gmean[i], s); });
        if (threadIdx.x < NR_TC) {
            cmean[threadIdx.x] = t;
            if (live) mean_out[j] = t;
        }
        __syncthreads();
        mean = cmean[c];
    }
    // pass 2: h_j = sum_i wnorm_i (x_ij - mean_j)^2
    {
        const double t = column_sum([&](double v, int i, double s) {
            const double d = v - mean;
            return fma(d * d, wnorm[i], s);
        });
        if (threadIdx.x < NR_TC) {
            const double nx = sqrt(t);
            cmul[threadIdx.x] = sn / nx;
            if (live) norm_out[j] = nx;
        }
        __syncthreads();
    }
    // pass 3: x_ij <- (x_ij - mean_j) * mul_j * rowmul_i
    const double mul = cmul[c];
    for (int i0 = ifirst; i0 < n; i0 += 8 * RS) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (live && i0 + u * RS < n) ? col[(size_t)(i0 + u * RS) * ldx] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (live && i0 + u * RS < n) {
                const double rm = rowmul ? rowmul[i0 + u * RS] : 1.0;
                col[(size_t)(i0 + u * RS) * ldx] = (v[u] - mean) * mul * rm;
            }
    }
}
void launch_normalize_resident(double *X, long long ldx, int n, int p, const double *gmean, const double *wnorm,
                               const double *rowmul, double sn, double *mean_out, double *norm_out, cudaStream_t st)
{
    normalize_resident_kernel<<<(p + NR_TC - 1) / NR_TC, NR_NT, 0, st>>>(X, ldx, n, p, gmean, wnorm, rowmul, sn, mean_out, norm_out);
    CUDA_CHECK(cudaGetLastError());
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
    // every element is its own 32-byte sector somewhere in a multi-GB design: what matters is how many of these reads are
    // in flight, so a thread issues eight before it stores any
    for (int i0 = r0; i0 < r1; i0 += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (i0 + u < r1) v[u] = __ldg(X + (size_t)(i0 + u) * ldx + src);
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (i0 + u < r1) Xn[(size_t)(i0 + u) * ldn + jn] = v[u];
    }
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
