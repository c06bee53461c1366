// Device-side helpers shared by kernels.cu and chain_fit.cu: deterministic block reductions and scans.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "engine.h"

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


}  // namespace bess
