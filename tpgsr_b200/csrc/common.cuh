// Shared device/host helpers for the tpgsr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define TPGSR_OK 0
#define TPGSR_ERR_BAD_ARG (-1)
#define TPGSR_ERR_LAUNCH (-2)
#define TPGSR_ERR_WORKSPACE (-3)
#define TPGSR_ERR_UNSUPPORTED (-4)
#define TPGSR_ERR_DRIVER (-5)

// activation codes shared by conv epilogues / BN-apply / act-bwd
#define TPGSR_ACT_NONE 0
#define TPGSR_ACT_RELU 1
#define TPGSR_ACT_PRELU 2   // one shared slope, passed by device pointer
#define TPGSR_ACT_MISH 3
#define TPGSR_ACT_TANH 4

// output layout of conv epilogues
#define TPGSR_OUT_NHWC 0
#define TPGSR_OUT_NCHW 1

static inline int tpgsr_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TPGSR_OK : TPGSR_ERR_LAUNCH;
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float softplus_f(float x) {
    // F.softplus(beta=1, threshold=20)
    return x > 20.f ? x : log1pf(expf(x));
}
// mish(x) = x * tanh(softplus(x)) (model/tsrn.py:480-488).  With e = exp(x): tanh(log(1 + e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = n / (n + 2),
// n = e (e + 2) — one exponential and one division instead of expf + log1pf + tanhf (the PixelShuffle+mish and BN+mish kernels were
// SFU / ALU-bound on those, ncu sm__throughput 74-80%, profiles/r02/ncu_hbm_kernels.csv).  No cancellation anywhere: relative error ~1e-7.
// Above the softplus threshold (x > 20) tanh(x) == 1 in fp32.
__device__ __forceinline__ float mish_f(float x) {
    if (x > 20.f) return x;
    const float e = __expf(x);
    const float n = e * (e + 2.f);
    return x * __fdividef(n, n + 2.f);
}
__device__ __forceinline__ float mish_grad_f(float x) {
    // d/dx [x * tanh(sp(x))] = tanh(sp) + x * (1 - tanh(sp)^2) * sigmoid(x)   (sp' = sigmoid for x<=20, 1 above)
    if (x > 20.f) return 1.f;
    const float e = __expf(x);
    const float n = e * (e + 2.f);
    const float t = __fdividef(n, n + 2.f);
    const float sg = __fdividef(e, 1.f + e);
    return t + x * (1.f - t * t) * sg;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    switch (act) {
        case TPGSR_ACT_RELU: return v > 0.f ? v : 0.f;
        case TPGSR_ACT_PRELU: return v >= 0.f ? v : slope * v;
        case TPGSR_ACT_MISH: return mish_f(v);
        case TPGSR_ACT_TANH: return tanhf(v);
        default: return v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum of one double per thread (blockDim.x = 256): fixed shuffle tree inside the warps, then the eight warp
// totals are added in order by thread 0, which alone receives the result.  Used by the scalar "finalize" tails of the two-stage
// reductions (loss values, PReLU slope gradient, gradient norm) — a single thread looping over the partials cost 30-60 us per scalar.
__device__ __forceinline__ double block_sum_d256(double v) {
    __shared__ double sh_block_sum[8];
    v = warp_sum_d(v);
    if ((threadIdx.x & 31) == 0) sh_block_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sh_block_sum[i];
    }
    __syncthreads();
    return s;
}
