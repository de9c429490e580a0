// Flat-buffer optimiser step: global-norm clipping of a sub-range + Adam, on one fp32 parameter/gradient/moment buffer.
//
// Reference ops replaced: torch.nn.utils.clip_grad_norm_(model.parameters(), 0.25) over the SR model only
// (interfaces/super_resolution.py:422-423) and optim.Adam(lr, betas=(beta1, 0.999)).step() over SR + student parameters
// (interfaces/base.py:449-450), i.e. the foreach ATen kernels.  HBM-bound: 28 B/param (read p,g,m,v; write p,m,v).
// The step counter and bias corrections live on the device so the whole train step can be replayed as a CUDA graph.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n, float* __restrict__ partial) {
    float s = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s = fmaf(g[i], g[i], s);
    __shared__ float sh[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += sh[i];
        partial[blockIdx.x] = a;
    }
}

// state[0] = step (as float), [1] = 1-beta1^t, [2] = 1-beta2^t, [3] = last grad norm, [4] = clip coefficient
__global__ void __launch_bounds__(256) optim_prepare_kernel(const float* __restrict__ partial, int nblocks, float max_norm, float beta1,
                                                            float beta2, float* __restrict__ state, int advance, float grad_scale) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[i];
    s = block_sum_d256(s);
    if (threadIdx.x) return;
    float norm = (float)sqrt(s) * grad_scale;
    float coef = max_norm / (norm + 1e-6f);
    state[3] = norm;
    state[4] = max_norm > 0.f ? (coef < 1.f ? coef : 1.f) : 1.f;
    if (advance) {
        float t = state[0] + 1.f;
        state[0] = t;
        state[1] = (float)(1.0 - pow((double)beta1, (double)t));
        state[2] = (float)(1.0 - pow((double)beta2, (double)t));
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
            long n_clipped, const float* __restrict__ state, float lr, float beta1, float beta2, float eps, float grad_scale) {
    const float bc1 = state[1], bc2 = state[2], coef = state[4];
    const float step_size = lr / bc1;
    const float bc2_sqrt = sqrtf(bc2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        if (i < n_clipped) gi *= coef;
        float mi = m[i], vi = v[i];
        mi = mi + (gi - mi) * (1.f - beta1);                 // exp_avg.lerp_(grad, 1-beta1)
        vi = vi * beta2 + gi * gi * (1.f - beta2);           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (mi / denom);
        m[i] = mi;
        v[i] = vi;
    }
}

// float4 version (16-byte aligned buffers): four 16-byte streams in flight per thread instead of four 4-byte ones; the arithmetic per element
// is the scalar kernel's, operation for operation (IEEE sqrt and division: the Adam update is compared element-wise against the reference)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, bool clipped, float coef, float step_size, float bc2_sqrt,
                                         float beta1, float beta2, float eps, float grad_scale) {
    float gi = g * grad_scale;
    if (clipped) gi *= coef;
    float mi = m, vi = v;
    mi = mi + (gi - mi) * (1.f - beta1);
    vi = vi * beta2 + gi * gi * (1.f - beta2);
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p = p - step_size * (mi / denom);
    m = mi;
    v = vi;
}
__global__ void __launch_bounds__(256)
adam_v4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4,
               long n_clipped, const float* __restrict__ state, float lr, float beta1, float beta2, float eps, float grad_scale) {
    const float bc1 = state[1], bc2 = state[2], coef = state[4];
    const float step_size = lr / bc1;
    const float bc2_sqrt = sqrtf(bc2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pi = p[i], mi = m[i], vi = v[i];
        const float4 gi = g[i];
        const long e = 4 * i;
        adam_one(pi.x, gi.x, mi.x, vi.x, e < n_clipped, coef, step_size, bc2_sqrt, beta1, beta2, eps, grad_scale);
        adam_one(pi.y, gi.y, mi.y, vi.y, e + 1 < n_clipped, coef, step_size, bc2_sqrt, beta1, beta2, eps, grad_scale);
        adam_one(pi.z, gi.z, mi.z, vi.z, e + 2 < n_clipped, coef, step_size, bc2_sqrt, beta1, beta2, eps, grad_scale);
        adam_one(pi.w, gi.w, mi.w, vi.w, e + 3 < n_clipped, coef, step_size, bc2_sqrt, beta1, beta2, eps, grad_scale);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

__global__ void scale_kernel(float* __restrict__ x, long n, const float* __restrict__ coef_p, float s) {
    float c = (coef_p ? __ldg(coef_p) : 1.f) * s;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] *= c;
}

inline int opt_grid(long n) {
    long b = (n + 255) / 256;
    return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" {

size_t tpgsr_optim_workspace_bytes(long n) { return (size_t)opt_grid(n) * sizeof(float); }

// Computes grad_scale*||g[0:n_clipped]||_2 -> state[3], clip coefficient -> state[4]; if advance, step counter/bias corrections.
int tpgsr_optim_prepare(const float* grads, long n_clipped, float max_norm, float beta1, float beta2, float grad_scale,
                        float* state, int advance, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!grads || !state || !workspace) return TPGSR_ERR_BAD_ARG;
    int grid = opt_grid(n_clipped > 0 ? n_clipped : 1);
    if (workspace_bytes < (size_t)grid * sizeof(float)) return TPGSR_ERR_WORKSPACE;
    sumsq_kernel<<<grid, 256, 0, stream>>>(grads, n_clipped, static_cast<float*>(workspace));
    int rc = tpgsr_launch_status();
    if (rc) return rc;
    optim_prepare_kernel<<<1, 256, 0, stream>>>(static_cast<float*>(workspace), grid, max_norm, beta1, beta2, state, advance, grad_scale);
    return tpgsr_launch_status();
}

// One Adam step over the flat buffers; gradients in [0, n_clipped) are scaled by the clip coefficient in state[4] first,
// every gradient by grad_scale (1/world_size after a sum all-reduce).
int tpgsr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, long n_clipped,
                    const float* state, float lr, float beta1, float beta2, float eps, float grad_scale, cudaStream_t stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !state || n <= 0) return TPGSR_ERR_BAD_ARG;
    if (!(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) && n >= 4) {
        const long n4 = n / 4;
        adam_v4_kernel<<<opt_grid(n4), 256, 0, stream>>>(reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads), reinterpret_cast<float4*>(exp_avg),
                                                         reinterpret_cast<float4*>(exp_avg_sq), n4, n_clipped, state, lr, beta1, beta2, eps, grad_scale);
        const long done = 4 * n4;
        if (done < n) {      // up to three tail elements
            const long nc = n_clipped > done ? n_clipped - done : 0;
            adam_kernel<<<1, 32, 0, stream>>>(params + done, grads + done, exp_avg + done, exp_avg_sq + done, n - done, nc, state, lr, beta1, beta2, eps, grad_scale);
        }
        return tpgsr_launch_status();
    }
    adam_kernel<<<opt_grid(n), 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, n_clipped, state, lr, beta1, beta2, eps, grad_scale);
    return tpgsr_launch_status();
}

// x *= coef[0] * s   (applies the clip coefficient to .grad in place, for callers that keep torch.optim)
int tpgsr_scale_inplace(float* x, long n, const float* coef, float s, cudaStream_t stream) {
    if (!x) return TPGSR_ERR_BAD_ARG;
    scale_kernel<<<opt_grid(n), 256, 0, stream>>>(x, n, coef, s);
    return tpgsr_launch_status();
}

}  // extern "C"
