// Fused image loss (MSE + gradient-prior L1) and semantic (distillation) loss, forward and analytic backward.
//
// Reference ops replaced: ImageLoss.forward / GradientPriorLoss (loss/image_loss.py:19-51, ~25 ATen launches) and
// SemanticLoss.forward (loss/semantic_loss.py:21-38) with their autograd.
//   image:    L = w0 * mean((o-t)^2) + w1 * mean(|gm(o[:, :3]) - gm(t[:, :3])|),
//             gm(x) = sqrt(((r-l)/2)^2 + ((t-b)/2)^2 + 1e-6), zero padding.
//   semantic: L = mean(|gt - p|) + mean((gt+1e-20) * (log(gt+1e-20) - log(p+1e-20)))
// Tensors are NCHW contiguous.  HBM-bound: 131 072 B/img forward (SR + HR read once), +65 536 B/img backward.
// Reductions are two-stage and deterministic (per-block partials, double-precision finalize).
#include "common.cuh"

namespace {

__device__ __forceinline__ float gm_at(const float* __restrict__ p, int h, int w, int H, int W, float& dxh, float& dyh) {
    // p points at the (b,c) plane; returns gm and the half-differences
    float r = w + 1 < W ? p[h * W + w + 1] : 0.f;
    float l = w > 0 ? p[h * W + w - 1] : 0.f;
    float t = h > 0 ? p[(h - 1) * W + w] : 0.f;
    float b = h + 1 < H ? p[(h + 1) * W + w] : 0.f;
    dxh = (r - l) * 0.5f;
    dyh = (t - b) * 0.5f;
    return sqrtf(dxh * dxh + dyh * dyh + 1e-6f);
}

__global__ void __launch_bounds__(256)
image_loss_fwd_kernel(const float* __restrict__ o, const float* __restrict__ tg, int B, int C, int H, int W, int gradient,
                      float* __restrict__ partial) {
    const long total = (long)B * C * H * W;
    float s_mse = 0.f, s_gp = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int w = (int)(i % W);
        long q = i / W;
        int h = (int)(q % H);
        long pl = q / H;              // plane index b*C + c
        int c = (int)(pl % C);
        float d = o[i] - tg[i];
        s_mse += d * d;
        if (gradient && c < 3) {
            float a, e;
            float g1 = gm_at(o + pl * H * W, h, w, H, W, a, e);
            float g2 = gm_at(tg + pl * H * W, h, w, H, W, a, e);
            s_gp += fabsf(g1 - g2);
        }
    }
    __shared__ float sh[2][8];
    s_mse = warp_sum(s_mse);
    s_gp = warp_sum(s_gp);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s_mse; sh[1][threadIdx.x >> 5] = s_gp; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += sh[0][i]; b += sh[1][i]; }
        partial[blockIdx.x * 2] = a;
        partial[blockIdx.x * 2 + 1] = b;
    }
}

__global__ void __launch_bounds__(256) image_loss_finalize_kernel(const float* __restrict__ partial, int nblocks, double n_mse, double n_gp,
                                           float w0, float w1, float* __restrict__ loss, int accumulate, float out_scale) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) { a += partial[i * 2]; b += partial[i * 2 + 1]; }
    a = block_sum_d256(a);
    b = block_sum_d256(b);
    if (threadIdx.x) return;
    float v = (float)((double)w0 * (a / n_mse) + (n_gp > 0 ? (double)w1 * (b / n_gp) : 0.0)) * out_scale;
    loss[0] = accumulate ? loss[0] + v : v;
}

// signed, normalised half-differences of the gradient-prior term at pixel q:  sgn(gm_o - gm_t) * (dx/2, dy/2) / gm_o
__device__ __forceinline__ void gp_terms(const float* __restrict__ po, const float* __restrict__ pt, int h, int w, int H,
                                         int W, float& ax, float& ay) {
    float dxo, dyo, dxt, dyt;
    float go = gm_at(po, h, w, H, W, dxo, dyo);
    float gt = gm_at(pt, h, w, H, W, dxt, dyt);
    float s = go > gt ? 1.f : (go < gt ? -1.f : 0.f);
    ax = s * 0.5f * dxo / go;
    ay = s * 0.5f * dyo / go;
}

__global__ void __launch_bounds__(256)
image_loss_bwd_kernel(const float* __restrict__ o, const float* __restrict__ tg, const float* __restrict__ upstream,
                      float scale, int B, int C, int H, int W, int gradient, float w0, float w1, float* __restrict__ dout,
                      int accumulate) {
    const long total = (long)B * C * H * W;
    const float up = (upstream ? __ldg(upstream) : 1.f) * scale;
    const float k_mse = up * w0 * 2.f / (float)total;
    const float k_gp = up * w1 / (float)((long)B * (C < 3 ? C : 3) * H * W);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int w = (int)(i % W);
        long q = i / W;
        int h = (int)(q % H);
        long pl = q / H;
        int c = (int)(pl % C);
        float g = k_mse * (o[i] - tg[i]);
        if (gradient && c < 3) {
            const float* po = o + pl * H * W;
            const float* pt = tg + pl * H * W;
            float ax, ay, s = 0.f;
            // x(p) is the right neighbour of (h,w-1), the left neighbour of (h,w+1), the top neighbour of (h+1,w),
            // the bottom neighbour of (h-1,w)
            if (w > 0) { gp_terms(po, pt, h, w - 1, H, W, ax, ay); s += ax; }
            if (w + 1 < W) { gp_terms(po, pt, h, w + 1, H, W, ax, ay); s -= ax; }
            if (h + 1 < H) { gp_terms(po, pt, h + 1, w, H, W, ax, ay); s += ay; }
            if (h > 0) { gp_terms(po, pt, h - 1, w, H, W, ax, ay); s -= ay; }
            g += k_gp * s;
        }
        dout[i] = accumulate ? dout[i] + g : g;
    }
}

// ---------------------------------------------------------------------------------------------- semantic loss
__global__ void __launch_bounds__(256)
semantic_loss_fwd_kernel(const float* __restrict__ p, const float* __restrict__ gt, long n, float* __restrict__ partial) {
    float s1 = 0.f, s2 = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float a = p[i], g = gt[i];
        s1 += fabsf(g - a);
        float tgv = g + 1e-20f;
        s2 += tgv * (logf(tgv) - logf(a + 1e-20f));
    }
    __shared__ float sh[2][8];
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += sh[0][i]; b += sh[1][i]; }
        partial[blockIdx.x * 2] = a;
        partial[blockIdx.x * 2 + 1] = b;
    }
}

__global__ void __launch_bounds__(256) semantic_loss_finalize_kernel(const float* __restrict__ partial, int nblocks, double n, float* __restrict__ loss,
                                              int accumulate, float out_scale) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) { a += partial[i * 2]; b += partial[i * 2 + 1]; }
    a = block_sum_d256(a);
    b = block_sum_d256(b);
    if (threadIdx.x) return;
    float v = (float)(a / n + b / n) * out_scale;
    loss[0] = accumulate ? loss[0] + v : v;
}

__global__ void semantic_loss_bwd_kernel(const float* __restrict__ p, const float* __restrict__ gt, long n,
                                         const float* __restrict__ upstream, float scale, float* __restrict__ dp, int accumulate) {
    const float up = (upstream ? __ldg(upstream) : 1.f) * scale / (float)n;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float a = p[i], g = gt[i];
        float d = a > g ? 1.f : (a < g ? -1.f : 0.f);          // d|gt-p|/dp
        d -= (g + 1e-20f) / (a + 1e-20f);                       // d[-tg*log(p+eps)]/dp
        float v = up * d;
        dp[i] = accumulate ? dp[i] + v : v;
    }
}

inline int loss_grid(long n) {
    long b = (n + 255) / 256;
    return (int)(b > 592 ? 592 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" {

size_t tpgsr_loss_workspace_bytes(long n) { return (size_t)loss_grid(n) * 2 * sizeof(float); }

// loss[0] (+)= out_scale * ImageLoss(out, target); NCHW contiguous.
int tpgsr_image_loss_fwd(const float* out, const float* target, int B, int C, int H, int W, int gradient, float w0, float w1,
                         float out_scale, float* loss, int accumulate, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
    if (!out || !target || !loss || !workspace) return TPGSR_ERR_BAD_ARG;
    long n = (long)B * C * H * W;
    int grid = loss_grid(n);
    if (workspace_bytes < (size_t)grid * 2 * sizeof(float)) return TPGSR_ERR_WORKSPACE;
    float* partial = static_cast<float*>(workspace);
    image_loss_fwd_kernel<<<grid, 256, 0, stream>>>(out, target, B, C, H, W, gradient, partial);
    int rc = tpgsr_launch_status();
    if (rc) return rc;
    int cg = C < 3 ? C : 3;
    image_loss_finalize_kernel<<<1, 256, 0, stream>>>(partial, grid, (double)n, gradient ? (double)B * cg * H * W : 0.0, w0, w1,
                                                     loss, accumulate, out_scale);
    return tpgsr_launch_status();
}

// dout (+)= upstream[0] * scale * dImageLoss/dout   (upstream may be NULL = 1)
int tpgsr_image_loss_bwd(const float* out, const float* target, const float* upstream, float scale, int B, int C, int H, int W,
                         int gradient, float w0, float w1, float* dout, int accumulate, cudaStream_t stream) {
    if (!out || !target || !dout) return TPGSR_ERR_BAD_ARG;
    long n = (long)B * C * H * W;
    long b = (n + 255) / 256;
    image_loss_bwd_kernel<<<(int)(b > 148 * 16 ? 148 * 16 : b), 256, 0, stream>>>(out, target, upstream, scale, B, C, H, W,
                                                                                   gradient, w0, w1, dout, accumulate);
    return tpgsr_launch_status();
}

int tpgsr_semantic_loss_fwd(const float* pred, const float* gt, long n, float out_scale, float* loss, int accumulate,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!pred || !gt || !loss || !workspace) return TPGSR_ERR_BAD_ARG;
    int grid = loss_grid(n);
    if (workspace_bytes < (size_t)grid * 2 * sizeof(float)) return TPGSR_ERR_WORKSPACE;
    float* partial = static_cast<float*>(workspace);
    semantic_loss_fwd_kernel<<<grid, 256, 0, stream>>>(pred, gt, n, partial);
    int rc = tpgsr_launch_status();
    if (rc) return rc;
    semantic_loss_finalize_kernel<<<1, 256, 0, stream>>>(partial, grid, (double)n, loss, accumulate, out_scale);
    return tpgsr_launch_status();
}

int tpgsr_semantic_loss_bwd(const float* pred, const float* gt, long n, const float* upstream, float scale, float* dpred,
                            int accumulate, cudaStream_t stream) {
    if (!pred || !gt || !dpred) return TPGSR_ERR_BAD_ARG;
    semantic_loss_bwd_kernel<<<loss_grid(n), 256, 0, stream>>>(pred, gt, n, upstream, scale, dpred, accumulate);
    return tpgsr_launch_status();
}

}  // extern "C"
