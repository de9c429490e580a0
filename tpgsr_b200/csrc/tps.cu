// Thin-plate-spline rectification fused with the bilinear sampler (forward) and its gradient w.r.t. the control points.
//
// Reference ops replaced: TPSSpatialTransformer.forward (model/tps_spatial_transformer.py:97-112: cat with padding,
// two matmuls, clamp, 2x-1) + F.grid_sample(bilinear, zeros, align_corners=False) (model/tps_spatial_transformer.py:10-11)
// and their autograd.  The image is data (no gradient); only d/d(control points) is produced.
//
// One CTA per image: the 23x2 mapping matrix is built in shared memory from inverse_kernel[:, :20] @ ctrl, each thread then
// evaluates the 23-term dot product for its pixels from the transposed coordinate representation reprT[23][HW] (coalesced),
// and gathers four float4 taps (NHWC, C=4).  HBM-bound in principle (40 960 B/img fwd) but L2-resident at B=48.
#include "common.cuh"

namespace {

constexpr int NK = 23;       // 20 control points + 3 affine terms
constexpr int NCTRL = 20;

__device__ __forceinline__ float4 ld4_or_zero(const float* img, int H, int W, int b, int yy, int xx) {
    if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W)
        return __ldg(reinterpret_cast<const float4*>(img + ((size_t)(b * H + yy) * W + xx) * 4));
    return make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void build_mapping(const float* inv_kernel, const float* ctrl_b, float (*Msm)[2]) {
    // M[k][c] = sum_{j<20} inv_kernel[k][j] * ctrl[j][c]   (the 3 padding rows of Y are zero)
    for (int i = threadIdx.x; i < NK * 2; i += blockDim.x) {
        int k = i >> 1, c = i & 1;
        float s = 0.f;
        for (int j = 0; j < NCTRL; ++j) s = fmaf(__ldg(inv_kernel + k * NK + j), ctrl_b[j * 2 + c], s);
        Msm[k][c] = s;
    }
}

__global__ void __launch_bounds__(256)
tps_sample_fwd_kernel(const float* __restrict__ img, const float* __restrict__ ctrl, const float* __restrict__ inv_kernel,
                      const float* __restrict__ reprT, float* __restrict__ out, float* __restrict__ src, int H, int W) {
    const int b = blockIdx.x, HW = H * W;
    __shared__ float Msm[NK][2];
    build_mapping(inv_kernel, ctrl + (size_t)b * NCTRL * 2, Msm);
    __syncthreads();
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            float r = __ldg(reprT + (size_t)k * HW + p);
            sx = fmaf(r, Msm[k][0], sx);
            sy = fmaf(r, Msm[k][1], sy);
        }
        src[((size_t)b * HW + p) * 2 + 0] = sx;
        src[((size_t)b * HW + p) * 2 + 1] = sy;
        float gx = 2.0f * fminf(fmaxf(sx, 0.f), 1.f) - 1.0f;
        float gy = 2.0f * fminf(fmaxf(sy, 0.f), 1.f) - 1.0f;
        float ix = ((gx + 1.f) * W - 1.f) / 2.f;
        float iy = ((gy + 1.f) * H - 1.f) / 2.f;
        float x0f = floorf(ix), y0f = floorf(iy);
        int x0 = (int)x0f, y0 = (int)y0f;
        float tx = ix - x0f, ty = iy - y0f;
        float4 nw = ld4_or_zero(img, H, W, b, y0, x0), ne = ld4_or_zero(img, H, W, b, y0, x0 + 1);
        float4 sw = ld4_or_zero(img, H, W, b, y0 + 1, x0), se = ld4_or_zero(img, H, W, b, y0 + 1, x0 + 1);
        float wnw = (1.f - tx) * (1.f - ty), wne = tx * (1.f - ty), wsw = (1.f - tx) * ty, wse = tx * ty;
        float4 o;
        o.x = nw.x * wnw + ne.x * wne + sw.x * wsw + se.x * wse;
        o.y = nw.y * wnw + ne.y * wne + sw.y * wsw + se.y * wse;
        o.z = nw.z * wnw + ne.z * wne + sw.z * wsw + se.z * wse;
        o.w = nw.w * wnw + ne.w * wne + sw.w * wsw + se.w * wse;
        *reinterpret_cast<float4*>(out + ((size_t)b * HW + p) * 4) = o;
    }
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// dctrl[b][j][c] = sum_k inv_kernel[k][j] * sum_p reprT[k][p] * dsrc[p][c]
__global__ void __launch_bounds__(256)
tps_sample_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ img, const float* __restrict__ src,
                      const float* __restrict__ inv_kernel, const float* __restrict__ reprT, float* __restrict__ dctrl,
                      int H, int W, int accumulate) {
    const int b = blockIdx.x, HW = H * W;
    float acc[NK][2];
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k][0] = acc[k][1] = 0.f;
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
        float sx = src[((size_t)b * HW + p) * 2 + 0], sy = src[((size_t)b * HW + p) * 2 + 1];
        float gx = 2.0f * fminf(fmaxf(sx, 0.f), 1.f) - 1.0f;
        float gy = 2.0f * fminf(fmaxf(sy, 0.f), 1.f) - 1.0f;
        float ix = ((gx + 1.f) * W - 1.f) / 2.f;
        float iy = ((gy + 1.f) * H - 1.f) / 2.f;
        float x0f = floorf(ix), y0f = floorf(iy);
        int x0 = (int)x0f, y0 = (int)y0f;
        float tx = ix - x0f, ty = iy - y0f;
        float4 nw = ld4_or_zero(img, H, W, b, y0, x0), ne = ld4_or_zero(img, H, W, b, y0, x0 + 1);
        float4 sw = ld4_or_zero(img, H, W, b, y0 + 1, x0), se = ld4_or_zero(img, H, W, b, y0 + 1, x0 + 1);
        float4 g = __ldg(reinterpret_cast<const float4*>(dout + ((size_t)b * HW + p) * 4));
        float gix = dot4(g, sub4(ne, nw)) * (1.f - ty) + dot4(g, sub4(se, sw)) * ty;
        float giy = dot4(g, sub4(sw, nw)) * (1.f - tx) + dot4(g, sub4(se, ne)) * tx;
        // unnormalise (W/2), grid = 2*clamp(s)-1 (x2, gradient passes on the closed interval [0,1])
        float dsx = (sx >= 0.f && sx <= 1.f) ? gix * (float)W : 0.f;
        float dsy = (sy >= 0.f && sy <= 1.f) ? giy * (float)H : 0.f;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            float r = __ldg(reprT + (size_t)k * HW + p);
            acc[k][0] = fmaf(r, dsx, acc[k][0]);
            acc[k][1] = fmaf(r, dsy, acc[k][1]);
        }
    }
    __shared__ float red[8][NK * 2];
    __shared__ float dM[NK * 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        float a = warp_sum(acc[k][0]), c = warp_sum(acc[k][1]);
        if (lane == 0) { red[warp][k * 2] = a; red[warp][k * 2 + 1] = c; }
    }
    __syncthreads();
    if (threadIdx.x < NK * 2) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
        dM[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x < NCTRL * 2) {
        int j = threadIdx.x >> 1, c = threadIdx.x & 1;
        float s = 0.f;
        for (int k = 0; k < NK; ++k) s = fmaf(__ldg(inv_kernel + k * NK + j), dM[k * 2 + c], s);
        float* d = dctrl + (size_t)b * NCTRL * 2 + threadIdx.x;
        *d = accumulate ? *d + s : s;
    }
}

}  // namespace

extern "C" {

// img: NHWC [B,H,W,4]; ctrl: [B,20,2]; inv_kernel: [23,23]; reprT: [23, H*W] (transposed target_coordinate_repr);
// out: NHWC [B,H,W,4]; src: [B,H*W,2] (source coordinates before the clamp — also TPS.forward's 2nd return value).
int tpgsr_tps_sample_fwd(const float* img, const float* ctrl, const float* inv_kernel, const float* reprT, float* out,
                         float* src, int B, int H, int W, int C, cudaStream_t stream) {
    if (!img || !ctrl || !inv_kernel || !reprT || !out || !src || B <= 0) return TPGSR_ERR_BAD_ARG;
    if (C != 4) return TPGSR_ERR_UNSUPPORTED;
    tps_sample_fwd_kernel<<<B, 256, 0, stream>>>(img, ctrl, inv_kernel, reprT, out, src, H, W);
    return tpgsr_launch_status();
}

int tpgsr_tps_sample_bwd(const float* dout, const float* img, const float* src, const float* inv_kernel, const float* reprT,
                         float* dctrl, int B, int H, int W, int C, int accumulate, cudaStream_t stream) {
    if (!dout || !img || !src || !inv_kernel || !reprT || !dctrl || B <= 0) return TPGSR_ERR_BAD_ARG;
    if (C != 4) return TPGSR_ERR_UNSUPPORTED;
    tps_sample_bwd_kernel<<<B, 256, 0, stream>>>(dout, img, src, inv_kernel, reprT, dctrl, H, W, accumulate);
    return tpgsr_launch_status();
}

}  // extern "C"
