// Thin-plate-spline rectification fused with the bilinear sampler (forward) and its gradient w.r.t. the control points.
//
// Reference ops replaced: TPSSpatialTransformer.forward (model/tps_spatial_transformer.py:97-112: cat with padding,
// two matmuls, clamp, 2x-1) + F.grid_sample(bilinear, zeros, align_corners=False) (model/tps_spatial_transformer.py:10-11)
// and their autograd.  The image is data (no gradient); only d/d(control points) is produced.
//
// Work decomposition (round 2; the first version ran one CTA per image and re-read the 94 KB coordinate representation for every image —
// 3.3x the algorithmic bytes, 51% / 27% of the measured HBM peak at B = 8192, profiles/r02/call1_stdout.log):
//   forward : CTA = (256-pixel chunk, group of G images).  A thread keeps the 23 representation values of ITS pixel in registers and walks
//             the G images: 46 FMAs from the per-image 23x2 mapping in shared memory, source coordinates out, four C-channel taps gathered
//             (NHWC: neighbouring pixels are neighbouring addresses), blended, stored.  HBM traffic per image = image in + image out +
//             coordinates = the 40 960 algorithmic bytes of SURVEY §8d (+ 94 KB / G of representation).
//   backward: CTA = group of G images, whole map (4 pixels per thread at 16x64), the representation staged ONCE per CTA in shared memory;
//             per image the 46 partial sums are reduced with warp shuffles and a fixed-order cross-warp sum (deterministic).
// G = 1 at the bench batch (192 CTAs forward, 48 backward) and 2 forward from B = 1024; the backward kernel always takes one image per CTA.
#include "common.cuh"

namespace {

constexpr int NK = 23;       // 20 control points + 3 affine terms
constexpr int NCTRL = 20;
constexpr int GMAX = 8;      // images per CTA (upper bound)

template <int C>
struct Px { float v[C]; };

template <int C>
__device__ __forceinline__ Px<C> ld_px(const float* img, int H, int W, long long b, int yy, int xx) {
    Px<C> o;
    if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) {
        const float* p = img + ((size_t)(b * H + yy) * W + xx) * C;
        if (C == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p));
            o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[C - 1] = t.w;
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) o.v[c] = __ldg(p + c);
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) o.v[c] = 0.f;
    }
    return o;
}

// M[g][k][c] = sum_{j<20} inv_kernel[k][j] * ctrl[b0+g][j][c]   (the 3 padding rows of Y are zero)
__device__ __forceinline__ void build_mappings(const float* inv_kernel, const float* ctrl, long long b0, int ng, float (*Msm)[NK][2]) {
    for (int i = threadIdx.x; i < ng * NK * 2; i += blockDim.x) {
        const int g = i / (NK * 2), r = i - g * NK * 2, k = r >> 1, c = r & 1;
        const float* cb = ctrl + (size_t)(b0 + g) * NCTRL * 2;
        float s = 0.f;
        for (int j = 0; j < NCTRL; ++j) s = fmaf(__ldg(inv_kernel + k * NK + j), __ldg(cb + j * 2 + c), s);
        Msm[g][k][c] = s;
    }
}

// sampling position of a source coordinate: grid = 2*clamp(s,0,1)-1, unnormalised with align_corners=False
__device__ __forceinline__ void sample_pos(float sx, float sy, int H, int W, int& x0, int& y0, float& tx, float& ty) {
    const float gx = 2.0f * fminf(fmaxf(sx, 0.f), 1.f) - 1.0f;
    const float gy = 2.0f * fminf(fmaxf(sy, 0.f), 1.f) - 1.0f;
    const float ix = ((gx + 1.f) * W - 1.f) / 2.f;
    const float iy = ((gy + 1.f) * H - 1.f) / 2.f;
    const float x0f = floorf(ix), y0f = floorf(iy);
    x0 = (int)x0f; y0 = (int)y0f;
    tx = ix - x0f; ty = iy - y0f;
}

template <int C>
__global__ void __launch_bounds__(256)
tps_sample_fwd_kernel(const float* __restrict__ img, const float* __restrict__ ctrl, const float* __restrict__ inv_kernel,
                      const float* __restrict__ reprT, float* __restrict__ out, float* __restrict__ src, int B, int H, int W, int G) {
    const int HW = H * W;
    const long long b0 = (long long)blockIdx.y * G;
    const int ng = (B - b0) < G ? (int)(B - b0) : G;
    __shared__ float Msm[GMAX][NK][2];
    build_mappings(inv_kernel, ctrl, b0, ng, Msm);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    float r[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) r[k] = p < HW ? __ldg(reprT + (size_t)k * HW + p) : 0.f;
    __syncthreads();
    if (p >= HW) return;
    for (int g = 0; g < ng; ++g) {
        const long long b = b0 + g;
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            sx = fmaf(r[k], Msm[g][k][0], sx);
            sy = fmaf(r[k], Msm[g][k][1], sy);
        }
        *reinterpret_cast<float2*>(src + ((size_t)b * HW + p) * 2) = make_float2(sx, sy);
        int x0, y0;
        float tx, ty;
        sample_pos(sx, sy, H, W, x0, y0, tx, ty);
        const Px<C> nw = ld_px<C>(img, H, W, b, y0, x0), ne = ld_px<C>(img, H, W, b, y0, x0 + 1);
        const Px<C> sw = ld_px<C>(img, H, W, b, y0 + 1, x0), se = ld_px<C>(img, H, W, b, y0 + 1, x0 + 1);
        const float wnw = (1.f - tx) * (1.f - ty), wne = tx * (1.f - ty), wsw = (1.f - tx) * ty, wse = tx * ty;
        float o[C];
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = nw.v[c] * wnw + ne.v[c] * wne + sw.v[c] * wsw + se.v[c] * wse;
        float* dst = out + ((size_t)b * HW + p) * C;
        if (C == 4) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[C - 1]);
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) dst[c] = o[c];
        }
    }
}

// dctrl[b][j][c] = sum_k inv_kernel[k][j] * sum_p reprT[k][p] * dsrc[p][c]
template <int C>
__global__ void __launch_bounds__(256)
tps_sample_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ img, const float* __restrict__ src,
                      const float* __restrict__ inv_kernel, const float* __restrict__ reprT, float* __restrict__ dctrl,
                      int B, int H, int W, int G, int accumulate) {
    extern __shared__ float rsm[];                       // [NK][HW] when it fits (G > 1), else unused
    const int HW = H * W;
    const long long b0 = (long long)blockIdx.x * G;
    const int ng = (B - b0) < G ? (int)(B - b0) : G;
    const bool staged = G > 1;
    if (staged) {
        for (int i = threadIdx.x; i < NK * HW; i += blockDim.x) rsm[i] = __ldg(reprT + i);
        __syncthreads();
    }
    __shared__ float red[8][NK * 2];
    __shared__ float dM[NK * 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = 0; g < ng; ++g) {
        const long long b = b0 + g;
        float acc[NK][2];
#pragma unroll
        for (int k = 0; k < NK; ++k) acc[k][0] = acc[k][1] = 0.f;
        for (int p = threadIdx.x; p < HW; p += blockDim.x) {
            const float2 s2 = *reinterpret_cast<const float2*>(src + ((size_t)b * HW + p) * 2);
            int x0, y0;
            float tx, ty;
            sample_pos(s2.x, s2.y, H, W, x0, y0, tx, ty);
            const Px<C> nw = ld_px<C>(img, H, W, b, y0, x0), ne = ld_px<C>(img, H, W, b, y0, x0 + 1);
            const Px<C> sw = ld_px<C>(img, H, W, b, y0 + 1, x0), se = ld_px<C>(img, H, W, b, y0 + 1, x0 + 1);
            const Px<C> gq = ld_px<C>(dout, 1, HW, b, 0, p);
            float gix = 0.f, giy = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                gix += gq.v[c] * ((ne.v[c] - nw.v[c]) * (1.f - ty) + (se.v[c] - sw.v[c]) * ty);
                giy += gq.v[c] * ((sw.v[c] - nw.v[c]) * (1.f - tx) + (se.v[c] - ne.v[c]) * tx);
            }
            // unnormalise (W/2), grid = 2*clamp(s)-1 (x2, gradient passes on the closed interval [0,1])
            const float dsx = (s2.x >= 0.f && s2.x <= 1.f) ? gix * (float)W : 0.f;
            const float dsy = (s2.y >= 0.f && s2.y <= 1.f) ? giy * (float)H : 0.f;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const float r = staged ? rsm[k * HW + p] : __ldg(reprT + (size_t)k * HW + p);
                acc[k][0] = fmaf(r, dsx, acc[k][0]);
                acc[k][1] = fmaf(r, dsy, acc[k][1]);
            }
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const float a = warp_sum(acc[k][0]), c = warp_sum(acc[k][1]);
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
            const int j = threadIdx.x >> 1, c = threadIdx.x & 1;
            float s = 0.f;
            for (int k = 0; k < NK; ++k) s = fmaf(__ldg(inv_kernel + k * NK + j), dM[k * 2 + c], s);
            float* d = dctrl + (size_t)b * NCTRL * 2 + threadIdx.x;
            *d = accumulate ? *d + s : s;
        }
        __syncthreads();
    }
}

inline int tps_group(int B) { return B >= 1024 ? 2 : 1; }

}  // namespace

extern "C" {

// img: NHWC [B,H,W,C], C = 3 (mask=False) or 4; ctrl: [B,20,2]; inv_kernel: [23,23]; reprT: [23, H*W] (transposed target_coordinate_repr);
// out: NHWC [B,H,W,C]; src: [B,H*W,2] (source coordinates before the clamp — also TPS.forward's 2nd return value).
int tpgsr_tps_sample_fwd(const float* img, const float* ctrl, const float* inv_kernel, const float* reprT, float* out,
                         float* src, int B, int H, int W, int C, cudaStream_t stream) {
    if (!img || !ctrl || !inv_kernel || !reprT || !out || !src || B <= 0) return TPGSR_ERR_BAD_ARG;
    if (C != 4 && C != 3) return TPGSR_ERR_UNSUPPORTED;
    const int G = tps_group(B);
    dim3 grid((H * W + 255) / 256, (B + G - 1) / G);
    if (C == 4) tps_sample_fwd_kernel<4><<<grid, 256, 0, stream>>>(img, ctrl, inv_kernel, reprT, out, src, B, H, W, G);
    else tps_sample_fwd_kernel<3><<<grid, 256, 0, stream>>>(img, ctrl, inv_kernel, reprT, out, src, B, H, W, G);
    return tpgsr_launch_status();
}

int tpgsr_tps_sample_bwd(const float* dout, const float* img, const float* src, const float* inv_kernel, const float* reprT,
                         float* dctrl, int B, int H, int W, int C, int accumulate, cudaStream_t stream) {
    if (!dout || !img || !src || !inv_kernel || !reprT || !dctrl || B <= 0) return TPGSR_ERR_BAD_ARG;
    if (C != 4 && C != 3) return TPGSR_ERR_UNSUPPORTED;
    // one image per CTA: staging the representation in shared memory for several images (94 KB, 2 CTAs / SM) was measured SLOWER than
    // re-reading it through L2 (426 us against 224 us at B = 8192, profiles/r02/call6_stdout.log); the staged path stays for maps it helps
    const int G = 1;
    const size_t smem = 0;
    auto kern = C == 4 ? tps_sample_bwd_kernel<4> : tps_sample_bwd_kernel<3>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TPGSR_ERR_LAUNCH;
    kern<<<(B + G - 1) / G, 256, smem, stream>>>(dout, img, src, inv_kernel, reprT, dctrl, B, H, W, G, accumulate);
    return tpgsr_launch_status();
}

}  // extern "C"
