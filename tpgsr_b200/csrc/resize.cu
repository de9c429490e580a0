// Resampling kernels of the text-prior branch (HBM/L2-bound elementwise gathers, deterministic gather-form backward).
//
// Reference ops replaced:
//   * TextBase.parse_crnn_data (interfaces/base.py:806-829): F.interpolate(bicubic, align_corners=False, A=-0.75, clamped
//     neighbour indices) of RGB to (32,100) followed by 0.299 R + 0.587 G + 0.114 B        -> tpgsr_bicubic_gray_fwd/bwd
//   * F.interpolate(mode='bilinear', align_corners=True) of the InfoGen map [B,32,1,203] to (16,64) (model/tsrn.py:196): every
//     output row equals the single input row                                                  -> tpgsr_bilinear_rows_fwd/bwd
//   * the stride-2 ConvTranspose2d layers of InfoGen (model/tsrn.py:89-106) on H=1 maps are run as stride-1 convolutions
//     over a zero-upsampled input with the flipped middle kernel row                          -> tpgsr_zero_upsample_w(+bwd),
//                                                                                               tpgsr_tconv_weight_fwd/bwd
//   * softmax + [T,B,C] -> [B,1,T,C] prior formatting + prior dropout (interfaces/super_resolution.py:319-321,376-382)
//                                                                                             -> tpgsr_softmax_prior_fwd/bwd
#include "common.cuh"

namespace {

__device__ __forceinline__ void cubic_coeffs(float t, float* w) {
    const float A = -0.75f;
    float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
    w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
    w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
    w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
    w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

__device__ __forceinline__ void cubic_src(int o, float scale, int n_in, int* idx, float* w) {
    // ATen area_pixel_compute_source_index(align_corners=False, cubic=True): src = scale*(dst+0.5)-0.5 (no clamp to 0)
    float src = scale * (o + 0.5f) - 0.5f;
    float fl = floorf(src);
    cubic_coeffs(src - fl, w);
    int i0 = (int)fl;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int i = i0 - 1 + k;
        idx[k] = i < 0 ? 0 : (i > n_in - 1 ? n_in - 1 : i);
    }
}

// in: NCHW view (strides in floats), channels 0..2 used; out: [B,1,Ho,Wo] dense
__global__ void __launch_bounds__(256)
bicubic_gray_fwd_kernel(const float* __restrict__ in, long sb, long sc, long sh, long sw, int B, int Hi, int Wi,
                        float* __restrict__ out, int Ho, int Wo) {
    const long total = (long)B * Ho * Wo;
    const float sy = (float)Hi / Ho, sx = (float)Wi / Wo;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int wo = (int)(i % Wo);
        long q = i / Wo;
        int ho = (int)(q % Ho);
        int b = (int)(q / Ho);
        int iy[4], ix[4];
        float wy[4], wx[4];
        cubic_src(ho, sy, Hi, iy, wy);
        cubic_src(wo, sx, Wi, ix, wx);
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* p = in + b * sb + c * sc;
            float s = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float r = 0.f;
#pragma unroll
                for (int d = 0; d < 4; ++d) r += wx[d] * __ldg(p + iy[a] * sh + ix[d] * sw);
                s += wy[a] * r;
            }
            acc[c] = s;
        }
        out[i] = 0.299f * acc[0] + 0.587f * acc[1] + 0.114f * acc[2];
    }
}

// gather-form backward: every input pixel (b,c<3,h,w) sums the output gradients whose 4x4 footprint touches it.
// Output rows/cols that can touch input index i:  o in [ (i-2)/scale - 1, (i+3)/scale + 1 ]  (conservative window, exact test inside)
__global__ void __launch_bounds__(256)
bicubic_gray_bwd_kernel(const float* __restrict__ dout, int B, int Hi, int Wi, int Ho, int Wo, float* __restrict__ din,
                        long sb, long sc, long sh, long sw, int accumulate) {
    const long total = (long)B * 3 * Hi * Wi;
    const float sy = (float)Hi / Ho, sx = (float)Wi / Wo;
    const float gray[3] = {0.299f, 0.587f, 0.114f};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int w = (int)(i % Wi);
        long q = i / Wi;
        int h = (int)(q % Hi);
        long r2 = q / Hi;
        int c = (int)(r2 % 3);
        int b = (int)(r2 / 3);
        int ho_lo = (int)floorf((h - 2.5f) / sy) - 1, ho_hi = (int)ceilf((h + 2.5f) / sy) + 1;
        int wo_lo = (int)floorf((w - 2.5f) / sx) - 1, wo_hi = (int)ceilf((w + 2.5f) / sx) + 1;
        // clamped border indices collect every footprint that runs off the edge
        if (h == 0) ho_lo = 0;
        if (h == Hi - 1) ho_hi = Ho - 1;
        if (w == 0) wo_lo = 0;
        if (w == Wi - 1) wo_hi = Wo - 1;
        ho_lo = ho_lo < 0 ? 0 : ho_lo; ho_hi = ho_hi > Ho - 1 ? Ho - 1 : ho_hi;
        wo_lo = wo_lo < 0 ? 0 : wo_lo; wo_hi = wo_hi > Wo - 1 ? Wo - 1 : wo_hi;
        float s = 0.f;
        for (int ho = ho_lo; ho <= ho_hi; ++ho) {
            int iy[4];
            float wy[4];
            cubic_src(ho, sy, Hi, iy, wy);
            float cy = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) cy += iy[a] == h ? wy[a] : 0.f;
            if (cy == 0.f) continue;
            float rs = 0.f;
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                int ix[4];
                float wx[4];
                cubic_src(wo, sx, Wi, ix, wx);
                float cx = 0.f;
#pragma unroll
                for (int d = 0; d < 4; ++d) cx += ix[d] == w ? wx[d] : 0.f;
                if (cx != 0.f) rs += cx * __ldg(dout + ((size_t)b * Ho + ho) * Wo + wo);
            }
            s += cy * rs;
        }
        float* d = din + b * sb + c * sc + h * sh + w * sw;
        float v = gray[c] * s;
        *d = accumulate ? *d + v : v;
    }
}

// in: NHWC [B,1,Wi,C]; out: NHWC [B,Ho,Wo,C], out[b,h,w,:] = lerp(in[w0], in[w1]) with src = w*(Wi-1)/(Wo-1)
__global__ void __launch_bounds__(256)
bilinear_rows_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Wi, int C, int Ho, int Wo, int ldo) {
    const long total = (long)B * Ho * Wo * C;
    const float scale = Wo > 1 ? (float)(Wi - 1) / (Wo - 1) : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int w = (int)(p % Wo);
        long q = p / Wo;
        int b = (int)(q / Ho);
        float src = scale * w;
        int w0 = (int)src;
        if (w0 > Wi - 1) w0 = Wi - 1;
        int w1 = w0 + 1 < Wi ? w0 + 1 : Wi - 1;
        float t = src - w0;
        const float* r = in + (size_t)b * Wi * C;
        out[(size_t)p * ldo + c] = (1.f - t) * r[(size_t)w0 * C + c] + t * r[(size_t)w1 * C + c];
    }
}

// din[b, wi, c] = sum_h sum_{w: w0(w)==wi} (1-t) dout + sum_{w: w1(w)==wi} t dout
__global__ void __launch_bounds__(256)
bilinear_rows_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int B, int Wi, int C, int Ho, int Wo, int ldo) {
    const long total = (long)B * Wi * C;
    const float scale = Wo > 1 ? (float)(Wi - 1) / (Wo - 1) : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int wi = (int)(p % Wi);
        int b = (int)(p / Wi);
        // outputs w with src in (wi-1, wi+1)
        int lo = scale > 0.f ? (int)floorf((wi - 1) / scale) : 0, hi = scale > 0.f ? (int)ceilf((wi + 1) / scale) : Wo - 1;
        lo = lo < 0 ? 0 : lo;
        hi = hi > Wo - 1 ? Wo - 1 : hi;
        float s = 0.f;
        for (int w = lo; w <= hi; ++w) {
            float src = scale * w;
            int w0 = (int)src;
            if (w0 > Wi - 1) w0 = Wi - 1;
            int w1 = w0 + 1 < Wi ? w0 + 1 : Wi - 1;
            float t = src - w0;
            float coef = (w0 == wi ? 1.f - t : 0.f) + (w1 == wi ? t : 0.f);
            if (coef == 0.f) continue;
            float g = 0.f;
            for (int h = 0; h < Ho; ++h) g += dout[((size_t)(b * Ho + h) * Wo + w) * ldo + c];
            s += coef * g;
        }
        din[i] = s;
    }
}

// y[b, 0, stride*w, :] = x[b, 0, w, :], zeros elsewhere (Wout = (W-1)*stride + 1)
__global__ void zero_upsample_w_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int W, int C, int stride) {
    const int Wo = (W - 1) * stride + 1;
    const long total = (long)B * Wo * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int wo = (int)(p % Wo);
        int b = (int)(p / Wo);
        y[i] = (wo % stride == 0) ? x[((size_t)b * W + wo / stride) * C + c] : 0.f;
    }
}
__global__ void zero_upsample_w_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int W, int C, int stride) {
    const int Wo = (W - 1) * stride + 1;
    const long total = (long)B * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int w = (int)(p % W);
        int b = (int)(p / W);
        dx[i] = dy[((size_t)b * Wo + w * stride) * C + c];
    }
}

// ConvTranspose2d weight [Ci][Co][KH][KW] -> equivalent stride-1 conv weight OIHW [Co][Ci][1][KW] = Wt[ci][co][kh_mid][KW-1-kw]
__global__ void tconv_weight_fwd_kernel(const float* __restrict__ wt, float* __restrict__ wc, int Ci, int Co, int KH, int KW, int kh_mid) {
    const int total = Co * Ci * KW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int kw = i % KW;
        int ci = (i / KW) % Ci;
        int co = i / (KW * Ci);
        wc[i] = wt[(((size_t)ci * Co + co) * KH + kh_mid) * KW + (KW - 1 - kw)];
    }
}
// gradient scatter back (rows kh != kh_mid never contribute on H=1 maps: their gradient is exactly zero)
__global__ void tconv_weight_bwd_kernel(const float* __restrict__ dwc, float* __restrict__ dwt, int Ci, int Co, int KH, int KW,
                                        int kh_mid, int accumulate) {
    const int total = Ci * Co * KH * KW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int kw = i % KW;
        int kh = (i / KW) % KH;
        int co = (i / (KW * KH)) % Co;
        int ci = i / (KW * KH * Co);
        float v = kh == kh_mid ? dwc[((size_t)co * Ci + ci) * KW + (KW - 1 - kw)] : 0.f;
        dwt[i] = accumulate ? dwt[i] + v : v;
    }
}

// logits [T,B,C] -> probs [T,B,C] (softmax over C) and prior NHWC [B,1,T,C] = probs[t,b,:] * (b >= drop_first ? 1 : 0)
__global__ void softmax_prior_fwd_kernel(const float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ prior,
                                         int T, int B, int C, int drop_first) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per (t,b)
    const int lane = threadIdx.x & 31;
    if (row >= T * B) return;
    const int t = row / B, b = row % B;
    const float* x = logits + (size_t)row * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(x[c] - mx);
    s = warp_sum(s);
    const float keep = b >= drop_first ? 1.f : 0.f;
    for (int c = lane; c < C; c += 32) {
        float p = expf(x[c] - mx) / s;
        probs[(size_t)row * C + c] = p;
        if (prior) prior[((size_t)b * T + t) * C + c] = p * keep;
    }
}

// dlogits = p * (g - sum_c p g), g = dprobs[t,b,:] + keep * dprior[b,t,:]
__global__ void softmax_prior_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ dprobs,
                                         const float* __restrict__ dprior, float* __restrict__ dlogits, int T, int B, int C,
                                         int drop_first) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= T * B) return;
    const int t = row / B, b = row % B;
    const float keep = b >= drop_first ? 1.f : 0.f;
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) {
        float g = (dprobs ? dprobs[(size_t)row * C + c] : 0.f) + (dprior ? keep * dprior[((size_t)b * T + t) * C + c] : 0.f);
        dot += probs[(size_t)row * C + c] * g;
    }
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) {
        float g = (dprobs ? dprobs[(size_t)row * C + c] : 0.f) + (dprior ? keep * dprior[((size_t)b * T + t) * C + c] : 0.f);
        dlogits[(size_t)row * C + c] = probs[(size_t)row * C + c] * (g - dot);
    }
}

inline int rs_grid(long n) {
    long b = (n + 255) / 256;
    return (int)(b > 148L * 16 ? 148L * 16 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" {

int tpgsr_bicubic_gray_fwd(const float* in, long sb, long sc, long sh, long sw, int B, int Hi, int Wi, float* out, int Ho, int Wo,
                           cudaStream_t stream) {
    if (!in || !out) return TPGSR_ERR_BAD_ARG;
    bicubic_gray_fwd_kernel<<<rs_grid((long)B * Ho * Wo), 256, 0, stream>>>(in, sb, sc, sh, sw, B, Hi, Wi, out, Ho, Wo);
    return tpgsr_launch_status();
}

int tpgsr_bicubic_gray_bwd(const float* dout, int B, int Hi, int Wi, int Ho, int Wo, float* din, long sb, long sc, long sh, long sw,
                           int accumulate, cudaStream_t stream) {
    if (!dout || !din) return TPGSR_ERR_BAD_ARG;
    bicubic_gray_bwd_kernel<<<rs_grid((long)B * 3 * Hi * Wi), 256, 0, stream>>>(dout, B, Hi, Wi, Ho, Wo, din, sb, sc, sh, sw, accumulate);
    return tpgsr_launch_status();
}

int tpgsr_bilinear_rows_fwd(const float* in, float* out, int B, int Wi, int C, int Ho, int Wo, int ldo, cudaStream_t stream) {
    if (!in || !out) return TPGSR_ERR_BAD_ARG;
    bilinear_rows_fwd_kernel<<<rs_grid((long)B * Ho * Wo * C), 256, 0, stream>>>(in, out, B, Wi, C, Ho, Wo, ldo > 0 ? ldo : C);
    return tpgsr_launch_status();
}

int tpgsr_bilinear_rows_bwd(const float* dout, float* din, int B, int Wi, int C, int Ho, int Wo, int ldo, cudaStream_t stream) {
    if (!dout || !din) return TPGSR_ERR_BAD_ARG;
    bilinear_rows_bwd_kernel<<<rs_grid((long)B * Wi * C), 256, 0, stream>>>(dout, din, B, Wi, C, Ho, Wo, ldo > 0 ? ldo : C);
    return tpgsr_launch_status();
}

int tpgsr_zero_upsample_w(const float* x, float* y, int B, int W, int C, int stride, cudaStream_t stream) {
    if (!x || !y || stride < 1) return TPGSR_ERR_BAD_ARG;
    zero_upsample_w_kernel<<<rs_grid((long)B * ((W - 1) * stride + 1) * C), 256, 0, stream>>>(x, y, B, W, C, stride);
    return tpgsr_launch_status();
}

int tpgsr_zero_upsample_w_bwd(const float* dy, float* dx, int B, int W, int C, int stride, cudaStream_t stream) {
    if (!dy || !dx || stride < 1) return TPGSR_ERR_BAD_ARG;
    zero_upsample_w_bwd_kernel<<<rs_grid((long)B * W * C), 256, 0, stream>>>(dy, dx, B, W, C, stride);
    return tpgsr_launch_status();
}

int tpgsr_tconv_weight_fwd(const float* wt, float* wc, int Ci, int Co, int KH, int KW, int kh_mid, cudaStream_t stream) {
    if (!wt || !wc) return TPGSR_ERR_BAD_ARG;
    tconv_weight_fwd_kernel<<<rs_grid((long)Co * Ci * KW), 256, 0, stream>>>(wt, wc, Ci, Co, KH, KW, kh_mid);
    return tpgsr_launch_status();
}

int tpgsr_tconv_weight_bwd(const float* dwc, float* dwt, int Ci, int Co, int KH, int KW, int kh_mid, int accumulate,
                           cudaStream_t stream) {
    if (!dwc || !dwt) return TPGSR_ERR_BAD_ARG;
    tconv_weight_bwd_kernel<<<rs_grid((long)Ci * Co * KH * KW), 256, 0, stream>>>(dwc, dwt, Ci, Co, KH, KW, kh_mid, accumulate);
    return tpgsr_launch_status();
}

int tpgsr_softmax_prior_fwd(const float* logits, float* probs, float* prior, int T, int B, int C, int drop_first,
                            cudaStream_t stream) {
    if (!logits || !probs) return TPGSR_ERR_BAD_ARG;
    softmax_prior_fwd_kernel<<<(T * B + 7) / 8, 256, 0, stream>>>(logits, probs, prior, T, B, C, drop_first);
    return tpgsr_launch_status();
}

int tpgsr_softmax_prior_bwd(const float* probs, const float* dprobs, const float* dprior, float* dlogits, int T, int B, int C,
                            int drop_first, cudaStream_t stream) {
    if (!probs || !dlogits || (!dprobs && !dprior)) return TPGSR_ERR_BAD_ARG;
    softmax_prior_bwd_kernel<<<(T * B + 7) / 8, 256, 0, stream>>>(probs, dprobs, dprior, dlogits, T, B, C, drop_first);
    return tpgsr_launch_status();
}

}  // extern "C"
