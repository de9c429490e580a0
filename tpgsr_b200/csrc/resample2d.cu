// General 2-D resampling helpers for the SRResNet_TL flavour of the text-prior branch (BASELINE config 5; model/srresnet.py:139-195):
//   * nn.ConvTranspose2d(k3, stride (2,2)/(2,1), padding 0/1) on maps with H > 1  =  zero-upsample (stride) + stride-1 convolution with the
//     flipped, in/out-transposed kernel and padding k-1-p: tpgsr_zero_upsample2d(_bwd), tpgsr_tconv_weight2d_fwd/bwd; the convolution
//     itself runs on the conv kernels of this library;
//   * F.interpolate(mode='bilinear', align_corners=True) [B,25,213,32] -> [B,16,64,32] (model/srresnet.py:150): tpgsr_bilinear2d_fwd/bwd;
//   * nn.PixelShuffle(2) without a fused activation (UpsampleBLock with PReLU, model/srresnet.py:223-234; PReLU with one shared slope
//     commutes with the shuffle and is applied before it): tpgsr_pixelshuffle_fwd/bwd.
// All tensors NHWC fp32.
#include "common.cuh"

namespace {

static inline int ew_grid_h(long n) {
    long b = (n + 255) / 256;
    return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

// y[b][h*sh][w*sw][c] = x[b][h][w][c], zeros elsewhere; Ho = (H-1)*sh + 1, Wo = (W-1)*sw + 1
__global__ void zero_upsample2d_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C, int sh, int sw) {
    const int Ho = (H - 1) * sh + 1, Wo = (W - 1) * sw + 1;
    const long total = (long)B * Ho * Wo * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long p = i / C;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        y[i] = (ho % sh == 0 && wo % sw == 0) ? x[(((size_t)b * H + ho / sh) * W + wo / sw) * C + c] : 0.f;
    }
}
__global__ void zero_upsample2d_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C, int sh, int sw) {
    const int Ho = (H - 1) * sh + 1, Wo = (W - 1) * sw + 1;
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long p = i / C;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        dx[i] = dy[(((size_t)b * Ho + h * sh) * Wo + w * sw) * C + c];
    }
}

// ConvTranspose2d weight [Ci][Co][KH][KW] <-> equivalent stride-1 conv weight OIHW [Co][Ci][KH][KW]: wc[co][ci][kh][kw] = wt[ci][co][KH-1-kh][KW-1-kw]
__global__ void tconv_weight2d_fwd_kernel(const float* __restrict__ wt, float* __restrict__ wc, int Ci, int Co, int KH, int KW) {
    const int total = Co * Ci * KH * KW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kw = i % KW, kh = (i / KW) % KH, ci = (i / (KW * KH)) % Ci, co = i / (KW * KH * Ci);
        wc[i] = wt[(((size_t)ci * Co + co) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)];
    }
}
__global__ void tconv_weight2d_bwd_kernel(const float* __restrict__ dwc, float* __restrict__ dwt, int Ci, int Co, int KH, int KW, int accumulate) {
    const int total = Ci * Co * KH * KW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kw = i % KW, kh = (i / KW) % KH, co = (i / (KW * KH)) % Co, ci = i / (KW * KH * Co);
        const float v = dwc[(((size_t)co * Ci + ci) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)];
        dwt[i] = accumulate ? dwt[i] + v : v;
    }
}

__device__ __forceinline__ void ac_coords(int o, float scale, int n_in, int& i0, int& i1, float& t) {
    const float src = scale * o;
    i0 = (int)src;
    if (i0 > n_in - 1) i0 = n_in - 1;
    i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
    t = src - i0;
}

// out[b][ho][wo][c] = bilinear(in) with align_corners=True: src = o * (n_in - 1) / (n_out - 1)
__global__ void __launch_bounds__(256)
bilinear2d_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Hi, int Wi, int C, int Ho, int Wo, int ldo) {
    const long total = (long)B * Ho * Wo * C;
    const float sh = Ho > 1 ? (float)(Hi - 1) / (Ho - 1) : 0.f, sw = Wo > 1 ? (float)(Wi - 1) / (Wo - 1) : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long p = i / C;
        const int wo = (int)(p % Wo);
        long q = p / Wo;
        const int ho = (int)(q % Ho);
        const int b = (int)(q / Ho);
        int h0, h1, w0, w1;
        float th, tw;
        ac_coords(ho, sh, Hi, h0, h1, th);
        ac_coords(wo, sw, Wi, w0, w1, tw);
        const float* base = in + (size_t)b * Hi * Wi * C + c;
        const float v00 = base[((size_t)h0 * Wi + w0) * C], v01 = base[((size_t)h0 * Wi + w1) * C];
        const float v10 = base[((size_t)h1 * Wi + w0) * C], v11 = base[((size_t)h1 * Wi + w1) * C];
        // same association as ATen's upsample_bilinear2d: rows first, then columns
        out[(size_t)p * ldo + c] = (1.f - th) * ((1.f - tw) * v00 + tw * v01) + th * ((1.f - tw) * v10 + tw * v11);
    }
}

// din[b][hi][wi][c] = sum over outputs (ho, wo) whose 2x2 footprint contains (hi, wi) of weight * dout   (gather form: deterministic)
__global__ void __launch_bounds__(256)
bilinear2d_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int B, int Hi, int Wi, int C, int Ho, int Wo, int ldo) {
    const long total = (long)B * Hi * Wi * C;
    const float sh = Ho > 1 ? (float)(Hi - 1) / (Ho - 1) : 0.f, sw = Wo > 1 ? (float)(Wi - 1) / (Wo - 1) : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long p = i / C;
        const int wi = (int)(p % Wi);
        long q = p / Wi;
        const int hi = (int)(q % Hi);
        const int b = (int)(q / Hi);
        int hlo = sh > 0.f ? (int)floorf((hi - 1) / sh) : 0, hhi = sh > 0.f ? (int)ceilf((hi + 1) / sh) : Ho - 1;
        int wlo = sw > 0.f ? (int)floorf((wi - 1) / sw) : 0, whi = sw > 0.f ? (int)ceilf((wi + 1) / sw) : Wo - 1;
        hlo = hlo < 0 ? 0 : hlo; hhi = hhi > Ho - 1 ? Ho - 1 : hhi;
        wlo = wlo < 0 ? 0 : wlo; whi = whi > Wo - 1 ? Wo - 1 : whi;
        float s = 0.f;
        for (int ho = hlo; ho <= hhi; ++ho) {
            int h0, h1; float th;
            ac_coords(ho, sh, Hi, h0, h1, th);
            const float ch = (h0 == hi ? 1.f - th : 0.f) + (h1 == hi ? th : 0.f);
            if (ch == 0.f) continue;
            for (int wo = wlo; wo <= whi; ++wo) {
                int w0, w1; float tw;
                ac_coords(wo, sw, Wi, w0, w1, tw);
                const float cw = (w0 == wi ? 1.f - tw : 0.f) + (w1 == wi ? tw : 0.f);
                if (cw == 0.f) continue;
                s += ch * cw * dout[((size_t)(b * Ho + ho) * Wo + wo) * ldo + c];
            }
        }
        din[i] = s;
    }
}

// PixelShuffle(2) on NHWC: y[b][2h+i][2w+j][c] = x[b][h][w][4c + 2i + j]
__global__ void __launch_bounds__(256) pixelshuffle_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long p = i / C;
        const int w = (int)(p % W);
        const long q = p / W;
        const int h = (int)(q % H);
        const int b = (int)(q / H);
        const float4 v = *reinterpret_cast<const float4*>(x + (size_t)p * 4 * C + 4 * c);
        const size_t o = (((size_t)b * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
        y[o] = v.x; y[o + C] = v.y; y[o + (size_t)2 * W * C] = v.z; y[o + (size_t)2 * W * C + C] = v.w;
    }
}
__global__ void __launch_bounds__(256) pixelshuffle_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long p = i / C;
        const int w = (int)(p % W);
        const long q = p / W;
        const int h = (int)(q % H);
        const int b = (int)(q / H);
        const size_t o = (((size_t)b * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
        *reinterpret_cast<float4*>(dx + (size_t)p * 4 * C + 4 * c) = make_float4(dy[o], dy[o + C], dy[o + (size_t)2 * W * C], dy[o + (size_t)2 * W * C + C]);
    }
}

}  // namespace

extern "C" {

int tpgsr_zero_upsample2d(const float* x, float* y, int B, int H, int W, int C, int sh, int sw, cudaStream_t stream) {
    if (!x || !y || sh < 1 || sw < 1) return TPGSR_ERR_BAD_ARG;
    zero_upsample2d_kernel<<<ew_grid_h((long)B * ((H - 1) * sh + 1) * ((W - 1) * sw + 1) * C), 256, 0, stream>>>(x, y, B, H, W, C, sh, sw);
    return tpgsr_launch_status();
}
int tpgsr_zero_upsample2d_bwd(const float* dy, float* dx, int B, int H, int W, int C, int sh, int sw, cudaStream_t stream) {
    if (!dy || !dx || sh < 1 || sw < 1) return TPGSR_ERR_BAD_ARG;
    zero_upsample2d_bwd_kernel<<<ew_grid_h((long)B * H * W * C), 256, 0, stream>>>(dy, dx, B, H, W, C, sh, sw);
    return tpgsr_launch_status();
}
int tpgsr_tconv_weight2d_fwd(const float* wt, float* wc, int Ci, int Co, int KH, int KW, cudaStream_t stream) {
    if (!wt || !wc) return TPGSR_ERR_BAD_ARG;
    tconv_weight2d_fwd_kernel<<<ew_grid_h((long)Ci * Co * KH * KW), 256, 0, stream>>>(wt, wc, Ci, Co, KH, KW);
    return tpgsr_launch_status();
}
int tpgsr_tconv_weight2d_bwd(const float* dwc, float* dwt, int Ci, int Co, int KH, int KW, int accumulate, cudaStream_t stream) {
    if (!dwc || !dwt) return TPGSR_ERR_BAD_ARG;
    tconv_weight2d_bwd_kernel<<<ew_grid_h((long)Ci * Co * KH * KW), 256, 0, stream>>>(dwc, dwt, Ci, Co, KH, KW, accumulate);
    return tpgsr_launch_status();
}
int tpgsr_bilinear2d_fwd(const float* in, float* out, int B, int Hi, int Wi, int C, int Ho, int Wo, int ldo, cudaStream_t stream) {
    if (!in || !out) return TPGSR_ERR_BAD_ARG;
    bilinear2d_fwd_kernel<<<ew_grid_h((long)B * Ho * Wo * C), 256, 0, stream>>>(in, out, B, Hi, Wi, C, Ho, Wo, ldo > 0 ? ldo : C);
    return tpgsr_launch_status();
}
int tpgsr_bilinear2d_bwd(const float* dout, float* din, int B, int Hi, int Wi, int C, int Ho, int Wo, int ldo, cudaStream_t stream) {
    if (!dout || !din) return TPGSR_ERR_BAD_ARG;
    bilinear2d_bwd_kernel<<<ew_grid_h((long)B * Hi * Wi * C), 256, 0, stream>>>(dout, din, B, Hi, Wi, C, Ho, Wo, ldo > 0 ? ldo : C);
    return tpgsr_launch_status();
}
int tpgsr_pixelshuffle_fwd(const float* x, float* y, int B, int H, int W, int C, cudaStream_t stream) {
    if (!x || !y) return TPGSR_ERR_BAD_ARG;
    pixelshuffle_fwd_kernel<<<ew_grid_h((long)B * H * W * C), 256, 0, stream>>>(x, y, B, H, W, C);
    return tpgsr_launch_status();
}
int tpgsr_pixelshuffle_bwd(const float* dy, float* dx, int B, int H, int W, int C, cudaStream_t stream) {
    if (!dy || !dx) return TPGSR_ERR_BAD_ARG;
    pixelshuffle_bwd_kernel<<<ew_grid_h((long)B * H * W * C), 256, 0, stream>>>(dy, dx, B, H, W, C);
    return tpgsr_launch_status();
}

}  // extern "C"
