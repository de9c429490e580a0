// "kw-stacking": the 9x9 convolutions with a 4-channel side (TSRN block1: 4 -> 64, final conv: 64 -> 4; model/tsrn.py:26,43) as
// 9x1 tensor-core convolutions on 64-channel operands.
//
// A KxK conv whose narrow side has C <= 64/K channels is hostile to the tensor cores as it stands: N = 4 (padded to 16) or K_gemm = 4 per
// tap.  Fold the kernel COLUMN index into the channel dimension instead:
//
//   narrow INPUT  (block1 forward / its weight gradient; the final conv's data gradient):
//       XS[q][kw*C + i] = X[q + sign*(kw - pw)][i]        ("stacked" planes, 64 channels of which K*C are used)
//       Y[q][o] = sum_kh sum_j XS[q + (kh-ph)*Wp][j] * Wpk[o][kh*64 + j]                     -> a KHx1 conv, Ci = 64
//   narrow OUTPUT (final conv forward; block1's data gradient):
//       P[q][kw*C + o] = sum_kh sum_i X[q + (kh-ph)*Wp][i] * Wpk[kw*C + o][kh*64 + i]        -> a KHx1 conv, Co = K*C
//       Y[q][o] = bias[o] + sum_kw P[q + (kw-pw)][kw*C + o]                                   -> tpgsr_kwsum (shift-and-add)
//
// Both run on the padded-flat kernels of conv_flat.cu with KH taps instead of KH*KW: 9x fewer MMAs, each with a useful N (36 -> 48 or 64)
// instead of N = 16 — the final conv's three kernels were 15% of the train step (profiles/r01).  This file holds the four small helper
// kernels: weight packing, the stacked split, the shift-and-add epilogue and the weight-gradient scatter.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

// Wpk rows x (KH*64) bf16 hi/lo, K-major (what tpgsr_conv2d_flat expects as packed weights).
//   narrow_out = 1: row n = kw*C + o (n < KW*C), column kh*64 + i (i < 64 = wide input channels)
//   narrow_out = 0: row o (o < 64 wide outputs), column kh*64 + (kw*C + i)   (zero for kw*C + i >= KW*C)
// element = W[a][b][kh'][kw'] with (a, b) = transposed ? (i, o) : (o, i), kh' = flip_kh ? KH-1-kh : kh, kw' likewise.
__global__ void kwstack_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Co_w,
                                    int Ci_w, int KH, int KW, int C, int narrow_out, int transposed, int flip_kh, int flip_kw, int rows) {
    const int Kp = KH * 64;
    const int total = rows * Kp;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int n = idx / Kp, kk = idx - n * Kp;
        const int kh = kk / 64, j = kk - kh * 64;
        int o, i, kw;
        bool valid;
        if (narrow_out) { kw = n / C; o = n - kw * C; i = j; valid = kw < KW; }
        else { o = n; kw = j / C; i = j - kw * C; valid = kw < KW; }
        float v = 0.f;
        if (valid) {
            const int a = transposed ? i : o, b = transposed ? o : i;
            const int khw = flip_kh ? KH - 1 - kh : kh, kww = flip_kw ? KW - 1 - kw : kw;
            if (a < Co_w && b < Ci_w) v = w[(((size_t)a * Ci_w + b) * KH + khw) * KW + kww];
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[idx] = h;
        if (lo) lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// x fp32 dense NHWC [B,H,W,C] (pitch ldx) -> stacked padded-flat bf16 planes [B][H+P][W+P][64]:
//   plane[q(b,r,c)][kw*C + i] = x[b][r][c + sign*(kw - pw)][i]   (0 outside the row, in pad rows, and for kw*C + i >= KW*C)
__global__ void __launch_bounds__(256)
kwstack_split_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int B, int H,
                     int W, int C, int KW, int pw, int P, int sign) {
    const int Hp = H + P, Wp = W + P;
    const long long total = (long long)B * Hp * Wp * 16;        // 16 groups of 4 channels per row
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long q = idx >> 4;
        const int j0 = (int)(idx & 15) * 4;
        const int c = (int)(q % Wp);
        const long long t = q / Wp;
        const int r = (int)(t % Hp);
        const int b = (int)(t / Hp);
        float f[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < H) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                const int kw = j / C, i = j - kw * C;
                const int cs = c + sign * (kw - pw);
                if (kw < KW && cs >= 0 && cs < W) f[u] = __ldg(x + ((size_t)(b * H + r) * W + cs) * ldx + i);
            }
        }
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            h[u] = __float2bfloat16_rn(f[u]);
            l[u] = __float2bfloat16_rn(f[u] - __bfloat162float(h[u]));
        }
        *reinterpret_cast<uint2*>(hi + (size_t)q * 64 + j0) = *reinterpret_cast<uint2*>(h);
        if (lo) *reinterpret_cast<uint2*>(lo + (size_t)q * 64 + j0) = *reinterpret_cast<uint2*>(l);
    }
}

// y[b][r][c][o] = act(bias[o] + sum_kw p[b][r][c + kw - pw][kw*C + o])    p: fp32 dense NHWC [B,H,W,ldp]; y NHWC (pitch ldy) or NCHW
__global__ void __launch_bounds__(256)
kwsum_kernel(const float* __restrict__ p, int ldp, const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W, int C, int KW,
             int pw, int ldy, int act, const float* __restrict__ act_param, int out_mode, int accumulate) {
    const long long total = (long long)B * H * W * C;
    const float slope = (act == TPGSR_ACT_PRELU && act_param) ? __ldg(act_param) : 0.f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int o, c, r, b;
        if (out_mode == TPGSR_OUT_NCHW) {      // idx enumerates the NCHW output: consecutive threads = consecutive columns
            c = (int)(idx % W); long long t = idx / W; r = (int)(t % H); t /= H; o = (int)(t % C); b = (int)(t / C);
        } else {
            o = (int)(idx % C); long long t = idx / C; c = (int)(t % W); t /= W; r = (int)(t % H); b = (int)(t / H);
        }
        float s = bias ? __ldg(bias + o) : 0.f;
        const float* row = p + (size_t)(b * H + r) * W * ldp;
        for (int kw = 0; kw < KW; ++kw) {
            const int cs = c + kw - pw;
            if (cs >= 0 && cs < W) s += __ldg(row + (size_t)cs * ldp + kw * C + o);
        }
        s = apply_act(s, act, slope);
        float* dst = out_mode == TPGSR_OUT_NCHW ? y + (((size_t)b * C + o) * H + r) * W + c : y + ((size_t)(b * H + r) * W + c) * ldy + o;
        *dst = accumulate ? *dst + s : s;
    }
}

// dwp: OIHW weight gradient of the KHx1 stacked conv.  layout 0: dwp[(kw*C + co)][ci (64)][kh]  (narrow output, C = Co)
//                                                       layout 1: dwp[co (64)][(kw*C + ci)][kh]  (narrow input,  C = Ci; dwp has 64 columns)
// dw[co][ci][kh][kw] (+)= dwp[...]
__global__ void kwstack_wgrad_scatter_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Co, int Ci, int KH, int KW, int layout,
                                             int accumulate) {
    const int total = Co * Ci * KH * KW;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int kw = idx % KW;
        const int kh = (idx / KW) % KH;
        const int ci = (idx / (KW * KH)) % Ci;
        const int co = idx / (KW * KH * Ci);
        const float v = layout == 0 ? dwp[((size_t)(kw * Co + co) * 64 + ci) * KH + kh] : dwp[((size_t)co * 64 + kw * Ci + ci) * KH + kh];
        dw[idx] = accumulate ? dw[idx] + v : v;
    }
}

}  // namespace

extern "C" {

int tpgsr_kwstack_pack_weights(const float* w_oihw, void* hi, void* lo, int Co, int Ci, int KH, int KW, int narrow_out, int transposed,
                               int flip_kh, int flip_kw, int rows, cudaStream_t stream) {
    if (!w_oihw || !hi || rows < 1) return TPGSR_ERR_BAD_ARG;
    // C = channel count of the narrow side of the EFFECTIVE conv (after the optional transposition)
    const int C = narrow_out ? (transposed ? Ci : Co) : (transposed ? Co : Ci);
    if (C * KW > 64) return TPGSR_ERR_UNSUPPORTED;
    const int total = rows * KH * 64;
    int blocks = (total + 255) / 256;
    kwstack_pack_kernel<<<blocks > 592 ? 592 : blocks, 256, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo),
                                                                          Co, Ci, KH, KW, C, narrow_out, transposed, flip_kh, flip_kw, rows);
    return tpgsr_launch_status();
}

int tpgsr_kwstack_split(const float* x, int ldx, void* hi, void* lo, int B, int H, int W, int C, int KW, int pw, int P, int sign,
                        cudaStream_t stream) {
    if (!x || !hi || C < 1 || C * KW > 64 || (sign != 1 && sign != -1)) return TPGSR_ERR_BAD_ARG;
    long long blocks = ((long long)B * (H + P) * (W + P) * 16 + 255) / 256;
    kwstack_split_kernel<<<(int)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, stream>>>(
        x, ldx, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), B, H, W, C, KW, pw, P, sign);
    return tpgsr_launch_status();
}

int tpgsr_kwsum(const float* p, int ldp, const float* bias, float* y, int B, int H, int W, int C, int KW, int pw, int ldy, int act,
                const float* act_param, int out_mode, int accumulate, cudaStream_t stream) {
    if (!p || !y || C < 1 || KW < 1) return TPGSR_ERR_BAD_ARG;
    long long blocks = ((long long)B * H * W * C + 255) / 256;
    kwsum_kernel<<<(int)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, stream>>>(p, ldp, bias, y, B, H, W, C, KW, pw, ldy > 0 ? ldy : C, act,
                                                                                    act_param, out_mode, accumulate);
    return tpgsr_launch_status();
}

int tpgsr_kwstack_wgrad_scatter(const float* dwp, float* dw_oihw, int Co, int Ci, int KH, int KW, int layout, int accumulate,
                                cudaStream_t stream) {
    if (!dwp || !dw_oihw || (layout != 0 && layout != 1)) return TPGSR_ERR_BAD_ARG;
    const int total = Co * Ci * KH * KW;
    kwstack_wgrad_scatter_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(dwp, dw_oihw, Co, Ci, KH, KW, layout, accumulate);
    return tpgsr_launch_status();
}

}  // extern "C"
