// fp32 CUDA-core implicit-GEMM convolution family on NHWC tensors (forward / data-gradient / weight-gradient).
//
// This is the exact-fp32 path: it serves every contraction whose shape is hostile to tensor cores
// (Cin=4 / Cin=1 / Cout=4 9x9 convs, the STN head's tiny maps, linears, GRU/LSTM input projections) and is the
// numerical baseline the tcgen05 path (conv_tc.cu) is validated against.  Reference ops replaced:
// nn.Conv2d / nn.Linear forward+backward as used by model/tsrn.py:26-43,376-381,467,495, model/stn_head.py:12-53,
// model/crnn/crnn.py:35-67 (cuDNN/cuBLAS in the reference).
//
// GEMM view (stride-1 convolution):  Y[m, n] = act(bias[n] + sum_k A[m,k] * Wk[k,n])
//   m = (b, ho, wo),  k = (kh, kw, ci),  A[m,k] = X[b, ho+kh-ph, wo+kw-pw, ci]  (0 outside),  Wk = packed weights.
// The data-gradient is the same gather with the flipped/transposed packing (see pack kernel) and ph' = KH-1-ph.
#include "common.cuh"
#include "conv_geom.h"

namespace {

// ------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> Wk_fwd [K=(kh,kw,ci)][Co]  and  Wk_dgrad [K'=(kh',kw',co)][Ci] (flipped taps)
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, float* __restrict__ wk_fwd,
                                         float* __restrict__ wk_dgrad, int Co, int Ci, int KH, int KW) {
    int total = Co * Ci * KH * KW;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int kw = idx % KW;
        int kh = (idx / KW) % KH;
        int ci = (idx / (KW * KH)) % Ci;
        int co = idx / (KW * KH * Ci);
        float v = w[idx];
        if (wk_fwd) wk_fwd[(size_t)((kh * KW + kw) * Ci + ci) * Co + co] = v;
        if (wk_dgrad) wk_dgrad[(size_t)(((KH - 1 - kh) * KW + (KW - 1 - kw)) * Co + co) * Ci + ci] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// forward / dgrad implicit GEMM.  128 x BN output tile, BK = 16, 256 threads, 8 x TN per thread,
// register-prefetch double buffering.
// ------------------------------------------------------------------------------------------------
template <int BN, int TN, bool VEC>
__global__ void __launch_bounds__(256)
conv_igemm_kernel(const float* __restrict__ X, const float* __restrict__ Wk, const float* __restrict__ bias,
                  float* __restrict__ Y, ConvGeom g, int act, const float* __restrict__ act_param, int out_mode,
                  int accumulate, int kt_per_split) {
    constexpr int BM = 128, BK = 16, TM = 8;
    constexpr int NTN = BN / TN;
    static_assert((BM / TM) * NTN == 256, "thread layout");
    constexpr int B_PER_T = BK * BN / 256;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int t = threadIdx.x;
    const int tn = t % NTN, tm = t / NTN;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int K = g.K, N = g.Co;
    const int nkt_all = (K + BK - 1) / BK;
    // split-K (tiny-M, large-K layers such as the STN head): blockIdx.z owns a k-tile range and writes a raw partial tile
    int kt0 = 0, nkt = nkt_all;
    if (kt_per_split > 0) {
        kt0 = blockIdx.z * kt_per_split;
        nkt = kt0 + kt_per_split < nkt_all ? kt0 + kt_per_split : nkt_all;
        Y += (size_t)blockIdx.z * (size_t)g.M * N;
    }

    // ---- per-thread A-row bookkeeping
    constexpr int A_ROWS = VEC ? 2 : 8;
    int a_b[A_ROWS], a_hw[A_ROWS];          // hw packs (ho<<16)|wo ; b = -1 marks m out of range
#pragma unroll
    for (int i = 0; i < A_ROWS; ++i) {
        int idx = t + i * 256;
        int row = VEC ? (idx >> 2) : (idx >> 4);
        long m = m0 + row;
        if (m < g.M) {
            int wo = (int)(m % g.Wo);
            long r = m / g.Wo;
            int ho = (int)(r % g.Ho);
            a_b[i] = (int)(r / g.Ho);
            a_hw[i] = (ho << 16) | wo;
        } else {
            a_b[i] = -1;
            a_hw[i] = 0;
        }
    }
    const int a_kq = VEC ? (t & 3) * 4 : (t & 15);

    float4 a_reg4[VEC ? 2 : 1];
    float a_reg1[VEC ? 1 : 8];
    float b_reg[B_PER_T];

    auto load_tile = [&](int kt) {
        int k = kt * BK + a_kq;
        int tap = k / g.Ci;
        int ci = k - tap * g.Ci;
        int kh = tap / g.KW;
        int kw = tap - kh * g.KW;
        bool kvalid = k < K;
#pragma unroll
        for (int i = 0; i < A_ROWS; ++i) {
            int ho = a_hw[i] >> 16, wo = a_hw[i] & 0xffff;
            int hi = ho + kh - g.ph, wi = wo + kw - g.pw;
            bool ok = kvalid && a_b[i] >= 0 && (unsigned)hi < (unsigned)g.Hi && (unsigned)wi < (unsigned)g.Wi;
            const float* p = X + ((size_t)(a_b[i] * g.Hi + hi) * g.Wi + wi) * g.ldx + ci;
            if (VEC) {
                a_reg4[i] = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                a_reg1[i] = ok ? __ldg(p) : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            int idx = t + i * 256;
            int kk = idx / BN, nn = idx % BN;
            int kg = kt * BK + kk, ng = n0 + nn;
            b_reg[i] = (kg < K && ng < N) ? __ldg(Wk + (size_t)kg * N + ng) : 0.f;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_ROWS; ++i) {
            int idx = t + i * 256;
            if (VEC) {
                int row = idx >> 2;
                As[buf][a_kq + 0][row] = a_reg4[i].x;
                As[buf][a_kq + 1][row] = a_reg4[i].y;
                As[buf][a_kq + 2][row] = a_reg4[i].z;
                As[buf][a_kq + 3][row] = a_reg4[i].w;
            } else {
                int row = idx >> 4;
                As[buf][a_kq][row] = a_reg1[i];
            }
        }
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            int idx = t + i * 256;
            Bs[buf][idx / BN][idx % BN] = b_reg[i];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tile(kt0);
    store_tile(0);
    __syncthreads();
    for (int kt = kt0; kt < nkt; ++kt) {
        int buf = (kt - kt0) & 1;
        if (kt + 1 < nkt) load_tile(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * TM]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * TM + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if (TN == 4) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * TN]);
                b[0] = b0.x; b[1 % TN] = b0.y; b[2 % TN] = b0.z; b[3 % TN] = b0.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[buf][k][tn * TN + j];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nkt) store_tile(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue
    const float slope = (act == TPGSR_ACT_PRELU && act_param) ? __ldg(act_param) : 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        long m = m0 + tm * TM + i;
        if (m >= g.M) continue;
        int wo = (int)(m % g.Wo);
        long r = m / g.Wo;
        int ho = (int)(r % g.Ho);
        int b = (int)(r / g.Ho);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tn * TN + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            v = apply_act(v, act, slope);
            size_t off = out_mode == TPGSR_OUT_NCHW
                             ? ((size_t)(b * N + n) * g.Ho + ho) * g.Wo + wo
                             : (size_t)m * g.ldy + n;
            if (accumulate) v += Y[off];
            Y[off] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dWk[k, n] = sum_m A[m,k] * dY[m,n], split over M (deterministic 2-stage reduction).
// 64(k) x BN(n) tile, 16-row m chunks, 128 threads, 8 x TN per thread.
// ------------------------------------------------------------------------------------------------
template <int BN, int TN, bool VEC>
__global__ void __launch_bounds__(128)
conv_wgrad_kernel(const float* __restrict__ X, const float* __restrict__ dY, float* __restrict__ ws, ConvGeom g,
                  long m_per_split) {
    constexpr int BKT = 64, BMC = 16, TK = 8;
    constexpr int NTN = BN / TN;
    static_assert((BKT / TK) * NTN == 128, "thread layout");
    constexpr int B_PER_T = BMC * BN / 128;
    __shared__ __align__(16) float As[2][BMC][BKT + 4];
    __shared__ __align__(16) float Bs[2][BMC][BN];

    const int t = threadIdx.x;
    const int tn = t % NTN, tk = t / NTN;
    const int k0 = blockIdx.x * BKT, n0 = blockIdx.y * BN;
    const int K = g.K, N = g.Co;
    const long m_begin = (long)blockIdx.z * m_per_split;
    long m_end = m_begin + m_per_split;
    if (m_end > g.M) m_end = g.M;
    const int nchunks = m_begin < m_end ? (int)((m_end - m_begin + BMC - 1) / BMC) : 0;

    // A bookkeeping: k is fixed per thread across chunks
    constexpr int A_PER_T = VEC ? 2 : 8;
    const int a_k = VEC ? (k0 + (t & 15) * 4) : (k0 + (t & 63));
    const bool a_kvalid = a_k < K;
    int a_ci, a_kh, a_kw;
    {
        int tap = a_k / g.Ci;
        a_ci = a_k - tap * g.Ci;
        a_kh = tap / g.KW;
        a_kw = tap - a_kh * g.KW;
    }
    float4 a_reg4[VEC ? 2 : 1];
    float a_reg1[VEC ? 1 : 8];
    float b_reg[B_PER_T];

    auto load_chunk = [&](int c) {
        long mbase = m_begin + (long)c * BMC;
#pragma unroll
        for (int i = 0; i < A_PER_T; ++i) {
            int idx = t + i * 128;
            int mm = VEC ? (idx >> 4) : (idx >> 6);
            long m = mbase + mm;
            bool ok = a_kvalid && m < m_end;
            int wo = 0, ho = 0, b = 0;
            if (ok) {
                wo = (int)(m % g.Wo);
                long r = m / g.Wo;
                ho = (int)(r % g.Ho);
                b = (int)(r / g.Ho);
            }
            int hi = ho + a_kh - g.ph, wi = wo + a_kw - g.pw;
            ok = ok && (unsigned)hi < (unsigned)g.Hi && (unsigned)wi < (unsigned)g.Wi;
            const float* p = X + ((size_t)(b * g.Hi + hi) * g.Wi + wi) * g.ldx + a_ci;
            if (VEC) a_reg4[i] = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
            else a_reg1[i] = ok ? __ldg(p) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            int idx = t + i * 128;
            int mm = idx / BN, nn = idx % BN;
            long m = mbase + mm;
            int ng = n0 + nn;
            b_reg[i] = (m < m_end && ng < N) ? __ldg(dY + (size_t)m * g.ldy + ng) : 0.f;
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER_T; ++i) {
            int idx = t + i * 128;
            if (VEC) *reinterpret_cast<float4*>(&As[buf][idx >> 4][(idx & 15) * 4]) = a_reg4[i];
            else As[buf][idx >> 6][idx & 63] = a_reg1[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER_T; ++i) {
            int idx = t + i * 128;
            Bs[buf][idx / BN][idx % BN] = b_reg[i];
        }
    };

    float acc[TK][TN];
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (nchunks > 0) {
        load_chunk(0);
        store_chunk(0);
    }
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        int buf = c & 1;
        if (c + 1 < nchunks) load_chunk(c + 1);
#pragma unroll
        for (int mm = 0; mm < BMC; ++mm) {
            float a[TK], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][mm][tk * TK]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][mm][tk * TK + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if (TN == 4) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][mm][tn * TN]);
                b[0] = b0.x; b[1 % TN] = b0.y; b[2 % TN] = b0.z; b[3 % TN] = b0.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[buf][mm][tn * TN + j];
            }
#pragma unroll
            for (int i = 0; i < TK; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (c + 1 < nchunks) store_chunk(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TK; ++i) {
        int k = k0 + tk * TK + i;
        if (k >= K) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tn * TN + j;
            if (n < N) ws[((size_t)blockIdx.z * K + k) * N + n] = acc[i][j];
        }
    }
}

// split-K finalize: y = act(sum_z ws[z][m][n] + bias) with the usual epilogue options
__global__ void conv_splitk_finalize_kernel(const float* __restrict__ ws, int splits, const float* __restrict__ bias,
                                            float* __restrict__ Y, ConvGeom g, int act, const float* __restrict__ act_param,
                                            int out_mode, int accumulate) {
    const long total = g.M * g.Co;
    const float slope = (act == TPGSR_ACT_PRELU && act_param) ? __ldg(act_param) : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long m = i / g.Co;
        int n = (int)(i - m * g.Co);
        float v = bias ? __ldg(bias + n) : 0.f;
        for (int z = 0; z < splits; ++z) v += ws[(size_t)z * total + i];
        v = apply_act(v, act, slope);
        size_t off;
        if (out_mode == TPGSR_OUT_NCHW) {
            int wo = (int)(m % g.Wo);
            long r = m / g.Wo;
            int ho = (int)(r % g.Ho);
            int b = (int)(r / g.Ho);
            off = ((size_t)(b * g.Co + n) * g.Ho + ho) * g.Wo + wo;
        } else {
            off = (size_t)m * g.ldy + n;
        }
        if (accumulate) v += Y[off];
        Y[off] = v;
    }
}

// sum the split partials and scatter to OIHW:  dW[n][ci][kh][kw] (+)= sum_z ws[z][k][n]
// blockDim = (32, 8): 32 consecutive (k,n) elements x 8 lanes over the splits, fixed-order shared-memory tree.
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dW, int splits, int K, int N,
                                         int Ci, int KH, int KW, int accumulate) {
    const int total = K * N;
    const int idx = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (idx < total) {
        // independent partial sums keep 4 loads in flight per thread (fixed association order: deterministic)
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int z = threadIdx.y;
        for (; z + 24 < splits; z += 32) {
            s0 += __ldg(ws + (size_t)z * total + idx);
            s1 += __ldg(ws + (size_t)(z + 8) * total + idx);
            s2 += __ldg(ws + (size_t)(z + 16) * total + idx);
            s3 += __ldg(ws + (size_t)(z + 24) * total + idx);
        }
        for (; z < splits; z += 8) s0 += __ldg(ws + (size_t)z * total + idx);
        s = (s0 + s1) + (s2 + s3);
    }
    __shared__ float sh[8][33];
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y != 0 || idx >= total) return;
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][threadIdx.x];
    int n = idx % N, k = idx / N;
    int tap = k / Ci, ci = k - tap * Ci;
    int kh = tap / KW, kw = tap - kh * KW;
    size_t o = ((size_t)(n * Ci + ci) * KH + kh) * KW + kw;
    dW[o] = accumulate ? dW[o] + s : s;
}

int wgrad_splits(const ConvGeom& g, int BN) {
    int tiles = ceil_div(g.K, 64) * ceil_div(g.Co, BN);
    long want = (148L * 6 + tiles - 1) / tiles;
    long maxs = (g.M + 63) / 64;
    long s = want < maxs ? want : maxs;
    if (s < 1) s = 1;
    if (s > 1024) s = 1024;
    return (int)s;
}

int fill_geom(ConvGeom& g, int B, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int KH, int KW, int ph, int pw,
              int ldx, int ldy) {
    if (B <= 0 || Hi <= 0 || Wi <= 0 || Ci <= 0 || Ho <= 0 || Wo <= 0 || Co <= 0 || KH <= 0 || KW <= 0) return TPGSR_ERR_BAD_ARG;
    if (Ho >= 65536 || Wo >= 65536) return TPGSR_ERR_UNSUPPORTED;
    g.B = B; g.Hi = Hi; g.Wi = Wi; g.Ci = Ci; g.Ho = Ho; g.Wo = Wo; g.Co = Co; g.KH = KH; g.KW = KW; g.ph = ph; g.pw = pw;
    g.ldx = ldx > 0 ? ldx : Ci;
    g.ldy = ldy > 0 ? ldy : Co;
    g.K = KH * KW * Ci;
    g.M = (long)B * Ho * Wo;
    return TPGSR_OK;
}

}  // namespace

extern "C" {

int tpgsr_pack_conv_weights(const float* w_oihw, float* wk_fwd, float* wk_dgrad, int Co, int Ci, int KH, int KW,
                            cudaStream_t stream) {
    if (!w_oihw || (!wk_fwd && !wk_dgrad)) return TPGSR_ERR_BAD_ARG;
    int total = Co * Ci * KH * KW;
    pack_conv_weights_kernel<<<ceil_div(total, 256) > 592 ? 592 : ceil_div(total, 256), 256, 0, stream>>>(
        w_oihw, wk_fwd, wk_dgrad, Co, Ci, KH, KW);
    return tpgsr_launch_status();
}

// 3x3 'same' convolution of a ONE-channel image (CRNN conv0: model/crnn/crnn.py:35, 1 -> 64 on the 32 x 100 grey crop).  K = 9: nothing for a
// GEMM to do — the op is the 39 MB output write.  A thread owns 4 output channels of 4 horizontally adjacent pixels: 18 input loads (3 x 6
// window, L1-resident), 144 FMAs, four float4 stores (the 16 threads of a pixel write its 256-byte row: whole lines).  The gather-GEMM
// kernel took 73 us for this layer at B = 48 (profiles/r02/launches_c3_summary_call21.txt).
__global__ void __launch_bounds__(256)
conv_c1_3x3_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ bias, float* __restrict__ y,
                   int B, int H, int W, int Co, int ldy, int act, const float* __restrict__ act_param) {
    const unsigned cq = Co >> 2, WG = (W + 3) >> 2;
    const unsigned total = (unsigned)B * H * WG * cq;
    const unsigned stride = gridDim.x * blockDim.x;          // a multiple of cq (host): the channel quad of a thread never changes
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned cqi = i % cq;
    float4 wv[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wv[t] = __ldg(reinterpret_cast<const float4*>(wk + (size_t)t * Co) + cqi);
    const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cqi) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float slope = (act == TPGSR_ACT_PRELU && act_param) ? __ldg(act_param) : 0.f;
    for (; i < total; i += stride) {
        const unsigned g = i / cq;
        const unsigned t = g / WG, wg = g - t * WG;
        const unsigned b = t / (unsigned)H, h = t - b * (unsigned)H;
        const int w0 = (int)wg * 4;
        float xw[3][6];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int hh = (int)h - 1 + a;
            const bool rok = (unsigned)hh < (unsigned)H;
            const float* row = x + ((size_t)b * H + (rok ? hh : 0)) * W;
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                const int ww = w0 - 1 + d;
                xw[a][d] = (rok && (unsigned)ww < (unsigned)W) ? __ldg(row + ww) : 0.f;
            }
        }
        float* yrow = y + (((size_t)b * H + h) * W + w0) * ldy + 4 * cqi;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (w0 + j >= W) break;
            float4 o = bv;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float xv = xw[a][j + d];
                    const float4 ww = wv[a * 3 + d];
                    o.x = fmaf(xv, ww.x, o.x); o.y = fmaf(xv, ww.y, o.y); o.z = fmaf(xv, ww.z, o.z); o.w = fmaf(xv, ww.w, o.w);
                }
            if (act != TPGSR_ACT_NONE) { o.x = apply_act(o.x, act, slope); o.y = apply_act(o.y, act, slope); o.z = apply_act(o.z, act, slope); o.w = apply_act(o.w, act, slope); }
            *reinterpret_cast<float4*>(yrow + (size_t)j * ldy) = o;
        }
    }
}

// Generic gather-GEMM used for both the forward conv (wk = Wk_fwd, geometry as given) and the data gradient
// (x := dY, wk = Wk_dgrad, Ci := conv Co, Co := conv Ci, ph := KH-1-ph, pw := KW-1-pw, (Hi,Wi) := dY dims, (Ho,Wo) := dX dims).
static int igemm_splits(const ConvGeom& g, bool small_n) {
    long tiles = (long)ceil_div(g.M, 128) * ceil_div(g.Co, small_n ? 16 : 64);
    int nkt = (g.K + 15) / 16;
    if (tiles >= 74 || nkt < 16) return 1;
    long s = (148L * 2) / tiles;
    if (s > nkt / 4) s = nkt / 4;
    return s < 2 ? 1 : (int)s;
}

size_t tpgsr_conv2d_igemm_workspace_bytes(int B, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int KH, int KW) {
    ConvGeom g;
    if (fill_geom(g, B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, 0, 0, 0, 0)) return 0;
    int s = igemm_splits(g, Co <= 16);
    return s > 1 ? (size_t)s * g.M * g.Co * sizeof(float) : 0;
}

int tpgsr_conv2d_igemm_ws(const float* x, const float* wk, const float* bias, float* y, int B, int Hi, int Wi, int Ci,
                          int Ho, int Wo, int Co, int KH, int KW, int ph, int pw, int ldx, int ldy, int act,
                          const float* act_param, int out_mode, int accumulate, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream) {
    ConvGeom g;
    int rc = fill_geom(g, B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, ph, pw, ldx, ldy);
    if (rc) return rc;
    if (!x || !wk || !y) return TPGSR_ERR_BAD_ARG;
    if (Ci == 1 && KH == 3 && KW == 3 && ph == 1 && pw == 1 && Ho == Hi && Wo == Wi && (Co & 3) == 0 && Co <= 256 && 256 % (Co >> 2) == 0 &&
        out_mode == TPGSR_OUT_NHWC && !accumulate && g.ldx == 1 && (g.ldy & 3) == 0 && (long)B * Hi * Wi * Co < (1L << 31) &&
        !((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(wk) | reinterpret_cast<uintptr_t>(bias)) & 15)) {
        const long threads = (long)B * Hi * ((Wi + 3) / 4) * (Co / 4);
        const long blocks = (threads + 255) / 256;
        conv_c1_3x3_kernel<<<(int)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, stream>>>(x, wk, bias, y, B, Hi, Wi, Co, g.ldy, act, act_param);
        return tpgsr_launch_status();
    }
    bool vec = (Ci % 4 == 0) && (g.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    bool small_n = Co <= 16;
    int splits = igemm_splits(g, small_n);
    if (splits > 1 && (!workspace || workspace_bytes < (size_t)splits * g.M * g.Co * sizeof(float))) splits = 1;
    const int nkt = (g.K + 15) / 16;
    const int kps = splits > 1 ? (nkt + splits - 1) / splits : 0;
    if (splits > 1) splits = (nkt + kps - 1) / kps;
    dim3 grid(ceil_div(g.M, 128), ceil_div(Co, small_n ? 16 : 64), splits > 1 ? splits : 1);
    float* out = splits > 1 ? static_cast<float*>(workspace) : y;
    ConvGeom gk = g;
    if (splits > 1) gk.ldy = Co;
    const float* kb = splits > 1 ? nullptr : bias;
    const int kact = splits > 1 ? TPGSR_ACT_NONE : act;
    const int kom = splits > 1 ? TPGSR_OUT_NHWC : out_mode;
    const int kacc = splits > 1 ? 0 : accumulate;
    if (small_n) {
        if (vec) conv_igemm_kernel<16, 1, true><<<grid, 256, 0, stream>>>(x, wk, kb, out, gk, kact, act_param, kom, kacc, kps);
        else conv_igemm_kernel<16, 1, false><<<grid, 256, 0, stream>>>(x, wk, kb, out, gk, kact, act_param, kom, kacc, kps);
    } else {
        if (vec) conv_igemm_kernel<64, 4, true><<<grid, 256, 0, stream>>>(x, wk, kb, out, gk, kact, act_param, kom, kacc, kps);
        else conv_igemm_kernel<64, 4, false><<<grid, 256, 0, stream>>>(x, wk, kb, out, gk, kact, act_param, kom, kacc, kps);
    }
    rc = tpgsr_launch_status();
    if (rc || splits <= 1) return rc;
    long total = g.M * g.Co;
    conv_splitk_finalize_kernel<<<ceil_div(total, 256) > 1184 ? 1184 : ceil_div(total, 256), 256, 0, stream>>>(
        static_cast<float*>(workspace), splits, bias, y, g, act, act_param, out_mode, accumulate);
    return tpgsr_launch_status();
}

// Generic gather-GEMM used for both the forward conv (wk = Wk_fwd, geometry as given) and the data gradient
// (x := dY, wk = Wk_dgrad, Ci := conv Co, Co := conv Ci, ph := KH-1-ph, pw := KW-1-pw, (Hi,Wi) := dY dims, (Ho,Wo) := dX dims).
int tpgsr_conv2d_igemm(const float* x, const float* wk, const float* bias, float* y, int B, int Hi, int Wi, int Ci,
                       int Ho, int Wo, int Co, int KH, int KW, int ph, int pw, int ldx, int ldy, int act,
                       const float* act_param, int out_mode, int accumulate, cudaStream_t stream) {
    return tpgsr_conv2d_igemm_ws(x, wk, bias, y, B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, ph, pw, ldx, ldy, act, act_param, out_mode,
                                 accumulate, nullptr, 0, stream);
}

// shared by the CUDA-core and tensor-core weight-gradient paths: dW (OIHW) (+)= sum over splits of ws[z][k][n]
int tpgsr_conv2d_wgrad_reduce(const float* ws, float* dw_oihw, int splits, int K, int N, int Ci, int KH, int KW, int accumulate,
                              cudaStream_t stream) {
    if (!ws || !dw_oihw || splits <= 0) return TPGSR_ERR_BAD_ARG;
    conv_wgrad_reduce_kernel<<<ceil_div((long)K * N, 32), dim3(32, 8), 0, stream>>>(ws, dw_oihw, splits, K, N, Ci, KH, KW, accumulate);
    return tpgsr_launch_status();
}

size_t tpgsr_conv2d_wgrad_workspace_bytes(int B, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int KH, int KW) {
    ConvGeom g;
    if (fill_geom(g, B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, 0, 0, 0, 0)) return 0;
    int s = wgrad_splits(g, Co <= 16 ? 16 : 64);
    return (size_t)s * g.K * g.Co * sizeof(float);
}

// dW (OIHW) (+)= conv weight gradient; x is the forward input (NHWC, pitch ldx), dy the output gradient (pitch ldy).
int tpgsr_conv2d_wgrad(const float* x, const float* dy, float* dw_oihw, int B, int Hi, int Wi, int Ci, int Ho, int Wo,
                       int Co, int KH, int KW, int ph, int pw, int ldx, int ldy, int accumulate, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
    ConvGeom g;
    int rc = fill_geom(g, B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, ph, pw, ldx, ldy);
    if (rc) return rc;
    if (!x || !dy || !dw_oihw || !workspace) return TPGSR_ERR_BAD_ARG;
    bool small_n = Co <= 16;
    int BN = small_n ? 16 : 64;
    int splits = wgrad_splits(g, BN);
    if (workspace_bytes < (size_t)splits * g.K * g.Co * sizeof(float)) return TPGSR_ERR_WORKSPACE;
    long mps = (g.M + splits - 1) / splits;
    mps = (mps + 15) / 16 * 16;
    bool vec = (Ci % 4 == 0) && (g.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    dim3 grid(ceil_div(g.K, 64), ceil_div(Co, BN), splits);
    float* ws = static_cast<float*>(workspace);
    if (small_n) {
        if (vec) conv_wgrad_kernel<16, 1, true><<<grid, 128, 0, stream>>>(x, dy, ws, g, mps);
        else conv_wgrad_kernel<16, 1, false><<<grid, 128, 0, stream>>>(x, dy, ws, g, mps);
    } else {
        if (vec) conv_wgrad_kernel<64, 4, true><<<grid, 128, 0, stream>>>(x, dy, ws, g, mps);
        else conv_wgrad_kernel<64, 4, false><<<grid, 128, 0, stream>>>(x, dy, ws, g, mps);
    }
    rc = tpgsr_launch_status();
    if (rc) return rc;
    int total = g.K * g.Co;
    conv_wgrad_reduce_kernel<<<ceil_div(total, 32), dim3(32, 8), 0, stream>>>(ws, dw_oihw, splits, g.K, g.Co, Ci, KH, KW, accumulate);
    return tpgsr_launch_status();
}

}  // extern "C"
