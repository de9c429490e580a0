// Bidirectional LSTM recurrence (hidden 256 per direction, time-first) for the CRNN text-prior branch.
//
// Reference op replaced: nn.LSTM(nIn, 256, bidirectional=True) inside BidirectionalLSTM (model/crnn/crnn.py:5-26; cuDNN RNN in
// the reference).  The input projection x W_ih^T + b_ih for both directions is one GEMM (N = 2*4*256) done by the conv/GEMM
// kernels; this file is the serial part.  Gate order i,f,g,o; h0 = c0 = 0.
//
// One CTA per (batch chunk of 8, direction), thread j = hidden unit j.  Per step every thread forms its 4 gate pre-activations
// for the 8 batch items: 256 x 4 coalesced loads of W_hh^T (1 MiB per direction, L2-resident) against h broadcast from shared
// memory.  Latency/L2-bound by construction (T = 26 serial steps); reported as steps/us, not against the tensor roofline.
#include "common.cuh"

namespace {

constexpr int LH = 256;      // hidden units per direction
constexpr int BC = 8;        // batch items per CTA

// row of (t, b) in the [T*B, ...] matrices: time-first (nn.LSTM's layout) or batch-first (what the NHWC conv stack produces)
#define LROW(t, b) (bf ? ((size_t)(b) * T + (t)) : ((size_t)(t) * B + (b)))

// whh [2][4H][H] -> whhT [2][H][4H]
__global__ void lstm_transpose_whh_kernel(const float* __restrict__ whh, float* __restrict__ whhT) {
    __shared__ float tile[32][33];
    const int dir = blockIdx.z;
    const float* src = whh + (size_t)dir * 4 * LH * LH;
    float* dst = whhT + (size_t)dir * 4 * LH * LH;
    int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;            // src is [4H rows][H cols]
    for (int i = threadIdx.y; i < 32; i += 8) tile[i][threadIdx.x] = src[(size_t)(r0 + i) * LH + c0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) dst[(size_t)(c0 + i) * 4 * LH + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

// gx: [T,B,2,4H] (input projection incl. bias_ih); bhh: [2][4H]; y: [T,B,2H]; stash: [T,B,2,5,H] = i,f,g,o,c
__global__ void __launch_bounds__(LH)
lstm_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ whhT, const float* __restrict__ bhh,
                float* __restrict__ y, float* __restrict__ stash, int T, int B, int bf) {
    const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * BC;
    const int nb = (B - b0) < BC ? (B - b0) : BC;
    __shared__ __align__(16) float hs[LH][BC];
    const float* WT = whhT + (size_t)dir * 4 * LH * LH;
    float bias[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = __ldg(bhh + dir * 4 * LH + g * LH + j);
    float c[BC], h[BC];
#pragma unroll
    for (int b = 0; b < BC; ++b) { c[b] = 0.f; h[b] = 0.f; hs[j][b] = 0.f; }
    __syncthreads();
    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        float acc[4][BC];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int b = 0; b < BC; ++b)
                acc[g][b] = bias[g] + (b < nb ? __ldg(gx + (LROW(t, b0 + b) * 2 + dir) * 4 * LH + g * LH + j) : 0.f);
#pragma unroll 4
        for (int k = 0; k < LH; ++k) {
            float4 h0 = *reinterpret_cast<const float4*>(&hs[k][0]);
            float4 h1 = *reinterpret_cast<const float4*>(&hs[k][4]);
            float hv[BC] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float w = __ldg(WT + (size_t)k * 4 * LH + g * LH + j);
#pragma unroll
                for (int b = 0; b < BC; ++b) acc[g][b] = fmaf(w, hv[b], acc[g][b]);
            }
        }
#pragma unroll
        for (int b = 0; b < BC; ++b) {
            float ig = sigmoid_f(acc[0][b]), fg = sigmoid_f(acc[1][b]), gg = tanhf(acc[2][b]), og = sigmoid_f(acc[3][b]);
            c[b] = fg * c[b] + ig * gg;
            h[b] = og * tanhf(c[b]);
            if (b < nb) {
                y[LROW(t, b0 + b) * 2 * LH + dir * LH + j] = h[b];
                if (stash) {
                    float* sp = stash + ((LROW(t, b0 + b) * 2 + dir) * 5) * LH + j;
                    sp[0] = ig; sp[LH] = fg; sp[2 * LH] = gg; sp[3 * LH] = og; sp[4 * LH] = c[b];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < BC; ++b) hs[j][b] = h[b];
        __syncthreads();
    }
}

// BPTT: dy [T,B,2H] -> dgx [T,B,2,4H] (gradient w.r.t. the gate pre-activations).  whh: original [2][4H][H] layout
// (dh_prev[j] = sum_rows W_hh[row][j] * dgate[row] reads it coalesced over j).
__global__ void __launch_bounds__(LH)
lstm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ stash, const float* __restrict__ whh,
                float* __restrict__ dgx, int T, int B, int bf) {
    const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * BC;
    const int nb = (B - b0) < BC ? (B - b0) : BC;
    __shared__ __align__(16) float dgs[4 * LH][BC];      // 32 KiB
    const float* W = whh + (size_t)dir * 4 * LH * LH;
    float dh[BC], dc[BC];
#pragma unroll
    for (int b = 0; b < BC; ++b) { dh[b] = 0.f; dc[b] = 0.f; }
    for (int s = T - 1; s >= 0; --s) {                   // walk the processing order backwards
        const int t = dir ? T - 1 - s : s;
        const int tprev = dir ? t + 1 : t - 1;           // the step processed just before t (holds c_{prev})
#pragma unroll
        for (int b = 0; b < BC; ++b) {
            float di_p = 0.f, df_p = 0.f, dg_p = 0.f, do_p = 0.f;
            if (b < nb) {
                const float* sp = stash + ((LROW(t, b0 + b) * 2 + dir) * 5) * LH + j;
                float ig = sp[0], fg = sp[LH], gg = sp[2 * LH], og = sp[3 * LH], cc = sp[4 * LH];
                float cprev = s > 0 ? stash[((LROW(tprev, b0 + b) * 2 + dir) * 5 + 4) * LH + j] : 0.f;
                float dht = dh[b] + __ldg(dy + LROW(t, b0 + b) * 2 * LH + dir * LH + j);
                float tc = tanhf(cc);
                float dct = dc[b] + dht * og * (1.f - tc * tc);
                do_p = dht * tc * og * (1.f - og);
                di_p = dct * gg * ig * (1.f - ig);
                dg_p = dct * ig * (1.f - gg * gg);
                df_p = dct * cprev * fg * (1.f - fg);
                dc[b] = dct * fg;
                float* gp = dgx + (LROW(t, b0 + b) * 2 + dir) * 4 * LH + j;
                gp[0] = di_p; gp[LH] = df_p; gp[2 * LH] = dg_p; gp[3 * LH] = do_p;
            }
            dgs[j][b] = di_p; dgs[LH + j][b] = df_p; dgs[2 * LH + j][b] = dg_p; dgs[3 * LH + j][b] = do_p;
        }
        __syncthreads();
        float acc[BC];
#pragma unroll
        for (int b = 0; b < BC; ++b) acc[b] = 0.f;
#pragma unroll 4
        for (int r = 0; r < 4 * LH; ++r) {
            float w = __ldg(W + (size_t)r * LH + j);
            float4 g0 = *reinterpret_cast<const float4*>(&dgs[r][0]);
            float4 g1 = *reinterpret_cast<const float4*>(&dgs[r][4]);
            acc[0] = fmaf(w, g0.x, acc[0]); acc[1] = fmaf(w, g0.y, acc[1]); acc[2] = fmaf(w, g0.z, acc[2]); acc[3] = fmaf(w, g0.w, acc[3]);
            acc[4] = fmaf(w, g1.x, acc[4]); acc[5] = fmaf(w, g1.y, acc[5]); acc[6] = fmaf(w, g1.z, acc[6]); acc[7] = fmaf(w, g1.w, acc[7]);
        }
#pragma unroll
        for (int b = 0; b < BC; ++b) dh[b] = acc[b];
        __syncthreads();
    }
}

// hprev[t,b,dir,:] = y[t_prev,b,dir*H:] (the hidden state that entered step t), zero at each direction's first step
__global__ void lstm_shift_h_kernel(const float* __restrict__ y, float* __restrict__ hprev, int T, int B, int bf) {
    const long total = (long)T * B * 2 * LH;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int j = (int)(i % LH);
        long q = i / LH;
        int dir = (int)(q % 2);
        long r = q / 2;
        int b = bf ? (int)(r / T) : (int)(r % B);
        int t = bf ? (int)(r % T) : (int)(r / B);
        int tp = dir ? t + 1 : t - 1;
        hprev[i] = (tp >= 0 && tp < T) ? y[LROW(tp, b) * 2 * LH + dir * LH + j] : 0.f;
    }
}

}  // namespace

extern "C" {

int tpgsr_lstm_transpose_whh(const float* whh, float* whhT, cudaStream_t stream) {
    if (!whh || !whhT) return TPGSR_ERR_BAD_ARG;
    lstm_transpose_whh_kernel<<<dim3(LH / 32, 4 * LH / 32, 2), dim3(32, 8), 0, stream>>>(whh, whhT);
    return tpgsr_launch_status();
}

int tpgsr_lstm_fwd(const float* gx, const float* whhT, const float* bhh, float* y, float* stash, int T, int B, int hidden,
                   int batch_first, cudaStream_t stream) {
    if (!gx || !whhT || !bhh || !y || T <= 0 || B <= 0) return TPGSR_ERR_BAD_ARG;
    if (hidden != LH) return TPGSR_ERR_UNSUPPORTED;
    lstm_fwd_kernel<<<dim3((B + BC - 1) / BC, 2), LH, 0, stream>>>(gx, whhT, bhh, y, stash, T, B, batch_first);
    return tpgsr_launch_status();
}

int tpgsr_lstm_bwd(const float* dy, const float* stash, const float* whh, float* dgx, int T, int B, int hidden,
                   int batch_first, cudaStream_t stream) {
    if (!dy || !stash || !whh || !dgx || T <= 0 || B <= 0) return TPGSR_ERR_BAD_ARG;
    if (hidden != LH) return TPGSR_ERR_UNSUPPORTED;
    lstm_bwd_kernel<<<dim3((B + BC - 1) / BC, 2), LH, 0, stream>>>(dy, stash, whh, dgx, T, B, batch_first);
    return tpgsr_launch_status();
}

int tpgsr_lstm_shift_h(const float* y, float* hprev, int T, int B, int hidden, int batch_first, cudaStream_t stream) {
    if (!y || !hprev) return TPGSR_ERR_BAD_ARG;
    if (hidden != LH) return TPGSR_ERR_UNSUPPORTED;
    long total = (long)T * B * 2 * LH;
    long blocks = (total + 255) / 256;
    lstm_shift_h_kernel<<<(int)(blocks > 2368 ? 2368 : blocks), 256, 0, stream>>>(y, hprev, T, B, batch_first);
    return tpgsr_launch_status();
}

}  // extern "C"
