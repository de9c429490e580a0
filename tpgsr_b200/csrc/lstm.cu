// Bidirectional LSTM recurrence (hidden 256 per direction, time-first) for the CRNN text-prior branch.
//
// Reference op replaced: nn.LSTM(nIn, 256, bidirectional=True) inside BidirectionalLSTM (model/crnn/crnn.py:5-26; cuDNN RNN in
// the reference).  The input projection x W_ih^T + b_ih for both directions is one GEMM (N = 2*4*256) done by the conv/GEMM
// kernels; this file is the serial part.  Gate order i,f,g,o; h0 = c0 = 0.
//
// Latency-bound by construction (T = 26 serial steps): reported as us per step, not against the tensor roofline.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int LH = 256;      // hidden units per direction

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

// ---------------------------------------------------------------------------------------------------------------------------------
// Cluster recurrence.  The serial part of an LSTM layer is T dependent steps of h W_hh^T ([B,256] x [256,1024] per direction): tiny
// GEMMs whose weight matrix (1 MiB fp32 per direction) does not fit one SM.  A thread-block CLUSTER of 8 CTAs owns one (batch group,
// direction): CTA r keeps the 128 gate rows of hidden units [32r, 32r+32) resident in shared memory (128 KB fp32) for all T steps, forms
// its units' gates for the group's batch items from the full h vector, and pushes its 32 new h values per batch item into the h buffers
// of all 8 CTAs through distributed shared memory (one cluster barrier per step, h double-buffered).  Nothing but gx / y / stash touches
// HBM or L2 inside the loop.  The round-1 kernel streamed the whole W_hh from L2 in every CTA at every step (766 us per layer at B = 48,
// profiles/r02/call1_stdout.log).
// FMA phase: warp w covers k in [32w, 32w+32), lane u = hidden unit: 4 gates x BG batch accumulators, W as one LDS.128 per k (layout
// [k][u][gate]), h as broadcast LDS.128; the 8 warps' partial sums meet in shared memory.  Gate phase: thread (b, u).
// ---------------------------------------------------------------------------------------------------------------------------------
// split cluster barrier: work that the peers do not wait for (HBM stores, next-step loads) goes between arrive and wait
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int LS = 8;             // CTAs per cluster = hidden slices
constexpr int LU = LH / LS;       // hidden units per CTA (32)

template <int BG>
struct LstmSmem {
    static constexpr int W_FLOATS = LH * LU * 4;              // fwd: [k][u][gate]; bwd: [gate][u][j]
    static constexpr int HS_FLOATS = 2 * LH * BG;             // fwd: h double buffer [2][k][b]; bwd: partial dh [2 halves][j][b]
    static constexpr int RED_FLOATS = 8 * 4 * BG * LU;        // fwd: [warp][gate*BG + b][u]; bwd: receive slots [2][src][u][b] (+ dgate [128][b])
    static constexpr int LOC_FLOATS = LU * BG;
    static constexpr size_t BYTES = (size_t)(W_FLOATS + HS_FLOATS + RED_FLOATS + LOC_FLOATS) * sizeof(float);
};

// gx: [T,B,2,4H] (input projection incl. bias_ih); bhh: [2][4H]; y: [T,B,2H]; stash: [T,B,2,5,H] = i,f,g,o,c
template <int BG>
__global__ void __cluster_dims__(LS, 1, 1) __launch_bounds__(256, 1)
lstm_fwd_cluster_kernel(const float* __restrict__ gx, const float* __restrict__ whhT, const float* __restrict__ bhh,
                        float* __restrict__ y, float* __restrict__ stash, int T, int B, int bf, long long* dbg) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float lsm[];
    float* Ws = lsm;
    float* hs = Ws + LstmSmem<BG>::W_FLOATS;
    float* red = hs + LstmSmem<BG>::HS_FLOATS;
    float* hloc = red + LstmSmem<BG>::RED_FLOATS;
    const int r = (int)cluster.block_rank();
    const int dir = blockIdx.z, b0 = blockIdx.y * BG;
    const int nb = (B - b0) < BG ? (B - b0) : BG;
    const int tid = threadIdx.x, w = tid >> 5, u = tid & 31;
    constexpr int Q = BG / 4;                                 // float4s per h row
    const float* WT = whhT + (size_t)dir * 4 * LH * LH;       // [k][gate*256 + j]
    for (int idx = tid; idx < LH * 4 * LU; idx += 256) {
        const int k = idx >> 7, rem = idx & 127, g = rem >> 5, uu = rem & 31;
        Ws[(k * LU + uu) * 4 + g] = __ldg(WT + (size_t)k * 4 * LH + g * LH + LU * r + uu);
    }
    for (int idx = tid; idx < LH * BG; idx += 256) hs[idx] = 0.f;
    // gate-phase role: thread (pb, u) owns hidden unit 32r + u of batch item b0 + pb
    const int pb = w;
    const bool gate_thread = pb < BG;
    float bias[4], c = 0.f, ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, hval = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = __ldg(bhh + dir * 4 * LH + g * LH + LU * r + u);
    cluster.sync();                                           // every CTA of the cluster is running and has cleared its h buffer
    const bool stamp = dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;      // developer timeline (scripts/time_lstm.py)
    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        const float* hb = hs + (s & 1) * LH * BG;
        if (stamp) dbg[s * 6] = clock64();
        float gxv[4] = {0.f, 0.f, 0.f, 0.f};
        if (gate_thread && pb < nb) {
            const float* gp = gx + (LROW(t, b0 + pb) * 2 + dir) * 4 * LH + LU * r + u;
#pragma unroll
            for (int g = 0; g < 4; ++g) gxv[g] = __ldg(gp + g * LH);
        }
        float acc[4][BG];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int b = 0; b < BG; ++b) acc[g][b] = 0.f;
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
            const int k = 32 * w + kk;
            const float4 wv = *reinterpret_cast<const float4*>(Ws + (k * LU + u) * 4);
            float hv[BG];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float4 h4 = *reinterpret_cast<const float4*>(hb + k * BG + 4 * q);
                hv[4 * q] = h4.x; hv[4 * q + 1] = h4.y; hv[4 * q + 2] = h4.z; hv[4 * q + 3] = h4.w;
            }
#pragma unroll
            for (int b = 0; b < BG; ++b) {
                acc[0][b] = fmaf(wv.x, hv[b], acc[0][b]);
                acc[1][b] = fmaf(wv.y, hv[b], acc[1][b]);
                acc[2][b] = fmaf(wv.z, hv[b], acc[2][b]);
                acc[3][b] = fmaf(wv.w, hv[b], acc[3][b]);
            }
        }
        if (stamp) dbg[s * 6 + 1] = clock64() + (long long)(acc[0][0] * 0.f);
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int b = 0; b < BG; ++b) red[(w * 4 * BG + g * BG + b) * LU + u] = acc[g][b];
        __syncthreads();
        if (stamp) dbg[s * 6 + 2] = clock64();
        if (gate_thread) {
            float pre[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float sacc = bias[g] + gxv[g];
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) sacc += red[(ww * 4 * BG + g * BG + pb) * LU + u];
                pre[g] = sacc;
            }
            ig = sigmoid_f(pre[0]); fg = sigmoid_f(pre[1]); gg = tanhf(pre[2]); og = sigmoid_f(pre[3]);
            c = fg * c + ig * gg;
            hval = og * tanhf(c);
            hloc[u * BG + pb] = pb < nb ? hval : 0.f;
        }
        __syncthreads();
        if (stamp) dbg[s * 6 + 3] = clock64();
        // push this CTA's [32 units][BG] block of h_t into the next-step h buffer of every CTA of the cluster (itself included)
        const int nxt = (s + 1) & 1;
        for (int i = tid; i < LS * LU * Q; i += 256) {
            const int dst = i / (LU * Q), q = i - dst * (LU * Q);
            const float4 v = *reinterpret_cast<const float4*>(hloc + 4 * q);
            float* remote = cluster.map_shared_rank(hs, dst);
            *reinterpret_cast<float4*>(remote + nxt * LH * BG + (LU * r) * BG + 4 * q) = v;
        }
        cluster_arrive_release();                             // h_t is on its way to every CTA ...
        if (stamp) dbg[s * 6 + 4] = clock64();
        if (gate_thread && pb < nb) {                         // ... the HBM outputs are written behind the arrive: the release need not wait for them
            y[LROW(t, b0 + pb) * 2 * LH + dir * LH + LU * r + u] = hval;
            if (stash) {
                float* sp = stash + ((LROW(t, b0 + pb) * 2 + dir) * 5) * LH + LU * r + u;
                sp[0] = ig; sp[LH] = fg; sp[2 * LH] = gg; sp[3 * LH] = og; sp[4 * LH] = c;
            }
        }
        cluster_wait_acquire();                               // h_t complete everywhere (release / acquire at cluster scope)
        if (stamp) dbg[s * 6 + 5] = clock64();
    }
    cluster.sync();                                           // no CTA leaves while a peer may still be writing into its shared memory
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Tensor-core recurrence (round 2).  The FMA phase above is 54 % of a step (3 160 of 5 840 cycles, profiles/r02/time_lstm.log).  Here
// the step's GEMM runs on mma.sync m16n8k16 with the gate rows on the M side and the 8 batch items of the group as N:
//   gates^T[128 gate rows of this CTA][8] = W_slice[128][256] . h^T[256][8]
// warp w owns gate rows 16w .. 16w+15 and keeps their A fragments (bf16 hi + lo, 16 k-steps: 128 registers) for the whole launch — the
// weights never leave the register file, there is no 128 KB weight tile in shared memory any more; B = h as bf16 hi / lo rows [b][k] in
// shared memory (ldmatrix), three products, fp32 accumulate in four independent chains.  Each CTA still pushes its 32 new h values per
// batch item (now as bf16 hi / lo, the same 1 KB) into the next-step buffers of all 8 CTAs through distributed shared memory.
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int MB = 8;             // batch items per cluster (the N of the MMAs)
constexpr int HKP = 264;          // bf16 per row of an h tile: 256 + 8 (ldmatrix rows 528 bytes apart fall into distinct bank groups)
constexpr int PRP = 132;          // floats per row of the pre-activation exchange tile [b][gate row]

__device__ __forceinline__ void lstm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void lstm_ldsm4(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}
__device__ __forceinline__ void lstm_split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    const __nv_bfloat16 la = __float2bfloat16_rn(a - __bfloat162float(ha)), lb = __float2bfloat16_rn(b - __bfloat162float(hb));
    hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(la) | ((uint32_t)__bfloat16_as_ushort(lb) << 16);
}
// gate non-linearities on the SFU (ex2.approx + rcp.approx, absolute error ~1e-7; the libm expf / tanhf sequences were ~770 cycles of a step)
__device__ __forceinline__ float lstm_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float lstm_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

constexpr int HBP = 40;           // bf16 per (plane, batch item) row of a source block: 32 units + 8 pad (80-byte rows: conflict-free ldmatrix)
constexpr int HBLK = 2 * MB * HBP;      // bf16 per source block [hi, lo][b][HBP] = 1280 bytes = what one CTA contributes per step

struct LstmMmaSmem {
    static constexpr int H_ELEMS = 2 * LS * HBLK;             // [step parity][source CTA][hi, lo][b][HBP] bf16
    static constexpr size_t BYTES = (size_t)H_ELEMS * 2 + (size_t)2 * HBLK * 2 + (size_t)MB * PRP * 4 + 64;
};

// mbarrier wait that traps instead of hanging the GPU if the exchange ever deadlocks
__device__ __forceinline__ void lstm_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
#pragma unroll 1
    for (int it = 0; it < 400; ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(0x989680u)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// The h exchange.  v1 pushed with per-thread st.shared::cluster + barrier.cluster arrive.release / wait.acquire: 1 160 + 240 cycles of a
// 2 780-cycle step (profiles/r02/time_lstm_v2.log) — the release has to wait for the remote stores.  Here the h tiles are laid out by SOURCE
// CTA ([parity][source][hi, lo][b][32 + 8]: a k-window of 32 = one source block, contiguous 1 280 bytes), each CTA sends its block to every
// peer with ONE bulk copy (cp.async.bulk shared::cta -> shared::cluster) that completes on the receiver's mbarrier, and a step starts when
// the local mbarrier has seen all 8 blocks: data and signal travel together, no cluster barrier inside the loop.
__global__ void __cluster_dims__(LS, 1, 1) __launch_bounds__(256, 1)
lstm_fwd_mma_kernel(const float* __restrict__ gx, const float* __restrict__ whhT, const float* __restrict__ bhh,
                    float* __restrict__ y, float* __restrict__ stash, int T, int B, int bf, long long* dbg) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) uint8_t lsm_raw[];
    __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(lsm_raw);                    // [2][LS][HBLK]
    __nv_bfloat16* hloc = hbuf + LstmMmaSmem::H_ELEMS;                                  // [2][HBLK]: this CTA's new h values (double buffered)
    float* pre = reinterpret_cast<float*>(hloc + 2 * HBLK);                             // [MB][PRP]: W h of this CTA's 128 gate rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(pre + MB * PRP);                       // [2]: all 8 blocks of a parity buffer have landed
    const int r = (int)cluster.block_rank();
    const int dir = blockIdx.z, b0 = blockIdx.y * MB;
    const int nb = (B - b0) < MB ? (B - b0) : MB;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, u = lane;
    const int gid = lane >> 2, tig = lane & 3;
    // A fragments of this warp's 16 gate rows (local row n = gate * 32 + unit -> column gate * 256 + 32 r + unit of whhT [k][4H]):
    //   a0 = (row 16w + gid, k = 16 ks + 2 tig, +1), a1 = (row + 8), a2 = (k + 8), a3 = (row + 8, k + 8)
    uint32_t a_hi[16][4], a_lo[16][4];
    {
        const float* WT = whhT + (size_t)dir * 4 * LH * LH;
        const int n0 = 16 * w + gid, n1 = n0 + 8;
        const int c0 = (n0 >> 5) * LH + LU * r + (n0 & 31), c1 = (n1 >> 5) * LH + LU * r + (n1 & 31);
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
            const int k0 = 16 * ks + 2 * tig;
            lstm_split_pair(__ldg(WT + (size_t)k0 * 4 * LH + c0), __ldg(WT + (size_t)(k0 + 1) * 4 * LH + c0), a_hi[ks][0], a_lo[ks][0]);
            lstm_split_pair(__ldg(WT + (size_t)k0 * 4 * LH + c1), __ldg(WT + (size_t)(k0 + 1) * 4 * LH + c1), a_hi[ks][1], a_lo[ks][1]);
            lstm_split_pair(__ldg(WT + (size_t)(k0 + 8) * 4 * LH + c0), __ldg(WT + (size_t)(k0 + 9) * 4 * LH + c0), a_hi[ks][2], a_lo[ks][2]);
            lstm_split_pair(__ldg(WT + (size_t)(k0 + 8) * 4 * LH + c1), __ldg(WT + (size_t)(k0 + 9) * 4 * LH + c1), a_hi[ks][3], a_lo[ks][3]);
        }
    }
    for (int i = tid; i < (LstmMmaSmem::H_ELEMS + 2 * HBLK) / 2; i += 256) reinterpret_cast<uint32_t*>(hbuf)[i] = 0u;      // h_0 = 0 (and the pads)
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_barrier_init();
    }
    // gate-phase role: thread (pb, u) owns hidden unit 32r + u of batch item b0 + pb
    const int pb = w;
    float bias[4], c = 0.f, ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, hval = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = __ldg(bhh + dir * 4 * LH + g * LH + LU * r + u);
    // ldmatrix source of this lane: matrix j = lane / 8 -> row (batch item) lane % 8, unit offset 8 j inside a 32-unit source block
    const int lrow = lane & 7, lk = (lane >> 3) * 8;
    // push role (threads 0..7): this CTA's block -> slot r of CTA tid's buffers; shared::cluster addresses of the peer's buffer / barriers
    uint32_t peer_buf = 0, peer_bar = 0;
    if (tid < LS) {
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_buf) : "r"((uint32_t)__cvta_generic_to_shared(hbuf + r * HBLK)), "r"(tid));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_bar) : "r"((uint32_t)__cvta_generic_to_shared(bars)), "r"(tid));
    }
    cluster.sync();                                           // every CTA of the cluster is running, has cleared its buffers and armed its barriers
    const bool stamp = dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        const int par = s & 1;
        if (stamp) dbg[s * 6] = clock64();
        float gxv[4] = {0.f, 0.f, 0.f, 0.f};
        if (pb < nb) {
            const float* gp = gx + (LROW(t, b0 + pb) * 2 + dir) * 4 * LH + LU * r + u;
#pragma unroll
            for (int g = 0; g < 4; ++g) gxv[g] = __ldg(gp + g * LH);
        }
        if (s > 0) lstm_mbar_wait(&bars[par], ((s - 1) >> 1) & 1);      // the 8 blocks of h_{s-1} have landed in buffer `par`
        if (stamp) dbg[s * 6 + 5] = clock64();
        const __nv_bfloat16* hb = hbuf + par * LS * HBLK;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 8; ++kp) {                      // source block kp = units 32 kp .. = two k-steps: one ldmatrix.x4 per plane
            uint32_t bh[4], bl[4];
            lstm_ldsm4(bh, hb + kp * HBLK + lrow * HBP + lk);
            lstm_ldsm4(bl, hb + kp * HBLK + MB * HBP + lrow * HBP + lk);
            lstm_mma(acc[(2 * kp) & 3], a_hi[2 * kp], bh[0], bh[1]);
            lstm_mma(acc[(2 * kp + 1) & 3], a_hi[2 * kp + 1], bh[2], bh[3]);
            lstm_mma(acc[(2 * kp) & 3], a_lo[2 * kp], bh[0], bh[1]);
            lstm_mma(acc[(2 * kp + 1) & 3], a_lo[2 * kp + 1], bh[2], bh[3]);
            lstm_mma(acc[(2 * kp) & 3], a_hi[2 * kp], bl[0], bl[1]);
            lstm_mma(acc[(2 * kp + 1) & 3], a_hi[2 * kp + 1], bl[2], bl[3]);
        }
        {
            // D fragment: (row 16w + gid, batch 2 tig, +1), (row + 8, ...) -> pre[b][gate row]
            const int n0 = 16 * w + gid;
            const float d0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]), d1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
            const float d2 = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]), d3 = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
            pre[(2 * tig) * PRP + n0] = d0; pre[(2 * tig + 1) * PRP + n0] = d1;
            pre[(2 * tig) * PRP + n0 + 8] = d2; pre[(2 * tig + 1) * PRP + n0 + 8] = d3;
        }
        if (stamp) dbg[s * 6 + 1] = clock64();
        __syncthreads();
        if (stamp) dbg[s * 6 + 2] = clock64();
        {
            float p4[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) p4[g] = pre[pb * PRP + g * LU + u] + (bias[g] + gxv[g]);
            ig = lstm_sigmoid(p4[0]); fg = lstm_sigmoid(p4[1]); gg = lstm_tanh(p4[2]); og = lstm_sigmoid(p4[3]);
            c = fg * c + ig * gg;
            hval = og * lstm_tanh(c);
            const float hv = pb < nb ? hval : 0.f;
            const __nv_bfloat16 hh = __float2bfloat16_rn(hv);
            __nv_bfloat16* hl = hloc + par * HBLK;
            hl[pb * HBP + u] = hh;
            hl[MB * HBP + pb * HBP + u] = __float2bfloat16_rn(hv - __bfloat162float(hh));
        }
        if (s + 1 < T) {
            tc::fence_proxy_async();                          // the generic-proxy writes of hloc are ordered before the bulk copies read them
            __syncthreads();
            if (stamp) dbg[s * 6 + 3] = clock64();
            const int nxt = par ^ 1;
            if (tid < LS) {
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(peer_buf + (uint32_t)(nxt * LS * HBLK * 2)), "r"((uint32_t)__cvta_generic_to_shared(hloc + par * HBLK)),
                               "r"((uint32_t)(HBLK * 2)), "r"(peer_bar + (uint32_t)(nxt * 8))
                             : "memory");
            } else if (tid == LS) {
                tc::mbar_expect_tx(&bars[nxt], (uint32_t)(LS * HBLK * 2));      // this CTA expects the 8 blocks of h_s in buffer `nxt`
            }
            if (stamp) dbg[s * 6 + 4] = clock64();
        }
        if (pb < nb) {
            y[LROW(t, b0 + pb) * 2 * LH + dir * LH + LU * r + u] = hval;
            if (stash) {
                float* sp = stash + ((LROW(t, b0 + pb) * 2 + dir) * 5) * LH + LU * r + u;
                sp[0] = ig; sp[LH] = fg; sp[2 * LH] = gg; sp[3 * LH] = og; sp[4 * LH] = c;
            }
        }
    }
    cluster.sync();                                           // no CTA leaves while a peer's copy may still be landing in its shared memory
}

// BPTT: dy [T,B,2H] -> dgx [T,B,2,4H] (gradient w.r.t. the gate pre-activations).  whh: original [2][4H][H] layout.  Same cluster
// decomposition: CTA r owns the gates of units [32r, 32r+32) (its 128 rows of W_hh, all 256 columns, resident); per step it forms its
// dgates, multiplies them into a partial dh_prev for ALL 256 units, and reduce-scatters the partials through distributed shared memory
// (slot [source CTA] in the owner's receive buffer, summed in a fixed order: deterministic).
template <int BG>
__global__ void __cluster_dims__(LS, 1, 1) __launch_bounds__(256, 1)
lstm_bwd_cluster_kernel(const float* __restrict__ dy, const float* __restrict__ stash, const float* __restrict__ whh,
                        float* __restrict__ dgx, int T, int B, int bf) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float lsm[];
    float* Wb = lsm;                                          // [gate*32 + u][j]
    float* pbuf = Wb + LstmSmem<BG>::W_FLOATS;                // [2 row halves][j][b]
    float* recv = pbuf + LstmSmem<BG>::HS_FLOATS;             // [2][src][u][b]  (2 * 8 * 32 * BG floats)
    float* dgs = recv + 2 * LS * LU * BG;                     // [gate*32 + u][b]  (128 * BG floats; RED_FLOATS = 1024 * BG covers both)
    const int r = (int)cluster.block_rank();
    const int dir = blockIdx.z, b0 = blockIdx.y * BG;
    const int nb = (B - b0) < BG ? (B - b0) : BG;
    const int tid = threadIdx.x, w = tid >> 5, u = tid & 31;
    constexpr int Q = BG / 4;
    const float* W = whh + (size_t)dir * 4 * LH * LH;         // [gate*256 + unit][j]
    for (int idx = tid; idx < 4 * LU * LH; idx += 256) {
        const int row = idx >> 8, j = idx & 255, g = row >> 5, uu = row & 31;
        Wb[row * LH + j] = __ldg(W + (size_t)(g * LH + LU * r + uu) * LH + j);
    }
    const int pb = w;
    const bool gate_thread = pb < BG;
    const int half = tid >> 7, jp = tid & 127;                // FMA phase: columns 2jp, 2jp+1 over rows [64*half, 64*half + 64)
    float dc = 0.f;
    cluster.sync();
    for (int s = T - 1; s >= 0; --s) {                        // walk the processing order backwards
        const int t = dir ? T - 1 - s : s;
        const int tprev = dir ? t + 1 : t - 1;                // the step processed just before t (holds c_prev)
        if (gate_thread) {
            float di_p = 0.f, df_p = 0.f, dg_p = 0.f, do_p = 0.f;
            if (pb < nb) {
                float dh = 0.f;
                if (s < T - 1) {
                    const float* rv = recv + ((s + 1) & 1) * LS * LU * BG + u * BG + pb;
#pragma unroll
                    for (int src = 0; src < LS; ++src) dh += rv[src * LU * BG];
                }
                const size_t col = (size_t)LU * r + u;
                const float* sp = stash + ((LROW(t, b0 + pb) * 2 + dir) * 5) * LH + col;
                const float ig = sp[0], fg = sp[LH], gg = sp[2 * LH], og = sp[3 * LH], cc = sp[4 * LH];
                const float cprev = s > 0 ? stash[((LROW(tprev, b0 + pb) * 2 + dir) * 5 + 4) * LH + col] : 0.f;
                const float dht = dh + __ldg(dy + LROW(t, b0 + pb) * 2 * LH + dir * LH + col);
                const float tcc = tanhf(cc);
                const float dct = dc + dht * og * (1.f - tcc * tcc);
                do_p = dht * tcc * og * (1.f - og);
                di_p = dct * gg * ig * (1.f - ig);
                dg_p = dct * ig * (1.f - gg * gg);
                df_p = dct * cprev * fg * (1.f - fg);
                dc = dct * fg;
                float* gp = dgx + (LROW(t, b0 + pb) * 2 + dir) * 4 * LH + col;
                gp[0] = di_p; gp[LH] = df_p; gp[2 * LH] = dg_p; gp[3 * LH] = do_p;
            }
            dgs[(0 * LU + u) * BG + pb] = di_p; dgs[(1 * LU + u) * BG + pb] = df_p;
            dgs[(2 * LU + u) * BG + pb] = dg_p; dgs[(3 * LU + u) * BG + pb] = do_p;
        }
        if (s == 0) break;                                    // nothing consumes dh before the first processed step (uniform: no barrier skipped unevenly)
        __syncthreads();
        float acc[2][BG];
#pragma unroll
        for (int b = 0; b < BG; ++b) { acc[0][b] = 0.f; acc[1][b] = 0.f; }
#pragma unroll 4
        for (int rr = 0; rr < 64; ++rr) {
            const int row = 64 * half + rr;
            const float2 wv = *reinterpret_cast<const float2*>(Wb + row * LH + 2 * jp);
            float gv[BG];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float4 g4 = *reinterpret_cast<const float4*>(dgs + row * BG + 4 * q);
                gv[4 * q] = g4.x; gv[4 * q + 1] = g4.y; gv[4 * q + 2] = g4.z; gv[4 * q + 3] = g4.w;
            }
#pragma unroll
            for (int b = 0; b < BG; ++b) {
                acc[0][b] = fmaf(wv.x, gv[b], acc[0][b]);
                acc[1][b] = fmaf(wv.y, gv[b], acc[1][b]);
            }
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int q = 0; q < Q; ++q)
                *reinterpret_cast<float4*>(pbuf + (half * LH + 2 * jp + jj) * BG + 4 * q) =
                    make_float4(acc[jj][4 * q], acc[jj][4 * q + 1], acc[jj][4 * q + 2], acc[jj][4 * q + 3]);
        __syncthreads();
        // reduce-scatter: the partial dh_prev of units [32 dst, 32 dst + 32) goes to slot r of CTA dst's receive buffer (parity s & 1)
        for (int i = tid; i < LS * LU * Q; i += 256) {
            const int dst = i / (LU * Q), q = i - dst * (LU * Q);
            const float4 a = *reinterpret_cast<const float4*>(pbuf + (LU * dst) * BG + 4 * q);
            const float4 b4 = *reinterpret_cast<const float4*>(pbuf + (LH + LU * dst) * BG + 4 * q);
            float* remote = cluster.map_shared_rank(recv, dst);
            *reinterpret_cast<float4*>(remote + ((s & 1) * LS + r) * LU * BG + 4 * q) = make_float4(a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w);
        }
        cluster.sync();
    }
    cluster.sync();                                           // no CTA leaves while a peer may still be writing into its shared memory
}

// Tensor-core BPTT (round 2): the transposed product on the same decomposition.  CTA r owns the gates of units [32r, 32r+32); per step
// its gate threads form the 128 x 8 gate derivatives (bf16 hi / lo tile [b][gate row], the B operand), the MMAs give the partial
//   dh_prev^T[256 units][8] = W_slice^T[256][128] . dgates^T[128][8]
// (warp w: units 32w .. 32w+31 = exactly the units CTA w owns, A fragments of W^T resident in 128 registers), and every lane stores its
// accumulator fragment straight into slot r of CTA w's receive buffer through distributed shared memory: no partial-sum staging, one
// __syncthreads and one cluster barrier per step.  The owner sums the 8 slots in a fixed order (deterministic).
constexpr int DGP = 136;          // bf16 per row of the gate-derivative tile [b][128 gate rows + 8]

struct LstmBwdMmaSmem {
    // dgate tile (hi, lo) + receive slots [2][src][u][b] + send staging [2][dst][u][b] + 2 mbarriers
    static constexpr size_t BYTES = (size_t)2 * MB * DGP * 2 + (size_t)2 * LS * LU * MB * 4 + (size_t)2 * LS * LU * MB * 4 + 64;
};

__global__ void __cluster_dims__(LS, 1, 1) __launch_bounds__(256, 1)
lstm_bwd_mma_kernel(const float* __restrict__ dy, const float* __restrict__ stash, const float* __restrict__ whh,
                    float* __restrict__ dgx, int T, int B, int bf) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) uint8_t lsm_raw[];
    __nv_bfloat16* dgs = reinterpret_cast<__nv_bfloat16*>(lsm_raw);                     // [hi, lo][MB][DGP]
    float* recv = reinterpret_cast<float*>(dgs + 2 * MB * DGP);                         // [2][src][u][b]
    float* pst = recv + 2 * LS * LU * MB;                                               // [2][dst][u][b]: partial dh_prev on its way to the owners
    uint64_t* bars = reinterpret_cast<uint64_t*>(pst + 2 * LS * LU * MB);               // [2]: the 8 partial blocks of a parity have landed
    const int r = (int)cluster.block_rank();
    const int dir = blockIdx.z, b0 = blockIdx.y * MB;
    const int nb = (B - b0) < MB ? (B - b0) : MB;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, u = lane;
    const int gid = lane >> 2, tig = lane & 3;
    // A fragments of W^T for this warp's two 16-unit tiles: a0 = (unit j = 32w + 16mt + gid, k = gate row 16ks + 2 tig, +1), a1 = (j + 8),
    // a2 = (k + 8), a3 = (j + 8, k + 8); gate row n (local) = gate * 32 + unit  ->  row gate * 256 + 32 r + unit of whh [4H][H]
    uint32_t a_hi[2][8][4], a_lo[2][8][4];
    {
        const float* W = whh + (size_t)dir * 4 * LH * LH;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int j0 = 32 * w + 16 * mt + gid, j1 = j0 + 8;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
                for (int hk = 0; hk < 2; ++hk) {
                    const int n = 16 * ks + 8 * hk + 2 * tig;                  // n and n + 1 lie in the same gate (n even, 32 | gate size)
                    const size_t g0 = (size_t)((n >> 5) * LH + LU * r + (n & 31)) * LH;
                    lstm_split_pair(__ldg(W + g0 + j0), __ldg(W + g0 + LH + j0), a_hi[mt][ks][2 * hk], a_lo[mt][ks][2 * hk]);
                    lstm_split_pair(__ldg(W + g0 + j1), __ldg(W + g0 + LH + j1), a_hi[mt][ks][2 * hk + 1], a_lo[mt][ks][2 * hk + 1]);
                }
            }
        }
    }
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_barrier_init();
    }
    const int pb = w;
    const int lrow = lane & 7, lk = (lane >> 3) * 8;
    // exchange (as in the forward kernel): warp w's accumulator fragments are the partial dh_prev of the units CTA w owns; they are staged
    // as block w, and thread j < 8 sends block j to slot r of CTA j's receive buffer with one bulk copy that completes on CTA j's mbarrier
    uint32_t peer_buf = 0, peer_bar = 0;
    if (tid < LS) {
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_buf) : "r"((uint32_t)__cvta_generic_to_shared(recv + r * LU * MB)), "r"(tid));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_bar) : "r"((uint32_t)__cvta_generic_to_shared(bars)), "r"(tid));
    }
    float dc = 0.f;
    cluster.sync();
    for (int s = T - 1; s >= 0; --s) {                        // walk the processing order backwards
        const int t = dir ? T - 1 - s : s;
        const int tprev = dir ? t + 1 : t - 1;                // the step processed just before t (holds c_prev)
        {
            float di_p = 0.f, df_p = 0.f, dg_p = 0.f, do_p = 0.f;
            float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, cc = 0.f, cprev = 0.f, dyv = 0.f;
            const size_t col = (size_t)LU * r + u;
            if (pb < nb) {                                    // operands first: they do not depend on the exchange
                const float* sp = stash + ((LROW(t, b0 + pb) * 2 + dir) * 5) * LH + col;
                ig = sp[0]; fg = sp[LH]; gg = sp[2 * LH]; og = sp[3 * LH]; cc = sp[4 * LH];
                cprev = s > 0 ? stash[((LROW(tprev, b0 + pb) * 2 + dir) * 5 + 4) * LH + col] : 0.f;
                dyv = __ldg(dy + LROW(t, b0 + pb) * 2 * LH + dir * LH + col);
            }
            if (s < T - 1) lstm_mbar_wait(&bars[(s + 1) & 1], (((T - 2) - s) >> 1) & 1);      // the partials pushed during step s + 1 have landed
            if (pb < nb) {
                float dh = 0.f;
                if (s < T - 1) {
                    const float* rv = recv + ((s + 1) & 1) * LS * LU * MB + u * MB + pb;
#pragma unroll
                    for (int src = 0; src < LS; ++src) dh += rv[src * LU * MB];
                }
                const float dht = dh + dyv;
                const float tcc = lstm_tanh(cc);
                const float dct = dc + dht * og * (1.f - tcc * tcc);
                do_p = dht * tcc * og * (1.f - og);
                di_p = dct * gg * ig * (1.f - ig);
                dg_p = dct * ig * (1.f - gg * gg);
                df_p = dct * cprev * fg * (1.f - fg);
                dc = dct * fg;
                float* gp = dgx + (LROW(t, b0 + pb) * 2 + dir) * 4 * LH + col;
                gp[0] = di_p; gp[LH] = df_p; gp[2 * LH] = dg_p; gp[3 * LH] = do_p;
            }
            const float dv[4] = {di_p, df_p, dg_p, do_p};
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const __nv_bfloat16 hh = __float2bfloat16_rn(dv[g]);
                dgs[pb * DGP + g * LU + u] = hh;
                dgs[MB * DGP + pb * DGP + g * LU + u] = __float2bfloat16_rn(dv[g] - __bfloat162float(hh));
            }
        }
        if (s == 0) break;                                    // nothing consumes dh before the first processed step (uniform)
        __syncthreads();
        float acc[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) acc[mt][c2][0] = acc[mt][c2][1] = acc[mt][c2][2] = acc[mt][c2][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) {                      // two k-steps (32 gate rows) per ldmatrix.x4
            uint32_t bh[4], bl[4];
            lstm_ldsm4(bh, dgs + lrow * DGP + 32 * kp + lk);
            lstm_ldsm4(bl, dgs + MB * DGP + lrow * DGP + 32 * kp + lk);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                lstm_mma(acc[mt][0], a_hi[mt][2 * kp], bh[0], bh[1]);
                lstm_mma(acc[mt][1], a_hi[mt][2 * kp + 1], bh[2], bh[3]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                lstm_mma(acc[mt][0], a_lo[mt][2 * kp], bh[0], bh[1]);
                lstm_mma(acc[mt][1], a_lo[mt][2 * kp + 1], bh[2], bh[3]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                lstm_mma(acc[mt][0], a_hi[mt][2 * kp], bl[0], bl[1]);
                lstm_mma(acc[mt][1], a_hi[mt][2 * kp + 1], bl[2], bl[3]);
            }
        }
        // reduce-scatter: block w of the staging area = this warp's partial dh_prev of the units CTA w owns: rows u = 16 mt + gid (+ 8), b = 2 tig, +1
        {
            const int par = s & 1;
            float* dst = pst + (size_t)(par * LS + w) * LU * MB + gid * MB + 2 * tig;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                *reinterpret_cast<float2*>(dst + (16 * mt) * MB) = make_float2(acc[mt][0][0] + acc[mt][1][0], acc[mt][0][1] + acc[mt][1][1]);
                *reinterpret_cast<float2*>(dst + (16 * mt + 8) * MB) = make_float2(acc[mt][0][2] + acc[mt][1][2], acc[mt][0][3] + acc[mt][1][3]);
            }
            tc::fence_proxy_async();
            __syncthreads();
            if (tid < LS) {
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(peer_buf + (uint32_t)(par * LS * LU * MB * 4)), "r"((uint32_t)__cvta_generic_to_shared(pst + (size_t)(par * LS + tid) * LU * MB)),
                               "r"((uint32_t)(LU * MB * 4)), "r"(peer_bar + (uint32_t)(par * 8))
                             : "memory");
            } else if (tid == LS) {
                tc::mbar_expect_tx(&bars[par], (uint32_t)(LS * LU * MB * 4));
            }
        }
    }
    cluster.sync();                                           // no CTA leaves while a peer's copy may still be landing in its shared memory
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

long long* g_lstm_dbg = nullptr;
int g_lstm_mma = 1;       // 1: tensor-core recurrence kernels; 0: the FFMA cluster kernels (tpgsr_lstm_set_mode: cross-check tests, A/B)

}  // namespace

extern "C" {

// developer instrumentation: clock64 stamps of the first CTA of tpgsr_lstm_fwd, 6 per step
int tpgsr_lstm_set_debug(long long* dbg) { g_lstm_dbg = dbg; return TPGSR_OK; }
// 1 (default): tensor-core recurrence kernels; 0: the FFMA cluster kernels
int tpgsr_lstm_set_mode(int mma) { g_lstm_mma = mma ? 1 : 0; return TPGSR_OK; }

int tpgsr_lstm_transpose_whh(const float* whh, float* whhT, cudaStream_t stream) {
    if (!whh || !whhT) return TPGSR_ERR_BAD_ARG;
    lstm_transpose_whh_kernel<<<dim3(LH / 32, 4 * LH / 32, 2), dim3(32, 8), 0, stream>>>(whh, whhT);
    return tpgsr_launch_status();
}

int tpgsr_lstm_fwd(const float* gx, const float* whhT, const float* bhh, float* y, float* stash, int T, int B, int hidden,
                   int batch_first, cudaStream_t stream) {
    if (!gx || !whhT || !bhh || !y || T <= 0 || B <= 0) return TPGSR_ERR_BAD_ARG;
    if (hidden != LH) return TPGSR_ERR_UNSUPPORTED;
    if (g_lstm_mma) {
        if (cudaFuncSetAttribute(lstm_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LstmMmaSmem::BYTES) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        lstm_fwd_mma_kernel<<<dim3(LS, (B + MB - 1) / MB, 2), 256, LstmMmaSmem::BYTES, stream>>>(gx, whhT, bhh, y, stash, T, B, batch_first, g_lstm_dbg);
        return tpgsr_launch_status();
    }
    if (B <= 4) {
        constexpr int BG = 4;
        if (cudaFuncSetAttribute(lstm_fwd_cluster_kernel<BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LstmSmem<BG>::BYTES) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        lstm_fwd_cluster_kernel<BG><<<dim3(LS, (B + BG - 1) / BG, 2), 256, LstmSmem<BG>::BYTES, stream>>>(gx, whhT, bhh, y, stash, T, B, batch_first, g_lstm_dbg);
    } else {
        constexpr int BG = 8;
        if (cudaFuncSetAttribute(lstm_fwd_cluster_kernel<BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LstmSmem<BG>::BYTES) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        lstm_fwd_cluster_kernel<BG><<<dim3(LS, (B + BG - 1) / BG, 2), 256, LstmSmem<BG>::BYTES, stream>>>(gx, whhT, bhh, y, stash, T, B, batch_first, g_lstm_dbg);
    }
    return tpgsr_launch_status();
}

int tpgsr_lstm_bwd(const float* dy, const float* stash, const float* whh, float* dgx, int T, int B, int hidden,
                   int batch_first, cudaStream_t stream) {
    if (!dy || !stash || !whh || !dgx || T <= 0 || B <= 0) return TPGSR_ERR_BAD_ARG;
    if (hidden != LH) return TPGSR_ERR_UNSUPPORTED;
    if (g_lstm_mma) {
        if (cudaFuncSetAttribute(lstm_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LstmBwdMmaSmem::BYTES) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        lstm_bwd_mma_kernel<<<dim3(LS, (B + MB - 1) / MB, 2), 256, LstmBwdMmaSmem::BYTES, stream>>>(dy, stash, whh, dgx, T, B, batch_first);
        return tpgsr_launch_status();
    }
    if (B <= 4) {
        constexpr int BG = 4;
        if (cudaFuncSetAttribute(lstm_bwd_cluster_kernel<BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LstmSmem<BG>::BYTES) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        lstm_bwd_cluster_kernel<BG><<<dim3(LS, (B + BG - 1) / BG, 2), 256, LstmSmem<BG>::BYTES, stream>>>(dy, stash, whh, dgx, T, B, batch_first);
    } else {
        constexpr int BG = 8;
        if (cudaFuncSetAttribute(lstm_bwd_cluster_kernel<BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LstmSmem<BG>::BYTES) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        lstm_bwd_cluster_kernel<BG><<<dim3(LS, (B + BG - 1) / BG, 2), 256, LstmSmem<BG>::BYTES, stream>>>(dy, stash, whh, dgx, T, B, batch_first);
    }
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
