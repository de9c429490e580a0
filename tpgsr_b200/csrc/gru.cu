// Bidirectional GRU recurrence (hidden 32 per direction) over rows or columns of an NHWC feature map.
//
// Reference op replaced: nn.GRU(64, 32, bidirectional=True, batch_first=True) inside GruBlock
// (model/tsrn.py:491-508; cuDNN RNN in the reference).  The input projection W_ih x + b_ih is a plain GEMM done by the
// conv/GEMM kernels; this file is the serial part: one warp per (sequence, direction), lane = hidden unit, the recurrent
// matrix held in registers (rows of W_hh forward, columns backward), h / gate gradients broadcast through shared memory.
// Gate order r,z,n;  n = tanh(gx_n + r * (W_hn h + b_hn));  h' = (1-z) n + z h   (SURVEY Appendix B).
//
// Sequences are addressed on the pixel-row index m of the NHWC map so that the column-wise GRU (gru1, which the reference
// runs on a transposed copy) needs no transpose:  m(s,t) = (s / inner) * outer_stride + (s % inner) * inner_stride + t * t_stride.
//
// Round-1 profile (profiles/r01): the first version prefetched one step ahead (forward) / not at all (backward) and ran at
// ~1 700 cycles per serial step = DRAM latency.  This version keeps a PF-deep register ring of every per-step operand, sizes
// the grid to one resident wave, issues the 32x96 mat-vec as packed fp32 FMAs (FFMA2, sm_100) and moves the dW_hh outer products
// out of the serial loop into their own reduction kernel, so the recurrences are issue/HBM bound instead of latency bound.
// Traffic per (pixel, direction): forward reads 96 gate pre-activations, writes 32 outputs (+32 fused residual sum) + 128 stash
// floats; backward reads 32 dy + 32 h + 128 stash, writes 96.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

struct SeqGeom {
    int S, T, inner;
    long outer_stride, inner_stride, t_stride;
};

__device__ __forceinline__ long seq_base(const SeqGeom& g, int s) {
    return (long)(s / g.inner) * g.outer_stride + (long)(s % g.inner) * g.inner_stride;
}

constexpr int H = 32;          // hidden units per direction
constexpr int WARPS = 4;
constexpr int PF = 8;          // prefetch depth (steps) of the per-warp cp.async rings in shared memory

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// 4-byte asynchronous global -> shared copy (LDGSTS): the per-step operands of step t + PF are in flight while step t computes.
// (A register ring did not work: ptxas loaded into one temporary and moved it into the ring slot right away, so every step
// still waited for its own load — profiles/r01 source view.)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// gate non-linearities on the SFU (ex2.approx + rcp.approx: ~2 ulp): sigma(x) = 1 / (1 + e^-x),  tanh(x) = 1 - 2 / (1 + e^2x)
// (absolute error ~1e-7; the exact expf / IEEE divide / tanhf sequences were ~100 of the ~265 instructions of a step)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

constexpr int PFB = 6;         // backward ring depth (6 operands x 2 sequences per step: 48 KB static shared memory limit)
// NS (template parameter below): sequences advanced together by one warp = independent dependency chains (instruction-level
// parallelism).  2 when there are enough sequences to fill the chip with half as many warps, 1 otherwise (the T=64 row GRU).

// gx: [M, 192] = [dir][gate r,z,n][32] pre-activations from the input side (bias_ih included)
// y:  [M, ldy] GRU output (dir0 -> cols 0..31, dir1 -> cols 32..63); ysum = y + residual (optional)
// stash: [M][2][4][32] = r, z, n, (W_hn h + b_hn)   (training only; may be NULL)
template <int NS>
__global__ void __launch_bounds__(WARPS * 32, 3)
gru_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const float* __restrict__ bhh,
               float* __restrict__ y, int ldy, const float* __restrict__ residual, int ldr, float* __restrict__ ysum,
               int ldys, float* __restrict__ stash, SeqGeom g) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    // row `lane` of W_hr / W_hz / W_hn as 16 (even, odd) pairs: one FFMA2 covers two recurrent inputs
    float2 wr[H / 2], wz[H / 2], wn[H / 2];
    const float* W = whh + (size_t)dir * 3 * H * H;
#pragma unroll
    for (int i = 0; i < H / 2; ++i) {
        wr[i] = __ldg(reinterpret_cast<const float2*>(W + (0 * H + lane) * H) + i);
        wz[i] = __ldg(reinterpret_cast<const float2*>(W + (1 * H + lane) * H) + i);
        wn[i] = __ldg(reinterpret_cast<const float2*>(W + (2 * H + lane) * H) + i);
    }
    const float bhr = __ldg(bhh + dir * 3 * H + lane), bhz = __ldg(bhh + dir * 3 * H + H + lane),
                bhn = __ldg(bhh + dir * 3 * H + 2 * H + lane);
    const bool has_res = ysum != nullptr;
    __shared__ float ring[WARPS][PF][NS][4][H];       // [step slot][sequence][gx_r, gx_z, gx_n, residual][lane]
    __shared__ __align__(16) float hb[WARPS][NS][H];

    for (int s0 = (blockIdx.x * WARPS + warp) * NS; s0 < g.S; s0 += gridDim.x * WARPS * NS) {
        long m0[NS];
        bool live[NS];
#pragma unroll
        for (int v = 0; v < NS; ++v) {
            live[v] = s0 + v < g.S;
            m0[v] = seq_base(g, live[v] ? s0 + v : s0) + (dir ? (long)(g.T - 1) * g.t_stride : 0);
        }
        const long step = dir ? -g.t_stride : g.t_stride;
        // one commit group per step (empty beyond the end) keeps the wait_group arithmetic uniform
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (u < g.T) {
#pragma unroll
                for (int v = 0; v < NS; ++v) {
                    const long m = m0[v] + u * step;
                    const float* gp = gx + (size_t)m * 6 * H + dir * 3 * H + lane;
                    cp_async4(&ring[warp][u][v][0][lane], gp);
                    cp_async4(&ring[warp][u][v][1][lane], gp + H);
                    cp_async4(&ring[warp][u][v][2][lane], gp + 2 * H);
                    if (has_res) cp_async4(&ring[warp][u][v][3][lane], residual + (size_t)m * ldr + dir * H + lane);
                }
            }
            cp_async_commit();
        }
        float h[NS];
        __syncwarp();
#pragma unroll
        for (int v = 0; v < NS; ++v) { h[v] = 0.f; hb[warp][v][lane] = 0.f; }
        __syncwarp();
        for (int t = 0; t < g.T; ++t) {
            const int slot = t % PF;
            cp_async_wait<PF - 1>();               // this lane's copies for step t have landed (each lane reads only its own)
            float gxr[NS], gxz[NS], gxn[NS], res[NS];
#pragma unroll
            for (int v = 0; v < NS; ++v) {
                gxr[v] = ring[warp][slot][v][0][lane]; gxz[v] = ring[warp][slot][v][1][lane]; gxn[v] = ring[warp][slot][v][2][lane];
                res[v] = has_res ? ring[warp][slot][v][3][lane] : 0.f;
            }
            if (t + PF < g.T) {                    // refill the slot just consumed with the operands of step t + PF
#pragma unroll
                for (int v = 0; v < NS; ++v) {
                    const long mn = m0[v] + (long)(t + PF) * step;
                    const float* np = gx + (size_t)mn * 6 * H + dir * 3 * H + lane;
                    cp_async4(&ring[warp][slot][v][0][lane], np);
                    cp_async4(&ring[warp][slot][v][1][lane], np + H);
                    cp_async4(&ring[warp][slot][v][2][lane], np + 2 * H);
                    if (has_res) cp_async4(&ring[warp][slot][v][3][lane], residual + (size_t)mn * ldr + dir * H + lane);
                }
            }
            cp_async_commit();
            float2 ar[NS], az[NS], an[NS];
#pragma unroll
            for (int v = 0; v < NS; ++v) { ar[v] = make_float2(bhr, 0.f); az[v] = make_float2(bhz, 0.f); an[v] = make_float2(bhn, 0.f); }
#pragma unroll
            for (int q = 0; q < H / 4; ++q) {
#pragma unroll
                for (int v = 0; v < NS; ++v) {
                    const float4 hv = *reinterpret_cast<const float4*>(&hb[warp][v][q * 4]);
                    const float2 h01 = make_float2(hv.x, hv.y), h23 = make_float2(hv.z, hv.w);
                    ar[v] = ffma2(wr[2 * q], h01, ar[v]); az[v] = ffma2(wz[2 * q], h01, az[v]); an[v] = ffma2(wn[2 * q], h01, an[v]);
                    ar[v] = ffma2(wr[2 * q + 1], h23, ar[v]); az[v] = ffma2(wz[2 * q + 1], h23, az[v]); an[v] = ffma2(wn[2 * q + 1], h23, an[v]);
                }
            }
            __syncwarp();                          // every lane has read the old h before it is overwritten
#pragma unroll
            for (int v = 0; v < NS; ++v) {
                const float hn = an[v].x + an[v].y;
                const float r = fast_sigmoid(gxr[v] + (ar[v].x + ar[v].y));
                const float z = fast_sigmoid(gxz[v] + (az[v].x + az[v].y));
                const float n = fast_tanh(gxn[v] + r * hn);
                h[v] = (1.f - z) * n + z * h[v];
                hb[warp][v][lane] = h[v];
                if (live[v]) {
                    const long m = m0[v] + (long)t * step;
                    y[(size_t)m * ldy + dir * H + lane] = h[v];
                    if (has_res) ysum[(size_t)m * ldys + dir * H + lane] = h[v] + res[v];
                    if (stash) {
                        float* sp = stash + ((size_t)m * 2 + dir) * 4 * H + lane;
                        sp[0] = r; sp[H] = z; sp[2 * H] = n; sp[3 * H] = hn;
                    }
                }
            }
            __syncwarp();
        }
        cp_async_wait<0>();
    }
}

// ============================================================================================================ tensor-core recurrence
// Round 2.  The warp-per-sequence kernels above are instruction-issue bound: ~265 instructions per (sequence, step), the same 51 us
// for the row (T = 64) and the column (T = 16) recurrence at B = 48 (profiles/r02/hbm_kernels_call11.log).  Here a CTA of 4 warps
// advances 16 sequences of one direction together and the 16 x 96 x 32 mat-vec batch of a step runs on the tensor cores
// (mma.sync m16n8k16, bf16 operands split hi + lo, three products, fp32 accumulate — the same split arithmetic as the convolutions):
//   gates^T[16 seq][96] = h^T[16 seq][32] . W_hh^T[32][96]
//   * A = h^T as bf16 hi / lo tiles in shared memory (double buffered, one __syncthreads per step), read with ldmatrix;
//   * B = W_hh^T: warp w owns hidden units 8w .. 8w+7 of the three gates — 24 registers of fragments, loaded once;
//   * the accumulator fragment of a lane is (sequences g, g+8) x (units 8w + 2q, +1): exactly the elements whose gx / stash / y
//     it loads and stores as float2 (8 rows x 32 bytes per warp request: whole sectors), and whose h it keeps in fp32 registers.
// ~48 instructions per (sequence, step) instead of 265; the per-step operands arrive through a 4-deep cp.async ring as before.
constexpr int GS = 16;          // sequences per CTA
constexpr int HP = 40;          // bf16 per row of the h tile (32 + 8 pad: the 8 rows of an ldmatrix phase fall into distinct banks)
constexpr int PFM = 4;          // cp.async ring depth (steps)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// (a, b) -> packed bf16 pair of the high parts (a in the low half) and of the residuals
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    const __nv_bfloat16 la = __float2bfloat16_rn(a - __bfloat162float(ha)), lb = __float2bfloat16_rn(b - __bfloat162float(hb));
    hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(la) | ((uint32_t)__bfloat16_as_ushort(lb) << 16);
}

// RES: a residual is added into ysum; STASH: training (the gates are kept for BPTT).
// A CTA has about one warp per scheduler, so every dependent instruction's latency is exposed and the step time is the length of the
// instruction stream of the slowest warp (measured: ~1 000 cycles per step at ~300 instructions, whatever their mix —
// profiles/r02/time_gru_v1.log, gru_probe.log).  Hence warp specialisation:
//   * warps 0-3 (compute) run ONLY the critical chain: ldmatrix -> 18 MMAs from zero accumulators (145 cycles) -> + gx from the staged
//     tile -> gates (154-cycle MUFU chains, profiles/r02/hmma_lat.log) -> bf16 split -> shared stores of h_{t+1} and of the step's results;
//   * warps 4-7 (I/O) move everything else: cp.async of the gx / residual tiles PFM - 1 steps ahead and the previous step's y, ysum and
//     stash from the output tiles to global memory, with 16-byte accesses on whole 128-byte lines (a lane's own fragment elements are
//     8 rows x 32 bytes per warp request: 640 sector requests per step against 160 line requests through the tiles).
// One __syncthreads (256 threads) per step hands h_{t+1}, the output tiles and the landed gx tile over.
constexpr int GXP = 104;        // floats per row of a staged gx tile (96 + 8: the 8 rows of a half-warp's float2 reads hit distinct banks)
constexpr int OP = 40;          // floats per row of the res / output tiles
constexpr int SLOT_F = GS * GXP + GS * OP;      // floats per ring slot: gx tile + residual tile
constexpr size_t gru_mma_smem(bool stash) { return (size_t)PFM * SLOT_F * 4 + (size_t)2 * (stash ? 5 : 1) * GS * OP * 4 + (size_t)2 * 2 * GS * HP * 2; }

template <bool RES, bool STASH>
__global__ void __launch_bounds__(256)
gru_fwd_mma_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const float* __restrict__ bhh, float* __restrict__ y,
                   int ldy, const float* __restrict__ residual, int ldr, float* __restrict__ ysum, int ldys,
                   float* __restrict__ stash, SeqGeom g, long long* dbg) {
    constexpr int NQ = STASH ? 5 : 1;                  // output tiles per step: h, r, z, n, W_hn h + b_hn
    extern __shared__ __align__(16) uint8_t gru_smem[];
    float* ring = reinterpret_cast<float*>(gru_smem);                                   // [PFM][gx 16 x GXP | res 16 x OP]
    float* otile = ring + PFM * SLOT_F;                                                 // [2][NQ][16][OP]
    __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(otile + 2 * NQ * GS * OP);   // [2][hi, lo][16][HP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;

    if (warp < 4) {
        // ======================================================= compute warps
        const int gid = lane >> 2, tig = lane & 3;
        const int uu = warp * 8 + 2 * tig;             // this lane's two hidden units (of each gate)
        // B fragments of W_hh^T for gate q: b0 = (k = 2 tig, +1; n = gid), b1 = (k + 8); k-step ks covers k = 16 ks ..; n = unit 8 warp + gid
        uint32_t w_hi[3][2][2], w_lo[3][2][2];
        const float* W = whh + (size_t)dir * 3 * H * H;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float2 wv = __ldg(reinterpret_cast<const float2*>(W + (size_t)(q * H + warp * 8 + gid) * H + 16 * ks + 8 * j + 2 * tig));
                    split_pair(wv.x, wv.y, w_hi[q][ks][j], w_lo[q][ks][j]);
                }
        const float2 bhr = __ldg(reinterpret_cast<const float2*>(bhh + dir * 3 * H + uu));
        const float2 bhz = __ldg(reinterpret_cast<const float2*>(bhh + dir * 3 * H + H + uu));
        const float2 bhn = __ldg(reinterpret_cast<const float2*>(bhh + dir * 3 * H + 2 * H + uu));
        // ldmatrix row address of this lane: matrix j = lane / 8 -> rows (j & 1) * 8 + lane % 8, k offset (j >> 1) * 8
        const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lk = (lane >> 4) * 8;
        for (int grp = blockIdx.x; grp * GS < g.S; grp += gridDim.x) {
            for (int i = threadIdx.x; i < 2 * GS * HP / 2; i += 128) reinterpret_cast<uint32_t*>(hbuf)[i] = 0u;      // h_0 = 0 (both planes of buffer 0)
            float h[4] = {0.f, 0.f, 0.f, 0.f};           // (seq gid, unit uu), (gid, uu+1), (gid+8, uu), (gid+8, uu+1)
            __syncthreads();
            const bool stamp = dbg && blockIdx.x == 0 && dir == 0 && threadIdx.x == 0 && grp == 0;      // developer timeline (scripts/time_gru.py)
            for (int t = 0; t < g.T; ++t) {
                const int slot = t % PFM, cur = t & 1;
                if (stamp) dbg[t * 4] = clock64();
                uint32_t a_hi[2][4], a_lo[2][4];
                const __nv_bfloat16* hb = hbuf + cur * 2 * GS * HP;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    ldmatrix_x4(a_hi[ks], hb + lrow * HP + 16 * ks + lk);
                    ldmatrix_x4(a_lo[ks], hb + GS * HP + lrow * HP + 16 * ks + lk);
                }
                // zero accumulators, one set per (gate, k-step): chains of 3 dependent MMAs
                float aR[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                float aZ[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                float aN[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    mma_bf16_16816(aR[ks], a_hi[ks], w_hi[0][ks][0], w_hi[0][ks][1]);
                    mma_bf16_16816(aZ[ks], a_hi[ks], w_hi[1][ks][0], w_hi[1][ks][1]);
                    mma_bf16_16816(aN[ks], a_hi[ks], w_hi[2][ks][0], w_hi[2][ks][1]);
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    mma_bf16_16816(aR[ks], a_lo[ks], w_hi[0][ks][0], w_hi[0][ks][1]);
                    mma_bf16_16816(aZ[ks], a_lo[ks], w_hi[1][ks][0], w_hi[1][ks][1]);
                    mma_bf16_16816(aN[ks], a_lo[ks], w_hi[2][ks][0], w_hi[2][ks][1]);
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    mma_bf16_16816(aR[ks], a_hi[ks], w_lo[0][ks][0], w_lo[0][ks][1]);
                    mma_bf16_16816(aZ[ks], a_hi[ks], w_lo[1][ks][0], w_lo[1][ks][1]);
                    mma_bf16_16816(aN[ks], a_hi[ks], w_lo[2][ks][0], w_lo[2][ks][1]);
                }
                // this step's gx (in the shadow of the MMAs)
                const float* gs = ring + slot * SLOT_F + uu;
                const float2 gr[2] = {*reinterpret_cast<const float2*>(gs + gid * GXP), *reinterpret_cast<const float2*>(gs + (gid + 8) * GXP)};
                const float2 gz[2] = {*reinterpret_cast<const float2*>(gs + gid * GXP + H), *reinterpret_cast<const float2*>(gs + (gid + 8) * GXP + H)};
                const float2 gn[2] = {*reinterpret_cast<const float2*>(gs + gid * GXP + 2 * H), *reinterpret_cast<const float2*>(gs + (gid + 8) * GXP + 2 * H)};
                float rr[4], zz[4], nn[4], hh[4];
                if (stamp) dbg[t * 4 + 1] = clock64() + (long long)(aN[0][0] * 0.f);      // (depends on an MMA result: stamps their completion)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int sq = e >> 1;
                    const float bR = (e & 1) ? bhr.y : bhr.x, bZ = (e & 1) ? bhz.y : bhz.x, bN = (e & 1) ? bhn.y : bhn.x;
                    const float gR = (e & 1) ? gr[sq].y : gr[sq].x, gZ = (e & 1) ? gz[sq].y : gz[sq].x, gN = (e & 1) ? gn[sq].y : gn[sq].x;
                    hh[e] = (aN[0][e] + aN[1][e]) + bN;
                    rr[e] = fast_sigmoid((aR[0][e] + aR[1][e]) + (gR + bR));
                    zz[e] = fast_sigmoid((aZ[0][e] + aZ[1][e]) + (gZ + bZ));
                    nn[e] = fast_tanh(gN + rr[e] * hh[e]);
                    h[e] = (1.f - zz[e]) * nn[e] + zz[e] * h[e];
                }
                if (stamp) dbg[t * 4 + 2] = clock64() + (long long)(h[3] * 0.f);
                {
                    __nv_bfloat16* hn_ = hbuf + (cur ^ 1) * 2 * GS * HP;
                    uint32_t hi, lo;
                    split_pair(h[0], h[1], hi, lo);
                    *reinterpret_cast<uint32_t*>(hn_ + gid * HP + uu) = hi;
                    *reinterpret_cast<uint32_t*>(hn_ + GS * HP + gid * HP + uu) = lo;
                    split_pair(h[2], h[3], hi, lo);
                    *reinterpret_cast<uint32_t*>(hn_ + (gid + 8) * HP + uu) = hi;
                    *reinterpret_cast<uint32_t*>(hn_ + GS * HP + (gid + 8) * HP + uu) = lo;
                    float* ot = otile + cur * NQ * GS * OP + uu;
                    *reinterpret_cast<float2*>(ot + gid * OP) = make_float2(h[0], h[1]);
                    *reinterpret_cast<float2*>(ot + (gid + 8) * OP) = make_float2(h[2], h[3]);
                    if (STASH) {
                        *reinterpret_cast<float2*>(ot + 1 * GS * OP + gid * OP) = make_float2(rr[0], rr[1]);
                        *reinterpret_cast<float2*>(ot + 1 * GS * OP + (gid + 8) * OP) = make_float2(rr[2], rr[3]);
                        *reinterpret_cast<float2*>(ot + 2 * GS * OP + gid * OP) = make_float2(zz[0], zz[1]);
                        *reinterpret_cast<float2*>(ot + 2 * GS * OP + (gid + 8) * OP) = make_float2(zz[2], zz[3]);
                        *reinterpret_cast<float2*>(ot + 3 * GS * OP + gid * OP) = make_float2(nn[0], nn[1]);
                        *reinterpret_cast<float2*>(ot + 3 * GS * OP + (gid + 8) * OP) = make_float2(nn[2], nn[3]);
                        *reinterpret_cast<float2*>(ot + 4 * GS * OP + gid * OP) = make_float2(hh[0], hh[1]);
                        *reinterpret_cast<float2*>(ot + 4 * GS * OP + (gid + 8) * OP) = make_float2(hh[2], hh[3]);
                    }
                }
                if (stamp) dbg[t * 4 + 3] = clock64();
                __syncthreads();                       // h_{t+1}, the output tiles of step t and the gx tile of step t + 1 are complete
            }
            __syncthreads();                           // (the I/O warps drain the last output tile; the tiles are free for the next group)
        }
    } else {
        // ======================================================= I/O warps
        const int tid = threadIdx.x - 128;
        // this thread copies 16-byte chunk tid + 128 j (j < 3) of the gx tile (24 chunks per row) and chunk tid of the residual / output
        // tiles (8 chunks per row)
        int crow[3], ccol[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { const int c = tid + 128 * j; crow[j] = c / 24; ccol[j] = (c - crow[j] * 24) * 4; }
        const int srow = tid >> 3, scol = (tid & 7) * 4;
        for (int grp = blockIdx.x; grp * GS < g.S; grp += gridDim.x) {
            const long start = dir ? (long)(g.T - 1) * g.t_stride : 0;
            const long step = dir ? -g.t_stride : g.t_stride;
            auto row_m0 = [&](int row) { const int sq = grp * GS + row; return seq_base(g, sq < g.S ? sq : g.S - 1) + start; };
            const float* lp[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) lp[j] = gx + (size_t)row_m0(crow[j]) * 6 * H + dir * 3 * H + ccol[j];
            const long fstep = step * 6 * H;
            const long ms = row_m0(srow);
            const bool slive = grp * GS + srow < g.S;
            const float* rp = RES ? residual + (size_t)ms * ldr + dir * H + scol : nullptr;
            const long rstep = step * ldr;
            float* yp = y + (size_t)ms * ldy + dir * H + scol;
            const long ystep = step * ldy;
            float* qp = RES ? ysum + (size_t)ms * ldys + dir * H + scol : nullptr;
            const long qstep = step * ldys;
            float* tp = STASH ? stash + ((size_t)ms * 2 + dir) * 4 * H + scol : nullptr;
            const long tstep = step * 8 * H;
            auto fetch = [&](int slot) {
                float* sl = ring + slot * SLOT_F;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sl + crow[j] * GXP + ccol[j])), "l"(lp[j]) : "memory");
                    lp[j] += fstep;
                }
                if (RES) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sl + GS * GXP + srow * OP + scol)), "l"(rp) : "memory");
                    rp += rstep;
                }
            };
            // the previous step's results leave through the output tiles: thread -> 16 bytes of row srow of every tile
            auto store_prev = [&](int par, int slot_prev) {
                const float* ot = otile + par * NQ * GS * OP + srow * OP + scol;
                const float4 hv = *reinterpret_cast<const float4*>(ot);
                float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (RES) rv = *reinterpret_cast<const float4*>(ring + slot_prev * SLOT_F + GS * GXP + srow * OP + scol);
                float4 sv[4];
                if (STASH) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) sv[q] = *reinterpret_cast<const float4*>(ot + (q + 1) * GS * OP);
                }
                if (slive) {
                    *reinterpret_cast<float4*>(yp) = hv;
                    if (RES) *reinterpret_cast<float4*>(qp) = make_float4(hv.x + rv.x, hv.y + rv.y, hv.z + rv.z, hv.w + rv.w);
                    if (STASH) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(tp + q * H) = sv[q];
                    }
                }
                yp += ystep;
                if (RES) qp += qstep;
                if (STASH) tp += tstep;
            };
            // groups 0 .. PFM-2 in flight; step t refills slot (t + PFM - 1) % PFM (consumed during step t - 1) and waits for group t + 1
#pragma unroll
            for (int u = 0; u < PFM - 1; ++u) {
                if (u < g.T) fetch(u);
                cp_async_commit();
            }
            cp_async_wait<PFM - 2>();                  // group 0 (own copies), then the barrier makes everybody's visible
            __syncthreads();
            for (int t = 0; t < g.T; ++t) {
                if (t > 0) store_prev((t & 1) ^ 1, (t - 1) % PFM);
                if (t + PFM - 1 < g.T) fetch((t + PFM - 1) % PFM);
                cp_async_commit();
                cp_async_wait<PFM - 2>();              // own copies of step t + 1 have landed ...
                __syncthreads();                       // ... and are visible to the compute warps
            }
            store_prev((g.T - 1) & 1, (g.T - 1) % PFM);
            cp_async_wait<0>();
            __syncthreads();
        }
    }
}

// workspace layout shared by the two backward kernels and the reduction: part[block][dir][PART]
//   [0, 3072)      dW_hh partial  [g*32 + j][i]          (gru_dwhh_kernel)
//   [3072, 3168)   db_hh partial  [g*32 + j]             (gru_bwd_kernel)
//   [3168, 3264)   db_ih partial  [g*32 + j]             (gru_bwd_kernel)
constexpr int PART = 3 * H * H + 6 * H;

// BPTT, serial part.  dy: grad wrt GRU output [M, lddy]; y: forward output (h values); stash from forward.
// Writes dgx [M,192] = (dr_pre, dz_pre, dn_pre) and per-block bias-gradient partials.
template <int NS>
__global__ void __launch_bounds__(WARPS * 32, 3)
gru_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
               const float* __restrict__ stash, const float* __restrict__ whh, float* __restrict__ dgx,
               float* __restrict__ part, SeqGeom g) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    __shared__ __align__(16) float dg[WARPS][NS][3][H];      // broadcast of (dr_pre, dz_pre, d_hn) per sequence
    __shared__ float red[WARPS][6][H];

    // column `lane` of W_hr / W_hz / W_hn as (j, j+1) pairs:  dh_prev[lane] += sum_j W[g*32+j][lane] * dgate_g[j]
    float2 wr[H / 2], wz[H / 2], wn[H / 2];
    const float* W = whh + (size_t)dir * 3 * H * H;
#pragma unroll
    for (int j = 0; j < H / 2; ++j) {
        wr[j] = make_float2(__ldg(W + (0 * H + 2 * j) * H + lane), __ldg(W + (0 * H + 2 * j + 1) * H + lane));
        wz[j] = make_float2(__ldg(W + (1 * H + 2 * j) * H + lane), __ldg(W + (1 * H + 2 * j + 1) * H + lane));
        wn[j] = make_float2(__ldg(W + (2 * H + 2 * j) * H + lane), __ldg(W + (2 * H + 2 * j + 1) * H + lane));
    }
    float dbh[3] = {0.f, 0.f, 0.f}, dbi[3] = {0.f, 0.f, 0.f};
    __shared__ float ring[WARPS][PFB][NS][6][H];       // [step slot][sequence][r, z, n, hn, dy, h_prev][lane]

    for (int s0 = (blockIdx.x * WARPS + warp) * NS; s0 < g.S; s0 += gridDim.x * WARPS * NS) {
        // BPTT walks the processing order backwards: dir0 processed t=0..T-1 -> walk T-1..0; dir1 processed T-1..0 -> walk 0..T-1
        long m0[NS];
        bool live[NS];
#pragma unroll
        for (int v = 0; v < NS; ++v) {
            live[v] = s0 + v < g.S;
            m0[v] = seq_base(g, live[v] ? s0 + v : s0) + (dir ? 0 : (long)(g.T - 1) * g.t_stride);
        }
        const long step = dir ? g.t_stride : -g.t_stride;      // towards the previously processed element
#pragma unroll
        for (int u = 0; u < PFB; ++u) {
            if (u < g.T) {
#pragma unroll
                for (int v = 0; v < NS; ++v) {
                    const long m = m0[v] + u * step;
                    const float* sp = stash + ((size_t)m * 2 + dir) * 4 * H + lane;
                    cp_async4(&ring[warp][u][v][0][lane], sp);
                    cp_async4(&ring[warp][u][v][1][lane], sp + H);
                    cp_async4(&ring[warp][u][v][2][lane], sp + 2 * H);
                    cp_async4(&ring[warp][u][v][3][lane], sp + 3 * H);
                    cp_async4(&ring[warp][u][v][4][lane], dy + (size_t)m * lddy + dir * H + lane);
                    if (u + 1 < g.T) cp_async4(&ring[warp][u][v][5][lane], y + (size_t)(m + step) * ldy + dir * H + lane);
                }
            }
            cp_async_commit();
        }
        float dh[NS];
#pragma unroll
        for (int v = 0; v < NS; ++v) dh[v] = 0.f;
        for (int t = 0; t < g.T; ++t) {
            const int slot = t % PFB;
            cp_async_wait<PFB - 1>();
            float r[NS], z[NS], n[NS], hn[NS], hprev[NS];
#pragma unroll
            for (int v = 0; v < NS; ++v) {
                r[v] = ring[warp][slot][v][0][lane]; z[v] = ring[warp][slot][v][1][lane]; n[v] = ring[warp][slot][v][2][lane];
                hn[v] = ring[warp][slot][v][3][lane];
                dh[v] += ring[warp][slot][v][4][lane];
                hprev[v] = (t + 1 < g.T) ? ring[warp][slot][v][5][lane] : 0.f;
            }
            if (t + PFB < g.T) {
#pragma unroll
                for (int v = 0; v < NS; ++v) {
                    const long mn = m0[v] + (long)(t + PFB) * step;
                    const float* sp = stash + ((size_t)mn * 2 + dir) * 4 * H + lane;
                    cp_async4(&ring[warp][slot][v][0][lane], sp);
                    cp_async4(&ring[warp][slot][v][1][lane], sp + H);
                    cp_async4(&ring[warp][slot][v][2][lane], sp + 2 * H);
                    cp_async4(&ring[warp][slot][v][3][lane], sp + 3 * H);
                    cp_async4(&ring[warp][slot][v][4][lane], dy + (size_t)mn * lddy + dir * H + lane);
                    if (t + PFB + 1 < g.T) cp_async4(&ring[warp][slot][v][5][lane], y + (size_t)(mn + step) * ldy + dir * H + lane);
                }
            }
            cp_async_commit();
            float dhz[NS];
            __syncwarp();
#pragma unroll
            for (int v = 0; v < NS; ++v) {
                const float dn = dh[v] * (1.f - z[v]);
                const float dz = dh[v] * (hprev[v] - n[v]);
                const float dn_pre = dn * (1.f - n[v] * n[v]);
                const float dr_pre = dn_pre * hn[v] * r[v] * (1.f - r[v]);
                const float dz_pre = dz * z[v] * (1.f - z[v]);
                const float dhn = dn_pre * r[v];
                dhz[v] = dh[v] * z[v];
                if (live[v]) {
                    const long m = m0[v] + (long)t * step;
                    float* gp = dgx + (size_t)m * 6 * H + dir * 3 * H + lane;
                    gp[0] = dr_pre; gp[H] = dz_pre; gp[2 * H] = dn_pre;
                    dbi[0] += dr_pre; dbi[1] += dz_pre; dbi[2] += dn_pre;
                    dbh[0] += dr_pre; dbh[1] += dz_pre; dbh[2] += dhn;
                }
                dg[warp][v][0][lane] = dr_pre; dg[warp][v][1][lane] = dz_pre; dg[warp][v][2][lane] = dhn;
            }
            __syncwarp();
#pragma unroll
            for (int v = 0; v < NS; ++v) {
                float2 c0 = make_float2(dhz[v], 0.f), c1 = make_float2(0.f, 0.f), c2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < H / 4; ++q) {
                    const float4 vr = *reinterpret_cast<const float4*>(&dg[warp][v][0][q * 4]);
                    const float4 vz = *reinterpret_cast<const float4*>(&dg[warp][v][1][q * 4]);
                    const float4 vn = *reinterpret_cast<const float4*>(&dg[warp][v][2][q * 4]);
                    c0 = ffma2(wr[2 * q], make_float2(vr.x, vr.y), c0); c0 = ffma2(wr[2 * q + 1], make_float2(vr.z, vr.w), c0);
                    c1 = ffma2(wz[2 * q], make_float2(vz.x, vz.y), c1); c1 = ffma2(wz[2 * q + 1], make_float2(vz.z, vz.w), c1);
                    c2 = ffma2(wn[2 * q], make_float2(vn.x, vn.y), c2); c2 = ffma2(wn[2 * q + 1], make_float2(vn.z, vn.w), c2);
                }
                dh[v] = (c0.x + c0.y) + (c1.x + c1.y) + (c2.x + c2.y);
            }
        }
        cp_async_wait<0>();
    }
    // deterministic block reduction of the bias partials (fixed warp order)
#pragma unroll
    for (int a = 0; a < 3; ++a) { red[warp][a][lane] = dbh[a]; red[warp][3 + a][lane] = dbi[a]; }
    __syncthreads();
    float* out = part + ((size_t)blockIdx.x * 2 + dir) * PART + 3 * H * H;
    for (int i = threadIdx.x; i < 6 * H; i += blockDim.x) {
        const int a = i / H, l = i % H;
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) sum += red[w][a][l];
        out[i] = sum;
    }
}

// ---- tensor-core BPTT (round 2): the forward kernel's structure with the transposed recurrence
//   dh_{t-1}[16 seq][32] = dh_t z_t + G_t[16 seq][96] . W_hh[96][32],   G = (dr_pre, dz_pre, dn_pre r)
// A = G as bf16 hi / lo tiles in shared memory (6 k-steps), B = W_hh (warp w owns the 8 output units 8w ..), 18 MMAs per step in six
// chains of three; the element-wise gate derivatives need no MUFU.  I/O warps stage the stash / dy / h_prev tiles PFM - 1 steps ahead
// and move the previous step's dgx tile out.  Bias-gradient partials: per-lane sums, reduced over the 8 sequence rows of a fragment by
// shuffles at the end, one slot per CTA in the workspace (same layout as the FFMA kernel, fixed order: deterministic).
constexpr int SP = 136;         // floats per row of a staged stash tile (128 + 8)
constexpr int GP = 104;         // elements per row of the G tiles (bf16) and of the dgx output tile (fp32): 96 + 8
constexpr int BSLOT_F = GS * SP + 2 * GS * OP;      // ring slot: stash tile + dy tile + h_prev tile
constexpr size_t gru_bwd_mma_smem() { return (size_t)PFM * BSLOT_F * 4 + (size_t)2 * GS * GP * 4 + (size_t)2 * 2 * GS * GP * 2; }

__global__ void __launch_bounds__(256)
gru_bwd_mma_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy, const float* __restrict__ stash,
                   const float* __restrict__ whh, float* __restrict__ dgx, float* __restrict__ part, SeqGeom g) {
    extern __shared__ __align__(16) uint8_t gru_smem[];
    float* ring = reinterpret_cast<float*>(gru_smem);                                   // [PFM][stash 16 x SP | dy 16 x OP | h_prev 16 x OP]
    float* otile = ring + PFM * BSLOT_F;                                                // [2][16][GP] dgx of the step
    __nv_bfloat16* gbuf = reinterpret_cast<__nv_bfloat16*>(otile + 2 * GS * GP);        // [2][hi, lo][16][GP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;

    if (warp < 4) {
        // ======================================================= compute warps
        const int gid = lane >> 2, tig = lane & 3;
        const int uu = warp * 8 + 2 * tig;
        // B fragments of W_hh: k = gate row 16 ks + (2 tig, +1 | +8), n = unit 8 warp + gid
        uint32_t w_hi[6][2], w_lo[6][2];
        const float* W = whh + (size_t)dir * 3 * H * H;
#pragma unroll
        for (int ks = 0; ks < 6; ++ks)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = 16 * ks + 8 * j + 2 * tig;
                split_pair(__ldg(W + (size_t)k * H + warp * 8 + gid), __ldg(W + (size_t)(k + 1) * H + warp * 8 + gid), w_hi[ks][j], w_lo[ks][j]);
            }
        const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lk = (lane >> 4) * 8;
        float dbi[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}}, dbh[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
        for (int grp = blockIdx.x; grp * GS < g.S; grp += gridDim.x) {
            const float liveA = grp * GS + gid < g.S ? 1.f : 0.f, liveB = grp * GS + gid + 8 < g.S ? 1.f : 0.f;
            for (int i = threadIdx.x; i < 2 * GS * GP / 2; i += 128) reinterpret_cast<uint32_t*>(gbuf)[i] = 0u;      // G_{-1} = 0 (both planes of buffer 0)
            float dhz[4] = {0.f, 0.f, 0.f, 0.f};
            __syncthreads();
            for (int t = 0; t < g.T; ++t) {
                const int slot = t % PFM, cur = t & 1;
                const __nv_bfloat16* gb = gbuf + cur * 2 * GS * GP;
                float acc[6][4];
#pragma unroll
                for (int ks = 0; ks < 6; ++ks) {
                    uint32_t a_hi[4], a_lo[4];
                    ldmatrix_x4(a_hi, gb + lrow * GP + 16 * ks + lk);
                    ldmatrix_x4(a_lo, gb + GS * GP + lrow * GP + 16 * ks + lk);
                    acc[ks][0] = acc[ks][1] = acc[ks][2] = acc[ks][3] = 0.f;
                    mma_bf16_16816(acc[ks], a_hi, w_hi[ks][0], w_hi[ks][1]);
                    mma_bf16_16816(acc[ks], a_lo, w_hi[ks][0], w_hi[ks][1]);
                    mma_bf16_16816(acc[ks], a_hi, w_lo[ks][0], w_lo[ks][1]);
                }
                // this step's operands (in the shadow of the MMAs)
                const float* sl = ring + slot * BSLOT_F;
                float2 r2[2], z2[2], n2[2], hn2[2], dy2[2], hp2[2];
#pragma unroll
                for (int sq = 0; sq < 2; ++sq) {
                    const float* sr = sl + (gid + 8 * sq) * SP + uu;
                    r2[sq] = *reinterpret_cast<const float2*>(sr); z2[sq] = *reinterpret_cast<const float2*>(sr + H);
                    n2[sq] = *reinterpret_cast<const float2*>(sr + 2 * H); hn2[sq] = *reinterpret_cast<const float2*>(sr + 3 * H);
                    dy2[sq] = *reinterpret_cast<const float2*>(sl + GS * SP + (gid + 8 * sq) * OP + uu);
                    hp2[sq] = *reinterpret_cast<const float2*>(sl + GS * SP + GS * OP + (gid + 8 * sq) * OP + uu);
                }
                float drp[4], dzp[4], dnp[4], dhn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int sq = e >> 1;
                    const float r = (e & 1) ? r2[sq].y : r2[sq].x, z = (e & 1) ? z2[sq].y : z2[sq].x, n = (e & 1) ? n2[sq].y : n2[sq].x;
                    const float hn = (e & 1) ? hn2[sq].y : hn2[sq].x, dyv = (e & 1) ? dy2[sq].y : dy2[sq].x, hp = (e & 1) ? hp2[sq].y : hp2[sq].x;
                    const float dh = dhz[e] + ((acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]) + (acc[4][e] + acc[5][e])) + dyv;
                    const float dn = dh * (1.f - z);
                    const float dz = dh * (hp - n);
                    dnp[e] = dn * (1.f - n * n);
                    drp[e] = dnp[e] * hn * r * (1.f - r);
                    dzp[e] = dz * z * (1.f - z);
                    dhn[e] = dnp[e] * r;
                    dhz[e] = dh * z;
                    const float lv = sq ? liveB : liveA;
                    dbi[0][e & 1] += lv * drp[e]; dbi[1][e & 1] += lv * dzp[e]; dbi[2][e & 1] += lv * dnp[e];
                    dbh[2][e & 1] += lv * dhn[e];
                }
                {
                    __nv_bfloat16* gn = gbuf + (cur ^ 1) * 2 * GS * GP;
                    float* ot = otile + cur * GS * GP;
#pragma unroll
                    for (int sq = 0; sq < 2; ++sq) {
                        const int row = gid + 8 * sq;
                        uint32_t hi, lo;
                        split_pair(drp[2 * sq], drp[2 * sq + 1], hi, lo);
                        *reinterpret_cast<uint32_t*>(gn + row * GP + uu) = hi; *reinterpret_cast<uint32_t*>(gn + GS * GP + row * GP + uu) = lo;
                        split_pair(dzp[2 * sq], dzp[2 * sq + 1], hi, lo);
                        *reinterpret_cast<uint32_t*>(gn + row * GP + H + uu) = hi; *reinterpret_cast<uint32_t*>(gn + GS * GP + row * GP + H + uu) = lo;
                        split_pair(dhn[2 * sq], dhn[2 * sq + 1], hi, lo);
                        *reinterpret_cast<uint32_t*>(gn + row * GP + 2 * H + uu) = hi; *reinterpret_cast<uint32_t*>(gn + GS * GP + row * GP + 2 * H + uu) = lo;
                        *reinterpret_cast<float2*>(ot + row * GP + uu) = make_float2(drp[2 * sq], drp[2 * sq + 1]);
                        *reinterpret_cast<float2*>(ot + row * GP + H + uu) = make_float2(dzp[2 * sq], dzp[2 * sq + 1]);
                        *reinterpret_cast<float2*>(ot + row * GP + 2 * H + uu) = make_float2(dnp[2 * sq], dnp[2 * sq + 1]);
                    }
                }
                __syncthreads();
            }
            __syncthreads();
        }
        // bias partials: db_hh = (dr_pre, dz_pre, dn_pre r) sums, db_ih = (dr_pre, dz_pre, dn_pre) sums; the first two gates coincide
#pragma unroll
        for (int j = 0; j < 2; ++j) { dbh[0][j] = dbi[0][j]; dbh[1][j] = dbi[1][j]; }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    dbi[a][j] += __shfl_xor_sync(0xffffffffu, dbi[a][j], o);
                    dbh[a][j] += __shfl_xor_sync(0xffffffffu, dbh[a][j], o);
                }
            }
        if (gid == 0) {
            float* out = part + ((size_t)blockIdx.x * 2 + dir) * PART + 3 * H * H;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < 2; ++j) { out[a * H + uu + j] = dbh[a][j]; out[3 * H + a * H + uu + j] = dbi[a][j]; }
        }
    } else {
        // ======================================================= I/O warps
        const int tid = threadIdx.x - 128;
        const int srow = tid >> 3, scol = (tid & 7) * 4;          // 8 chunks per row: dy, h_prev tiles; stash tile: 32 chunks per row
        int trow[4], tcol[4], orow[3], ocol[3];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int c = tid + 128 * j; trow[j] = c >> 5; tcol[j] = (c & 31) * 4; }
#pragma unroll
        for (int j = 0; j < 3; ++j) { const int c = tid + 128 * j; orow[j] = c / 24; ocol[j] = (c - orow[j] * 24) * 4; }
        for (int grp = blockIdx.x; grp * GS < g.S; grp += gridDim.x) {
            // BPTT walks the processing order backwards: dir0 processed t = 0..T-1 -> walk T-1..0; dir1 the other way
            const long start = dir ? 0 : (long)(g.T - 1) * g.t_stride;
            const long step = dir ? g.t_stride : -g.t_stride;      // towards the previously processed element
            auto row_m0 = [&](int row) { const int sq = grp * GS + row; return seq_base(g, sq < g.S ? sq : g.S - 1) + start; };
            const float* tp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) tp[j] = stash + ((size_t)row_m0(trow[j]) * 2 + dir) * 4 * H + tcol[j];
            const long tstep = step * 8 * H;
            const long ms = row_m0(srow);
            const float* dp = dy + (size_t)ms * lddy + dir * H + scol;
            const long dstep = step * lddy;
            const float* hp = y + (size_t)(ms + step) * ldy + dir * H + scol;      // output of the previously processed element
            const long hstep = step * ldy;
            float* op[3];
            bool olive[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                op[j] = dgx + (size_t)row_m0(orow[j]) * 6 * H + dir * 3 * H + ocol[j];
                olive[j] = grp * GS + orow[j] < g.S;
            }
            const long ostep = step * 6 * H;
            int fetched = 0;
            auto fetch = [&](int slot) {
                float* sl = ring + slot * BSLOT_F;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sl + trow[j] * SP + tcol[j])), "l"(tp[j]) : "memory");
                    tp[j] += tstep;
                }
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sl + GS * SP + srow * OP + scol)), "l"(dp) : "memory");
                dp += dstep;
                float* hdst = sl + GS * SP + GS * OP + srow * OP + scol;
                if (fetched + 1 < g.T) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(hdst)), "l"(hp) : "memory");
                else *reinterpret_cast<float4*>(hdst) = make_float4(0.f, 0.f, 0.f, 0.f);      // the first processed element has h_prev = 0
                hp += hstep;
                ++fetched;
            };
            auto store_prev = [&](int par) {
                const float* ot = otile + par * GS * GP;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(ot + orow[j] * GP + ocol[j]);
                    if (olive[j]) *reinterpret_cast<float4*>(op[j]) = v;
                    op[j] += ostep;
                }
            };
#pragma unroll
            for (int u = 0; u < PFM - 1; ++u) {
                if (u < g.T) fetch(u);
                cp_async_commit();
            }
            cp_async_wait<PFM - 2>();
            __syncthreads();
            for (int t = 0; t < g.T; ++t) {
                if (t > 0) store_prev((t & 1) ^ 1);
                if (t + PFM - 1 < g.T) fetch((t + PFM - 1) % PFM);
                cp_async_commit();
                cp_async_wait<PFM - 2>();
                __syncthreads();
            }
            store_prev((g.T - 1) & 1);
            cp_async_wait<0>();
            __syncthreads();
        }
    }
}

// dW_hh[dir][g*32+j][i] = sum over (sequence, step) of G_g[j] * h_prev[i],  G = (dr_pre, dz_pre, dn_pre * r), h_prev = the
// output of the previously processed step (0 at the first).  256 threads: thread (i = tid % 32, jg = tid / 32) owns the 12 rows
// jg*12 .. jg*12+11 of column i; 32 (sequence, step) rows are staged per iteration.
constexpr int DW_ROWS = 32;
__global__ void __launch_bounds__(256)
gru_dwhh_kernel(const float* __restrict__ dgx, const float* __restrict__ stash, const float* __restrict__ y, int ldy,
                float* __restrict__ part, SeqGeom g) {
    const int dir = blockIdx.y;
    const int tid = threadIdx.x, i = tid & 31, jg = tid >> 5;
    __shared__ __align__(16) float Gs[DW_ROWS][3 * H];
    __shared__ float Hs[DW_ROWS][H];
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    const int total = g.S * g.T;                      // (sequence, step) rows; < 2^31
    const int chunks = (total + DW_ROWS - 1) / DW_ROWS;
    // Software pipeline: the 12 (+12 r factors) + 4 global loads of chunk i+1 are issued into registers (independent, all in flight)
    // before the outer products of chunk i run from shared memory; the row -> pixel-row maps (the only divisions) are computed once per
    // row by 32 threads, one chunk ahead.  (v1 loaded inside the staging loop: one DRAM latency per element, 123 us; v2 did two 64-bit
    // divisions per element: 137 us.)
    __shared__ int rowM[2][DW_ROWS], rowP[2][DW_ROWS];
    float gv[12], rv[12], hv4[4];
    auto map_rows = [&](int ch, int buf) {
        if (tid < DW_ROWS) {
            const int rho = ch * DW_ROWS + tid;
            int m = -1, mp = -1;
            if (rho < total) {
                const int sq = rho / g.T, t = rho - sq * g.T;
                const int base = (int)seq_base(g, sq);
                m = base + t * (int)g.t_stride;
                const int tp = dir ? t + 1 : t - 1;           // previously processed position
                if (tp >= 0 && tp < g.T) mp = base + tp * (int)g.t_stride;
            }
            rowM[buf][tid] = m; rowP[buf][tid] = mp;
        }
    };
    auto load_chunk = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const int e = tid + u * 256;
            const int row = e / (3 * H), c = e - row * (3 * H);
            const int m = rowM[buf][row];
            const bool ok = m >= 0;
            const size_t mm = ok ? (size_t)m : 0;
            gv[u] = ok ? __ldg(dgx + mm * 6 * H + dir * 3 * H + c) : 0.f;
            rv[u] = (ok && c >= 2 * H) ? __ldg(stash + (mm * 2 + dir) * 4 * H + (c - 2 * H)) : 1.f;      // third gate: dn_pre * r
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * 256;
            const int row = e / H, c = e - row * H;
            const int mp = rowP[buf][row];
            hv4[u] = mp >= 0 ? __ldg(y + (size_t)mp * ldy + dir * H + c) : 0.f;
        }
    };
    int ch = blockIdx.x, it = 0;
    if (ch < chunks) map_rows(ch, 0);
    __syncthreads();
    if (ch < chunks) load_chunk(0);
    for (; ch < chunks; ch += gridDim.x, ++it) {
        const int cur = it & 1;
        const bool more = ch + (int)gridDim.x < chunks;
#pragma unroll
        for (int u = 0; u < 12; ++u) { const int e = tid + u * 256; Gs[e / (3 * H)][e % (3 * H)] = gv[u] * rv[u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int e = tid + u * 256; Hs[e / H][e % H] = hv4[u]; }
        if (more) map_rows(ch + gridDim.x, cur ^ 1);
        __syncthreads();
        if (more) load_chunk(cur ^ 1);
#pragma unroll 4
        for (int row = 0; row < DW_ROWS; ++row) {
            const float hv = Hs[row][i];
            const float4 g0 = *reinterpret_cast<const float4*>(&Gs[row][jg * 12]);
            const float4 g1 = *reinterpret_cast<const float4*>(&Gs[row][jg * 12 + 4]);
            const float4 g2 = *reinterpret_cast<const float4*>(&Gs[row][jg * 12 + 8]);
            acc[0] = fmaf(g0.x, hv, acc[0]); acc[1] = fmaf(g0.y, hv, acc[1]); acc[2] = fmaf(g0.z, hv, acc[2]); acc[3] = fmaf(g0.w, hv, acc[3]);
            acc[4] = fmaf(g1.x, hv, acc[4]); acc[5] = fmaf(g1.y, hv, acc[5]); acc[6] = fmaf(g1.z, hv, acc[6]); acc[7] = fmaf(g1.w, hv, acc[7]);
            acc[8] = fmaf(g2.x, hv, acc[8]); acc[9] = fmaf(g2.y, hv, acc[9]); acc[10] = fmaf(g2.z, hv, acc[10]); acc[11] = fmaf(g2.w, hv, acc[11]);
        }
        __syncthreads();
    }
    float* out = part + ((size_t)blockIdx.x * 2 + dir) * PART;
#pragma unroll
    for (int k = 0; k < 12; ++k) out[(jg * 12 + k) * H + i] = acc[k];
}

// sums the per-block partials in a fixed order.  blockDim (32, 8): 32 consecutive elements x 8 lanes over the blocks.
__global__ void gru_bwd_reduce_kernel(const float* __restrict__ part, int nb_w, int nb_b, float* __restrict__ dwhh,
                                      float* __restrict__ dbhh, float* __restrict__ dbih, int accumulate) {
    const int idx = blockIdx.x * 32 + threadIdx.x;   // over 2 * PART
    __shared__ float sh[8][33];
    float s = 0.f;
    int dir = 0, e = 0;
    if (idx < 2 * PART) {
        dir = idx / PART; e = idx % PART;
        const int nb = e < 3 * H * H ? nb_w : nb_b;
        float s0 = 0.f, s1 = 0.f;
        int b = threadIdx.y;
        for (; b + 8 < nb; b += 16) {
            s0 += part[((size_t)b * 2 + dir) * PART + e];
            s1 += part[((size_t)(b + 8) * 2 + dir) * PART + e];
        }
        if (b < nb) s0 += part[((size_t)b * 2 + dir) * PART + e];
        s = s0 + s1;
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y != 0 || idx >= 2 * PART) return;
    s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sh[k][threadIdx.x];
    float* dst;
    if (e < 3 * H * H) dst = dwhh + (size_t)dir * 3 * H * H + e;
    else if (e < 3 * H * H + 3 * H) dst = dbhh + dir * 3 * H + (e - 3 * H * H);
    else dst = dbih + dir * 3 * H + (e - 3 * H * H - 3 * H);
    *dst = accumulate ? *dst + s : s;
}

// one resident wave: 3 CTAs of 4 warps per SM, two directions
int gru_ns(int S) { return 2 * S >= 2 * 148 * 3 * WARPS * 2 ? 2 : 1; }      // two directions; 2 per warp only if that still fills 3 CTAs per SM
int gru_blocks(int S) {
    const int NS = gru_ns(S);
    int b = (S + WARPS * NS - 1) / (WARPS * NS);
    int cap = 148 * 3 / 2;
    return b < cap ? b : cap;
}
long long* g_gru_dbg = nullptr;
int g_gru_mma = 1;      // 1: pick per shape; 0: always the warp-per-sequence FFMA kernels; 2: always the tensor-core kernels (tpgsr_gru_set_mode; cross-check tests, A/B measurements)
int dw_blocks(int S, int T) {
    long chunks = ((long)S * T + DW_ROWS - 1) / DW_ROWS;
    return (int)(chunks < 148 ? chunks : 148);
}


// BPTT recurrence on nb CTAs per direction (the bias partials of every one of the nb workspace slots are written: the reduction sums nb)
int launch_gru_bwd_recurrence(const float* dy, int lddy, const float* y, int ldy, const float* stash, const float* whh, float* dgx,
                              float* part, const SeqGeom& g, int nb, cudaStream_t stream) {
    const bool al16 = !(lddy & 3) && !(ldy & 3) && !(((uintptr_t)dy | (uintptr_t)y | (uintptr_t)stash | (uintptr_t)dgx) & 15);
    // measured (profiles/r02/time_gru_v5.log): without MUFU work in the step the FFMA kernel is already short; the tensor-core kernel wins
    // only for many short sequences (column GRU at B = 48: 29.9 vs 36.6 us; row GRU 54 vs 38 us, B = 5 column 11.2 vs 9.0 us)
    if (g_gru_mma && al16 && (g_gru_mma == 2 || (g.T <= 32 && g.S >= 1024))) {
        const size_t smem = gru_bwd_mma_smem();
        if (cudaFuncSetAttribute(gru_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TPGSR_ERR_LAUNCH;
        gru_bwd_mma_kernel<<<dim3(nb, 2), 256, smem, stream>>>(dy, lddy, y, ldy, stash, whh, dgx, part, g);
    } else if (gru_ns(g.S) == 2) {
        gru_bwd_kernel<2><<<dim3(nb, 2), WARPS * 32, 0, stream>>>(dy, lddy, y, ldy, stash, whh, dgx, part, g);
    } else {
        gru_bwd_kernel<1><<<dim3(nb, 2), WARPS * 32, 0, stream>>>(dy, lddy, y, ldy, stash, whh, dgx, part, g);
    }
    return tpgsr_launch_status();
}

}  // namespace

extern "C" {

// 1 (default): tensor-core or FFMA recurrence kernels by shape; 0: the FFMA kernels of round 1; 2: the tensor-core kernels
int tpgsr_gru_set_mode(int mma) { g_gru_mma = mma < 0 ? 0 : (mma > 2 ? 2 : mma); return TPGSR_OK; }
// developer instrumentation: clock64 stamps of the first CTA's compute warp 0, 4 per step (step start, MMAs done, gates done, stores issued)
int tpgsr_gru_set_debug(long long* dbg) { g_gru_dbg = dbg; return TPGSR_OK; }

size_t tpgsr_gru_bwd_workspace_bytes(int S) {
    // the second kernel's block count depends on T too; 148 is its cap
    int nb = gru_blocks(S) > 148 ? gru_blocks(S) : 148;
    return (size_t)nb * 2 * PART * sizeof(float);
}

// whh: [2][96][32] (weight_hh_l0, weight_hh_l0_reverse), bhh: [2][96].
int tpgsr_gru_fwd(const float* gx, const float* whh, const float* bhh, float* y, int ldy, const float* residual, int ldr,
                  float* ysum, int ldys, float* stash, int S, int T, int inner, long outer_stride, long inner_stride,
                  long t_stride, cudaStream_t stream) {
    if (!gx || !whh || !bhh || !y || S <= 0 || T <= 0 || inner <= 0) return TPGSR_ERR_BAD_ARG;
    if (ysum && !residual) return TPGSR_ERR_BAD_ARG;
    SeqGeom g{S, T, inner, outer_stride, inner_stride, t_stride};
    // the tensor-core kernel moves 16-byte chunks: pitches and base pointers must keep them aligned
    const bool al16 = !(ldy & 3) && !(ldr & 3) && !(ldys & 3) && !(((uintptr_t)gx | (uintptr_t)y | (uintptr_t)residual | (uintptr_t)ysum | (uintptr_t)stash) & 15);
    // measured (profiles/r02/time_gru_v4.log, B200): the tensor-core kernel wins whenever there are enough 16-sequence groups to occupy the
    // chip or the sequences are short — column GRU (T = 16) 24 vs 45 us at B = 48, 11 vs 15 us at B = 5; row GRU (T = 64) 40 vs 46 us at
    // B = 48 (S = 768) — and loses for a few long sequences (S = 80, T = 64: 38 vs 25 us), where a warp per sequence spreads wider
    if (g_gru_mma && al16 && (g_gru_mma == 2 || S >= 512 || T <= 32)) {
        int groups = (S + GS - 1) / GS;
        const dim3 grid(groups < 148 * 4 ? groups : 148 * 4, 2);
        const size_t smem = gru_mma_smem(stash != nullptr);
        auto kern = ysum ? (stash ? gru_fwd_mma_kernel<true, true> : gru_fwd_mma_kernel<true, false>)
                         : (stash ? gru_fwd_mma_kernel<false, true> : gru_fwd_mma_kernel<false, false>);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TPGSR_ERR_LAUNCH;
        kern<<<grid, 256, smem, stream>>>(gx, whh, bhh, y, ldy, residual, ldr, ysum, ldys, stash, g, g_gru_dbg);
        return tpgsr_launch_status();
    }
    dim3 grid(gru_blocks(S), 2);
    if (gru_ns(S) == 2) gru_fwd_kernel<2><<<grid, WARPS * 32, 0, stream>>>(gx, whh, bhh, y, ldy, residual, ldr, ysum, ldys, stash, g);
    else gru_fwd_kernel<1><<<grid, WARPS * 32, 0, stream>>>(gx, whh, bhh, y, ldy, residual, ldr, ysum, ldys, stash, g);
    return tpgsr_launch_status();
}

// dwhh [2][96][32], dbhh [2][96], dbih [2][96] are (+)= ; dgx [M,192] is overwritten.
int tpgsr_gru_bwd(const float* dy, int lddy, const float* y, int ldy, const float* stash, const float* whh, float* dgx,
                  float* dwhh, float* dbhh, float* dbih, int accumulate, int S, int T, int inner, long outer_stride,
                  long inner_stride, long t_stride, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!dy || !y || !stash || !whh || !dgx || !dwhh || !dbhh || !dbih || !workspace) return TPGSR_ERR_BAD_ARG;
    if (S <= 0 || T <= 0 || inner <= 0) return TPGSR_ERR_BAD_ARG;
    if (workspace_bytes < tpgsr_gru_bwd_workspace_bytes(S)) return TPGSR_ERR_WORKSPACE;
    SeqGeom g{S, T, inner, outer_stride, inner_stride, t_stride};
    float* part = static_cast<float*>(workspace);
    const int nb = gru_blocks(S), nbw = dw_blocks(S, T);
    int rc = launch_gru_bwd_recurrence(dy, lddy, y, ldy, stash, whh, dgx, part, g, nb, stream);
    if (rc) return rc;
    gru_dwhh_kernel<<<dim3(nbw, 2), 256, 0, stream>>>(dgx, stash, y, ldy, part, g);
    rc = tpgsr_launch_status();
    if (rc) return rc;
    gru_bwd_reduce_kernel<<<ceil_div(2 * PART, 32), dim3(32, 8), 0, stream>>>(part, nbw, nb, dwhh, dbhh, dbih, accumulate);
    return tpgsr_launch_status();
}

// The same backward in two calls, so that a caller can queue the weight-gradient half on another stream:
//   tpgsr_gru_bwd_recurrence  — serial BPTT: dgx and the per-block bias partials (into `workspace`)
//   tpgsr_gru_bwd_weights     — dW_hh partials (into the same `workspace`) + the reduction into dwhh / dbhh / dbih
int tpgsr_gru_bwd_recurrence(const float* dy, int lddy, const float* y, int ldy, const float* stash, const float* whh, float* dgx, int S,
                             int T, int inner, long outer_stride, long inner_stride, long t_stride, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
    if (!dy || !y || !stash || !whh || !dgx || !workspace || S <= 0 || T <= 0 || inner <= 0) return TPGSR_ERR_BAD_ARG;
    if (workspace_bytes < tpgsr_gru_bwd_workspace_bytes(S)) return TPGSR_ERR_WORKSPACE;
    SeqGeom g{S, T, inner, outer_stride, inner_stride, t_stride};
    float* part = static_cast<float*>(workspace);
    return launch_gru_bwd_recurrence(dy, lddy, y, ldy, stash, whh, dgx, part, g, gru_blocks(S), stream);
}

int tpgsr_gru_bwd_weights(const float* dgx, const float* y, int ldy, const float* stash, float* dwhh, float* dbhh, float* dbih, int accumulate,
                          int S, int T, int inner, long outer_stride, long inner_stride, long t_stride, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream) {
    if (!dgx || !y || !stash || !dwhh || !dbhh || !dbih || !workspace || S <= 0 || T <= 0 || inner <= 0) return TPGSR_ERR_BAD_ARG;
    if (workspace_bytes < tpgsr_gru_bwd_workspace_bytes(S)) return TPGSR_ERR_WORKSPACE;
    SeqGeom g{S, T, inner, outer_stride, inner_stride, t_stride};
    float* part = static_cast<float*>(workspace);
    const int nb = gru_blocks(S), nbw = dw_blocks(S, T);
    gru_dwhh_kernel<<<dim3(nbw, 2), 256, 0, stream>>>(dgx, stash, y, ldy, part, g);
    int rc = tpgsr_launch_status();
    if (rc) return rc;
    gru_bwd_reduce_kernel<<<ceil_div(2 * PART, 32), dim3(32, 8), 0, stream>>>(part, nbw, nb, dwhh, dbhh, dbih, accumulate);
    return tpgsr_launch_status();
}

}  // extern "C"
