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
int dw_blocks(int S, int T) {
    long chunks = ((long)S * T + DW_ROWS - 1) / DW_ROWS;
    return (int)(chunks < 148 ? chunks : 148);
}

}  // namespace

extern "C" {

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
    if (gru_ns(S) == 2) gru_bwd_kernel<2><<<dim3(nb, 2), WARPS * 32, 0, stream>>>(dy, lddy, y, ldy, stash, whh, dgx, part, g);
    else gru_bwd_kernel<1><<<dim3(nb, 2), WARPS * 32, 0, stream>>>(dy, lddy, y, ldy, stash, whh, dgx, part, g);
    int rc = tpgsr_launch_status();
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
    const int nb = gru_blocks(S);
    if (gru_ns(S) == 2) gru_bwd_kernel<2><<<dim3(nb, 2), WARPS * 32, 0, stream>>>(dy, lddy, y, ldy, stash, whh, dgx, part, g);
    else gru_bwd_kernel<1><<<dim3(nb, 2), WARPS * 32, 0, stream>>>(dy, lddy, y, ldy, stash, whh, dgx, part, g);
    return tpgsr_launch_status();
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
