// Bidirectional GRU recurrence (hidden 32 per direction) over rows or columns of an NHWC feature map.
//
// Reference op replaced: nn.GRU(64, 32, bidirectional=True, batch_first=True) inside GruBlock
// (model/tsrn.py:491-508; cuDNN RNN in the reference).  The input projection W_ih x + b_ih is a plain GEMM done by the
// conv/GEMM kernels; this file is the serial part: one warp per (sequence, direction), lane = hidden unit, the 96x32
// recurrent matrix held in registers (forward) / shared memory (backward), h broadcast through shared memory.
// Gate order r,z,n;  n = tanh(gx_n + r * (W_hn h + b_hn));  h' = (1-z) n + z h   (SURVEY Appendix B).
//
// Sequences are addressed on the pixel-row index m of the NHWC map so that the column-wise GRU (gru1, which the reference
// runs on a transposed copy) needs no transpose:  m(s,t) = (s / inner) * outer_stride + (s % inner) * inner_stride + t * t_stride.
//
// Latency-bound (serial T); traffic per (pixel, direction): read 96 gate pre-activations + write 32 outputs + 128 stash
// floats.  Reported in DESIGN.md as achieved bytes/s and steps/us, not against the tensor roofline.
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

// gx: [M, 192] = [dir][gate r,z,n][32] pre-activations from the input side (bias_ih included)
// y:  [M, ldy] GRU output (dir0 -> cols 0..31, dir1 -> cols 32..63); ysum = y + residual (optional)
// stash: [M][2][4][32] = r, z, n, (W_hn h + b_hn)   (training only; may be NULL)
__global__ void __launch_bounds__(WARPS * 32)
gru_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const float* __restrict__ bhh,
               float* __restrict__ y, int ldy, const float* __restrict__ residual, int ldr, float* __restrict__ ysum,
               int ldys, float* __restrict__ stash, SeqGeom g) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    __shared__ __align__(16) float hbuf[WARPS][H];

    float wr[H], wz[H], wn[H];
    const float* W = whh + (size_t)dir * 3 * H * H;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        wr[i] = __ldg(W + (0 * H + lane) * H + i);
        wz[i] = __ldg(W + (1 * H + lane) * H + i);
        wn[i] = __ldg(W + (2 * H + lane) * H + i);
    }
    const float bhr = __ldg(bhh + dir * 3 * H + lane), bhz = __ldg(bhh + dir * 3 * H + H + lane),
                bhn = __ldg(bhh + dir * 3 * H + 2 * H + lane);

    for (int s = blockIdx.x * WARPS + warp; s < g.S; s += gridDim.x * WARPS) {
        const long base = seq_base(g, s);
        float h = 0.f;
        hbuf[warp][lane] = 0.f;
        __syncwarp();
        long m = base + (dir ? (long)(g.T - 1) * g.t_stride : 0);
        const long step = dir ? -g.t_stride : g.t_stride;
        const float* gp = gx + (size_t)m * 6 * H + dir * 3 * H + lane;
        float gxr = __ldg(gp), gxz = __ldg(gp + H), gxn = __ldg(gp + 2 * H);
        for (int t = 0; t < g.T; ++t) {
            float nr = 0.f, nz = 0.f, nn = 0.f;
            if (t + 1 < g.T) {   // prefetch next step's pre-activations
                const float* np = gx + (size_t)(m + step) * 6 * H + dir * 3 * H + lane;
                nr = __ldg(np); nz = __ldg(np + H); nn = __ldg(np + 2 * H);
            }
            float ar = bhr, az = bhz, an = bhn;
#pragma unroll
            for (int q = 0; q < H / 4; ++q) {
                float4 hv = *reinterpret_cast<const float4*>(&hbuf[warp][q * 4]);
                ar = fmaf(wr[q * 4 + 0], hv.x, ar); az = fmaf(wz[q * 4 + 0], hv.x, az); an = fmaf(wn[q * 4 + 0], hv.x, an);
                ar = fmaf(wr[q * 4 + 1], hv.y, ar); az = fmaf(wz[q * 4 + 1], hv.y, az); an = fmaf(wn[q * 4 + 1], hv.y, an);
                ar = fmaf(wr[q * 4 + 2], hv.z, ar); az = fmaf(wz[q * 4 + 2], hv.z, az); an = fmaf(wn[q * 4 + 2], hv.z, an);
                ar = fmaf(wr[q * 4 + 3], hv.w, ar); az = fmaf(wz[q * 4 + 3], hv.w, az); an = fmaf(wn[q * 4 + 3], hv.w, an);
            }
            float r = sigmoid_f(gxr + ar);
            float z = sigmoid_f(gxz + az);
            float n = tanhf(gxn + r * an);
            h = (1.f - z) * n + z * h;
            y[(size_t)m * ldy + dir * H + lane] = h;
            if (ysum) ysum[(size_t)m * ldys + dir * H + lane] = h + residual[(size_t)m * ldr + dir * H + lane];
            if (stash) {
                float* sp = stash + ((size_t)m * 2 + dir) * 4 * H + lane;
                sp[0] = r; sp[H] = z; sp[2 * H] = n; sp[3 * H] = an;
            }
            __syncwarp();
            hbuf[warp][lane] = h;
            __syncwarp();
            gxr = nr; gxz = nz; gxn = nn;
            m += step;
        }
    }
}

// BPTT.  dy: grad wrt GRU output [M, lddy]; y: forward output (h values); stash from forward.
// Writes dgx [M,192] and per-block partials of dW_hh [2][96][32], db_hh [2][96]; db_ih partials [2][96].
// partial layout: part[block][dir][ 3072 (dWhh) + 96 (dbhh) + 96 (dbih) ]
constexpr int PART = 3 * H * H + 6 * H;

__global__ void __launch_bounds__(WARPS * 32)
gru_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
               const float* __restrict__ stash, const float* __restrict__ whh, float* __restrict__ dgx,
               float* __restrict__ part, SeqGeom g) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    __shared__ __align__(16) float Wsm[3 * H * H];          // [g*32+j][i]
    __shared__ __align__(16) float dg[WARPS][3][H];          // broadcast of dgh = (dr_pre, dz_pre, d_hn)
    __shared__ float red[3 * H * H];
    for (int i = threadIdx.x; i < 3 * H * H; i += blockDim.x) Wsm[i] = __ldg(whh + (size_t)dir * 3 * H * H + i);
    __syncthreads();

    float acc[3][H];   // dW_hh[g*32+j][lane]
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < H; ++j) acc[a][j] = 0.f;
    float dbh[3] = {0.f, 0.f, 0.f}, dbi[3] = {0.f, 0.f, 0.f};

    for (int s = blockIdx.x * WARPS + warp; s < g.S; s += gridDim.x * WARPS) {
        const long base = seq_base(g, s);
        // BPTT walks the processing order backwards: dir0 processed t=0..T-1 -> walk T-1..0; dir1 processed T-1..0 -> walk 0..T-1
        long m = base + (dir ? 0 : (long)(g.T - 1) * g.t_stride);
        const long step = dir ? g.t_stride : -g.t_stride;      // towards the previously processed element
        float dh = 0.f;
        for (int t = 0; t < g.T; ++t) {
            const bool has_prev = t + 1 < g.T;
            float hprev = has_prev ? __ldg(y + (size_t)(m + step) * ldy + dir * H + lane) : 0.f;
            const float* sp = stash + ((size_t)m * 2 + dir) * 4 * H + lane;
            float r = __ldg(sp), z = __ldg(sp + H), n = __ldg(sp + 2 * H), hn = __ldg(sp + 3 * H);
            dh += __ldg(dy + (size_t)m * lddy + dir * H + lane);
            float dn = dh * (1.f - z);
            float dz = dh * (hprev - n);
            float dn_pre = dn * (1.f - n * n);
            float dr_pre = dn_pre * hn * r * (1.f - r);
            float dz_pre = dz * z * (1.f - z);
            float dhn = dn_pre * r;
            float* gp = dgx + (size_t)m * 6 * H + dir * 3 * H + lane;
            gp[0] = dr_pre; gp[H] = dz_pre; gp[2 * H] = dn_pre;
            dbi[0] += dr_pre; dbi[1] += dz_pre; dbi[2] += dn_pre;
            dbh[0] += dr_pre; dbh[1] += dz_pre; dbh[2] += dhn;
            __syncwarp();
            dg[warp][0][lane] = dr_pre; dg[warp][1][lane] = dz_pre; dg[warp][2][lane] = dhn;
            __syncwarp();
            float carry = dh * z;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int q = 0; q < H / 4; ++q) {
                    float4 v = *reinterpret_cast<const float4*>(&dg[warp][a][q * 4]);
                    const float* wp = &Wsm[(a * H + q * 4) * H + lane];
                    carry = fmaf(wp[0], v.x, carry);
                    carry = fmaf(wp[H], v.y, carry);
                    carry = fmaf(wp[2 * H], v.z, carry);
                    carry = fmaf(wp[3 * H], v.w, carry);
                    acc[a][q * 4 + 0] = fmaf(v.x, hprev, acc[a][q * 4 + 0]);
                    acc[a][q * 4 + 1] = fmaf(v.y, hprev, acc[a][q * 4 + 1]);
                    acc[a][q * 4 + 2] = fmaf(v.z, hprev, acc[a][q * 4 + 2]);
                    acc[a][q * 4 + 3] = fmaf(v.w, hprev, acc[a][q * 4 + 3]);
                }
            }
            dh = carry;
            m += step;
        }
    }
    // deterministic block reduction: warps add their accumulators one after another
    for (int wsel = 0; wsel < WARPS; ++wsel) {
        if (warp == wsel) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    int idx = (a * H + j) * H + lane;
                    red[idx] = (wsel == 0 ? 0.f : red[idx]) + acc[a][j];
                }
        }
        __syncthreads();
    }
    float* out = part + ((size_t)blockIdx.x * 2 + dir) * PART;
    for (int i = threadIdx.x; i < 3 * H * H; i += blockDim.x) out[i] = red[i];
    __syncthreads();
    // bias partials through the same buffer
    for (int wsel = 0; wsel < WARPS; ++wsel) {
        if (warp == wsel) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                red[a * H + lane] = (wsel == 0 ? 0.f : red[a * H + lane]) + dbh[a];
                red[3 * H + a * H + lane] = (wsel == 0 ? 0.f : red[3 * H + a * H + lane]) + dbi[a];
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 6 * H; i += blockDim.x) out[3 * H * H + i] = red[i];
}

__global__ void gru_bwd_reduce_kernel(const float* __restrict__ part, int nblocks, float* __restrict__ dwhh,
                                      float* __restrict__ dbhh, float* __restrict__ dbih, int accumulate) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over 2 * PART
    if (idx >= 2 * PART) return;
    int dir = idx / PART, e = idx % PART;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += part[((size_t)b * 2 + dir) * PART + e];
    float* dst;
    if (e < 3 * H * H) dst = dwhh + (size_t)dir * 3 * H * H + e;
    else if (e < 3 * H * H + 3 * H) dst = dbhh + dir * 3 * H + (e - 3 * H * H);
    else dst = dbih + dir * 3 * H + (e - 3 * H * H - 3 * H);
    *dst = accumulate ? *dst + s : s;
}

int gru_blocks(int S) {
    int b = (S + WARPS - 1) / WARPS;
    int cap = 148 * 3;
    return b < cap ? b : cap;
}

}  // namespace

extern "C" {

size_t tpgsr_gru_bwd_workspace_bytes(int S) { return (size_t)gru_blocks(S) * 2 * PART * sizeof(float); }

// whh: [2][96][32] (weight_hh_l0, weight_hh_l0_reverse), bhh: [2][96].
int tpgsr_gru_fwd(const float* gx, const float* whh, const float* bhh, float* y, int ldy, const float* residual, int ldr,
                  float* ysum, int ldys, float* stash, int S, int T, int inner, long outer_stride, long inner_stride,
                  long t_stride, cudaStream_t stream) {
    if (!gx || !whh || !bhh || !y || S <= 0 || T <= 0 || inner <= 0) return TPGSR_ERR_BAD_ARG;
    if (ysum && !residual) return TPGSR_ERR_BAD_ARG;
    SeqGeom g{S, T, inner, outer_stride, inner_stride, t_stride};
    dim3 grid(gru_blocks(S), 2);
    gru_fwd_kernel<<<grid, WARPS * 32, 0, stream>>>(gx, whh, bhh, y, ldy, residual, ldr, ysum, ldys, stash, g);
    return tpgsr_launch_status();
}

// dwhh [2][96][32], dbhh [2][96], dbih [2][96] are (+)= ; dgx [M,192] is overwritten.
int tpgsr_gru_bwd(const float* dy, int lddy, const float* y, int ldy, const float* stash, const float* whh, float* dgx,
                  float* dwhh, float* dbhh, float* dbih, int accumulate, int S, int T, int inner, long outer_stride,
                  long inner_stride, long t_stride, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!dy || !y || !stash || !whh || !dgx || !dwhh || !dbhh || !dbih || !workspace) return TPGSR_ERR_BAD_ARG;
    if (workspace_bytes < tpgsr_gru_bwd_workspace_bytes(S)) return TPGSR_ERR_WORKSPACE;
    SeqGeom g{S, T, inner, outer_stride, inner_stride, t_stride};
    int nb = gru_blocks(S);
    dim3 grid(nb, 2);
    gru_bwd_kernel<<<grid, WARPS * 32, 0, stream>>>(dy, lddy, y, ldy, stash, whh, dgx, static_cast<float*>(workspace), g);
    int rc = tpgsr_launch_status();
    if (rc) return rc;
    gru_bwd_reduce_kernel<<<ceil_div(2 * PART, 256), 256, 0, stream>>>(static_cast<float*>(workspace), nb, dwhh, dbhh, dbih, accumulate);
    return tpgsr_launch_status();
}

}  // extern "C"
