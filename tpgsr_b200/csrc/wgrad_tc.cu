// Tensor-core convolution weight gradient for sm_100a:  dW[co][ci][kh][kw] = sum_p dY[p, co] * X[p + shift(kh,kw), ci].
//
// Reference op replaced: the weight-gradient half of nn.Conv2d.backward for the 64-input-channel convolutions of TSRN
// (model/tsrn.py:37,376,381,467,495,43; cuDNN wgrad in the reference).
//
// GEMM view per tap:  D_tap[ci (M = 64), co (N)] += A'[ci, p] * B'[co, p]   with the PIXELS as the reduction dimension.
// Both operands come straight from the NHWC bf16 planes: a TMA box of [64 pixels x 64 channels] lands as 64 rows of 128 bytes
// = the SWIZZLE_128B *MN-major* canonical layout (row = K index, 128 bytes = 64 consecutive M/N elements), so no transpose
// is ever materialised.  A' for tap (kh,kw) is the X box shifted by (kh-ph, kw-pw) with TMA zero fill = the conv padding.
// One CTA = (pixel split, kh, kw-group, co tile): the accumulators of all its taps live side by side in TMEM (<= 512 cols).
// Split partials go to the same workspace layout as the CUDA-core path and are reduced by conv_wgrad_reduce (deterministic).
//
// Precision: NPROD = 3 (hi*hi + hi*lo + lo*hi, fp32-class) or 1 (plain bf16), as in conv_tc.cu.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int KB_PIX = 64;                       // pixels per K block
constexpr int TILE_BYTES = KB_PIX * 64 * 2;      // one [64 px][64 ch] bf16 tile = 8 KiB

struct WgradTcParams {
    int B, H, W, Ci, Co;
    int KH, KW, ph, pw;
    int kb_w, kb_h;             // K block = kb_h rows x kb_w cols of one image (kb_w * kb_h == 64)
    int kbs_per_row, kbs_per_img;
    int kw_groups, taps_per_group;
    int n_tile;                 // co handled per CTA (multiple of 64)
    int total_kb, kb_per_split;
    float* ws;                  // [splits][K = KH*KW*Ci][Co]
};

template <int NPROD>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                     const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo, WgradTcParams p,
                     int stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int planes = NPROD == 3 ? 2 : 1;
    const int kh = blockIdx.y / p.kw_groups, kwg = blockIdx.y % p.kw_groups;
    const int kw0 = kwg * p.taps_per_group;
    const int ntaps = (p.KW - kw0) < p.taps_per_group ? (p.KW - kw0) : p.taps_per_group;
    const int n0 = blockIdx.z * p.n_tile;
    const int ngroups = p.n_tile / 64;
    const int stage_bytes = planes * (ngroups + p.taps_per_group) * TILE_BYTES;
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared-space provenance (LDS/STS, not generic LD/ST)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    uint64_t* tmem_full = empty + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int kb_begin = blockIdx.x * p.kb_per_split;
    int kb_end = kb_begin + p.kb_per_split;
    if (kb_end > p.total_kb) kb_end = p.total_kb;
    const int nkb = kb_end > kb_begin ? kb_end - kb_begin : 0;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.taps_per_group * p.n_tile) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmX_hi);
        tc::prefetch_tmap(&tmD_hi);
        if (NPROD == 3) { tc::prefetch_tmap(&tmX_lo); tc::prefetch_tmap(&tmD_lo); }
        for (int s = 0; s < stages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tmem_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, tmem_cols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // stage layout: [dY hi groups][dY lo groups][X hi taps][X lo taps]
    const int dy_plane_bytes = ngroups * TILE_BYTES;
    const int x_plane_bytes = p.taps_per_group * TILE_BYTES;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % stages;
                const uint32_t ph = (i / stages) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                const int kb = kb_begin + i;
                const int b = kb / p.kbs_per_img;
                const int rem = kb - b * p.kbs_per_img;
                const int h0 = (rem / p.kbs_per_row) * p.kb_h;
                const int w0 = (rem % p.kbs_per_row) * p.kb_w;
                uint8_t* st = smem + (size_t)s * stage_bytes;
                tc::mbar_expect_tx(&full[s], (uint32_t)(planes * (ngroups + ntaps) * TILE_BYTES));
                for (int gI = 0; gI < ngroups; ++gI) {
                    tc::tma_load_4d(st + gI * TILE_BYTES, &tmD_hi, &full[s], n0 + gI * 64, w0, h0, b);
                    if (NPROD == 3) tc::tma_load_4d(st + dy_plane_bytes + gI * TILE_BYTES, &tmD_lo, &full[s], n0 + gI * 64, w0, h0, b);
                }
                uint8_t* xs = st + planes * dy_plane_bytes;
                for (int t = 0; t < ntaps; ++t) {
                    const int c1 = w0 + kw0 + t - p.pw, c2 = h0 + kh - p.ph;
                    tc::tma_load_4d(xs + t * TILE_BYTES, &tmX_hi, &full[s], 0, c1, c2, b);
                    if (NPROD == 3) tc::tma_load_4d(xs + x_plane_bytes + t * TILE_BYTES, &tmX_lo, &full[s], 0, c1, c2, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_bf16(64, p.n_tile, 1, 1);      // both operands MN-major
            for (int i = 0; i < nkb; ++i) {
                const int s = i % stages;
                const uint32_t ph = (i / stages) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::tc_fence_after();
                const uint32_t st = tc::smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t dy_hi = st, dy_lo = st + dy_plane_bytes;
                const uint32_t x_hi = st + planes * dy_plane_bytes, x_lo = x_hi + x_plane_bytes;
                for (int t = 0; t < ntaps; ++t) {
#pragma unroll
                    for (int k = 0; k < KB_PIX / 16; ++k) {
                        const uint32_t koff = k * 2048;          // 16 K rows = two 1024-byte swizzle atoms
                        const uint64_t a_hi = tc::smem_desc_sw128(x_hi + t * TILE_BYTES + koff, TILE_BYTES, 1024);
                        const uint64_t b_hi = tc::smem_desc_sw128(dy_hi + koff, TILE_BYTES, 1024);
                        const uint32_t d = tmem_base + (uint32_t)(t * p.n_tile);
                        tc::umma_bf16(d, a_hi, b_hi, idesc, (i | k) != 0);
                        if (NPROD == 3) {
                            const uint64_t a_lo = tc::smem_desc_sw128(x_lo + t * TILE_BYTES + koff, TILE_BYTES, 1024);
                            const uint64_t b_lo = tc::smem_desc_sw128(dy_lo + koff, TILE_BYTES, 1024);
                            tc::umma_bf16(d, a_hi, b_lo, idesc, 1);
                            tc::umma_bf16(d, a_lo, b_hi, idesc, 1);
                        }
                    }
                }
                tc::tc_commit(&empty[s]);
            }
            tc::tc_commit(tmem_full);
        }
    } else {
        // epilogue.  M = 64 accumulators use the half-subpartition layout: row ci = 16*q + l lives in TMEM lane 32*q + l (l < 16)
        const int q = warp & 3;
        tc::mbar_wait(tmem_full, 0);
        tc::tc_fence_after();
        const int ci = 16 * q + lane;
        const int K = p.KH * p.KW * p.Ci;
        for (int t = 0; t < ntaps; ++t) {
            const int tap = kh * p.KW + kw0 + t;
            for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
                float v[32];
                tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.n_tile + c0), v);
                tc::tmem_ld_wait();
                if (lane < 16) {
                    float* dst = p.ws + ((size_t)blockIdx.x * K + (size_t)tap * p.Ci + ci) * p.Co + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + c0 + j < p.Co) dst[j] = nkb > 0 ? v[j] : 0.f;
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, tmem_cols);
    }
}

// x (fp32, pitch ldx, C channels) -> bf16 hi/lo planes with Cout >= C channels per pixel (extra channels zero)
__global__ void __launch_bounds__(256) split_bf16_pad_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ hi,
                                                              __nv_bfloat16* __restrict__ lo, long M, int C, int Cout) {
    const long total = M * Cout;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i / Cout;
        int c = (int)(i - r * Cout);
        float f = c < C ? x[(size_t)r * ldx + c] : 0.f;
        __nv_bfloat16 h = __float2bfloat16_rn(f);
        hi[i] = h;
        if (lo) lo[i] = __float2bfloat16_rn(f - __bfloat162float(h));
    }
}

struct WgradPlan {
    int kb_w, kb_h, taps_per_group, kw_groups, n_tile, n_tiles, splits, kb_per_split, total_kb, stages, stage_bytes;
};

int plan_wgrad(int B, int H, int W, int Ci, int Co, int Cpad, int KH, int KW, int nprod, WgradPlan& pl) {
    if (Ci != 64 || Cpad % 64 != 0 || Cpad < Co) return TPGSR_ERR_UNSUPPORTED;
    pl.kb_w = W < 64 ? W : 64;
    if (64 % pl.kb_w != 0 || W % pl.kb_w != 0) return TPGSR_ERR_UNSUPPORTED;
    pl.kb_h = 64 / pl.kb_w;
    if (H % pl.kb_h != 0) return TPGSR_ERR_UNSUPPORTED;
    pl.n_tile = Cpad <= 256 ? Cpad : 256;
    if (Cpad % pl.n_tile != 0) return TPGSR_ERR_UNSUPPORTED;
    int max_taps = 512 / pl.n_tile;
    if (pl.n_tile == 256 && KW > 1) { pl.n_tile = 128; max_taps = 4; }
    if (Cpad % pl.n_tile != 0) return TPGSR_ERR_UNSUPPORTED;
    pl.n_tiles = Cpad / pl.n_tile;
    pl.kw_groups = (KW + max_taps - 1) / max_taps;
    pl.taps_per_group = (KW + pl.kw_groups - 1) / pl.kw_groups;
    const int planes = nprod == 3 ? 2 : 1;
    pl.stage_bytes = planes * (pl.n_tile / 64 + pl.taps_per_group) * TILE_BYTES;
    pl.stages = (200 * 1024) / pl.stage_bytes;
    if (pl.stages > 4) pl.stages = 4;
    if (pl.stages < 2) return TPGSR_ERR_UNSUPPORTED;
    pl.total_kb = B * H * W / 64;
    int ctas_per_split = KH * pl.kw_groups * pl.n_tiles;
    int want = (148 * 2 + ctas_per_split - 1) / ctas_per_split;
    if (want < 1) want = 1;
    if (want > pl.total_kb) want = pl.total_kb;
    pl.kb_per_split = (pl.total_kb + want - 1) / want;
    pl.splits = (pl.total_kb + pl.kb_per_split - 1) / pl.kb_per_split;
    return TPGSR_OK;
}

}  // namespace

// defined in conv_simt.cu
extern "C" int tpgsr_conv2d_wgrad_reduce(const float* ws, float* dw_oihw, int splits, int K, int N, int Ci, int KH, int KW,
                                         int accumulate, cudaStream_t stream);

extern "C" {

int tpgsr_split_bf16_pad(const float* x, int ldx, void* hi, void* lo, long M, int C, int Cout, cudaStream_t stream) {
    if (!x || !hi || Cout < C) return TPGSR_ERR_BAD_ARG;
    long blocks = (M * Cout + 255) / 256;
    split_bf16_pad_kernel<<<(int)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, stream>>>(
        x, ldx, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), M, C, Cout);
    return tpgsr_launch_status();
}

int tpgsr_conv2d_wgrad_tc_supported(int B, int H, int W, int Ci, int Co, int Cpad, int KH, int KW, int nprod) {
    WgradPlan pl;
    return plan_wgrad(B, H, W, Ci, Co, Cpad, KH, KW, nprod, pl) == TPGSR_OK ? 1 : 0;
}

size_t tpgsr_conv2d_wgrad_tc_workspace_bytes(int B, int H, int W, int Ci, int Co, int Cpad, int KH, int KW, int nprod) {
    WgradPlan pl;
    if (plan_wgrad(B, H, W, Ci, Co, Cpad, KH, KW, nprod, pl)) return 0;
    return (size_t)pl.splits * KH * KW * Ci * Co * sizeof(float);
}

// dW (OIHW) (+)= weight gradient of a 'same' conv.  x planes: [B,H,W,Ci] bf16; dy planes: [B,H,W,Cpad] bf16 (channels >= Co zero).
int tpgsr_conv2d_wgrad_tc(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw_oihw, int B, int H,
                          int W, int Ci, int Co, int Cpad, int KH, int KW, int ph, int pw, int accumulate, int nprod,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!x_hi || !dy_hi || !dw_oihw || !workspace || (nprod != 1 && nprod != 3) || (nprod == 3 && (!x_lo || !dy_lo)))
        return TPGSR_ERR_BAD_ARG;
    WgradPlan pl;
    int rc = plan_wgrad(B, H, W, Ci, Co, Cpad, KH, KW, nprod, pl);
    if (rc) return rc;
    if (workspace_bytes < (size_t)pl.splits * KH * KW * Ci * Co * sizeof(float)) return TPGSR_ERR_WORKSPACE;
    WgradTcParams p;
    p.B = B; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.KH = KH; p.KW = KW; p.ph = ph; p.pw = pw;
    p.kb_w = pl.kb_w; p.kb_h = pl.kb_h; p.kbs_per_row = W / pl.kb_w; p.kbs_per_img = (H / pl.kb_h) * (W / pl.kb_w);
    p.kw_groups = pl.kw_groups; p.taps_per_group = pl.taps_per_group; p.n_tile = pl.n_tile;
    p.total_kb = pl.total_kb; p.kb_per_split = pl.kb_per_split; p.ws = static_cast<float*>(workspace);

    CUtensorMap tmX_hi, tmX_lo, tmD_hi, tmD_lo;
    uint64_t xd[4] = {(uint64_t)Ci, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t xs[3] = {(uint64_t)Ci * 2, (uint64_t)W * Ci * 2, (uint64_t)H * W * Ci * 2};
    uint64_t dd[4] = {(uint64_t)Cpad, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t ds[3] = {(uint64_t)Cpad * 2, (uint64_t)W * Cpad * 2, (uint64_t)H * W * Cpad * 2};
    uint32_t box[4] = {64, (uint32_t)pl.kb_w, (uint32_t)pl.kb_h, 1};
    if (tpgsr_make_tmap_bf16(&tmX_hi, x_hi, 4, xd, xs, box)) return TPGSR_ERR_DRIVER;
    if (tpgsr_make_tmap_bf16(&tmD_hi, dy_hi, 4, dd, ds, box)) return TPGSR_ERR_DRIVER;
    if (nprod == 3) {
        if (tpgsr_make_tmap_bf16(&tmX_lo, x_lo, 4, xd, xs, box)) return TPGSR_ERR_DRIVER;
        if (tpgsr_make_tmap_bf16(&tmD_lo, dy_lo, 4, dd, ds, box)) return TPGSR_ERR_DRIVER;
    } else {
        tmX_lo = tmX_hi;
        tmD_lo = tmD_hi;
    }
    const size_t smem = (size_t)pl.stages * pl.stage_bytes + 1024 + 256;
    dim3 grid(pl.splits, KH * pl.kw_groups, pl.n_tiles);
    if (nprod == 3) {
        if (cudaFuncSetAttribute(conv_wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        conv_wgrad_tc_kernel<3><<<grid, 192, smem, stream>>>(tmX_hi, tmX_lo, tmD_hi, tmD_lo, p, pl.stages);
    } else {
        if (cudaFuncSetAttribute(conv_wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        conv_wgrad_tc_kernel<1><<<grid, 192, smem, stream>>>(tmX_hi, tmX_lo, tmD_hi, tmD_lo, p, pl.stages);
    }
    rc = tpgsr_launch_status();
    if (rc) return rc;
    return tpgsr_conv2d_wgrad_reduce(static_cast<float*>(workspace), dw_oihw, pl.splits, KH * KW * Ci, Co, Ci, KH, KW, accumulate, stream);
}

}  // extern "C"
