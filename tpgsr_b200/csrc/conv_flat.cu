// "Padded-flat" tensor-core convolution family for sm_100a: forward / data-gradient / weight-gradient of the 'same' convolutions
// with 64 input channels that carry ~90% of the TSRN FLOPs (model/tsrn.py:37,376,381,467,495,43; cuDNN in the reference).
//
// Why a second tcgen05 conv.  conv_tc.cu fetches one shifted [128 px x 64 ch] A tile per tap through a 4-D TMA: 9 fetches of the
// same pixels for a 3x3, 81 for the 9x9.  ncu (profiles/r01) shows that kernel pinned at ~43 B/cycle/SM = the L2->SM ceiling with the
// tensor pipe 22% busy.  Here the bf16 operand planes are stored PADDED-FLAT in HBM:
//
//     plane[q][c],  q = (b*(H+P) + r)*(W+P) + col,   rows r >= H and columns col >= W are zero   (P = conv padding)
//
// so that the neighbour of pixel q for tap (kh,kw) is simply row  q + (kh-ph)*(W+P) + (kw-pw)  of the same 2-D matrix — the
// shared zero row/column plays the role of the left/right/top/bottom padding (and TMA zero-fills rows before the first / after the
// last image).  A CTA loads the halo rows of MT consecutive 128-row tiles ONCE into shared memory (SWIZZLE_128B, one pixel = one
// 128-byte row) and every tap's A operand is the same buffer at a row offset: the UMMA descriptor start address moves by
// (kh*(W+P)+kw)*128 bytes.  A traffic drops from KH*KW x to (1 + halo/tile) x, weights are fetched once per MT tiles.
// Outputs at pad positions are computed and dropped (8% for 16x64 maps with P=1).
//
// The weight gradient uses the same planes: dW_tap[ci][co] = sum_q X[q + shift(tap)][ci] * dY[q][co] is a plain GEMM over ALL rows
// (pads contribute zeros), both operands MN-major straight from the planes, the X window shared by the KW taps of one kernel row.
//
// Precision modes as in conv_tc.cu: NPROD = 3 (hi*hi + hi*lo + lo*hi, fp32-class) or 1 (bf16); NPROD = 2 (A one plane, B hi + lo:
// A_hi*B_hi + A_hi*B_lo) is the experimental data-gradient mode (dY rounded to bf16 once, exact weights).
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int ROW_BYTES = 128;      // one pixel row of a plane: 64 bf16 channels
constexpr int TILE_M = 128;

// mbarrier wait that traps instead of hanging the GPU if a pipeline bug ever deadlocks (try_wait suspends up to ~10 ms per poll)
__device__ __forceinline__ void mbar_wait_safe(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = tc::smem_u32(bar);
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

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {             // 32 lanes x 16 consecutive columns
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// UMMA shared-memory descriptors, SWIZZLE_128B, SBO = 1024 bytes (8 rows of 128 bytes).  Row-shifted operand windows (start address a
// multiple of 128 but not of 1024 bytes) need NO base_offset: the hardware applies the swizzle to the absolute shared-memory address
// (probed on B200: scripts/probe_flat.py, profiles/r01/probe_flat.log — base_offset = (addr >> 7) & 7 gives garbage).
// The descriptor is split in two 32-bit words so that the issue loop only ever adds small constants to the low word:
//   lo = (addr >> 4) | (LBO >> 4) << 16        hi = (SBO >> 4) | version 1 << 14 | SWIZZLE_128B << 29 = 0x40004040
constexpr uint32_t DESC_HI = 0x40004040u;
// one lane of a converged warp (the MMA warp runs its loops warp-uniformly and only the tcgen05 instructions are predicated on the
// elected lane: ptxas then keeps descriptors / TMEM addresses in uniform registers and emits straight UTCHMMA sequences instead of
// one ELECT + R2UR + branch loop per instruction — measured 3x fewer issue slots per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }

// D[tmem] (+)= A * B with descriptors given as (lo word, shared hi word); ACC is a compile-time accumulate flag
// as umma_lo, with the high word of the A descriptor given (tall plan: SBO = the pitch of the 2-D halo tile instead of 1024 bytes)
template <int ACC>
__device__ __forceinline__ void umma_ab(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %6};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "n"(ACC), "r"(DESC_HI), "r"(a_hi)
        : "memory");
}
template <int ACC>
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "n"(ACC), "r"(DESC_HI)
        : "memory");
}
// ============================================================================================================ forward / dgrad
constexpr int EPI_WARPS = 8;        // warps 2..9: two per TMEM lane quarter, alternating over the (tile, 32-column chunk) work items
constexpr int FWD_THREADS = (2 + EPI_WARPS) * 32;
constexpr int STG_PITCH = 36;       // floats per staged row: 32 columns + 4 (keeps 16-byte alignment, spreads banks)
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_PITCH * 4;

struct FlatConvParams {
    int B, H, W, Hp, Wp;
    int Ho, Wo;                // output map: H + 2 ph - KH + 1, W + 2 pw - KW + 1 ('same' convs: H, W); output (r, c) sits at plane position (r, c)
    long long Qtot;            // B*Hp*Wp rows in the planes
    int Co, KH, KW, ph, pw;
    int MT;                    // 128-row output tiles per CTA
    int n_tile;                // output channels per CTA (multiple of 16, <= 128)
    int tm_stride;             // TMEM columns between the accumulators of consecutive tiles (n_tile rounded up to 32)
    int Ci;                    // input channels (multiple of 64): the planes are [Qtot][Ci], an A stage holds ONE 64-channel block
    int G, kgroups, n_groups;  // kernel rows per A stage; kgroups = KH / G; n_groups = (Ci / 64) * kgroups A stages per CTA: group g =
                               // (channel block g / kgroups, kernel rows (g % kgroups) * G ..)   (1x1 conv: G = kgroups = 1, n_groups = Ci / 64)
    int a_stages;              // A ring depth (1 when n_groups == 1)
    int box_rows, n_box;       // one A stage = n_box TMA boxes of box_rows rows (box_rows % 8 == 0)
    int tb, b_stages;          // taps per B stage; B ring depth
    int nstack;                // 1 (bf16x3, n_tile <= 64): [B_hi; B_lo] is used as ONE N = 2*n_tile operand, 2 MMAs per K step instead of 3
    int tile_outer;            // always 0: taps outermost, the MT tiles innermost (a weight stage feeds MT x 8..12 MMAs).  Tile-after-tile order
                               //    with the weights re-streamed per tile was built twice (round 1, round 2 with per-box barriers + overlapped
                               //    epilogue) and measured slower both times: L2-bound on the weight refetch — DESIGN.md §4
    int tall;                  // TALL plan (KW == 1, tall kernels: the kw-stacked 9x9 convs): a CTA owns MT tiles of 16 image rows x 8 columns;
                               //   ONE 4-D TMA box stages the [16 + KH - 1][8 MT] pixel halo, tap kh is the same buffer kh * 8 MT rows further,
                               //   the 8-row groups of a tile are 8 MT rows apart (A descriptor SBO = 8 MT * 128 bytes)
    int tcols, n_hg, n_wg;     // tall: columns per CTA (8 MT); row groups / column groups per image
    int ring_bytes;            // bytes of the A+B rings
    int stg_off;               // byte offset of the epilogue staging area (== 0: aliases the rings, only legal when !tile_outer)
    int bar_off;               // byte offset of the barrier block
    int ldy, act, out_mode, accumulate;
    const float* bias;
    const float* act_param;
    float* y;
    __nv_bfloat16* yp_hi;      // optional: the (activated) output also as dense bf16 hi/lo planes [M][Co] = the P = 0 operand planes of a following
    __nv_bfloat16* yp_lo;      // 1x1 conv / weight gradient (saves a tpgsr_split_bf16_flat pass); requires Co % 32 == 0, NHWC output
    int dbg_skip;              // developer experiments: bit0 = skip the global stores, bit1 = skip TMEM load + staging
    long long* dbg;            // optional [grid][16] clock64 timestamps (developer instrumentation, scripts/flat_timeline.py)
};

// the 4 K steps (x 1 or 3 products) of one (tile, tap): 128 x n_tile x 64 MACs.  FIRST: the first MMA overwrites the accumulator.
template <int NPROD, int FIRST>
__device__ __forceinline__ void mma_tap(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, uint32_t idesc, uint32_t ahi) {
    if (FIRST) umma_ab<0>(d, ah, ahi, bh, idesc); else umma_ab<1>(d, ah, ahi, bh, idesc);
    if (NPROD >= 2) umma_ab<1>(d, ah, ahi, bl, idesc);
    if (NPROD == 3) umma_ab<1>(d, al, ahi, bh, idesc);
#pragma unroll
    for (int k = 1; k < 4; ++k) {          // 16 bf16 = 32 bytes = 2 descriptor units inside the 128-byte swizzle row
        umma_ab<1>(d, ah + 2 * k, ahi, bh + 2 * k, idesc);
        if (NPROD >= 2) umma_ab<1>(d, ah + 2 * k, ahi, bl + 2 * k, idesc);
        if (NPROD == 3) umma_ab<1>(d, al + 2 * k, ahi, bh + 2 * k, idesc);
    }
}

// split-bf16 with the weight planes stacked along N: the stage holds B_hi (n_tile rows) immediately followed by B_lo, i.e. one K-major
// operand of 2*n_tile rows.  D[:, 0:n] += A_hi B_hi^T and D[:, n:2n] += A_hi B_lo^T in ONE MMA (idesc2, N = 2n), then D[:, 0:n] += A_lo B_hi^T
// (idesc1).  The epilogue adds the two column halves.  Per K step the A_hi tile is read from shared memory once instead of twice: the MMA
// phase of an N = 64 conv is shared-memory-bandwidth bound (6 KB per 32-cycle MMA), so this is 112 instead of 144 smem cycles per K step.
// NPROD = 2 (A has no lo plane): only the stacked MMA — ONE 128 x 2n x 16 MMA per K step.
template <int NPROD, int FIRST>
__device__ __forceinline__ void mma_tap_stacked(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t idesc1, uint32_t idesc2, uint32_t ahi) {
    if (FIRST) umma_ab<0>(d, ah, ahi, bh, idesc2); else umma_ab<1>(d, ah, ahi, bh, idesc2);
    if (NPROD == 3) umma_ab<1>(d, al, ahi, bh, idesc1);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        umma_ab<1>(d, ah + 2 * k, ahi, bh + 2 * k, idesc2);
        if (NPROD == 3) umma_ab<1>(d, al + 2 * k, ahi, bh + 2 * k, idesc1);
    }
}

template <int NPROD>
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv_flat_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, FlatConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int a_planes = NPROD == 3 ? 2 : 1, b_planes = NPROD >= 2 ? 2 : 1;
    const int a_plane_bytes = p.n_box * p.box_rows * ROW_BYTES;
    const int a_stage_bytes = a_planes * a_plane_bytes;
    const int b_tap_bytes = p.n_tile * ROW_BYTES;                 // one tap, one plane: [n_tile co][64 ci] K-major
    const int b_plane_bytes = p.tb * b_tap_bytes;
    const int b_stage_bytes = b_planes * b_plane_bytes;
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared-space provenance (LDS/STS, not generic LD/ST)
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + (size_t)p.a_stages * a_stage_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + p.bar_off);
    uint64_t* a_empty = a_full + p.a_stages;
    uint64_t* b_full = a_empty + p.a_stages;
    uint64_t* b_empty = b_full + p.b_stages;
    uint64_t* tmem_full = b_empty + p.b_stages;                   // [MT] one per tile (tile_outer) / only [0] used otherwise
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 4);
    float* bias_s = reinterpret_cast<float*>(smem + p.bar_off + 256);      // [n_tile], 16-byte aligned (barrier block <= 256 bytes)

    const long long q0 = (long long)blockIdx.x * p.MT * TILE_M;   // first output row of this CTA
    const int n0 = blockIdx.y * p.n_tile;
    const int taps_per_group = p.G * p.KW;
    const int b_per_group = taps_per_group / p.tb;                // B stages consumed per A group (per tile when tile_outer)
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.MT * p.tm_stride) tmem_cols <<= 1;
    long long* dbg = p.dbg ? p.dbg + (size_t)(blockIdx.x + blockIdx.y * gridDim.x) * 16 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmA_hi);
        tc::prefetch_tmap(&tmB_hi);
        if (NPROD == 3) tc::prefetch_tmap(&tmA_lo);
        if (NPROD >= 2) tc::prefetch_tmap(&tmB_lo);
        for (int s = 0; s < p.a_stages; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < p.b_stages; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
        for (int t = 0; t < 4; ++t) tc::mbar_init(&tmem_full[t], 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, tmem_cols);
        tc::tmem_relinquish();
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < p.n_tile; i += EPI_WARPS * 32)
            bias_s[i] = (p.bias && n0 + i < p.Co) ? __ldg(p.bias + n0 + i) : 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            uint32_t b_it = 0;
            for (int g = 0; g < p.n_groups; ++g) {
                const int as = g % p.a_stages;
                mbar_wait_safe(&a_empty[as], ((g / p.a_stages) & 1) ^ 1);
                uint8_t* ast = a_ring + (size_t)as * a_stage_bytes;
                tc::mbar_expect_tx(&a_full[as], (uint32_t)a_stage_bytes);
                // first halo row of this group's kernel rows, in this group's 64-channel block
                const int cb = g / p.kgroups, gk = g - cb * p.kgroups;
                const long long row0 = q0 + (long long)(gk * p.G - p.ph) * p.Wp - p.pw;
                const int ch0 = cb * 64;
                if (p.tall) {          // one 4-D box [halo rows][8 MT columns][64 channels] per plane; rows / columns outside the plane grid arrive as zeros
                    const int wg = blockIdx.x % p.n_wg, hb = blockIdx.x / p.n_wg, hg = hb % p.n_hg, bimg = hb / p.n_hg;
                    tc::tma_load_4d(ast, &tmA_hi, &a_full[as], 0, wg * p.tcols, hg * 16 - p.ph, bimg);
                    if (NPROD == 3) tc::tma_load_4d(ast + a_plane_bytes, &tmA_lo, &a_full[as], 0, wg * p.tcols, hg * 16 - p.ph, bimg);
                }
                for (int bx = 0; bx < (p.tall ? 0 : p.n_box); ++bx) {
                    const int r = (int)(row0 + (long long)bx * p.box_rows);
                    tc::tma_load_2d(ast + (size_t)bx * p.box_rows * ROW_BYTES, &tmA_hi, &a_full[as], ch0, r);
                    if (NPROD == 3)
                        tc::tma_load_2d(ast + a_plane_bytes + (size_t)bx * p.box_rows * ROW_BYTES, &tmA_lo, &a_full[as], ch0, r);
                }
                const int reps = 1;
                for (int rep = 0; rep < reps; ++rep) {
                    for (int j = 0; j < b_per_group; ++j, ++b_it) {
                        const int bs = b_it % p.b_stages;
                        mbar_wait_safe(&b_empty[bs], ((b_it / p.b_stages) & 1) ^ 1);
                        uint8_t* bst = b_ring + (size_t)bs * b_stage_bytes;
                        tc::mbar_expect_tx(&b_full[bs], (uint32_t)b_stage_bytes);
                        for (int tt = 0; tt < p.tb; ++tt) {
                            const int tap = gk * taps_per_group + j * p.tb + tt;
                            const int k0 = tap * p.Ci + ch0;                   // K index of the packed weights [Co][KH*KW*Ci]
                            tc::tma_load_2d(bst + (size_t)tt * b_tap_bytes, &tmB_hi, &b_full[bs], k0, n0);
                            if (NPROD >= 2)
                                tc::tma_load_2d(bst + b_plane_bytes + (size_t)tt * b_tap_bytes, &tmB_lo, &b_full[bs], k0, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The warp runs the loops converged; only the tcgen05 instructions are predicated on the elected lane.  Per-MMA issue cost matters
        // (a 128x64x16 MMA is 32 tensor-pipe cycles): descriptors advance by adding constants to their low words.
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);      // same value, provably warp-uniform
        const bool leader = elect_one();
        const uint32_t idesc = tc::idesc_bf16(TILE_M, p.n_tile, 0, 0);
        const uint32_t idesc2 = tc::idesc_bf16(TILE_M, 2 * p.n_tile, 0, 0);
        const bool stacked = NPROD >= 2 && p.nstack;
        const uint32_t tile_lo = p.tall ? (8 * ROW_BYTES) >> 4 : (TILE_M * ROW_BYTES) >> 4;        // descriptor units (16 bytes) between consecutive tiles
        const uint32_t tap_rows = p.tall ? (uint32_t)p.tcols : (uint32_t)p.Wp;                       // rows between vertically adjacent taps
        const uint32_t ahi = p.tall ? (0x40004000u | (uint32_t)((p.tcols * ROW_BYTES) >> 4)) : DESC_HI;   // A descriptor: SBO = pitch of the 8-row groups
        uint32_t b_it = 0;
        {
            for (int g = 0; g < p.n_groups; ++g) {
                const int as = g % p.a_stages;
                mbar_wait_safe(&a_full[as], (g / p.a_stages) & 1);
                tc::tc_fence_after();
                if (dbg && g == 0 && leader) dbg[2] = clock64();
                const uint32_t a_hi0 = desc_lo(tc::smem_u32(a_ring + (size_t)as * a_stage_bytes), 16);
                const uint32_t a_lo0 = a_hi0 + (a_plane_bytes >> 4);
                for (int j = 0; j < b_per_group; ++j, ++b_it) {
                    const int bs = b_it % p.b_stages;
                    mbar_wait_safe(&b_full[bs], (b_it / p.b_stages) & 1);
                    tc::tc_fence_after();
                    if (dbg && b_it == 0 && leader) dbg[3] = clock64();
                    const uint32_t b_hi0 = desc_lo(tc::smem_u32(b_ring + (size_t)bs * b_stage_bytes), 16);
                    const uint32_t b_lo0 = b_hi0 + (b_plane_bytes >> 4);
                    for (int tt = 0; tt < p.tb; ++tt) {
                        const int tl = j * p.tb + tt;                  // tap index inside the group
                        const int khl = tl / p.KW, kw = tl - khl * p.KW;
                        const uint32_t row_off = (uint32_t)(khl * tap_rows + kw) * (ROW_BYTES >> 4);
                        const uint32_t bh = b_hi0 + tt * (b_tap_bytes >> 4), bl = b_lo0 + tt * (b_tap_bytes >> 4);
                        if (!leader) continue;
                        uint32_t ah = a_hi0 + row_off, al = a_lo0 + row_off, d = tmem_u;
                        if (stacked) {
                            if (g == 0 && tl == 0) {
                                for (int t = 0; t < p.MT; ++t, ah += tile_lo, al += tile_lo, d += p.tm_stride) mma_tap_stacked<NPROD, 1>(d, ah, al, bh, idesc, idesc2, ahi);
                            } else {
                                for (int t = 0; t < p.MT; ++t, ah += tile_lo, al += tile_lo, d += p.tm_stride) mma_tap_stacked<NPROD, 0>(d, ah, al, bh, idesc, idesc2, ahi);
                            }
                        } else if (g == 0 && tl == 0) {
                            for (int t = 0; t < p.MT; ++t, ah += tile_lo, al += tile_lo, d += p.tm_stride) mma_tap<NPROD, 1>(d, ah, al, bh, bl, idesc, ahi);
                        } else {
                            for (int t = 0; t < p.MT; ++t, ah += tile_lo, al += tile_lo, d += p.tm_stride) mma_tap<NPROD, 0>(d, ah, al, bh, bl, idesc, ahi);
                        }
                    }
                    if (leader) tc::tc_commit(&b_empty[bs]);
                }
                if (leader) tc::tc_commit(&a_empty[as]);
            }
            if (dbg && leader) dbg[4] = clock64();
            if (leader) tc::tc_commit(&tmem_full[0]);
        }
    } else {
        // ================= epilogue: warps 2..9.  Warp w owns TMEM lanes 32*(w % 4).. (= rows 32*(w%4).. of every tile) and
        // takes every second (tile, 32-column chunk) work item: TMEM -> registers -> (+bias) -> private staging rows -> coalesced stores.
        const int e = warp - 2, q = warp & 3, half = e >> 2;      // hardware rule: warp w may only read TMEM lanes 32*(w % 4)..+31
        float* stg = reinterpret_cast<float*>(smem + p.stg_off) + (size_t)e * 32 * STG_PITCH;
        const float slope = (p.act == TPGSR_ACT_PRELU && p.act_param) ? __ldg(p.act_param) : 0.f;
        const int nchunks = (p.n_tile + 31) / 32;
        const unsigned img = (unsigned)(p.Hp * p.Wp);
        const unsigned qtot = (unsigned)p.Qtot;               // < 2^31 (checked by the plan)
        int item = 0;
        for (int t = 0; t < p.MT; ++t) {
            bool waited = false;
            int pix = -1, bi = 0, r = 0, c = 0;
            for (int ch = 0; ch < nchunks; ++ch, ++item) {
                if ((item & 1) != half) continue;
                if (!waited) {
                    mbar_wait_safe(&tmem_full[0], 0);
                    tc::tc_fence_after();
                    waited = true;
                    if (dbg && e == 0 && lane == 0 && t == 0) dbg[5] = clock64();
                    // dense pixel index of this lane's output row (-1: pad position or beyond the last image)
                    if (p.tall) {          // tile row m = 8 i + j  <->  image row 16 hg + i, column 8 MT wg + 8 t + j
                        const int m = q * 32 + lane;
                        const int wg = blockIdx.x % p.n_wg, hb = blockIdx.x / p.n_wg, hg = hb % p.n_hg;
                        bi = hb / p.n_hg; r = hg * 16 + (m >> 3); c = wg * p.tcols + 8 * t + (m & 7);
                        pix = (r < p.Ho && c < p.Wo) ? (bi * p.Ho + r) * p.Wo + c : -1;
                    } else {
                    const unsigned qrow = (unsigned)q0 + (unsigned)(t * TILE_M + q * 32 + lane);
                    const unsigned ub = qrow / img, rem = qrow - ub * img;
                    const unsigned ur = rem / (unsigned)p.Wp, uc = rem - ur * (unsigned)p.Wp;
                    bi = (int)ub; r = (int)ur; c = (int)uc;
                    pix = (qrow < qtot && r < p.Ho && c < p.Wo) ? (bi * p.Ho + r) * p.Wo + c : -1;
                    }
                }
                const int c0 = ch * 32;
                const int cw = (p.n_tile - c0) >= 32 ? 32 : (p.n_tile - c0);       // 32 or 16 columns in this chunk
                int ncols = p.Co - n0 - c0;                                         // valid output channels in this chunk
                ncols = ncols < 0 ? 0 : (ncols > cw ? cw : ncols);
                float v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.tm_stride + c0);
                if (p.dbg_skip & 2) continue;
                if (cw == 32) tc::tmem_ld_32x32(taddr, v); else tmem_ld_32x16(taddr, v);
                tc::tmem_ld_wait();
                if (NPROD >= 2 && p.nstack) {          // stacked weights: columns [n_tile, 2 n_tile) hold the A_hi B_lo^T partial sums
                    float v2[32];
                    if (cw == 32) tc::tmem_ld_32x32(taddr + (uint32_t)p.n_tile, v2); else tmem_ld_32x16(taddr + (uint32_t)p.n_tile, v2);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += v2[j];
                }
                if (dbg && e == 0 && lane == 0 && item == 0) dbg[8] = clock64();
                float* drow = stg + (size_t)lane * STG_PITCH;
                const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = b4[j];
                    *reinterpret_cast<float4*>(drow + 4 * j) = make_float4(v[4 * j] + bb.x, v[4 * j + 1] + bb.y, v[4 * j + 2] + bb.z, v[4 * j + 3] + bb.w);
                }
                if (cw == 32) {
#pragma unroll
                    for (int j = 4; j < 8; ++j) {
                        const float4 bb = b4[j];
                        *reinterpret_cast<float4*>(drow + 4 * j) = make_float4(v[4 * j] + bb.x, v[4 * j + 1] + bb.y, v[4 * j + 2] + bb.z, v[4 * j + 3] + bb.w);
                    }
                }
                __syncwarp();
                if (dbg && e == 0 && lane == 0 && item == 0) dbg[9] = clock64();
                if (p.dbg_skip & 1) continue;
                if (p.out_mode == TPGSR_OUT_NHWC) {
                    if (ncols == cw && (p.ldy & 3) == 0 && ((n0 + c0) & 3) == 0 && (p.act == TPGSR_ACT_NONE || p.act == TPGSR_ACT_RELU)) {
                        // a lane owns one 16-byte column chunk; 4 (cw = 32) or 8 (cw = 16) rows per step.  All shuffles, then all shared-memory
                        // loads, then the (predicated) stores: three batches of independent instructions instead of a branchy dependent loop
                        // (the loop form cost ~200 cycles per row group, profiles/r01/flat_timeline_v4.log)
                        const int sh = cw == 32 ? 3 : 2;
                        const int lr = lane >> sh, c4 = lane & ((1 << sh) - 1);
                        const int R = 32 >> sh, nsteps = 32 / R;               // 8 or 4 steps
                        const float* src = stg + (size_t)lr * STG_PITCH + c4 * 4;
                        float* dcol = p.y + n0 + c0 + c4 * 4;
                        const size_t ldy = (size_t)p.ldy;
                        int pxs[8];
                        float4 vals[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) pxs[it] = it < nsteps ? __shfl_sync(0xffffffffu, pix, (lr + it * R) & 31) : -1;
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            if (it < nsteps) vals[it] = *reinterpret_cast<const float4*>(src + (size_t)it * R * STG_PITCH);
                        if (p.act == TPGSR_ACT_RELU) {          // (other activations take the generic path below)
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                if (it < nsteps) {
                                    vals[it].x = fmaxf(vals[it].x, 0.f); vals[it].y = fmaxf(vals[it].y, 0.f);
                                    vals[it].z = fmaxf(vals[it].z, 0.f); vals[it].w = fmaxf(vals[it].w, 0.f);
                                }
                            }
                        }
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            if (it < nsteps && pxs[it] >= 0) {
                                float4 val = vals[it];
                                float4* dst = reinterpret_cast<float4*>(dcol + (size_t)pxs[it] * ldy);
                                if (p.accumulate) {
                                    const float4 o = *dst;
                                    val.x += o.x; val.y += o.y; val.z += o.z; val.w += o.w;
                                }
                                *dst = val;
                                if (p.yp_hi) {
                                    const float f[4] = {val.x, val.y, val.z, val.w};
                                    __nv_bfloat16 h[4], l[4];
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        h[u] = __float2bfloat16_rn(f[u]);
                                        l[u] = __float2bfloat16_rn(f[u] - __bfloat162float(h[u]));
                                    }
                                    const size_t po = (size_t)pxs[it] * p.Co + n0 + c0 + c4 * 4;
                                    *reinterpret_cast<uint2*>(p.yp_hi + po) = *reinterpret_cast<uint2*>(h);
                                    if (p.yp_lo) *reinterpret_cast<uint2*>(p.yp_lo + po) = *reinterpret_cast<uint2*>(l);
                                }
                            }
                        }
                    } else if (ncols > 0) {
                        for (int i0 = 0; i0 < 32 * ncols; i0 += 32) {      // ragged channel counts (e.g. 37): scalar stores
                            const int i = i0 + lane;
                            const int row = i / ncols, cc = i - row * ncols;
                            const int px = __shfl_sync(0xffffffffu, pix, row & 31);
                            if (row < 32 && px >= 0) {
                                const float val = apply_act(stg[(size_t)row * STG_PITCH + cc], p.act, slope);
                                float* dst = p.y + (size_t)px * p.ldy + n0 + c0 + cc;
                                *dst = p.accumulate ? *dst + val : val;
                            }
                        }
                    }
                } else if (pix >= 0) {
                    // NCHW: for a fixed channel the 32 lanes hold 32 consecutive flat positions = (mostly) consecutive columns
                    for (int n = 0; n < ncols; ++n) {
                        const float val = apply_act(drow[n], p.act, slope);
                        float* dst = p.y + (((size_t)bi * p.Co + n0 + c0 + n) * p.Ho + r) * p.Wo + c;
                        *dst = p.accumulate ? *dst + val : val;
                    }
                }
                __syncwarp();
            }
            if (dbg && e == 0 && lane == 0 && t == 0) dbg[6] = clock64();
        }
        if (dbg && e == 0 && lane == 0) dbg[7] = clock64();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ============================================================================================================ weights-stationary plan
// 3x3 (up to 12 taps) 64 -> 64 convolutions, forward and data gradient: the dominant kernel of the TSRN (11 of 18 convs per SR pass).
// The plans above keep the PIXELS on the M side and stream the weights through shared memory once per CTA; with N = 64 outputs the MMA
// phase is bound by shared-memory operand reads (14 KB per K step) and the load / MMA / store phases of a CTA cannot overlap because the
// 144 KB of split weights do not fit beside the halo (tile-outer order re-streams them and becomes L2-bound).
// Here the GEMM is transposed:  D^T[co][pixel] = W[co][k] * X^T[k][pixel]
//   * A = the split weights, RESIDENT IN TENSOR MEMORY for the whole CTA: lanes 0..63 = W_hi rows, lanes 64..127 = W_lo rows,
//     KH*KW*32 columns (two bf16 per 32-bit cell).  Loaded once from L2 by the epilogue warps (tcgen05.st) while the TMA fetches the halo.
//   * B = the pixel halo in shared memory exactly as before (one 128-byte row per pixel, SWIZZLE_128B, tap = row offset), N = 96 or 64
//     pixels per MMA.  Shared-memory operand traffic per K step drops to N x 32 bytes (64 B/cycle at the tensor rate).
//   * two MMAs per K step (X_hi, X_lo) against the stacked [W_hi; W_lo] give all FOUR partial products; out[co] = D[co] + D[co + 64].
//   * the accumulator is double buffered (2 x N columns beside the weights: 288 + 192 = 480 of 512 columns for a 3x3), the MMA warp
//     runs tile after tile, the halo arrives in per-box barriers, and the 8 epilogue warps drain tile t while tile t + 1 is multiplied.
constexpr int WS_MAXBOX = 24;

struct WsParams {
    int Hp, Wp, Ho, Wo;
    long long Qtot;
    int Co, KH, KW, ph, pw;
    int NT;                    // pixel tiles per CTA
    int box_rows, n_box;       // halo = n_box TMA boxes of box_rows rows per plane, one mbarrier each
    int w_stages;              // depth of the weight staging ring of the prologue (one 16 KB tap per stage)
    int K;                     // KH*KW*64
    int stg_off, bar_off;
    int ldy, act, accumulate;
    const float* bias;
    const float* act_param;
    float* y;
    __nv_bfloat16* yp_hi;
    __nv_bfloat16* yp_lo;
    const __nv_bfloat16* w_hi; // packed K-major weights [64][K]
    const __nv_bfloat16* w_lo;
    long long* dbg;
};

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {              // 32 lanes x 16 consecutive columns
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// D[tmem] (+)= A[tmem] * B[smem descriptor]
template <int ACC>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "n"(ACC), "r"(DESC_HI)
        : "memory");
}

constexpr int WS_WSTAGES = 9;         // most stages of the weight staging ring: one tap (hi + lo tile, 16 KB) per stage; the ring starts in the
                                      // epilogue staging area and takes whatever shared memory the halo leaves (plan: w_stages >= 3)
constexpr int WS_MAXTAPS = 12;

template <int N, int TAPS>      // TAPS: compile-time tap count with a fully unrolled issue sequence (9 = 3x3), 0 = any
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv_ws_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, WsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int NH = N / 2;          // pixels (TMEM columns) an epilogue warp reads per tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int plane_bytes = p.n_box * p.box_rows * ROW_BYTES;
    uint8_t* halo = smem;                                          // [hi plane][lo plane]
    uint8_t* stg = smem + p.stg_off;                               // prologue: weight ring; afterwards: T_hi [N][64], T_lo [N][64] floats
    uint64_t* box_bar = reinterpret_cast<uint64_t*>(smem + p.bar_off);
    uint64_t* w_full = box_bar + WS_MAXBOX;                        // [WS_WSTAGES]
    uint64_t* w_empty = w_full + WS_WSTAGES;
    uint64_t* w_tap = w_empty + WS_WSTAGES;                        // [WS_MAXTAPS]: tap k of the weights is in tensor memory
    uint64_t* acc_full = w_tap + WS_MAXTAPS;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int* pixtab = reinterpret_cast<int*>(tmem_slot + 2);           // [N] dense pixel index of the current tile's rows (-1: dropped)

    const long long q0 = (long long)blockIdx.x * p.NT * N;        // first flat output row of this CTA
    int nt = (int)((p.Qtot - q0 + N - 1) / N);                     // tiles with at least one row inside the planes
    nt = nt > p.NT ? p.NT : nt;
    const int taps = p.KH * p.KW;
    const int w_cols = taps * 32;
    const int rot = blockIdx.x % taps;                             // this CTA's first weight tap
    int nb = (nt * N + (p.KH - 1) * p.Wp + p.KW - 1 + p.box_rows - 1) / p.box_rows;
    nb = nb > p.n_box ? p.n_box : nb;
    long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 32 : nullptr;      // developer timeline: 32 slots per CTA (scripts/ws_timeline.py)
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmX_hi);
        tc::prefetch_tmap(&tmX_lo);
        tc::prefetch_tmap(&tmW_hi);
        tc::prefetch_tmap(&tmW_lo);
        for (int s = 0; s < WS_MAXBOX; ++s) tc::mbar_init(&box_bar[s], 1);
        for (int s = 0; s < WS_WSTAGES; ++s) { tc::mbar_init(&w_full[s], 1); tc::mbar_init(&w_empty[s], EPI_WARPS); }
        for (int s = 0; s < WS_MAXTAPS; ++s) tc::mbar_init(&w_tap[s], EPI_WARPS);
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], EPI_WARPS); }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer: weight taps through the ring, the halo box after box.  Order = what tile 0 needs first.
            // Every CTA needs the same 144 KB of weights at the same moment: CTA i starts with tap i % taps, so that concurrent CTAs pull
            // different tiles through L2 (measured neutral: the prologue is bound by L2 -> SM throughput, profiles/r02/ws_timeline_history.md).
            const long long row0 = q0 - (long long)p.ph * p.Wp - p.pw;
            int bx = 0;
            const int early = ((nt >= 2 ? 2 : 1) * N + (p.KH - 1) * p.Wp + p.KW - 1 + p.box_rows - 1) / p.box_rows;      // boxes the first pass (tiles 0, 1) reads
            for (int i = 0, k = rot; i < taps; ++i, k = (k + 1 == taps ? 0 : k + 1)) {
                const int s = i % p.w_stages;
                mbar_wait_safe(&w_empty[s], ((i / p.w_stages) & 1) ^ 1);
                tc::mbar_expect_tx(&w_full[s], 2 * 64 * ROW_BYTES);
                tc::tma_load_2d(stg + (size_t)s * 2 * 64 * ROW_BYTES, &tmW_hi, &w_full[s], k * 64, 0);
                tc::tma_load_2d(stg + (size_t)s * 2 * 64 * ROW_BYTES + 64 * ROW_BYTES, &tmW_lo, &w_full[s], k * 64, 0);
                if (i == 0) {
                    for (; bx < early && bx < nb; ++bx) {
                        tc::mbar_expect_tx(&box_bar[bx], (uint32_t)(2 * p.box_rows * ROW_BYTES));
                        const int r = (int)(row0 + (long long)bx * p.box_rows);
                        tc::tma_load_2d(halo + (size_t)bx * p.box_rows * ROW_BYTES, &tmX_hi, &box_bar[bx], 0, r);
                        tc::tma_load_2d(halo + plane_bytes + (size_t)bx * p.box_rows * ROW_BYTES, &tmX_lo, &box_bar[bx], 0, r);
                    }
                }
            }
            for (; bx < nb; ++bx) {
                tc::mbar_expect_tx(&box_bar[bx], (uint32_t)(2 * p.box_rows * ROW_BYTES));
                const int r = (int)(row0 + (long long)bx * p.box_rows);
                tc::tma_load_2d(halo + (size_t)bx * p.box_rows * ROW_BYTES, &tmX_hi, &box_bar[bx], 0, r);
                tc::tma_load_2d(halo + plane_bytes + (size_t)bx * p.box_rows * ROW_BYTES, &tmX_lo, &box_bar[bx], 0, r);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (loops warp-uniform, no divisions per tap: an N = 96 MMA is 48 tensor cycles) =================
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
        const bool leader = elect_one();
        const uint32_t idesc = tc::idesc_bf16(128, N, 0, 0);
        const uint32_t x_hi0 = desc_lo(tc::smem_u32(halo), 16);
        const uint32_t x_lo0 = x_hi0 + (plane_bytes >> 4);
        int boxes_ready = 0;
        const int kh0 = rot / p.KW, kw0 = rot - kh0 * p.KW;
        uint32_t tap_off[TAPS > 0 ? TAPS : 1], tap_col[TAPS > 0 ? TAPS : 1];      // per tap (rotated order): B row offset in descriptor units, A column
        if (TAPS > 0) {
            int k = rot, kh = kh0, kw = kw0;
#pragma unroll
            for (int i = 0; i < (TAPS > 0 ? TAPS : 1); ++i) {
                tap_off[i] = (uint32_t)(kh * p.Wp + kw) * (ROW_BYTES >> 4);
                tap_col[i] = (uint32_t)(k * 32);
                ++k; ++kw;
                if (kw == p.KW) { kw = 0; ++kh; }
                if (k == taps) { k = 0; kh = 0; kw = 0; }
            }
        }
        // ---- first pass: tiles 0 and 1 together, tap after tap as the weights arrive.  The prologue is bound by L2 -> SM throughput
        // (~7 000 cycles for the weights + the first boxes) and one tile of MMAs is 4 500 cycles: two tiles keep the tensor pipe busy
        // through all of it, and both accumulators are free anyway (15.9 against 16.7 us per launch, profiles/r02/call16_stdout.log).
        const int first = nt >= 2 ? 2 : 1;
        {
            int need = ((first - 1) * N + (p.KH - 1) * p.Wp + p.KW - 1 + N - 1) / p.box_rows;
            need = need >= nb ? nb - 1 : need;
            while (boxes_ready <= need) {
                mbar_wait_safe(&box_bar[boxes_ready], 0);
                ++boxes_ready;
                if (dbg && leader && boxes_ready == 1) dbg[2] = clock64();
            }
            tc::tc_fence_after();
            int k = rot, kh = kh0, kw = kw0;
            for (int i = 0; i < taps; ++i) {
                mbar_wait_safe(&w_tap[k], 0);
                tc::tc_fence_after();
                if (dbg && leader && i == taps - 1) dbg[3] = clock64();
                for (int u = 0; u < first; ++u) {
                    if (leader) {
                        const uint32_t d = tmem_u + (uint32_t)(w_cols + u * N);
                        const uint32_t off = (uint32_t)(u * N + kh * p.Wp + kw) * (ROW_BYTES >> 4);
                        const uint32_t bh = x_hi0 + off, bl = x_lo0 + off;
                        const uint32_t a = tmem_u + (uint32_t)(k * 32);
                        if (i == 0) umma_ts<0>(d, a, bh, idesc); else umma_ts<1>(d, a, bh, idesc);
                        umma_ts<1>(d, a, bl, idesc);
#pragma unroll
                        for (int kk = 1; kk < 4; ++kk) {                   // 16 K elements = 8 TMEM columns of A = 32 bytes of the B rows
                            umma_ts<1>(d, a + 8 * kk, bh + 2 * kk, idesc);
                            umma_ts<1>(d, a + 8 * kk, bl + 2 * kk, idesc);
                        }
                        if (i == taps - 1) tc::tc_commit(&acc_full[u]);
                    }
                    if (dbg && leader && i == taps - 1) dbg[4 + u] = clock64();
                }
                ++k; ++kw;
                if (kw == p.KW) { kw = 0; ++kh; }
                if (k == taps) { k = 0; kh = 0; kw = 0; }
            }
        }
        // ---- the remaining tiles, one after the other (the epilogue of tile t - 2 has freed the accumulator)
        for (int t = first; t < nt; ++t) {
            const int buf = t & 1;
            mbar_wait_safe(&acc_empty[buf], ((t >> 1) & 1) ^ 1);
            int need = (t * N + (p.KH - 1) * p.Wp + p.KW - 1 + N - 1) / p.box_rows;      // last halo box this tile reads
            need = need >= nb ? nb - 1 : need;
            while (boxes_ready <= need) {
                mbar_wait_safe(&box_bar[boxes_ready], 0);
                ++boxes_ready;
            }
            tc::tc_fence_after();
            const uint32_t d = tmem_u + (uint32_t)(w_cols + buf * N);
            if (TAPS > 0) {
                // straight-line issue: 8 TAPS MMAs with precomputed operand offsets.  (The looped form below costs ~150-190 cycles of
                // address arithmetic and branches per tap and the tensor pipe does not queue deep enough to hide it: 66 instead of
                // 62 cycles per MMA, profiles/r02/ws_timeline_history.md.)
                if (leader) {
                    const uint32_t base_h = x_hi0 + (uint32_t)(t * N) * (ROW_BYTES >> 4), base_l = x_lo0 + (uint32_t)(t * N) * (ROW_BYTES >> 4);
#pragma unroll
                    for (int i = 0; i < (TAPS > 0 ? TAPS : 1); ++i) {
                        const uint32_t bh = base_h + tap_off[i], bl = base_l + tap_off[i], a = tmem_u + tap_col[i];
                        if (i == 0) umma_ts<0>(d, a, bh, idesc); else umma_ts<1>(d, a, bh, idesc);
                        umma_ts<1>(d, a, bl, idesc);
#pragma unroll
                        for (int kk = 1; kk < 4; ++kk) {
                            umma_ts<1>(d, a + 8 * kk, bh + 2 * kk, idesc);
                            umma_ts<1>(d, a + 8 * kk, bl + 2 * kk, idesc);
                        }
                    }
                }
            } else {
                int k = rot, kh = kh0, kw = kw0;
                for (int i = 0; i < taps; ++i) {
                    if (leader) {
                        const uint32_t off = (uint32_t)(t * N + kh * p.Wp + kw) * (ROW_BYTES >> 4);
                        const uint32_t bh = x_hi0 + off, bl = x_lo0 + off;
                        const uint32_t a = tmem_u + (uint32_t)(k * 32);
                        if (i == 0) umma_ts<0>(d, a, bh, idesc); else umma_ts<1>(d, a, bh, idesc);
                        umma_ts<1>(d, a, bl, idesc);
#pragma unroll
                        for (int kk = 1; kk < 4; ++kk) {
                            umma_ts<1>(d, a + 8 * kk, bh + 2 * kk, idesc);
                            umma_ts<1>(d, a + 8 * kk, bl + 2 * kk, idesc);
                        }
                    }
                    ++k; ++kw;
                    if (kw == p.KW) { kw = 0; ++kh; }
                    if (k == taps) { k = 0; kh = 0; kw = 0; }
                }
            }
            if (leader) tc::tc_commit(&acc_full[buf]);
            if (dbg && leader && t < 4) dbg[4 + t] = clock64();
        }
    } else {
        // ================= warps 2..9: weights -> TMEM, then the epilogue =================
        const int qq = warp & 3;                       // TMEM lane quarter this warp may access: lanes 32 qq ..
        const int half = (warp - 2) >> 2;              // which half of a tile's pixel columns / of a tap's weight columns
        const bool is_lo = qq >= 2;                    // lanes 64..127 hold the W_lo rows / their partial sums
        const int co = (qq & 1) * 32 + lane;           // output channel of this thread
        const int et = threadIdx.x - 64;               // 0..255 among the epilogue threads
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qq * 32) << 16);
        for (int i = 0, k = rot; i < taps; ++i, k = (k + 1 == taps ? 0 : k + 1)) {
            // one tap = [64 co][64 ci] bf16 per plane in the ring (SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7)); this
            // thread moves 64 bytes (32 K elements = 16 TMEM columns) of its weight row
            const int s = i % p.w_stages;
            mbar_wait_safe(&w_full[s], (i / p.w_stages) & 1);
            const uint8_t* row = stg + (size_t)s * 2 * 64 * ROW_BYTES + (is_lo ? 64 * ROW_BYTES : 0) + co * ROW_BYTES;
            uint4 w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const uint4*>(row + (((4 * half + j) ^ (co & 7)) << 4));
            tmem_st_32x16(lane_addr + (uint32_t)(k * 32 + 16 * half), reinterpret_cast<const uint32_t*>(w));
            tmem_st_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(&w_empty[s]); tc::mbar_arrive(&w_tap[k]); }
        }
        named_bar_sync(1, EPI_WARPS * 32);            // every warp is done with the ring: the area becomes the output staging tile
        float* T_hi = reinterpret_cast<float*>(stg);
        float* T_lo = T_hi + N * 64;
        const float slope = (p.act == TPGSR_ACT_PRELU && p.act_param) ? __ldg(p.act_param) : 0.f;
        const float bias = (p.bias && !is_lo) ? __ldg(p.bias + co) : 0.f;
        const unsigned img = (unsigned)(p.Hp * p.Wp);
        const unsigned qtot = (unsigned)p.Qtot;
        float* T_mine = (is_lo ? T_lo : T_hi) + (half * NH) * 64 + co;
        for (int t = 0; t < nt; ++t) {
            const int buf = t & 1;
            mbar_wait_safe(&acc_full[buf], (t >> 1) & 1);
            tc::tc_fence_after();
            if (dbg && warp == 2 && lane == 0 && t < 4) dbg[8 + t] = clock64();
            float v[NH];
            const uint32_t taddr = lane_addr + (uint32_t)(w_cols + buf * N + half * NH);
#pragma unroll
            for (int c = 0; c < NH / 16; ++c) tmem_ld_32x16(taddr + 16 * c, v + 16 * c);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);           // the accumulator is in registers: the MMA warp may overwrite it
            if (dbg && warp == 2 && lane == 0 && t < 4) dbg[12 + t] = clock64();
            if (t > 0) named_bar_sync(1, EPI_WARPS * 32);               // the previous tile has left the staging tile
            // transposed store: thread = channel, one row of T per pixel (32 lanes -> 32 consecutive floats: conflict-free)
#pragma unroll
            for (int j = 0; j < NH; ++j) T_mine[j * 64] = v[j] + bias;
            if (et < N) {                                               // dense pixel index of tile row et (-1 = pad position / beyond the last image)
                const unsigned qrow = (unsigned)q0 + (unsigned)(t * N + et);
                const unsigned ub = qrow / img, rem = qrow - ub * img;
                const unsigned ur = rem / (unsigned)p.Wp, uc = rem - ur * (unsigned)p.Wp;
                pixtab[et] = (qrow < qtot && (int)ur < p.Ho && (int)uc < p.Wo) ? ((int)ub * p.Ho + (int)ur) * p.Wo + (int)uc : -1;
            }
            named_bar_sync(1, EPI_WARPS * 32);
            if (dbg && warp == 2 && lane == 0 && t < 4) dbg[16 + t] = clock64();
            // out = act(T_hi + T_lo): 16 float4 per pixel row, a thread takes every 256th float4 (coalesced 512 bytes per warp)
#pragma unroll
            for (int i = 0; i < N * 16 / (EPI_WARPS * 32); ++i) {
                const int f = et + i * (EPI_WARPS * 32);
                const int px = pixtab[f >> 4];
                const float4 a4 = *reinterpret_cast<const float4*>(T_hi + 4 * f);
                const float4 b4 = *reinterpret_cast<const float4*>(T_lo + 4 * f);
                if (px < 0) continue;
                float4 val = make_float4(a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w);
                if (p.act != TPGSR_ACT_NONE) {
                    val.x = apply_act(val.x, p.act, slope); val.y = apply_act(val.y, p.act, slope);
                    val.z = apply_act(val.z, p.act, slope); val.w = apply_act(val.w, p.act, slope);
                }
                const int c4 = (f & 15) * 4;
                float4* dst = reinterpret_cast<float4*>(p.y + (size_t)px * p.ldy + c4);
                if (p.accumulate) {
                    const float4 o = *dst;
                    val.x += o.x; val.y += o.y; val.z += o.z; val.w += o.w;
                }
                *dst = val;
                if (p.yp_hi) {
                    const float fv[4] = {val.x, val.y, val.z, val.w};
                    __nv_bfloat16 h[4], l[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        h[u] = __float2bfloat16_rn(fv[u]);
                        l[u] = __float2bfloat16_rn(fv[u] - __bfloat162float(h[u]));
                    }
                    const size_t po = (size_t)px * p.Co + c4;
                    *reinterpret_cast<uint2*>(p.yp_hi + po) = *reinterpret_cast<uint2*>(h);
                    if (p.yp_lo) *reinterpret_cast<uint2*>(p.yp_lo + po) = *reinterpret_cast<uint2*>(l);
                }
            }
            if (dbg && warp == 2 && lane == 0 && t < 4) dbg[20 + t] = clock64();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[24] = clock64();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// fp32 dense NHWC [B,H,W,C] (pixel pitch ldx) -> padded-flat bf16 hi/lo planes [B][H+P][W+P][Cout] (pads and channels >= C zero)
// IDX = unsigned whenever the element count allows: the three divisions per thread were 64-bit (round 1) and cost more than the 26 MB of traffic
template <typename IDX>
__global__ void __launch_bounds__(256)
split_bf16_flat_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int B,
                       int H, int W, int C, int Cout, int P) {
    const IDX Hp = H + P, Wp = W + P;
    const IDX cq = Cout >> 2;
    const IDX total = (IDX)B * Hp * Wp * cq;
    const bool vec = ((ldx & 3) == 0) && ((C & 3) == 0);
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const IDX q = i / cq;
        const int c = (int)(i - q * cq) * 4;
        const IDX t = q / Wp;
        const int col = (int)(q - t * Wp);
        const IDX bb = t / Hp;
        const int r = (int)(t - bb * Hp);
        const int b = (int)bb;
        float f[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < H && col < W) {
            const float* src = x + ((size_t)(b * H + r) * W + col) * ldx + c;
            if (vec && c + 3 < C) {
                float4 v = *reinterpret_cast<const float4*>(src);
                f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c + j < C) f[j] = src[j];
            }
        }
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2bfloat16_rn(f[j]);
            l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
        }
        *reinterpret_cast<uint2*>(hi + (size_t)q * Cout + c) = *reinterpret_cast<uint2*>(h);
        if (lo) *reinterpret_cast<uint2*>(lo + (size_t)q * Cout + c) = *reinterpret_cast<uint2*>(l);
    }
}

struct FlatPlan {
    int MT, n_tile, n_tiles, tm_stride, nstack, pointwise, G, kgroups, n_groups, a_stages, box_rows, n_box, tb, b_stages, tile_outer, ring_bytes, stg_off, bar_off;
    int tall, tcols, n_hg, n_wg;
    size_t smem;
    int grid_x;
};

long long* g_flat_dbg = nullptr;
int g_flat_dbg_skip = 0;
int g_flat_no_nstack = 0;
int g_flat_no_tall = 0;     // developer switch (debug bit 3): disable the tall plan (A/B measurements)

// budget: dynamic shared memory a CTA may use; tmem_max: TMEM columns it may allocate; slots: CTAs resident on the GPU at once
int plan_flat_for(int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int nprod, size_t budget, int tmem_max, int slots,
                  FlatPlan& pl) {
    if (Co < 1 || B < 1 || H < 1 || W < 1 || Ci < 64 || Ci % 64 != 0) return TPGSR_ERR_UNSUPPORTED;
    const int pointwise = (KH == 1 && KW == 1) ? 1 : 0;
    const int cbs = Ci / 64;                                      // 64-channel blocks of the input
    // any padding 0 <= ph < KH, 0 <= pw < KW that the planes' shared zero rows / columns cover ('same': 2 ph = KH - 1; 'valid': ph = 0;
    // 'full' = the data gradient of a valid conv: ph = KH - 1); the output map must fit the plane grid
    if (ph < 0 || ph >= KH || pw < 0 || pw >= KW || P < ph || P < pw) return TPGSR_ERR_UNSUPPORTED;
    if (H + 2 * ph - KH + 1 < 1 || W + 2 * pw - KW + 1 < 1 || H + 2 * ph - KH + 1 > H + P || W + 2 * pw - KW + 1 > W + P) return TPGSR_ERR_UNSUPPORTED;
    const int a_planes = nprod == 3 ? 2 : 1, b_planes = nprod >= 2 ? 2 : 1;
    const int Wp = W + P;
    const int co16 = (Co + 15) / 16 * 16;
    pl.n_tiles = (co16 + 127) / 128;
    pl.n_tile = ((co16 + pl.n_tiles - 1) / pl.n_tiles + 15) / 16 * 16;
    // stacked weight planes (see mma_tap_stacked): bf16x3, one MMA of N = 2*n_tile <= 256; forces one tap per B stage (hi and lo tiles adjacent)
    pl.nstack = (nprod >= 2 && pl.n_tile <= 128) ? 1 : 0;
    if (g_flat_no_nstack) pl.nstack = 0;
    pl.tm_stride = ((pl.nstack ? 2 : 1) * pl.n_tile + 31) / 32 * 32;
    pl.pointwise = pointwise;
    const long long Qtot = (long long)B * (H + P) * Wp;
    if (Qtot + 4096 >= (1LL << 31)) return TPGSR_ERR_UNSUPPORTED;
    pl.tall = 0; pl.tcols = 0; pl.n_hg = 0; pl.n_wg = 0;
    // ---- TALL plan: KH x 1 kernels with KH >= 5 on 64 channels (the kw-stacked 9x9 convs, 9 x 1 after stacking).  In the flat plans a tap
    // shifts the window by a whole image row (W + P >= 68 flat rows), more than a 128-row tile: nothing is shared between the taps and
    // every pixel row is fetched KH times (525 MB through L2 for the final conv: 150 us).  Here a tile is 16 image rows x 8 columns, a CTA
    // owns MT such tiles side by side, and ONE 4-D TMA box [16 + KH - 1][8 MT][64] is the halo of all of them: 1.5x traffic instead of 9x.
    if (!g_flat_no_tall && KW == 1 && KH >= 5 && cbs == 1 && pw == 0) {
        const int Ho = H + 2 * ph - KH + 1, halo_rows = 16 + KH - 1;
        for (int MT = 2; MT >= 1; --MT) {
            if (MT * pl.tm_stride > tmem_max) continue;
            const int tcols = 8 * MT;
            const size_t a_b = (size_t)a_planes * halo_rows * tcols * ROW_BYTES;
            const size_t tap_bytes = (size_t)b_planes * pl.n_tile * ROW_BYTES;
            for (int want = 4; want >= 2; --want) {
                const int bs = want < KH ? want : KH;
                size_t ring = a_b + (size_t)bs * tap_bytes;
                if (ring < (size_t)STG_BYTES) ring = STG_BYTES;
                const size_t need = ring + 256 + (size_t)(pl.n_tile + 32) * 4 + 1024;
                if (need > budget) continue;
                pl.tall = 1; pl.tcols = tcols; pl.n_hg = (Ho + 15) / 16; pl.n_wg = (W + tcols - 1) / tcols;
                pl.MT = MT; pl.G = KH; pl.kgroups = 1; pl.n_groups = 1; pl.a_stages = 1; pl.box_rows = halo_rows * tcols; pl.n_box = 1;
                pl.tb = 1; pl.b_stages = bs; pl.smem = need; pl.ring_bytes = (int)ring; pl.stg_off = 0; pl.bar_off = (int)ring; pl.tile_outer = 0;
                pl.grid_x = B * pl.n_hg * pl.n_wg;
                return TPGSR_OK;
            }
        }
    }
    // candidate configurations, best first: whole kernel resident (G = KH) with MT = 3, 2, 1; then one kernel row per stage
    // among the feasible MT the one with the fewest (waves x tiles per CTA) wins; near-ties go to the larger MT (less weight traffic)
    const long long tiles = (Qtot + TILE_M - 1) / TILE_M;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && (KH == 1 || pointwise)) break;
        long long best_cost = -1;
        for (int MT = 3; MT >= 1; --MT) {
            if (MT * pl.tm_stride > tmem_max) continue;
            const int G = pointwise ? 1 : (pass == 0 ? KH : 1);
            const int kgroups = KH / G;
            const int n_groups = cbs * kgroups;
            const int rows = TILE_M * MT + (G - 1) * Wp + (KW - 1);
            int n_box = (rows + 255) / 256;
            int box_rows = ((rows + n_box - 1) / n_box + 7) / 8 * 8;
            if (box_rows > 256) { ++n_box; box_rows = ((rows + n_box - 1) / n_box + 7) / 8 * 8; }
            // streamed A groups (one kernel row / channel block per stage): up to 4 stages in flight — with 2 the 9x1 stacked convs waited
            // one L2 round trip per kernel row (profiles/r01: 154 us for the final conv)
            int a_stages = n_groups > 1 ? (n_groups < 4 ? n_groups : 4) : 1;
            const size_t a_stage_b = (size_t)a_planes * n_box * box_rows * ROW_BYTES;
            const size_t b_min = (size_t)2 * b_planes * pl.n_tile * ROW_BYTES * ((n_groups > 1 && KW > 1) ? KW : 1);
            while (a_stages > 2 && a_stages * a_stage_b + b_min + 4096 > budget) --a_stages;
            const size_t a_bytes = (size_t)a_stages * a_stage_b;
            // B ring: one tap per stage, 4..2 deep (deepest that fits) — or one kernel row of taps per stage, 2 deep, when the
            // tap tile is small and the kernel rows are streamed.  The epilogue staging area aliases the (drained) rings.
            const int taps_per_group = G * KW;
            const int tile_outer = 0;
            const size_t tap_bytes = (size_t)b_planes * pl.n_tile * ROW_BYTES;
            int tb = 1, b_stages = 0, ring_bytes = 0, stg_off = 0, bar_off = 0;
            size_t smem = 0;
            const bool row_stage = !tile_outer && n_groups > 1 && KW > 1 && tap_bytes * KW <= 40 * 1024 && !pl.nstack;
            for (int want = row_stage ? 2 : 4; want >= 2; --want) {
                tb = row_stage ? KW : 1;
                int bs = want;
                const int total_b = (tile_outer ? MT : n_groups) * (taps_per_group / tb);
                if (bs > total_b) bs = total_b;
                size_t ring = a_bytes + (size_t)bs * tb * tap_bytes;
                if (!tile_outer && ring < (size_t)STG_BYTES) ring = STG_BYTES;
                const size_t so = tile_outer ? ring : 0;
                const size_t bo = tile_outer ? ring + STG_BYTES : ring;
                const size_t need = bo + 256 /*barriers*/ + (size_t)(pl.n_tile + 32) * 4 /*bias*/ + 1024 /*align*/;
                if (need <= budget) { b_stages = bs; smem = need; ring_bytes = (int)ring; stg_off = (int)so; bar_off = (int)bo; break; }
            }
            if (b_stages == 0) continue;
            const long long ctas = (tiles + MT - 1) / MT * pl.n_tiles;
            const long long cost = ((ctas + slots - 1) / slots) * MT;
            if (best_cost >= 0 && cost * 23 >= best_cost * 20) continue;       // a smaller MT must win by > 15%
            best_cost = cost;
            pl.MT = MT; pl.G = G; pl.kgroups = kgroups; pl.n_groups = n_groups; pl.a_stages = a_stages; pl.box_rows = box_rows; pl.n_box = n_box;
            pl.tb = tb; pl.b_stages = b_stages; pl.smem = smem; pl.ring_bytes = ring_bytes; pl.stg_off = stg_off; pl.bar_off = bar_off;
            pl.tile_outer = tile_outer;
            pl.grid_x = (int)((tiles + MT - 1) / MT);
        }
        if (best_cost >= 0) return TPGSR_OK;
    }
    return TPGSR_ERR_UNSUPPORTED;
}

// One CTA per SM with up to 226 KB of the 227 KB a CTA may use and all 512 TMEM columns.  (A two-CTAs-per-SM tiling — 112 KB, 256 columns,
// one tile per CTA so that one CTA's loads / epilogue overlap the other's MMAs — was measured in round 2 and removed: 30.7 us against
// 26.4 us for the 3x3 64->64 conv, step time unchanged, profiles/r02/call1_stdout.log.)
int plan_flat(int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int nprod, FlatPlan& pl) {
    return plan_flat_for(B, H, W, P, Ci, Co, KH, KW, ph, pw, nprod, 226 * 1024, 512, 148, pl);
}

struct WsPlan {
    int N, NT, box_rows, n_box, w_stages, stg_off, bar_off, grid;
    size_t smem;
};
int g_flat_no_ws = 0;       // developer switch (debug bit 4): disable the weights-stationary plan (A/B measurements, cross-check tests)

// weights-stationary plan (conv_ws_kernel): split-bf16, exactly 64 -> 64 channels, 1..12 taps, NHWC output
int plan_ws(int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int nprod, size_t budget, int slots, WsPlan& pl) {
    if (g_flat_no_ws || nprod != 3 || Ci != 64 || Co != 64 || B < 1 || H < 1 || W < 1) return TPGSR_ERR_UNSUPPORTED;
    const int taps = KH * KW;
    if (taps < 1 || taps > 12) return TPGSR_ERR_UNSUPPORTED;      // (1x1: no halo; the kernel is then its double-buffered epilogue)
    if (KW == 1 && KH >= 5) return TPGSR_ERR_UNSUPPORTED;      // tall kernels: the taps share nothing in the flat row order — the tall plan's 2-D tile is the right shape
    if (ph < 0 || ph >= KH || pw < 0 || pw >= KW || P < ph || P < pw) return TPGSR_ERR_UNSUPPORTED;
    if (H + 2 * ph - KH + 1 < 1 || W + 2 * pw - KW + 1 < 1 || H + 2 * ph - KH + 1 > H + P || W + 2 * pw - KW + 1 > W + P) return TPGSR_ERR_UNSUPPORTED;
    const int Wp = W + P;
    const long long Qtot = (long long)B * (H + P) * Wp;
    if (Qtot + 4096 >= (1LL << 31)) return TPGSR_ERR_UNSUPPORTED;
    long long best = -1;
    const int cand[2] = {96, 64};
    for (int ci = 0; ci < 2; ++ci) {
        const int N = cand[ci];
        if (taps * 32 + 2 * N > 512) continue;
        const long long target = (Qtot + slots - 1) / slots;          // rows per CTA if the whole job were one wave
        int NT = (int)((target + N - 1) / N);
        const size_t tap_bytes = (size_t)2 * 64 * ROW_BYTES;
        size_t stg = (size_t)2 * N * 64 * 4;                             // output staging T_hi, T_lo [N][64] floats ...
        if (stg < 3 * tap_bytes) stg = 3 * tap_bytes;                    // ... aliasing the weight ring of the prologue (at least 3 taps deep)
        int n_box = 0;
        size_t smem = 0;
        for (; NT >= 1; --NT) {
            const int rows = NT * N + (KH - 1) * Wp + KW - 1;
            n_box = (rows + N / 2 - 1) / (N / 2);
            smem = (size_t)2 * n_box * (N / 2) * ROW_BYTES + stg + 1024 /*barriers + pixel table*/ + 1024 /*align*/;
            if (n_box <= WS_MAXBOX && smem <= budget) break;
        }
        if (NT < 1) continue;
        const long long grid = (Qtot + (long long)NT * N - 1) / ((long long)NT * N);
        const long long cost = ((grid + slots - 1) / slots) * NT * N;
        if (best >= 0 && cost >= best) continue;
        best = cost;
        // the weight ring takes the shared memory the halo leaves (up to one stage per tap)
        int w_stages = (int)((budget - (smem - stg)) / tap_bytes);
        w_stages = w_stages > taps ? taps : w_stages;
        w_stages = w_stages > WS_WSTAGES ? WS_WSTAGES : w_stages;
        if ((size_t)w_stages * tap_bytes > stg) { smem += (size_t)w_stages * tap_bytes - stg; stg = (size_t)w_stages * tap_bytes; }
        pl.N = N; pl.NT = NT; pl.box_rows = N / 2; pl.n_box = n_box; pl.grid = (int)grid; pl.w_stages = w_stages;
        pl.stg_off = 2 * n_box * (N / 2) * ROW_BYTES; pl.bar_off = pl.stg_off + (int)stg; pl.smem = smem;
    }
    return best >= 0 ? TPGSR_OK : TPGSR_ERR_UNSUPPORTED;
}

// ============================================================================================================ weight gradient
constexpr int WG_KB = 128;            // flat rows (pixels) per K block

struct FlatWgradParams {
    int Wp, Co, Ci, KH, KW, ph, pw;
    int cbs;                         // 64-channel blocks of the input: blockIdx.y = (kh * kw_groups + kwg) * cbs + cb
    int kw_groups, taps_per_group;   // the KW taps of a kernel row are split into groups that fit TMEM
    int n_tile;                      // co per CTA (multiple of 64)
    int x_rows;                      // rows of the X window per K block (WG_KB + taps_per_group - 1, rounded up to 8)
    int total_kb, kb_per_split;
    int stages;
    float* ws;                       // [splits][K = KH*KW*Ci][Co]
};

template <int NPROD>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_flat_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                       const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo, FlatWgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // NPROD = 2 (experimental): dY as ONE bf16 plane against X hi + lo — X_hi dY + X_lo dY
    constexpr int dy_planes = NPROD == 3 ? 2 : 1, x_planes = NPROD >= 2 ? 2 : 1;
    const int cb = blockIdx.y % p.cbs, khw = blockIdx.y / p.cbs;
    const int kh = khw / p.kw_groups, kwg = khw % p.kw_groups;
    const int kw0 = kwg * p.taps_per_group;
    const int ntaps = (p.KW - kw0) < p.taps_per_group ? (p.KW - kw0) : p.taps_per_group;
    const int n0 = blockIdx.z * p.n_tile;
    const int ngroups = p.n_tile / 64;
    const int dy_tile_bytes = WG_KB * ROW_BYTES;                  // one 64-channel group of dY rows
    const int dy_plane_bytes = ngroups * dy_tile_bytes;
    const int x_plane_bytes = p.x_rows * ROW_BYTES;
    const int stage_bytes = dy_planes == x_planes ? dy_planes * (dy_plane_bytes + x_plane_bytes) : dy_planes * dy_plane_bytes + x_planes * x_plane_bytes;
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared-space provenance (LDS/STS, not generic LD/ST)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tmem_full = empty + p.stages;
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
        if (NPROD >= 2) tc::prefetch_tmap(&tmX_lo);
        if (NPROD == 3) tc::prefetch_tmap(&tmD_lo);
        for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
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

    // stage layout: [dY hi groups][dY lo groups][X hi window][X lo window]
    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % p.stages;
                mbar_wait_safe(&empty[s], ((i / p.stages) & 1) ^ 1);
                const int qk = (kb_begin + i) * WG_KB;
                uint8_t* st = smem + (size_t)s * stage_bytes;
                tc::mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                for (int gI = 0; gI < ngroups; ++gI) {
                    tc::tma_load_2d(st + gI * dy_tile_bytes, &tmD_hi, &full[s], n0 + gI * 64, qk);
                    if (NPROD == 3) tc::tma_load_2d(st + dy_plane_bytes + gI * dy_tile_bytes, &tmD_lo, &full[s], n0 + gI * 64, qk);
                }
                uint8_t* xs = st + dy_planes * dy_plane_bytes;
                const int xr = qk + (kh - p.ph) * p.Wp + (kw0 - p.pw);
                tc::tma_load_2d(xs, &tmX_hi, &full[s], cb * 64, xr);
                if (NPROD >= 2) tc::tma_load_2d(xs + x_plane_bytes, &tmX_lo, &full[s], cb * 64, xr);
            }
        }
    } else if (warp == 1) {
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);      // same value, provably warp-uniform
        const bool leader = elect_one();
        {
            // NPROD = 3: the X_hi and X_lo windows are ONE M = 128 operand — MN-major, the two 64-channel chunks x_plane_bytes apart (LBO) —
            // so that [X_hi; X_lo]^T dY_hi and [X_hi; X_lo]^T dY_lo are two full-rate MMAs per K step (all four partial products) instead of
            // three half-rate M = 64 ones: 64 instead of 96 tensor cycles.  Lanes 0..63 of the accumulator hold the X_hi rows, 64..127 the X_lo rows.
            constexpr bool STK = NPROD == 3;
            const uint32_t idesc = tc::idesc_bf16(STK ? 128 : 64, p.n_tile, 1, 1);      // both operands MN-major
            for (int i = 0; i < nkb; ++i) {
                const int s = i % p.stages;
                mbar_wait_safe(&full[s], (i / p.stages) & 1);
                tc::tc_fence_after();
                const uint32_t st = tc::smem_u32(smem + (size_t)s * stage_bytes);
                // MN-major SWIZZLE_128B: LBO = bytes between 64-element M/N chunks (X hi / lo windows for the stacked A; dY groups for B), SBO = 1024
                const uint32_t dy_hi = desc_lo(st, dy_tile_bytes), dy_lo = dy_hi + (dy_plane_bytes >> 4);
                const uint32_t x_hi = desc_lo(st + dy_planes * dy_plane_bytes, STK ? x_plane_bytes : 1024), x_lo = x_hi + (x_plane_bytes >> 4);
                for (int t = 0; t < ntaps; ++t) {
                    if (!leader) continue;
                    const uint32_t d = tmem_u + (uint32_t)(t * p.n_tile);
                    const uint32_t xh = x_hi + (uint32_t)t * (ROW_BYTES >> 4), xl = x_lo + (uint32_t)t * (ROW_BYTES >> 4);   // tap t = window shifted by t rows
                    if (i == 0) {
                        umma_lo<0>(d, xh, dy_hi, idesc);
                        if (NPROD == 3) umma_lo<1>(d, xh, dy_lo, idesc);
                        if (NPROD == 2) umma_lo<1>(d, xl, dy_hi, idesc);
#pragma unroll
                        for (int k = 1; k < WG_KB / 16; ++k) {      // 16 K rows = 2048 bytes = 128 descriptor units
                            umma_lo<1>(d, xh + 128 * k, dy_hi + 128 * k, idesc);
                            if (NPROD == 3) umma_lo<1>(d, xh + 128 * k, dy_lo + 128 * k, idesc);
                            if (NPROD == 2) umma_lo<1>(d, xl + 128 * k, dy_hi + 128 * k, idesc);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < WG_KB / 16; ++k) {
                            umma_lo<1>(d, xh + 128 * k, dy_hi + 128 * k, idesc);
                            if (NPROD == 3) umma_lo<1>(d, xh + 128 * k, dy_lo + 128 * k, idesc);
                            if (NPROD == 2) umma_lo<1>(d, xl + 128 * k, dy_hi + 128 * k, idesc);
                        }
                    }
                }
                if (leader) tc::tc_commit(&empty[s]);
            }
            if (leader) tc::tc_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        mbar_wait_safe(tmem_full, 0);
        tc::tc_fence_after();
        const int K = p.KH * p.KW * p.Ci;
        if (NPROD == 3) {
            // M = 128 accumulators: lane = row; rows 0..63 = X_hi channels, 64..127 = X_lo channels of the same 64 input channels.  Warps q = 2, 3
            // hand their 32 x 32 blocks over through shared memory (the operand ring is drained), warps q = 0, 1 add theirs and store.
            float* scratch = reinterpret_cast<float*>(smem);            // [64 rows][33]
            const int ci3 = cb * 64 + (q & 1) * 32 + lane;
            for (int t = 0; t < ntaps; ++t) {
                const int tap = kh * p.KW + kw0 + t;
                for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
                    float v[32];
                    tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.n_tile + c0), v);
                    tc::tmem_ld_wait();
                    if (q >= 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) scratch[((q - 2) * 32 + lane) * 33 + j] = v[j];
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (q < 2) {
                        float* dst = p.ws + ((size_t)blockIdx.x * K + (size_t)tap * p.Ci + ci3) * p.Co + n0 + c0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + c0 + j < p.Co) dst[j] = nkb > 0 ? v[j] + scratch[(q * 32 + lane) * 33 + j] : 0.f;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
        } else {
        // M = 64 accumulators use the half-subpartition layout: row ci = 16*q + l lives in TMEM lane 32*q + l (l < 16)
        const int ci = cb * 64 + 16 * q + lane;
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
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, tmem_cols);
    }
}

struct FlatWgradPlan {
    int n_tile, n_tiles, cbs, kw_groups, taps_per_group, x_rows, stages, stage_bytes, total_kb, kb_per_split, splits;
};

int plan_wgrad_flat(int B, int H, int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int nprod, FlatWgradPlan& pl) {
    if (Ci < 64 || Ci % 64 != 0 || Cpad % 64 != 0 || Cpad < Co) return TPGSR_ERR_UNSUPPORTED;
    const int cbs = Ci / 64;
    const int dy_planes = nprod == 3 ? 2 : 1, x_planes = nprod >= 2 ? 2 : 1;
    // co tile: the candidate (multiple of 64 dividing Cpad, <= 256) with the fewest CTAs per pixel split whose ring is >= 2 deep
    int best = -1, best_cost = 1 << 30;
    for (int nt = 64; nt <= 256 && nt <= Cpad; nt += 64) {
        if (Cpad % nt != 0) continue;
        int max_taps = 512 / nt;
        int kwg = (KW + max_taps - 1) / max_taps;
        int tpg = (KW + kwg - 1) / kwg;
        int xr = (WG_KB + tpg - 1 + 7) / 8 * 8;
        int sb = (dy_planes * (nt / 64) * WG_KB + x_planes * xr) * ROW_BYTES;
        if ((212 * 1024) / sb < 2) continue;
        int cost = kwg * (Cpad / nt);
        if (cost < best_cost || (cost == best_cost && nt < best)) { best = nt; best_cost = cost; }
    }
    if (best < 0) return TPGSR_ERR_UNSUPPORTED;
    pl.n_tile = best;
    pl.cbs = cbs;
    pl.n_tiles = Cpad / pl.n_tile;
    const int max_taps = 512 / pl.n_tile;
    pl.kw_groups = (KW + max_taps - 1) / max_taps;
    pl.taps_per_group = (KW + pl.kw_groups - 1) / pl.kw_groups;
    pl.x_rows = (WG_KB + pl.taps_per_group - 1 + 7) / 8 * 8;
    pl.stage_bytes = (dy_planes * (pl.n_tile / 64) * WG_KB + x_planes * pl.x_rows) * ROW_BYTES;
    pl.stages = (212 * 1024) / pl.stage_bytes;
    if (pl.stages > 4) pl.stages = 4;
    const long long Qtot = (long long)B * (H + P) * (W + P);
    if (Qtot + 4096 >= (1LL << 31)) return TPGSR_ERR_UNSUPPORTED;
    pl.total_kb = (int)((Qtot + WG_KB - 1) / WG_KB);
    const int ctas_per_split = KH * pl.kw_groups * pl.n_tiles * cbs;
    // pixel splits: the count with the shortest modelled critical path = rounds of CTAs (1 CTA per SM: the ring takes the whole shared
    // memory) x (K blocks per CTA + 2 blocks' worth of prologue / epilogue) x MMA time per K block, + the split reduction, which reads
    // splits x K x Co floats at ~3 TB/s; ties go to fewer splits
    int want = 1;
    double best_path = -1.0;
    const double t_kb = pl.taps_per_group * (WG_KB / 16) * nprod * (pl.n_tile / 2) / 1965.0;           // us: M = 64 MMAs take N / 2 cycles
    const double t_red = (double)KH * KW * Ci * Co * 4.0 / 3.0e6;                                       // us per split
    for (int s = 1; s <= 148 && s <= pl.total_kb; ++s) {
        const long long rounds = ((long long)ctas_per_split * s + 147) / 148;
        const double path = rounds * ((pl.total_kb + s - 1) / s + 2) * t_kb + s * t_red;
        if (best_path < 0 || path < best_path - 1e-9) { best_path = path; want = s; }
    }
    pl.kb_per_split = (pl.total_kb + want - 1) / want;
    pl.splits = (pl.total_kb + pl.kb_per_split - 1) / pl.kb_per_split;
    return TPGSR_OK;
}

}  // namespace

// defined in conv_simt.cu
extern "C" int tpgsr_conv2d_wgrad_reduce(const float* ws, float* dw_oihw, int splits, int K, int N, int Ci, int KH, int KW,
                                         int accumulate, cudaStream_t stream);

extern "C" {

int tpgsr_split_bf16_flat(const float* x, int ldx, void* hi, void* lo, int B, int H, int W, int C, int Cout, int P,
                          cudaStream_t stream) {
    if (!x || !hi || (Cout & 3) || Cout < C || B < 1 || H < 1 || W < 1 || P < 0) return TPGSR_ERR_BAD_ARG;
    long long blocks = ((long long)B * (H + P) * (W + P) * (Cout / 4) + 255) / 256;
    const int grid = (int)(blocks > 148 * 16 ? 148 * 16 : blocks);
    if ((long long)B * (H + P) * (W + P) * (Cout / 4) + (long long)grid * 256 < (1LL << 32))
        split_bf16_flat_kernel<unsigned><<<grid, 256, 0, stream>>>(x, ldx, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), B, H, W, C, Cout, P);
    else
        split_bf16_flat_kernel<long long><<<grid, 256, 0, stream>>>(x, ldx, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), B, H, W, C, Cout, P);
    return tpgsr_launch_status();
}

int tpgsr_conv2d_flat_supported(int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int nprod) {
    FlatPlan pl;
    return plan_flat(B, H, W, P, Ci, Co, KH, KW, ph, pw, nprod, pl) == TPGSR_OK ? 1 : 0;
}

// developer instrumentation: per-CTA clock64 timestamps of the next conv_flat launches are written to dbg[grid][16]
int tpgsr_conv2d_flat_set_debug(long long* dbg) { g_flat_dbg = dbg; return TPGSR_OK; }
// bits 0-1: skip phases (timeline experiments); bit 2: no N-stacking of the weight planes; bit 3: no tall plan; bit 4: no weights-stationary plan
int tpgsr_conv2d_flat_set_debug_skip(int bits) {
    g_flat_dbg_skip = bits & 3;
    g_flat_no_nstack = (bits >> 2) & 1;
    g_flat_no_tall = (bits >> 3) & 1;
    g_flat_no_ws = (bits >> 4) & 1;
    return TPGSR_OK;
}

// tiling decided for a geometry (host-only; lets tests / docs inspect the plan without a GPU):
// out[13] = {MT, n_tile, n_tiles, tm_stride, G, n_groups, a_stages, box_rows, n_box, tb, b_stages, grid_x, tile_outer}; returns smem bytes (0 = unsupported)
size_t tpgsr_conv2d_flat_plan(int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int nprod, int* out) {
    WsPlan ws;
    if (plan_ws(B, H, W, P, Ci, Co, KH, KW, ph, pw, nprod, 226 * 1024, 148, ws) == TPGSR_OK) {      // mode 3: {NT, 64, 1, N, KH, 1, 1, box_rows, n_box, 0, 0, grid, 3}
        if (out) {
            const int v[13] = {ws.NT, 64, 1, ws.N, KH, 1, 1, ws.box_rows, ws.n_box, 0, 0, ws.grid, 3};
            for (int i = 0; i < 13; ++i) out[i] = v[i];
        }
        return ws.smem;
    }
    FlatPlan pl;
    if (plan_flat(B, H, W, P, Ci, Co, KH, KW, ph, pw, nprod, pl)) return 0;
    if (out) {
        const int v[13] = {pl.MT, pl.n_tile, pl.n_tiles, pl.tm_stride, pl.G, pl.n_groups, pl.a_stages, pl.box_rows, pl.n_box, pl.tb,
                           pl.b_stages, pl.grid_x, pl.tall ? 2 : pl.tile_outer};
        for (int i = 0; i < 13; ++i) out[i] = v[i];
    }
    return pl.smem;
}

// out[9] = {n_tile, n_tiles, kw_groups, taps_per_group, x_rows, stages, total_kb, kb_per_split, splits}; returns stage bytes (0 = unsupported)
size_t tpgsr_conv2d_wgrad_flat_plan(int B, int H, int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int nprod, int* out) {
    FlatWgradPlan pl;
    if (plan_wgrad_flat(B, H, W, P, Ci, Co, Cpad, KH, KW, nprod, pl)) return 0;
    if (out) {
        const int v[9] = {pl.n_tile, pl.n_tiles, pl.kw_groups, pl.taps_per_group, pl.x_rows, pl.stages, pl.total_kb, pl.kb_per_split,
                          pl.splits};
        for (int i = 0; i < 9; ++i) out[i] = v[i];
    }
    return (size_t)pl.stage_bytes;
}

// y = act(conv(x) + bias) for a stride-1 convolution with 64 k input channels and padding (ph, pw) <= P; y is [B, H + 2 ph - KH + 1, W + 2 pw - KW + 1, Co].  x_hi/x_lo: padded-flat bf16 planes of the input
// ([B][H+P][W+P][64], tpgsr_split_bf16_flat); w_hi/w_lo: K-major packed weights [Co][KH*KW*64] (tpgsr_pack_conv_weights_bf16);
// y: fp32 dense NHWC (pixel pitch ldy) or NCHW.  Data gradient: pass dY planes, the dgrad packing, Co := conv Ci.
int tpgsr_conv2d_flat_planes(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, float* y, void* yp_hi,
                             void* yp_lo, int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int ldy, int act,
                             const float* act_param, int out_mode, int accumulate, int nprod, cudaStream_t stream);

int tpgsr_conv2d_flat(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, float* y, int B,
                      int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int ldy, int act,
                      const float* act_param, int out_mode, int accumulate, int nprod, cudaStream_t stream) {
    return tpgsr_conv2d_flat_planes(x_hi, x_lo, w_hi, w_lo, bias, y, nullptr, nullptr, B, H, W, P, Ci, Co, KH, KW, ph, pw, ldy, act, act_param,
                                    out_mode, accumulate, nprod, stream);
}

// as tpgsr_conv2d_flat; additionally writes the output as dense bf16 hi/lo planes [B*H*W][Co] (yp_lo may be NULL in 1-product mode)
int tpgsr_conv2d_flat_planes(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, float* y, void* yp_hi,
                             void* yp_lo, int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int ldy, int act,
                             const float* act_param, int out_mode, int accumulate, int nprod, cudaStream_t stream) {
    if (!x_hi || !w_hi || !y || nprod < 1 || nprod > 3 || (nprod == 3 && !x_lo) || (nprod >= 2 && !w_lo)) return TPGSR_ERR_BAD_ARG;
    if (yp_hi && (Co % 32 != 0 || out_mode != TPGSR_OUT_NHWC || ((ldy > 0 ? ldy : Co) & 3) || (act != TPGSR_ACT_NONE && act != TPGSR_ACT_RELU)))
        return TPGSR_ERR_UNSUPPORTED;
    WsPlan ws;
    if (out_mode == TPGSR_OUT_NHWC && ((ldy > 0 ? ldy : Co) & 3) == 0 && plan_ws(B, H, W, P, Ci, Co, KH, KW, ph, pw, nprod, 226 * 1024, 148, ws) == TPGSR_OK) {
        WsParams q;
        q.Hp = H + P; q.Wp = W + P; q.Ho = H + 2 * ph - KH + 1; q.Wo = W + 2 * pw - KW + 1; q.Qtot = (long long)B * q.Hp * q.Wp;
        q.Co = Co; q.KH = KH; q.KW = KW; q.ph = ph; q.pw = pw; q.NT = ws.NT; q.box_rows = ws.box_rows; q.n_box = ws.n_box; q.w_stages = ws.w_stages; q.K = KH * KW * 64;
        q.stg_off = ws.stg_off; q.bar_off = ws.bar_off; q.ldy = ldy > 0 ? ldy : Co; q.act = act; q.accumulate = accumulate;
        q.bias = bias; q.act_param = act_param; q.y = y; q.yp_hi = static_cast<__nv_bfloat16*>(yp_hi); q.yp_lo = static_cast<__nv_bfloat16*>(yp_lo);
        q.w_hi = static_cast<const __nv_bfloat16*>(w_hi); q.w_lo = static_cast<const __nv_bfloat16*>(w_lo); q.dbg = g_flat_dbg;
        CUtensorMap tmX_hi, tmX_lo;
        uint64_t xd[2] = {64, (uint64_t)q.Qtot};
        uint64_t xs[1] = {128};
        uint32_t xb[2] = {64, (uint32_t)ws.box_rows};
        if (tpgsr_make_tmap_bf16(&tmX_hi, x_hi, 2, xd, xs, xb) || tpgsr_make_tmap_bf16(&tmX_lo, x_lo, 2, xd, xs, xb)) return TPGSR_ERR_DRIVER;
        CUtensorMap tmW_hi, tmW_lo;
        uint64_t wd[2] = {(uint64_t)q.K, 64};
        uint64_t wst[1] = {(uint64_t)q.K * 2};
        uint32_t wb[2] = {64, 64};
        if (tpgsr_make_tmap_bf16(&tmW_hi, w_hi, 2, wd, wst, wb) || tpgsr_make_tmap_bf16(&tmW_lo, w_lo, 2, wd, wst, wb)) return TPGSR_ERR_DRIVER;
        auto kern = ws.N == 96 ? (KH * KW == 9 ? conv_ws_kernel<96, 9> : conv_ws_kernel<96, 0>) : (KH * KW == 9 ? conv_ws_kernel<64, 9> : conv_ws_kernel<64, 0>);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws.smem) != cudaSuccess) return TPGSR_ERR_LAUNCH;
        kern<<<ws.grid, FWD_THREADS, ws.smem, stream>>>(tmX_hi, tmX_lo, tmW_hi, tmW_lo, q);
        return tpgsr_launch_status();
    }
    FlatPlan pl;
    int rc = plan_flat(B, H, W, P, Ci, Co, KH, KW, ph, pw, nprod, pl);
    if (rc) return rc;
    FlatConvParams p;
    p.B = B; p.H = H; p.W = W; p.Hp = H + P; p.Wp = W + P; p.Qtot = (long long)B * p.Hp * p.Wp;
    p.Ho = H + 2 * ph - KH + 1; p.Wo = W + 2 * pw - KW + 1;
    p.Co = Co; p.KH = KH; p.KW = KW; p.ph = ph; p.pw = pw;
    p.MT = pl.MT; p.n_tile = pl.n_tile; p.tm_stride = pl.tm_stride; p.Ci = Ci; p.G = pl.G; p.kgroups = pl.kgroups; p.n_groups = pl.n_groups;
    p.a_stages = pl.a_stages;
    p.box_rows = pl.box_rows; p.n_box = pl.n_box; p.tb = pl.tb; p.b_stages = pl.b_stages; p.ring_bytes = pl.ring_bytes;
    p.tile_outer = pl.tile_outer; p.stg_off = pl.stg_off; p.bar_off = pl.bar_off; p.nstack = pl.nstack;
    p.tall = pl.tall; p.tcols = pl.tcols; p.n_hg = pl.n_hg; p.n_wg = pl.n_wg;
    p.ldy = ldy > 0 ? ldy : Co; p.act = act; p.out_mode = out_mode; p.accumulate = accumulate;
    p.bias = bias; p.act_param = act_param; p.y = y; p.dbg = g_flat_dbg; p.dbg_skip = g_flat_dbg_skip;
    p.yp_hi = static_cast<__nv_bfloat16*>(yp_hi); p.yp_lo = static_cast<__nv_bfloat16*>(yp_lo);

    const uint64_t K = (uint64_t)KH * KW * Ci;
    CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
    uint64_t adims[2] = {(uint64_t)Ci, (uint64_t)p.Qtot};
    uint64_t astr[1] = {(uint64_t)Ci * 2};
    uint32_t abox[2] = {64, (uint32_t)pl.box_rows};
    uint64_t bdims[2] = {K, (uint64_t)Co};
    uint64_t bstr[1] = {K * 2};
    uint32_t bbox[2] = {64, (uint32_t)pl.n_tile};
    // tall plan: the planes as a 4-D tensor [B][H+P][W+P][Ci]; the box is the 2-D halo tile of MT x (16 x 8) outputs
    uint64_t tdims[4] = {(uint64_t)Ci, (uint64_t)p.Wp, (uint64_t)p.Hp, (uint64_t)B};
    uint64_t tstr[3] = {(uint64_t)Ci * 2, (uint64_t)p.Wp * Ci * 2, (uint64_t)p.Hp * p.Wp * Ci * 2};
    uint32_t tbox[4] = {64, (uint32_t)pl.tcols, (uint32_t)(16 + KH - 1), 1};
    if (pl.tall ? tpgsr_make_tmap_bf16(&tmA_hi, x_hi, 4, tdims, tstr, tbox) : tpgsr_make_tmap_bf16(&tmA_hi, x_hi, 2, adims, astr, abox)) return TPGSR_ERR_DRIVER;
    if (tpgsr_make_tmap_bf16(&tmB_hi, w_hi, 2, bdims, bstr, bbox)) return TPGSR_ERR_DRIVER;
    tmA_lo = tmA_hi;
    tmB_lo = tmB_hi;
    if (nprod == 3 && (pl.tall ? tpgsr_make_tmap_bf16(&tmA_lo, x_lo, 4, tdims, tstr, tbox) : tpgsr_make_tmap_bf16(&tmA_lo, x_lo, 2, adims, astr, abox)))
        return TPGSR_ERR_DRIVER;
    if (nprod >= 2 && tpgsr_make_tmap_bf16(&tmB_lo, w_lo, 2, bdims, bstr, bbox)) return TPGSR_ERR_DRIVER;
    dim3 grid(pl.grid_x, pl.n_tiles);
    if (nprod == 3) {
        if (cudaFuncSetAttribute(conv_flat_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        conv_flat_kernel<3><<<grid, FWD_THREADS, pl.smem, stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    } else if (nprod == 2) {
        if (cudaFuncSetAttribute(conv_flat_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        conv_flat_kernel<2><<<grid, FWD_THREADS, pl.smem, stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    } else {
        if (cudaFuncSetAttribute(conv_flat_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        conv_flat_kernel<1><<<grid, FWD_THREADS, pl.smem, stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    }
    return tpgsr_launch_status();
}

int tpgsr_conv2d_wgrad_flat_supported(int B, int H, int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int nprod) {
    FlatWgradPlan pl;
    return plan_wgrad_flat(B, H, W, P, Ci, Co, Cpad, KH, KW, nprod, pl) == TPGSR_OK ? 1 : 0;
}

size_t tpgsr_conv2d_wgrad_flat_workspace_bytes(int B, int H, int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int nprod) {
    FlatWgradPlan pl;
    if (plan_wgrad_flat(B, H, W, P, Ci, Co, Cpad, KH, KW, nprod, pl)) return 0;
    return (size_t)pl.splits * KH * KW * Ci * Co * sizeof(float);
}

// dW (OIHW) (+)= weight gradient of a 'same' conv.  x planes: padded-flat [B][H+P][W+P][64]; dy planes: padded-flat
// [B][H+P][W+P][Cpad] (channels >= Co zero), same P.
int tpgsr_conv2d_wgrad_flat(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw_oihw, int B, int H,
                            int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int ph, int pw, int accumulate, int nprod,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!x_hi || !dy_hi || !dw_oihw || !workspace || nprod < 1 || nprod > 3 || (nprod >= 2 && !x_lo) || (nprod == 3 && !dy_lo))
        return TPGSR_ERR_BAD_ARG;
    if (ph < 0 || ph >= KH || pw < 0 || pw >= KW || P < ph || P < pw) return TPGSR_ERR_UNSUPPORTED;
    FlatWgradPlan pl;
    int rc = plan_wgrad_flat(B, H, W, P, Ci, Co, Cpad, KH, KW, nprod, pl);
    if (rc) return rc;
    if (workspace_bytes < (size_t)pl.splits * KH * KW * Ci * Co * sizeof(float)) return TPGSR_ERR_WORKSPACE;
    const long long Qtot = (long long)B * (H + P) * (W + P);
    FlatWgradParams p;
    p.Wp = W + P; p.Co = Co; p.Ci = Ci; p.KH = KH; p.KW = KW; p.ph = ph; p.pw = pw;
    p.cbs = pl.cbs; p.kw_groups = pl.kw_groups; p.taps_per_group = pl.taps_per_group; p.n_tile = pl.n_tile; p.x_rows = pl.x_rows;
    p.total_kb = pl.total_kb; p.kb_per_split = pl.kb_per_split; p.stages = pl.stages;
    p.ws = static_cast<float*>(workspace);

    CUtensorMap tmX_hi, tmX_lo, tmD_hi, tmD_lo;
    uint64_t xd[2] = {(uint64_t)Ci, (uint64_t)Qtot};
    uint64_t xs[1] = {(uint64_t)Ci * 2};
    uint32_t xbox[2] = {64, (uint32_t)pl.x_rows};
    uint64_t dd[2] = {(uint64_t)Cpad, (uint64_t)Qtot};
    uint64_t ds[1] = {(uint64_t)Cpad * 2};
    uint32_t dbox[2] = {64, (uint32_t)WG_KB};
    if (tpgsr_make_tmap_bf16(&tmX_hi, x_hi, 2, xd, xs, xbox)) return TPGSR_ERR_DRIVER;
    if (tpgsr_make_tmap_bf16(&tmD_hi, dy_hi, 2, dd, ds, dbox)) return TPGSR_ERR_DRIVER;
    tmX_lo = tmX_hi;
    tmD_lo = tmD_hi;
    if (nprod >= 2 && tpgsr_make_tmap_bf16(&tmX_lo, x_lo, 2, xd, xs, xbox)) return TPGSR_ERR_DRIVER;
    if (nprod == 3 && tpgsr_make_tmap_bf16(&tmD_lo, dy_lo, 2, dd, ds, dbox)) return TPGSR_ERR_DRIVER;
    const size_t smem = (size_t)pl.stages * pl.stage_bytes + 1024 + 256;
    dim3 grid(pl.splits, KH * pl.kw_groups * pl.cbs, pl.n_tiles);
    if (nprod == 3) {
        if (cudaFuncSetAttribute(conv_wgrad_flat_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        conv_wgrad_flat_kernel<3><<<grid, 192, smem, stream>>>(tmX_hi, tmX_lo, tmD_hi, tmD_lo, p);
    } else if (nprod == 2) {
        if (cudaFuncSetAttribute(conv_wgrad_flat_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        conv_wgrad_flat_kernel<2><<<grid, 192, smem, stream>>>(tmX_hi, tmX_lo, tmD_hi, tmD_lo, p);
    } else {
        if (cudaFuncSetAttribute(conv_wgrad_flat_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TPGSR_ERR_LAUNCH;
        conv_wgrad_flat_kernel<1><<<grid, 192, smem, stream>>>(tmX_hi, tmX_lo, tmD_hi, tmD_lo, p);
    }
    rc = tpgsr_launch_status();
    if (rc) return rc;
    return tpgsr_conv2d_wgrad_reduce(static_cast<float*>(workspace), dw_oihw, pl.splits, KH * KW * Ci, Co, Ci, KH, KW, accumulate, stream);
}

}  // extern "C"
