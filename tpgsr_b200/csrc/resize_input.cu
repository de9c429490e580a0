// Per-image input transform on the device: dataset/dataset.py:615-632 resizeNormalize.__call__ —
//     img.resize((W, H), Image.BICUBIC) -> ToTensor -> [mask: img.convert('L'), 0 if x > mean else 255 -> ToTensor] -> cat
// for a whole batch of RAGGED uint8 RGB crops in one launch (the reference runs Pillow per image inside the DataLoader's collate, then copies
// fp32 tensors; here the raw bytes cross PCIe — 3 B/pixel of the crop instead of 16 B/pixel of the resized tensor — and the fp32 NCHW batch is
// produced where the train step reads it).
//
// Byte/integer work with a bit-exact bar, so the arithmetic is Pillow's (src/libImaging/Resample.c, 8 bits per channel), restated:
//   * weights per output coordinate in float64, normalised, rounded to 22-bit fixed point (precompute_coeffs + normalize_coeffs_8bpc).  Every
//     float64 operation is spelled with __dadd_rn / __dmul_rn / __ddiv_rn so that nvcc cannot contract a multiply-add into an FMA: the
//     rounding of k * 2^22 to an integer must match the CPU's bit for bit;
//   * horizontal pass, then vertical pass, each clip8((2^21 + sum(pixel * k)) >> 22) in int32, a uint8 image in between; a pass whose input and
//     output sizes agree is skipped (identity);
//   * L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16 (Convert.c L24); mask = (L * H*W > sum L) ? 0 : 1 — the reference compares the integer with
//     the float64 mean, which is the same predicate;
//   * ToTensor: uint8 -> float32, IEEE division by 255.
//
// One CTA per crop: the weight tables (W x ksize_x, H x ksize_y ints) and the resized H x W x 3 byte image live in shared memory, the
// horizontal-pass image (h_in x W x 3 bytes, h_in ragged) in a caller-provided scratch that stays in L2.  Reads of the crop are byte-granular
// but each crop is a few KB; stores of the fp32 planes are coalesced along W.  HBM-bound, latency-dominated at B = 48.
#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int PREC = 32 - 8 - 2;

// [host-testable begin] — tests/test_resize_host.py compiles this region (weight tables AND the kernel body, run as one sequential
// "thread" per CTA) with g++, the *_rn intrinsics mapped to plain IEEE operators with contraction off, and compares tables and outputs with
// the oracle / the reference's fixture: keep it free of anything but arithmetic, indexing, barriers and the one warp reduction.
__device__ __forceinline__ double bicubic_w(double x) {
    // a = -0.5:  ((a + 2) x - (a + 3)) x x + 1   for |x| < 1;   (((x - 5) x + 8) x - 4) a   for |x| < 2
    if (x < 0.0) x = -x;
    if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(1.5, x), -2.5), x), x), 1.0);
    if (x < 2.0) return __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(x, -5.0), x), 8.0), x), -4.0), -0.5);
    return 0.0;
}

__host__ __device__ inline int ksize_for(int in_size, int out_size) {
    double scale = (double)in_size / out_size;
    if (scale < 1.0) scale = 1.0;
    const double support = 2.0 * scale;
    int c = (int)support;
    if ((double)c < support) ++c;               // ceil
    return c * 2 + 1;
}

// one thread per output coordinate: bounds[xx] = (xmin, count), kk[xx * pitch + 0..count) = fixed-point weights
__device__ void build_coeffs(int in_size, int out_size, int pitch, int* __restrict__ bounds, int* __restrict__ kk) {
    const double scale = __ddiv_rn((double)(float)in_size, (double)out_size);      // (double)(in1 - in0) / outSize with float box coordinates
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = __dmul_rn(2.0, filterscale);
    const double ss = __ddiv_rn(1.0, filterscale);
    for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) {
        const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
        int xmin = (int)__dadd_rn(__dadd_rn(center, -support), 0.5);           // C cast: truncation toward zero
        if (xmin < 0) xmin = 0;
        int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        if (xmax > pitch) xmax = pitch;                                        // cannot happen when pitch = ksize_for(max in, out)
        int* k = kk + xx * pitch;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            const double w = bicubic_w(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss));
            ww = __dadd_rn(ww, w);
        }
        for (int x = 0; x < xmax; ++x) {
            double w = bicubic_w(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss));
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double f = __dmul_rn(w, (double)(1 << PREC));
            k[x] = w < 0.0 ? (int)__dadd_rn(-0.5, f) : (int)__dadd_rn(0.5, f);
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
}

__device__ __forceinline__ unsigned char clip8(int v) {
    v >>= PREC;                                   // arithmetic shift, like the C code on int
    return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct ResizeParams {
    const unsigned char* src;
    const long long* offsets;
    const int* hs;
    const int* ws;
    float* out;
    unsigned char* tmp;       // [B][max_h][W][3]
    int H, W, mask, max_h, max_w, ksx, ksy;
};

__global__ void __launch_bounds__(RS_THREADS) resize_normalize_kernel(ResizeParams p) {
    extern __shared__ __align__(16) unsigned char smem_rs[];
    const int H = p.H, W = p.W;
    int* kx = reinterpret_cast<int*>(smem_rs);            // [W][ksx]
    int* bx = kx + W * p.ksx;                              // [W][2]
    int* ky = bx + 2 * W;                                  // [H][ksy]
    int* by = ky + H * p.ksy;                              // [H][2]
    int* red = by + 2 * H;                                 // [RS_THREADS / 32 + 1]
    unsigned char* img = reinterpret_cast<unsigned char*>(red + RS_THREADS / 32 + 1);      // [H][W][3] resized bytes
    const int b = blockIdx.x;
    const int h_in = p.hs[b], w_in = p.ws[b];
    const int C = p.mask ? 4 : 3;
    float* out = p.out + (size_t)b * C * H * W;
    if (h_in < 1 || w_in < 1 || h_in > p.max_h || w_in > p.max_w) {      // sizes the launch was not provisioned for: poison, do not guess
        for (int i = threadIdx.x; i < C * H * W; i += blockDim.x) out[i] = __int_as_float(0x7fc00000);
        return;
    }
    const unsigned char* src = p.src + p.offsets[b];
    const bool horiz = w_in != W, vert = h_in != H;
    if (horiz) build_coeffs(w_in, W, p.ksx, bx, kx);
    if (vert) build_coeffs(h_in, H, p.ksy, by, ky);
    __syncthreads();
    // horizontal pass: src [h_in][w_in][3] -> mid [h_in][W][3]
    unsigned char* tmp = p.tmp + (size_t)b * p.max_h * W * 3;
    const unsigned char* mid = src;
    if (horiz) {
        for (int i = threadIdx.x; i < h_in * W; i += blockDim.x) {
            const int y = i / W, xx = i - y * W;
            const int xmin = bx[2 * xx], n = bx[2 * xx + 1];
            const int* k = kx + xx * p.ksx;
            const unsigned char* row = src + ((size_t)y * w_in + xmin) * 3;
            int s0 = 1 << (PREC - 1), s1 = s0, s2 = s0;
            for (int x = 0; x < n; ++x) {
                const int kv = k[x];
                s0 += row[3 * x] * kv;
                s1 += row[3 * x + 1] * kv;
                s2 += row[3 * x + 2] * kv;
            }
            unsigned char* d = tmp + (size_t)i * 3;
            d[0] = clip8(s0); d[1] = clip8(s1); d[2] = clip8(s2);
        }
        mid = tmp;
        __syncthreads();             // the block's own global writes are visible to the block after the barrier
    }
    // vertical pass: mid [h_in][W][3] -> img [H][W][3] (shared), luma sum for the mask
    int lsum = 0;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int yy = i / W, xx = i - yy * W;
        unsigned char r, g, bl;
        if (vert) {
            const int ymin = by[2 * yy], n = by[2 * yy + 1];
            const int* k = ky + yy * p.ksy;
            const unsigned char* col = mid + ((size_t)ymin * W + xx) * 3;
            int s0 = 1 << (PREC - 1), s1 = s0, s2 = s0;
            for (int y = 0; y < n; ++y) {
                const int kv = k[y];
                s0 += col[(size_t)y * W * 3] * kv;
                s1 += col[(size_t)y * W * 3 + 1] * kv;
                s2 += col[(size_t)y * W * 3 + 2] * kv;
            }
            r = clip8(s0); g = clip8(s1); bl = clip8(s2);
        } else {
            const unsigned char* px = mid + (size_t)i * 3;
            r = px[0]; g = px[1]; bl = px[2];
        }
        img[3 * i] = r; img[3 * i + 1] = g; img[3 * i + 2] = bl;
        lsum += (r * 19595 + g * 38470 + bl * 7471 + 0x8000) >> 16;
    }
    if (p.mask) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
    }
    __syncthreads();
    long long total = 0;
    if (p.mask)
        for (int w = 0; w < RS_THREADS / 32; ++w) total += red[w];
    // ToTensor: planes [C][H][W], consecutive threads = consecutive columns
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int r = img[3 * i], g = img[3 * i + 1], bl = img[3 * i + 2];
        out[i] = __fdiv_rn((float)r, 255.f);
        out[H * W + i] = __fdiv_rn((float)g, 255.f);
        out[2 * H * W + i] = __fdiv_rn((float)bl, 255.f);
        if (p.mask) {
            const int L = (r * 19595 + g * 38470 + bl * 7471 + 0x8000) >> 16;
            out[3 * H * W + i] = ((long long)L * (H * W) > total) ? 0.f : 1.f;       // 0 if L > mean else 255 -> /255
        }
    }
}

// [host-testable end]

size_t resize_smem_bytes(int H, int W, int ksx, int ksy) {
    return (size_t)(W * ksx + 2 * W + H * ksy + 2 * H + RS_THREADS / 32 + 1) * sizeof(int) + (size_t)H * W * 3 + 16;
}

}  // namespace

extern "C" {

// scratch for the horizontal-pass images: B * max_h * W * 3 bytes
size_t tpgsr_resize_normalize_workspace_bytes(int B, int max_h, int W) {
    return (size_t)(B > 0 ? B : 0) * (size_t)(max_h > 0 ? max_h : 0) * (size_t)(W > 0 ? W : 0) * 3;
}

// src: the crops' bytes back to back (each [h][w][3] uint8, RGB, tightly packed); offsets[b] = byte offset of crop b; hs/ws = its size — all in
// device memory.  max_h / max_w: upper bounds of hs / ws known to the host (they size the tables; a crop that exceeds them yields NaN, not a
// guess).  out: [B][3 + mask][H][W] fp32.
int tpgsr_resize_normalize(const unsigned char* src, const long long* offsets, const int* hs, const int* ws, int B, int max_h, int max_w, int H,
                           int W, int mask, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!src || !offsets || !hs || !ws || !out || B < 1 || H < 1 || W < 1 || max_h < 1 || max_w < 1) return TPGSR_ERR_BAD_ARG;
    if (max_h >= (1 << 24) || max_w >= (1 << 24)) return TPGSR_ERR_UNSUPPORTED;          // box coordinates are float32 in the algorithm
    if (workspace_bytes < tpgsr_resize_normalize_workspace_bytes(B, max_h, W) || (!workspace && workspace_bytes)) return TPGSR_ERR_WORKSPACE;
    ResizeParams p;
    p.src = src; p.offsets = offsets; p.hs = hs; p.ws = ws; p.out = out; p.tmp = static_cast<unsigned char*>(workspace);
    p.H = H; p.W = W; p.mask = mask ? 1 : 0; p.max_h = max_h; p.max_w = max_w;
    p.ksx = ksize_for(max_w, W);
    p.ksy = ksize_for(max_h, H);
    const size_t smem = resize_smem_bytes(H, W, p.ksx, p.ksy);
    if (smem > 200 * 1024) return TPGSR_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(resize_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return TPGSR_ERR_LAUNCH;
    resize_normalize_kernel<<<B, RS_THREADS, smem, stream>>>(p);
    return tpgsr_launch_status();
}

}  // extern "C"
