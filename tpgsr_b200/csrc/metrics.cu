// PSNR / SSIM of SR batches on the GPU (reference: utils/ssim_psnr.py:9-15 calculate_psnr, :18-78 SSIM/_ssim — the parity metrics of
// BASELINE.json, used by TextSR.eval at interfaces/base.py:230-231).  Both act on the first three channels (RGB) of NCHW images.
//
// One CTA per (image, channel): both 32x128-class planes are staged in shared memory once, every thread evaluates the 11x11 gaussian
// window (zero padding, sigma 1.5, as F.conv2d(padding=5, groups=C)) for its pixels, and the CTA writes two double partials
// (sum of squared 8-bit-scale error, sum of the SSIM map).  A one-warp finalize turns them into the two scalars in a fixed order.
#include "common.cuh"

namespace {

constexpr int WIN = 11, HALF = 5;

__global__ void __launch_bounds__(256)
psnr_ssim_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long sab, long sac, long sah, long saw, long sbb, long sbc,
                         long sbh, long sbw, int C, int H, int W, double* __restrict__ partial) {
    extern __shared__ float sm[];
    float* pa = sm;
    float* pb = sm + H * W;
    __shared__ float gw[WIN];
    __shared__ double red[2][8];
    const int img = blockIdx.x / C, c = blockIdx.x % C;
    if (threadIdx.x < WIN) {
        // gaussian(window_size=11, sigma=1.5) normalised in fp32 like the reference's torch.Tensor
        float s = 0.f;
        for (int x = 0; x < WIN; ++x) s += expf(-(float)((x - HALF) * (x - HALF)) / (2.f * 1.5f * 1.5f));
        gw[threadIdx.x] = expf(-(float)((threadIdx.x - HALF) * (threadIdx.x - HALF)) / (2.f * 1.5f * 1.5f)) / s;
    }
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int h = i / W, w = i - h * W;
        pa[i] = a[img * sab + c * sac + h * sah + w * saw];
        pb[i] = b[img * sbb + c * sbc + h * sbh + w * sbw];
    }
    __syncthreads();
    double se = 0.0, ss = 0.0;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int h = i / W, w = i - h * W;
        const float d = __fmul_rn(pa[i], 255.f) - __fmul_rn(pb[i], 255.f);      // no FMA contraction: identical images must give exactly 0
        se += (double)(d * d);
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
        for (int dy = -HALF; dy <= HALF; ++dy) {
            const int hh = h + dy;
            if (hh < 0 || hh >= H) continue;
            const float gy = gw[dy + HALF];
            for (int dx = -HALF; dx <= HALF; ++dx) {
                const int ww = w + dx;
                if (ww < 0 || ww >= W) continue;
                const float g = gy * gw[dx + HALF];
                const float x = pa[hh * W + ww], y = pb[hh * W + ww];
                m1 = fmaf(g, x, m1); m2 = fmaf(g, y, m2);
                s11 = fmaf(g, x * x, s11); s22 = fmaf(g, y * y, s22); s12 = fmaf(g, x * y, s12);
            }
        }
        const float mu11 = m1 * m1, mu22 = m2 * m2, mu12 = m1 * m2;
        const float v1 = s11 - mu11, v2 = s22 - mu22, v12 = s12 - mu12;
        ss += (double)(((2.f * mu12 + C1) * (2.f * v12 + C2)) / ((mu11 + mu22 + C1) * (v1 + v2 + C2)));
    }
    se = warp_sum_d(se);
    ss = warp_sum_d(ss);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = se; red[1][warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { t0 += red[0][i]; t1 += red[1][i]; }
        partial[2 * blockIdx.x] = t0;
        partial[2 * blockIdx.x + 1] = t1;
    }
}

// out[0] = PSNR = 20 log10(255 / sqrt(mse)) (inf when mse == 0), out[1] = mean SSIM
__global__ void psnr_ssim_finalize_kernel(const double* __restrict__ partial, int n, double count, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double se = 0.0, ss = 0.0;
    for (int i = 0; i < n; ++i) { se += partial[2 * i]; ss += partial[2 * i + 1]; }
    const double mse = se / count;
    out[0] = mse == 0.0 ? INFINITY : (float)(20.0 * log10(255.0 / sqrt(mse)));
    out[1] = (float)(ss / count);
}

}  // namespace

extern "C" {

size_t tpgsr_psnr_ssim_workspace_bytes(int B, int C) { return (size_t)B * C * 2 * sizeof(double); }

// a, b: NCHW fp32 images addressed through their strides (callers pass channel-sliced views); the first C (= 3) channels are compared.
// out: 2 floats on the device {psnr, ssim}.
int tpgsr_psnr_ssim(const float* a, const float* b, int B, int C, int H, int W, long sab, long sac, long sah, long saw, long sbb, long sbc,
                    long sbh, long sbw, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!a || !b || !out || !workspace || B < 1 || C < 1 || H < 1 || W < 1) return TPGSR_ERR_BAD_ARG;
    if (workspace_bytes < tpgsr_psnr_ssim_workspace_bytes(B, C)) return TPGSR_ERR_WORKSPACE;
    const size_t smem = (size_t)2 * H * W * sizeof(float);
    if (smem > 200 * 1024) return TPGSR_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(psnr_ssim_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return TPGSR_ERR_LAUNCH;
    double* part = static_cast<double*>(workspace);
    psnr_ssim_partial_kernel<<<B * C, 256, smem, stream>>>(a, b, sab, sac, sah, saw, sbb, sbc, sbh, sbw, C, H, W, part);
    int rc = tpgsr_launch_status();
    if (rc) return rc;
    psnr_ssim_finalize_kernel<<<1, 32, 0, stream>>>(part, B * C, (double)B * C * H * W, out);
    return tpgsr_launch_status();
}

}  // extern "C"
