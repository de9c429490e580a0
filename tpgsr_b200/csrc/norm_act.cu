// Normalisation, activation, pooling and reshuffle kernels on NHWC fp32 tensors (HBM/L2-bound elementwise + column reductions).
//
// Reference ops replaced: nn.BatchNorm2d/1d train+eval (model/tsrn.py:377,382,38; model/stn_head.py:19,51;
// model/crnn/crnn.py:48), mish (model/tsrn.py:480-488), nn.PReLU (model/tsrn.py:28), nn.ReLU, nn.MaxPool2d
// (model/stn_head.py:36-44, model/crnn/crnn.py:56-66), nn.PixelShuffle(2)+mish (model/tsrn.py:464-477), torch.tanh
// (model/tsrn.py:74), bias gradients, NCHW<->NHWC boundary transposes.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

// ------------------------------------------------------------------------------------------------
// generic 2-value column reduction over a [M, C] row-major matrix: partial[rb][v][c]
// ------------------------------------------------------------------------------------------------
// sum (x - p), sum (x - p)^2 with the per-channel pivot p = x[0][c]: the variance E[x^2] - mean^2 of a channel whose |mean| is much larger than
// its standard deviation (post-ReLU maps, large conv biases) cancels catastrophically in fp32 partial sums; around a pivot inside the data it
// does not (ATen uses Welford).  The finalize step adds the pivot back to the mean.
struct StatsF {
    const float* x; int ld;
    __device__ float2 operator()(long r, int c) const { float v = x[(size_t)r * ld + c] - __ldg(x + c); return make_float2(v, v * v); }
};
struct SumF {     // sum x (bias gradients)
    const float* x; int ld;
    __device__ float2 operator()(long r, int c) const { return make_float2(x[(size_t)r * ld + c], 0.f); }
};
struct BnBwdF {   // sum dy', sum dy'*xhat with dy' = dy * act'(y_pre)
    const float* dy; const float* x; const float* scale_shift; const float* mean_rstd; int C, ldx, lddy, act;
    __device__ float2 operator()(long r, int c) const {
        float xv = x[(size_t)r * ldx + c];
        float g = dy[(size_t)r * lddy + c];
        float yp = xv * scale_shift[c] + scale_shift[C + c];
        if (act == TPGSR_ACT_RELU) g = yp > 0.f ? g : 0.f;
        else if (act == TPGSR_ACT_MISH) g *= mish_grad_f(yp);
        float xh = (xv - mean_rstd[c]) * mean_rstd[C + c];
        return make_float2(g, g * xh);
    }
};

template <class F>
__global__ void __launch_bounds__(256) col_reduce2_kernel(F f, long M, int C, long rows_per_block, float* __restrict__ partial) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long r0 = (long)blockIdx.y * rows_per_block;
    long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float s1 = 0.f, s2 = 0.f;
    if (c < C)
        for (long r = r0 + threadIdx.y; r < r1; r += 8) {
            float2 v = f(r, c);
            s1 += v.x;
            s2 += v.y;
        }
    __shared__ float sh[2][8][33];
    sh[0][threadIdx.y][threadIdx.x] = s1;
    sh[1][threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += sh[0][i][threadIdx.x]; b += sh[1][i][threadIdx.x]; }
        partial[((size_t)blockIdx.y * 2 + 0) * C + c] = a;
        partial[((size_t)blockIdx.y * 2 + 1) * C + c] = b;
    }
}

inline void reduce_plan(long M, int C, int& row_blocks, long& rows_per_block) {
    int col_blocks = ceil_div(C, 32);
    long want = (148L * 4 + col_blocks - 1) / col_blocks;
    long maxb = (M + 63) / 64;
    long rb = want < maxb ? want : maxb;
    if (rb < 1) rb = 1;
    rows_per_block = (M + rb - 1) / rb;
    rows_per_block = (rows_per_block + 7) / 8 * 8;
    row_blocks = (int)((M + rows_per_block - 1) / rows_per_block);
}

template <class F>
int launch_col_reduce(F f, long M, int C, float* partial, int& row_blocks, cudaStream_t s) {
    long rpb;
    reduce_plan(M, C, row_blocks, rpb);
    dim3 grid(ceil_div(C, 32), row_blocks), block(32, 8);
    col_reduce2_kernel<F><<<grid, block, 0, s>>>(f, M, C, rpb, partial);
    return tpgsr_launch_status();
}

// Sum the [row_blocks][2][C] partials of column c.  blockDim = (32, 32): the 32 y-lanes split the row blocks (so the serial,
// latency-bound part is ~row_blocks/32 loads), results are valid on threadIdx.y == 0.  Deterministic (fixed summation tree).
constexpr int FIN_Y = 32;
__device__ __forceinline__ void sum_partials(const float* __restrict__ partial, int row_blocks, int C, int c, bool two,
                                             double& s1, double& s2) {
    float a = 0.f, b = 0.f;
    if (c < C)
        for (int rb = threadIdx.y; rb < row_blocks; rb += FIN_Y) {
            a += partial[((size_t)rb * 2 + 0) * C + c];
            if (two) b += partial[((size_t)rb * 2 + 1) * C + c];
        }
    __shared__ float shp[2][FIN_Y][33];
    shp[0][threadIdx.y][threadIdx.x] = a;
    shp[1][threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    s1 = 0.0; s2 = 0.0;
    if (threadIdx.y == 0)
#pragma unroll
        for (int i = 0; i < FIN_Y; ++i) { s1 += (double)shp[0][i][threadIdx.x]; s2 += (double)shp[1][i][threadIdx.x]; }
}

// BN forward finalize: mean / rstd / folded scale+shift / running stats
__global__ void bn_fwd_finalize_kernel(const float* __restrict__ partial, int row_blocks, long M, int C,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ mean_rstd, float* __restrict__ scale_shift,
                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                       float momentum, float eps, const float* __restrict__ pivot) {
    int c = blockIdx.x * 32 + threadIdx.x;
    double s1, s2;
    sum_partials(partial, row_blocks, C, c, true, s1, s2);
    if (c >= C || threadIdx.y != 0) return;
    double mean = s1 / (double)M;                    // of x - pivot
    double var = s2 / (double)M - mean * mean;
    mean += (double)pivot[c];
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    mean_rstd[c] = (float)mean;
    mean_rstd[C + c] = rstd;
    float sc = gamma[c] * rstd;
    scale_shift[c] = sc;
    scale_shift[C + c] = beta[c] - (float)mean * sc;
    if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

// eval mode: fold running stats into scale/shift (mean_rstd also written so the same bwd code could run)
__global__ void bn_eval_prepare_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                       float eps, float* __restrict__ mean_rstd, float* __restrict__ scale_shift) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float rstd = 1.f / sqrtf(running_var[c] + eps);
    float sc = gamma[c] * rstd;
    mean_rstd[c] = running_mean[c];
    mean_rstd[C + c] = rstd;
    scale_shift[c] = sc;
    scale_shift[C + c] = beta[c] - running_mean[c] * sc;
}

// y = act(x*scale + shift) (+ residual)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        const float* __restrict__ scale_shift, long M, int C, int ldx,
                                                        int ldy, int act, const float* __restrict__ residual, int ldr) {
    const int cq = C >> 2;
    const long total = M * cq;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i / cq;
        int c = (int)(i - r * cq) * 4;
        float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * ldx + c);
        float4 sc = *reinterpret_cast<const float4*>(scale_shift + c);
        float4 sh = *reinterpret_cast<const float4*>(scale_shift + C + c);
        v.x = apply_act(fmaf(v.x, sc.x, sh.x), act, 0.f);
        v.y = apply_act(fmaf(v.y, sc.y, sh.y), act, 0.f);
        v.z = apply_act(fmaf(v.z, sc.z, sh.z), act, 0.f);
        v.w = apply_act(fmaf(v.w, sc.w, sh.w), act, 0.f);
        if (residual) {
            float4 rr = *reinterpret_cast<const float4*>(residual + (size_t)r * ldr + c);
            v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        *reinterpret_cast<float4*>(y + (size_t)r * ldy + c) = v;
    }
}

// BN backward finalize: dgamma, dbeta (+)=, and per-channel coefficients for the apply pass
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int row_blocks, long M, int C,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                                       int accumulate) {
    int c = blockIdx.x * 32 + threadIdx.x;
    double s1, s2;
    sum_partials(partial, row_blocks, C, c, true, s1, s2);
    if (c >= C || threadIdx.y != 0) return;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s1 : (float)s1;
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)s2 : (float)s2;
    coef[c] = (float)(s1 / (double)M);
    coef[C + c] = (float)(s2 / (double)M);
}

// dx = scale * (dy' - mean(dy') - xhat * mean(dy'*xhat))
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            float* __restrict__ dx, const float* __restrict__ scale_shift,
                                                            const float* __restrict__ mean_rstd, const float* __restrict__ coef,
                                                            long M, int C, int lddy, int ldx, int lddx, int act, int train) {
    const long total = M * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i / C;
        int c = (int)(i - r * C);
        float xv = x[(size_t)r * ldx + c];
        float g = dy[(size_t)r * lddy + c];
        float sc = scale_shift[c];
        float yp = fmaf(xv, sc, scale_shift[C + c]);
        if (act == TPGSR_ACT_RELU) g = yp > 0.f ? g : 0.f;
        else if (act == TPGSR_ACT_MISH) g *= mish_grad_f(yp);
        float out;
        if (train) {
            float xh = (xv - mean_rstd[c]) * mean_rstd[C + c];
            out = sc * (g - coef[c] - xh * coef[C + c]);
        } else {
            out = sc * g;
        }
        dx[(size_t)r * lddx + c] = out;
    }
}

// column sums finalize (bias grads)
__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int row_blocks, int C, float* __restrict__ out,
                                       int accumulate) {
    int c = blockIdx.x * 32 + threadIdx.x;
    double s, unused;
    sum_partials(partial, row_blocks, C, c, false, s, unused);
    if (c >= C || threadIdx.y != 0) return;
    out[c] = accumulate ? out[c] + (float)s : (float)s;
}

// ------------------------------------------------------------------------------------------------
// Fused column reduction + finalize ("last block finalizes"): the blocks of the reduction take a ticket after publishing their partial
// sums; the block that draws the last ticket sums the partials in the same fixed order as the stand-alone finalize kernels (so results are
// bit-identical to the two-launch path) and resets the ticket, which therefore stays zero between launches and graph replays.  Saves one
// dependent launch (~6 us of a latency-bound chain) per BatchNorm forward and per BatchNorm backward.
// ------------------------------------------------------------------------------------------------
struct BnFwdFin {
    long M; const float* gamma; const float* beta; float* mean_rstd; float* scale_shift; float* running_mean; float* running_var;
    float momentum, eps; const float* pivot;
    __device__ void operator()(int c, int C, double s1, double s2) const {
        double mean = s1 / (double)M;                // of x - pivot
        double var = s2 / (double)M - mean * mean;
        mean += (double)pivot[c];
        if (var < 0.0) var = 0.0;
        float rstd = (float)(1.0 / sqrt(var + (double)eps));
        mean_rstd[c] = (float)mean;
        mean_rstd[C + c] = rstd;
        float sc = gamma[c] * rstd;
        scale_shift[c] = sc;
        scale_shift[C + c] = beta[c] - (float)mean * sc;
        if (running_mean) {
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
        }
    }
};
struct BnBwdFin {
    long M; float* dgamma; float* dbeta; float* coef; int accumulate;
    __device__ void operator()(int c, int C, double s1, double s2) const {
        if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s1 : (float)s1;
        if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)s2 : (float)s2;
        coef[c] = (float)(s1 / (double)M);
        coef[C + c] = (float)(s2 / (double)M);
    }
};

template <class F, class FIN>
__global__ void __launch_bounds__(256) col_reduce2_fin_kernel(F f, FIN fin, long M, int C, long rows_per_block, float* __restrict__ partial,
                                                              unsigned* __restrict__ ticket) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long r0 = (long)blockIdx.y * rows_per_block;
    long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float s1 = 0.f, s2 = 0.f;
    if (c < C)
        for (long r = r0 + threadIdx.y; r < r1; r += 8) {
            float2 v = f(r, c);
            s1 += v.x;
            s2 += v.y;
        }
    __shared__ float sh[2][8][33];
    __shared__ unsigned last;
    sh[0][threadIdx.y][threadIdx.x] = s1;
    sh[1][threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += sh[0][i][threadIdx.x]; b += sh[1][i][threadIdx.x]; }
        partial[((size_t)blockIdx.y * 2 + 0) * C + c] = a;
        partial[((size_t)blockIdx.y * 2 + 1) * C + c] = b;
    }
    __threadfence();                                      // partials visible device-wide before the ticket is taken
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1 ? 1u : 0u;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int row_blocks = gridDim.y;
    for (int c0 = 0; c0 < C; c0 += 32) {                  // the last block finalizes every channel, 32 at a time
        const int cc = c0 + threadIdx.x;
        // same summation tree as sum_partials (FIN_Y = 32 strided float sums, then doubles in order) so both paths agree bit for bit
        double d1 = 0.0, d2 = 0.0;
        __syncthreads();
        for (int grp = 0; grp < FIN_Y / 8; ++grp) {
            const int y = grp * 8 + threadIdx.y;
            float a = 0.f, b = 0.f;
            if (cc < C)
                for (int rb = y; rb < row_blocks; rb += FIN_Y) {
                    a += __ldcg(partial + ((size_t)rb * 2 + 0) * C + cc);
                    b += __ldcg(partial + ((size_t)rb * 2 + 1) * C + cc);
                }
            sh[0][threadIdx.y][threadIdx.x] = a;
            sh[1][threadIdx.y][threadIdx.x] = b;
            __syncthreads();
            if (threadIdx.y == 0)
#pragma unroll
                for (int i = 0; i < 8; ++i) { d1 += (double)sh[0][i][threadIdx.x]; d2 += (double)sh[1][i][threadIdx.x]; }
            __syncthreads();
        }
        if (threadIdx.y == 0 && cc < C) fin(cc, C, d1, d2);
    }
    if (threadIdx.x == 0 && threadIdx.y == 0) *ticket = 0u;
}

template <class F, class FIN>
int launch_col_reduce_fin(F f, FIN fin, long M, int C, float* partial, unsigned* ticket, cudaStream_t s) {
    long rpb;
    int row_blocks;
    reduce_plan(M, C, row_blocks, rpb);
    dim3 grid(ceil_div(C, 32), row_blocks), block(32, 8);
    col_reduce2_fin_kernel<F, FIN><<<grid, block, 0, s>>>(f, fin, M, C, rpb, partial, ticket);
    return tpgsr_launch_status();
}

// ------------------------------------------------------------------------------------------------
// BN apply / BN backward apply that ALSO emit the result as padded-flat bf16 hi/lo operand planes of the next tensor-core conv
// ([B][H+P][W+P][C], pad rows / columns zero: conv_flat.cu) — replaces a separate tpgsr_split_bf16_flat pass over the same tensor.
// The grid walks the FLAT rows, so the pad positions are written (zeros) by the same kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, float4 v) {
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(f[j]);
        l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
    }
    *reinterpret_cast<uint2*>(hi + off) = *reinterpret_cast<uint2*>(h);
    if (lo) *reinterpret_cast<uint2*>(lo + off) = *reinterpret_cast<uint2*>(l);
}

__global__ void __launch_bounds__(256) bn_apply_planes_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               const float* __restrict__ scale_shift, int B, int H, int W, int C, int act,
                                                               const float* __restrict__ residual, int P, __nv_bfloat16* __restrict__ hi,
                                                               __nv_bfloat16* __restrict__ lo) {
    const int cq = C >> 2, Hp = H + P, Wp = W + P;
    const long total = (long)B * Hp * Wp * cq;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long q = i / cq;
        const int c = (int)(i - q * cq) * 4;
        const int col = (int)(q % Wp);
        const long t = q / Wp;
        const int r = (int)(t % Hp);
        const int b = (int)(t / Hp);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < H && col < W) {
            const size_t row = ((size_t)b * H + r) * W + col;
            v = *reinterpret_cast<const float4*>(x + row * C + c);
            const float4 sc = *reinterpret_cast<const float4*>(scale_shift + c);
            const float4 sh = *reinterpret_cast<const float4*>(scale_shift + C + c);
            v.x = apply_act(fmaf(v.x, sc.x, sh.x), act, 0.f);
            v.y = apply_act(fmaf(v.y, sc.y, sh.y), act, 0.f);
            v.z = apply_act(fmaf(v.z, sc.z, sh.z), act, 0.f);
            v.w = apply_act(fmaf(v.w, sc.w, sh.w), act, 0.f);
            if (residual) {
                const float4 rr = *reinterpret_cast<const float4*>(residual + row * C + c);
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            *reinterpret_cast<float4*>(y + row * C + c) = v;
        }
        store_planes4(hi, lo, (size_t)q * C + c, v);
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_planes_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                   float* __restrict__ dx, const float* __restrict__ scale_shift,
                                                                   const float* __restrict__ mean_rstd, const float* __restrict__ coef,
                                                                   int B, int H, int W, int C, int act, int train, int P,
                                                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int cq = C >> 2, Hp = H + P, Wp = W + P;
    const long total = (long)B * Hp * Wp * cq;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long q = i / cq;
        const int c = (int)(i - q * cq) * 4;
        const int col = (int)(q % Wp);
        const long t = q / Wp;
        const int r = (int)(t % Hp);
        const int b = (int)(t / Hp);
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < H && col < W) {
            const size_t row = ((size_t)b * H + r) * W + col;
            const float4 xv4 = *reinterpret_cast<const float4*>(x + row * C + c);
            const float4 g4 = *reinterpret_cast<const float4*>(dy + row * C + c);
            const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float sc = scale_shift[c + j];
                const float yp = fmaf(xv[j], sc, scale_shift[C + c + j]);
                float g = gv[j];
                if (act == TPGSR_ACT_RELU) g = yp > 0.f ? g : 0.f;
                else if (act == TPGSR_ACT_MISH) g *= mish_grad_f(yp);
                if (train) {
                    const float xh = (xv[j] - mean_rstd[c + j]) * mean_rstd[C + c + j];
                    o[j] = sc * (g - coef[c + j] - xh * coef[C + c + j]);
                } else {
                    o[j] = sc * g;
                }
            }
            *reinterpret_cast<float4*>(dx + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
        }
        store_planes4(hi, lo, (size_t)q * C + c, make_float4(o[0], o[1], o[2], o[3]));
    }
}

// ------------------------------------------------------------------------------------------------
// elementwise activations
// ------------------------------------------------------------------------------------------------
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long n, int act,
                               const float* __restrict__ slope_p) {
    float slope = slope_p ? __ldg(slope_p) : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        y[i] = apply_act(x[i], act, slope);
}

// dx = dy * act'(.)  — `ref` is the pre-activation for PRELU/MISH, the OUTPUT for RELU/TANH.
// For PRELU also accumulates d(slope) = sum dy * x * [x<0] into partial[blockIdx.x].
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ ref,
                                                       float* __restrict__ dx, long n, int act,
                                                       const float* __restrict__ slope_p, float* __restrict__ partial) {
    float slope = slope_p ? __ldg(slope_p) : 0.f;
    float ds = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float g = dy[i], r = ref[i], o;
        switch (act) {
            case TPGSR_ACT_RELU: o = r > 0.f ? g : 0.f; break;
            case TPGSR_ACT_TANH: o = g * (1.f - r * r); break;
            case TPGSR_ACT_MISH: o = g * mish_grad_f(r); break;
            case TPGSR_ACT_PRELU: o = r >= 0.f ? g : slope * g; if (r < 0.f) ds += g * r; break;
            default: o = g;
        }
        dx[i] = o;
    }
    if (partial) {
        __shared__ float sh[8];
        ds = warp_sum(ds);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ds;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
            partial[blockIdx.x] = s;
        }
    }
}

__global__ void __launch_bounds__(256) scalar_finalize_kernel(const float* __restrict__ partial, int n, float* __restrict__ out, int accumulate) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    s = block_sum_d256(s);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + (float)s : (float)s;
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] = a[i] + b[i];
}

// ------------------------------------------------------------------------------------------------
// max-pool (NHWC), argmax stored as the flattened input pixel index within the image (h*W + w)
// ------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int* __restrict__ arg, int B,
                                   int Hi, int Wi, int C, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw) {
    const long total = (long)B * Ho * Wo * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int wo = (int)(p % Wo);
        long q = p / Wo;
        int ho = (int)(q % Ho);
        int b = (int)(q / Ho);
        float best = -INFINITY;
        int bi = -1;
        for (int a = 0; a < kh; ++a) {
            int h = ho * sh - ph + a;
            if ((unsigned)h >= (unsigned)Hi) continue;
            for (int d = 0; d < kw; ++d) {
                int w = wo * sw - pw + d;
                if ((unsigned)w >= (unsigned)Wi) continue;
                float v = x[((size_t)(b * Hi + h) * Wi + w) * C + c];
                if (v > best || bi < 0) { best = v; bi = h * Wi + w; }   // first max wins, like ATen
            }
        }
        y[i] = best;
        if (arg) arg[i] = bi;
    }
}

// gather formulation (deterministic): every input pixel sums the gradients of the windows that selected it
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const int* __restrict__ arg, float* __restrict__ dx,
                                   int B, int Hi, int Wi, int C, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph,
                                   int pw) {
    const long total = (long)B * Hi * Wi * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int w = (int)(p % Wi);
        long q = p / Wi;
        int h = (int)(q % Hi);
        int b = (int)(q / Hi);
        int me = h * Wi + w;
        float s = 0.f;
        int ho_lo = (h + ph - kh + sh) / sh;  if (h + ph - kh + 1 <= 0) ho_lo = 0;
        int wo_lo = (w + pw - kw + sw) / sw;  if (w + pw - kw + 1 <= 0) wo_lo = 0;
        int ho_hi = (h + ph) / sh, wo_hi = (w + pw) / sw;
        if (ho_hi > Ho - 1) ho_hi = Ho - 1;
        if (wo_hi > Wo - 1) wo_hi = Wo - 1;
        for (int ho = ho_lo; ho <= ho_hi; ++ho)
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                size_t o = ((size_t)(b * Ho + ho) * Wo + wo) * C + c;
                if (arg[o] == me) s += dy[o];
            }
        dx[i] = s;
    }
}

// float4 versions of the two kernels above (C % 4 == 0, fewer than 2^31 elements): 16-byte accesses and 32-bit index arithmetic.  The scalar
// kernels spent most of their time in three 64-bit divisions per element (CRNN pool0 at B = 48: 45 us forward, 60 us backward for 59 MB).
__global__ void __launch_bounds__(256) maxpool_fwd_v4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int4* __restrict__ arg,
                                                            unsigned total, int Hi, int Wi, int C4, int Ho, int Wo, int kh, int kw,
                                                            int sh, int sw, int ph, int pw) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned p = i / (unsigned)C4, c = i - p * (unsigned)C4;
        const unsigned q = p / (unsigned)Wo, wo = p - q * (unsigned)Wo;
        const unsigned b = q / (unsigned)Ho, ho = q - b * (unsigned)Ho;
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        int4 bi = make_int4(-1, -1, -1, -1);
        for (int a = 0; a < kh; ++a) {
            const int h = (int)ho * sh - ph + a;
            if ((unsigned)h >= (unsigned)Hi) continue;
            for (int d = 0; d < kw; ++d) {
                const int w = (int)wo * sw - pw + d;
                if ((unsigned)w >= (unsigned)Wi) continue;
                const int me = h * Wi + w;
                const float4 v = x[((size_t)(b * Hi + h) * Wi + w) * C4 + c];
                if (v.x > best.x || bi.x < 0) { best.x = v.x; bi.x = me; }   // first max wins, like ATen
                if (v.y > best.y || bi.y < 0) { best.y = v.y; bi.y = me; }
                if (v.z > best.z || bi.z < 0) { best.z = v.z; bi.z = me; }
                if (v.w > best.w || bi.w < 0) { best.w = v.w; bi.w = me; }
            }
        }
        y[i] = best;
        if (arg) arg[i] = bi;
    }
}

__global__ void __launch_bounds__(256) maxpool_bwd_v4_kernel(const float4* __restrict__ dy, const int4* __restrict__ arg, float4* __restrict__ dx,
                                                            unsigned total, int Hi, int Wi, int C4, int Ho, int Wo, int kh, int kw,
                                                            int sh, int sw, int ph, int pw) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned p = i / (unsigned)C4, c = i - p * (unsigned)C4;
        const unsigned q = p / (unsigned)Wi, w = p - q * (unsigned)Wi;
        const unsigned b = q / (unsigned)Hi, h = q - b * (unsigned)Hi;
        const int me = (int)(h * Wi + w);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        int ho_lo = ((int)h + ph - kh + sh) / sh;  if ((int)h + ph - kh + 1 <= 0) ho_lo = 0;
        int wo_lo = ((int)w + pw - kw + sw) / sw;  if ((int)w + pw - kw + 1 <= 0) wo_lo = 0;
        int ho_hi = ((int)h + ph) / sh, wo_hi = ((int)w + pw) / sw;
        if (ho_hi > Ho - 1) ho_hi = Ho - 1;
        if (wo_hi > Wo - 1) wo_hi = Wo - 1;
        for (int ho = ho_lo; ho <= ho_hi; ++ho)
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                const size_t o = ((size_t)(b * Ho + ho) * Wo + wo) * C4 + c;
                const int4 a = arg[o];
                const float4 g = dy[o];
                if (a.x == me) s.x += g.x;
                if (a.y == me) s.y += g.y;
                if (a.z == me) s.z += g.z;
                if (a.w == me) s.w += g.w;
            }
        dx[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// PixelShuffle(2) + mish on NHWC: x [B,H,W,4C] -> y [B,2H,2W,C];  y[b,2h+i,2w+j,c] = mish(x[b,h,w,4c+2i+j])
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixelshuffle_mish_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                      int B, int H, int W, int C) {
    // one thread per (input pixel, output channel c): reads a float4 (the 2x2 sub-pixels of channel c), writes 4 scalars
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int w = (int)(p % W);
        long q = p / W;
        int h = (int)(q % H);
        int b = (int)(q / H);
        float4 v = *reinterpret_cast<const float4*>(x + (size_t)p * 4 * C + 4 * c);
        size_t o = (((size_t)b * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
        y[o] = mish_f(v.x);
        y[o + C] = mish_f(v.y);
        y[o + (size_t)2 * W * C] = mish_f(v.z);
        y[o + (size_t)2 * W * C + C] = mish_f(v.w);
    }
}

__global__ void __launch_bounds__(256) pixelshuffle_mish_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                      float* __restrict__ dx, int B, int H, int W, int C) {
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int w = (int)(p % W);
        long q = p / W;
        int h = (int)(q % H);
        int b = (int)(q / H);
        float4 v = *reinterpret_cast<const float4*>(x + (size_t)p * 4 * C + 4 * c);
        size_t o = (((size_t)b * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
        float4 g;
        g.x = dy[o] * mish_grad_f(v.x);
        g.y = dy[o + C] * mish_grad_f(v.y);
        g.z = dy[o + (size_t)2 * W * C] * mish_grad_f(v.z);
        g.w = dy[o + (size_t)2 * W * C + C] * mish_grad_f(v.w);
        *reinterpret_cast<float4*>(dx + (size_t)p * 4 * C + 4 * c) = g;
    }
}

// ------------------------------------------------------------------------------------------------
// boundary transposes (strided NCHW view <-> dense NHWC), optional tanh-backward fused into the NCHW->NHWC direction
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int H, int W,
                                    long sb, long sc, long sh, long sw, int ldd, const float* __restrict__ tanh_out) {
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int w = (int)(p % W);
        long q = p / W;
        int h = (int)(q % H);
        int b = (int)(q / H);
        size_t so = (size_t)(b * sb + c * sc + h * sh + w * sw);
        float v = src[so];
        if (tanh_out) { float t = tanh_out[so]; v *= (1.f - t * t); }
        dst[(size_t)p * ldd + c] = v;
    }
}

// F.interpolate(x, scale_factor=s) (mode 'nearest', model/srcnn.py:138) fused with the NCHW -> NHWC boundary change:
// dst[b][h][w][c] = src[b][c][h / s][w / s]   (src addressed through its strides: callers pass channel-sliced views)
__global__ void upsample_nearest_nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int Ho,
                                                     int Wo, int scale, long sb, long sc, long sh, long sw) {
    const long total = (long)B * Ho * Wo * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long p = i / C;
        int w = (int)(p % Wo);
        long q = p / Wo;
        int h = (int)(q % Ho);
        int b = (int)(q / Ho);
        dst[i] = src[(size_t)(b * sb + c * sc + (h / scale) * sh + (w / scale) * sw)];
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int H, int W, int lds) {
    const long total = (long)B * C * H * W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int w = (int)(i % W);
        long q = i / W;
        int h = (int)(q % H);
        long r = q / H;
        int c = (int)(r % C);
        int b = (int)(r / C);
        dst[i] = src[((size_t)(b * H + h) * W + w) * lds + c];
    }
}

// dst[m, c] = (accumulate ? dst : 0) + alpha * src[m, c] for a [M, C] window of two pitched matrices
__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, long M, int C,
                                 float alpha, int accumulate) {
    const long total = M * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i / C;
        int c = (int)(i - r * C);
        float v = alpha * src[(size_t)r * lds + c];
        float* d = dst + (size_t)r * ldd + c;
        *d = accumulate ? *d + v : v;
    }
}

// wout[n][k] = sum_j w2[n][j] * w1[j][k];  bout[n] = b2[n] + sum_j w2[n][j] * b1[j]     (two chained linear maps without a non-linearity
// in between — GruBlock's 1x1 conv followed by the GRU input projection, model/tsrn.py:499-504 — collapse into one)
__global__ void fuse_linear_kernel(const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w1,
                                   const float* __restrict__ b1, float* __restrict__ wout, float* __restrict__ bout, int N2, int N1, int K1) {
    const int total = N2 * (K1 + 1);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int n = idx / (K1 + 1), k = idx - n * (K1 + 1);
        float s = 0.f;
        if (k < K1) {
            for (int j = 0; j < N1; ++j) s = fmaf(w2[(size_t)n * N1 + j], w1[(size_t)j * K1 + k], s);
            wout[(size_t)n * K1 + k] = s;
        } else {
            if (b1) for (int j = 0; j < N1; ++j) s = fmaf(w2[(size_t)n * N1 + j], b1[j], s);
            bout[n] = s + (b2 ? b2[n] : 0.f);
        }
    }
}

__global__ void add_i64_kernel(long long* __restrict__ x, long n, long long v) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] += v;
}

inline int ew_grid(long n) {
    long b = (n + 255) / 256;
    return (int)(b > 148L * 16 ? 148L * 16 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" {

size_t tpgsr_colreduce_workspace_bytes(long M, int C) {
    int rb; long rpb;
    reduce_plan(M, C, rb, rpb);
    return (size_t)rb * 2 * C * sizeof(float);
}

// Training-mode BN statistics of x[M,C] (+ running-stat update) -> mean_rstd[2C], scale_shift[2C]
int tpgsr_bn_stats(const float* x, long M, int C, int ldx, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float momentum, float eps, float* mean_rstd, float* scale_shift, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream) {
    if (!x || !gamma || !beta || !mean_rstd || !scale_shift || !workspace || M <= 0 || C <= 0) return TPGSR_ERR_BAD_ARG;
    if (workspace_bytes < tpgsr_colreduce_workspace_bytes(M, C)) return TPGSR_ERR_WORKSPACE;
    int rb;
    int rc = launch_col_reduce(StatsF{x, ldx > 0 ? ldx : C}, M, C, static_cast<float*>(workspace), rb, stream);
    if (rc) return rc;
    bn_fwd_finalize_kernel<<<ceil_div(C, 32), dim3(32, FIN_Y), 0, stream>>>(static_cast<float*>(workspace), rb, M, C, gamma, beta,
                                                                 mean_rstd, scale_shift, running_mean, running_var, momentum, eps, x);
    return tpgsr_launch_status();
}

int tpgsr_bn_eval_prepare(int C, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                          float eps, float* mean_rstd, float* scale_shift, cudaStream_t stream) {
    if (!gamma || !beta || !running_mean || !running_var || !mean_rstd || !scale_shift) return TPGSR_ERR_BAD_ARG;
    bn_eval_prepare_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(C, gamma, beta, running_mean, running_var, eps, mean_rstd, scale_shift);
    return tpgsr_launch_status();
}

int tpgsr_bn_apply(const float* x, float* y, const float* scale_shift, long M, int C, int ldx, int ldy, int act,
                   const float* residual, int ldr, cudaStream_t stream) {
    if (!x || !y || !scale_shift || (C & 3)) return TPGSR_ERR_BAD_ARG;
    ldx = ldx > 0 ? ldx : C; ldy = ldy > 0 ? ldy : C; ldr = ldr > 0 ? ldr : C;
    if ((ldx & 3) || (ldy & 3) || (ldr & 3)) return TPGSR_ERR_BAD_ARG;
    bn_apply_kernel<<<ew_grid(M * (C / 4)), 256, 0, stream>>>(x, y, scale_shift, M, C, ldx, ldy, act, residual, ldr);
    return tpgsr_launch_status();
}

// BN backward (train: full formula; eval: dx = scale*dy').  act' of the fused activation is applied to dy first.
int tpgsr_bn_bwd(const float* dy, const float* x, float* dx, const float* scale_shift, const float* mean_rstd, long M, int C,
                 int lddy, int ldx, int lddx, int act, int train, float* dgamma, float* dbeta, int accumulate,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!dy || !x || !dx || !scale_shift || !mean_rstd || !workspace) return TPGSR_ERR_BAD_ARG;
    lddy = lddy > 0 ? lddy : C; ldx = ldx > 0 ? ldx : C; lddx = lddx > 0 ? lddx : C;
    size_t need = tpgsr_colreduce_workspace_bytes(M, C) + 2 * (size_t)C * sizeof(float);
    if (workspace_bytes < need) return TPGSR_ERR_WORKSPACE;
    float* partial = static_cast<float*>(workspace);
    float* coef = partial + tpgsr_colreduce_workspace_bytes(M, C) / sizeof(float);
    int rb;
    int rc = launch_col_reduce(BnBwdF{dy, x, scale_shift, mean_rstd, C, ldx, lddy, act}, M, C, partial, rb, stream);
    if (rc) return rc;
    bn_bwd_finalize_kernel<<<ceil_div(C, 32), dim3(32, FIN_Y), 0, stream>>>(partial, rb, M, C, dgamma, dbeta, coef, accumulate);
    rc = tpgsr_launch_status();
    if (rc) return rc;
    bn_bwd_apply_kernel<<<ew_grid(M * C), 256, 0, stream>>>(dy, x, dx, scale_shift, mean_rstd, coef, M, C, lddy, ldx, lddx, act, train);
    return tpgsr_launch_status();
}

// Fused variants (round 2).  `ticket`: one zero-initialised unsigned in device memory, private to the stream the call is enqueued on
// (the kernel leaves it zero).  tpgsr_bn_stats_fused = tpgsr_bn_stats in ONE launch (bit-identical results).
int tpgsr_bn_stats_fused(const float* x, long M, int C, int ldx, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float momentum, float eps, float* mean_rstd, float* scale_shift, void* workspace,
                         size_t workspace_bytes, unsigned* ticket, cudaStream_t stream) {
    if (!x || !gamma || !beta || !mean_rstd || !scale_shift || !workspace || !ticket || M <= 0 || C <= 0) return TPGSR_ERR_BAD_ARG;
    if (workspace_bytes < tpgsr_colreduce_workspace_bytes(M, C)) return TPGSR_ERR_WORKSPACE;
    return launch_col_reduce_fin(StatsF{x, ldx > 0 ? ldx : C}, BnFwdFin{M, gamma, beta, mean_rstd, scale_shift, running_mean, running_var, momentum, eps, x},
                                 M, C, static_cast<float*>(workspace), ticket, stream);
}

// y = act(x*scale + shift) (+ residual) for a dense NHWC map [B,H,W,C], also written as padded-flat bf16 planes (pad P) for the next conv
int tpgsr_bn_apply_planes(const float* x, float* y, const float* scale_shift, int B, int H, int W, int C, int act, const float* residual,
                          int P, void* hi, void* lo, cudaStream_t stream) {
    if (!x || !y || !scale_shift || !hi || (C & 3) || B < 1 || H < 1 || W < 1 || P < 0) return TPGSR_ERR_BAD_ARG;
    bn_apply_planes_kernel<<<ew_grid((long)B * (H + P) * (W + P) * (C / 4)), 256, 0, stream>>>(
        x, y, scale_shift, B, H, W, C, act, residual, P, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo));
    return tpgsr_launch_status();
}

// tpgsr_bn_bwd in TWO launches (fused reduction + finalize, then the apply pass), dense tensors; when `hi` is given, dx is also written
// as padded-flat bf16 planes [B][H+P][W+P][C] (the dY operand of the preceding conv's data / weight gradient)
int tpgsr_bn_bwd_fused(const float* dy, const float* x, float* dx, const float* scale_shift, const float* mean_rstd, int B, int H, int W,
                       int C, int act, int train, float* dgamma, float* dbeta, int accumulate, int P, void* hi, void* lo, void* workspace,
                       size_t workspace_bytes, unsigned* ticket, cudaStream_t stream) {
    if (!dy || !x || !dx || !scale_shift || !mean_rstd || !workspace || (C & 3)) return TPGSR_ERR_BAD_ARG;
    const long M = (long)B * H * W;
    size_t need = tpgsr_colreduce_workspace_bytes(M, C) + 2 * (size_t)C * sizeof(float);
    if (workspace_bytes < need) return TPGSR_ERR_WORKSPACE;
    float* partial = static_cast<float*>(workspace);
    float* coef = partial + tpgsr_colreduce_workspace_bytes(M, C) / sizeof(float);
    int rc;
    if (ticket) {
        rc = launch_col_reduce_fin(BnBwdF{dy, x, scale_shift, mean_rstd, C, C, C, act}, BnBwdFin{M, dgamma, dbeta, coef, accumulate}, M, C, partial,
                                   ticket, stream);
    } else {                                              // ticket == NULL: the stand-alone finalize kernel (measured faster, see header)
        int rb;
        rc = launch_col_reduce(BnBwdF{dy, x, scale_shift, mean_rstd, C, C, C, act}, M, C, partial, rb, stream);
        if (rc) return rc;
        bn_bwd_finalize_kernel<<<ceil_div(C, 32), dim3(32, FIN_Y), 0, stream>>>(partial, rb, M, C, dgamma, dbeta, coef, accumulate);
        rc = tpgsr_launch_status();
    }
    if (rc) return rc;
    if (hi)
        bn_bwd_apply_planes_kernel<<<ew_grid((long)B * (H + P) * (W + P) * (C / 4)), 256, 0, stream>>>(
            dy, x, dx, scale_shift, mean_rstd, coef, B, H, W, C, act, train, P, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo));
    else
        bn_bwd_apply_kernel<<<ew_grid(M * C), 256, 0, stream>>>(dy, x, dx, scale_shift, mean_rstd, coef, M, C, C, C, C, act, train);
    return tpgsr_launch_status();
}

size_t tpgsr_bn_bwd_workspace_bytes(long M, int C) { return tpgsr_colreduce_workspace_bytes(M, C) + 2 * (size_t)C * sizeof(float); }

// out[C] (+)= column sums of x[M,C]  (bias gradients)
int tpgsr_colsum(const float* x, long M, int C, int ldx, float* out, int accumulate, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream) {
    if (!x || !out || !workspace) return TPGSR_ERR_BAD_ARG;
    if (workspace_bytes < tpgsr_colreduce_workspace_bytes(M, C)) return TPGSR_ERR_WORKSPACE;
    int rb;
    int rc = launch_col_reduce(SumF{x, ldx > 0 ? ldx : C}, M, C, static_cast<float*>(workspace), rb, stream);
    if (rc) return rc;
    colsum_finalize_kernel<<<ceil_div(C, 32), dim3(32, FIN_Y), 0, stream>>>(static_cast<float*>(workspace), rb, C, out, accumulate);
    return tpgsr_launch_status();
}

int tpgsr_act_fwd(const float* x, float* y, long n, int act, const float* slope, cudaStream_t stream) {
    if (!x || !y) return TPGSR_ERR_BAD_ARG;
    act_fwd_kernel<<<ew_grid(n), 256, 0, stream>>>(x, y, n, act, slope);
    return tpgsr_launch_status();
}

size_t tpgsr_act_bwd_workspace_bytes(long n) { return (size_t)ew_grid(n) * sizeof(float); }

// dx = dy*act'(ref); for PRELU also dslope (+)= sum(dy*x*[x<0]) (needs workspace)
int tpgsr_act_bwd(const float* dy, const float* ref, float* dx, long n, int act, const float* slope, float* dslope,
                  int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!dy || !ref || !dx) return TPGSR_ERR_BAD_ARG;
    int grid = ew_grid(n);
    float* partial = nullptr;
    if (act == TPGSR_ACT_PRELU && dslope) {
        if (!workspace || workspace_bytes < (size_t)grid * sizeof(float)) return TPGSR_ERR_WORKSPACE;
        partial = static_cast<float*>(workspace);
    }
    act_bwd_kernel<<<grid, 256, 0, stream>>>(dy, ref, dx, n, act, slope, partial);
    int rc = tpgsr_launch_status();
    if (rc || !partial) return rc;
    scalar_finalize_kernel<<<1, 256, 0, stream>>>(partial, grid, dslope, accumulate);
    return tpgsr_launch_status();
}

int tpgsr_add(const float* a, const float* b, float* y, long n, cudaStream_t stream) {
    if (!a || !b || !y) return TPGSR_ERR_BAD_ARG;
    add_kernel<<<ew_grid(n), 256, 0, stream>>>(a, b, y, n);
    return tpgsr_launch_status();
}

// dst window (+)= alpha * src window (pitched [M,C] matrices): concat / slice / scale / residual accumulation
int tpgsr_copy_cols(const float* src, int lds, float* dst, int ldd, long M, int C, float alpha, int accumulate,
                    cudaStream_t stream) {
    if (!src || !dst || M <= 0 || C <= 0) return TPGSR_ERR_BAD_ARG;
    copy_cols_kernel<<<ew_grid(M * C), 256, 0, stream>>>(src, lds, dst, ldd, M, C, alpha, accumulate);
    return tpgsr_launch_status();
}

int tpgsr_fuse_linear(const float* w2, const float* b2, const float* w1, const float* b1, float* wout, float* bout, int N2, int N1, int K1,
                      cudaStream_t stream) {
    if (!w2 || !w1 || !wout || !bout || N2 < 1 || N1 < 1 || K1 < 1) return TPGSR_ERR_BAD_ARG;
    fuse_linear_kernel<<<ew_grid((long)N2 * (K1 + 1)), 256, 0, stream>>>(w2, b2, w1, b1, wout, bout, N2, N1, K1);
    return tpgsr_launch_status();
}

// x[i] += v on an int64 buffer (BatchNorm num_batches_tracked counters kept in one flat buffer)
int tpgsr_add_i64(long long* x, long n, long long v, cudaStream_t stream) {
    if (!x || n <= 0) return TPGSR_ERR_BAD_ARG;
    add_i64_kernel<<<ew_grid(n), 256, 0, stream>>>(x, n, v);
    return tpgsr_launch_status();
}

int tpgsr_maxpool_fwd(const float* x, float* y, int* argmax, int B, int Hi, int Wi, int C, int kh, int kw, int sh, int sw,
                      int ph, int pw, cudaStream_t stream) {
    if (!x || !y) return TPGSR_ERR_BAD_ARG;
    int Ho = (Hi + 2 * ph - kh) / sh + 1, Wo = (Wi + 2 * pw - kw) / sw + 1;
    const long n4 = (long)B * Ho * Wo * (C / 4);
    if ((C & 3) == 0 && (long)B * Hi * Wi * C < (1L << 31) && !(((uintptr_t)x | (uintptr_t)y | (uintptr_t)argmax) & 15)) {
        maxpool_fwd_v4_kernel<<<ew_grid(n4), 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), reinterpret_cast<int4*>(argmax),
                                                                (unsigned)n4, Hi, Wi, C / 4, Ho, Wo, kh, kw, sh, sw, ph, pw);
        return tpgsr_launch_status();
    }
    maxpool_fwd_kernel<<<ew_grid((long)B * Ho * Wo * C), 256, 0, stream>>>(x, y, argmax, B, Hi, Wi, C, Ho, Wo, kh, kw, sh, sw, ph, pw);
    return tpgsr_launch_status();
}

int tpgsr_maxpool_bwd(const float* dy, const int* argmax, float* dx, int B, int Hi, int Wi, int C, int kh, int kw, int sh,
                      int sw, int ph, int pw, cudaStream_t stream) {
    if (!dy || !argmax || !dx) return TPGSR_ERR_BAD_ARG;
    int Ho = (Hi + 2 * ph - kh) / sh + 1, Wo = (Wi + 2 * pw - kw) / sw + 1;
    const long n4 = (long)B * Hi * Wi * (C / 4);
    if ((C & 3) == 0 && (long)B * Hi * Wi * C < (1L << 31) && !(((uintptr_t)dy | (uintptr_t)dx | (uintptr_t)argmax) & 15)) {
        maxpool_bwd_v4_kernel<<<ew_grid(n4), 256, 0, stream>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const int4*>(argmax), reinterpret_cast<float4*>(dx),
                                                                (unsigned)n4, Hi, Wi, C / 4, Ho, Wo, kh, kw, sh, sw, ph, pw);
        return tpgsr_launch_status();
    }
    maxpool_bwd_kernel<<<ew_grid((long)B * Hi * Wi * C), 256, 0, stream>>>(dy, argmax, dx, B, Hi, Wi, C, Ho, Wo, kh, kw, sh, sw, ph, pw);
    return tpgsr_launch_status();
}

int tpgsr_pixelshuffle_mish_fwd(const float* x, float* y, int B, int H, int W, int C, cudaStream_t stream) {
    if (!x || !y) return TPGSR_ERR_BAD_ARG;
    pixelshuffle_mish_fwd_kernel<<<ew_grid((long)B * H * W * C), 256, 0, stream>>>(x, y, B, H, W, C);
    return tpgsr_launch_status();
}

int tpgsr_pixelshuffle_mish_bwd(const float* dy, const float* x, float* dx, int B, int H, int W, int C, cudaStream_t stream) {
    if (!dy || !x || !dx) return TPGSR_ERR_BAD_ARG;
    pixelshuffle_mish_bwd_kernel<<<ew_grid((long)B * H * W * C), 256, 0, stream>>>(dy, x, dx, B, H, W, C);
    return tpgsr_launch_status();
}

// strided NCHW view -> dense NHWC (pitch ldd); if tanh_out != NULL, multiplies by (1 - tanh_out^2) (same strides as src)
int tpgsr_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, long sb, long sc, long sh, long sw, int ldd,
                       const float* tanh_out, cudaStream_t stream) {
    if (!src || !dst) return TPGSR_ERR_BAD_ARG;
    nchw_to_nhwc_kernel<<<ew_grid((long)B * C * H * W), 256, 0, stream>>>(src, dst, B, C, H, W, sb, sc, sh, sw, ldd > 0 ? ldd : C, tanh_out);
    return tpgsr_launch_status();
}

int tpgsr_upsample_nearest_nchw_to_nhwc(const float* src, float* dst, int B, int C, int Hi, int Wi, int scale, long sb, long sc,
                                        long sh, long sw, cudaStream_t stream) {
    if (!src || !dst || scale < 1 || B < 1 || C < 1 || Hi < 1 || Wi < 1) return TPGSR_ERR_BAD_ARG;
    upsample_nearest_nchw_to_nhwc_kernel<<<ew_grid((long)B * C * Hi * Wi * scale * scale), 256, 0, stream>>>(
        src, dst, B, C, Hi * scale, Wi * scale, scale, sb, sc, sh, sw);
    return tpgsr_launch_status();
}

int tpgsr_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, int lds, cudaStream_t stream) {
    if (!src || !dst) return TPGSR_ERR_BAD_ARG;
    nhwc_to_nchw_kernel<<<ew_grid((long)B * C * H * W), 256, 0, stream>>>(src, dst, B, C, H, W, lds > 0 ? lds : C);
    return tpgsr_launch_status();
}

}  // extern "C"
