// All weight re-packs of a train step in ONE launch.
//
// Every tensor-core convolution reads its weights as K-major bf16 hi/lo planes (forward and flipped/transposed data-gradient copies), the
// SIMT convolutions read fp32 [K][Co] copies and the two 9x9 convolutions read kw-stacked planes (kwstack.cu).  Packing them lazily costs
// one ~4 us launch per weight in front of the first convolution that needs it — 47 launches per TSRN step (profiles/r01), all on the
// critical path of the forward pass.  The packed copies only change when the optimiser moves the weights, so a train step keeps them in
// persistent buffers and refreshes all of them right after Adam with this kernel: blockIdx.y selects a job from a device-resident table,
// blockIdx.x strides over that job's elements.  Each job restates the element mapping of the per-weight kernel it replaces
// (pack_conv_weights_bf16_kernel in conv_tc.cu, pack_conv_weights_kernel in conv_simt.cu, kwstack_pack_kernel in kwstack.cu); the GPU tests
// compare the two bit for bit.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

// 80 bytes; mirrored by ops.PackStore (struct format '5Q10i')
struct PackJob {
    const float* w;      // OIHW fp32 source
    void* o0;            // kind 0: fwd hi      kind 1: wk_fwd      kind 2: hi
    void* o1;            // kind 0: fwd lo      kind 1: wk_dgrad    kind 2: lo
    void* o2;            // kind 0: dgrad hi
    void* o3;            // kind 0: dgrad lo
    int kind;            // 0 = bf16 K-major, 1 = fp32 [K][Co], 2 = kw-stacked bf16
    int Co, Ci, KH, KW;  // shape of w
    int C;               // kind 2: channels of the narrow side
    int narrow_out, transposed, flips /* bit0 = kh, bit1 = kw */, rows;   // kind 2
};

__device__ __forceinline__ void split(float v, __nv_bfloat16& h, __nv_bfloat16& l) {
    h = __float2bfloat16_rn(v);
    l = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void __launch_bounds__(256) pack_batched_kernel(const PackJob* __restrict__ jobs) {
    const PackJob j = jobs[blockIdx.y];
    const int Co = j.Co, Ci = j.Ci, KH = j.KH, KW = j.KW;
    const int stride = gridDim.x * blockDim.x;
    const int first = blockIdx.x * blockDim.x + threadIdx.x;
    if (j.kind == 2) {
        __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(j.o0);
        __nv_bfloat16* lo = static_cast<__nv_bfloat16*>(j.o1);
        const int C = j.C, Kp = KH * 64, total = j.rows * Kp;
        const bool flip_kh = j.flips & 1, flip_kw = j.flips & 2;
        for (int idx = first; idx < total; idx += stride) {
            const int n = idx / Kp, kk = idx - n * Kp;
            const int kh = kk / 64, c = kk - kh * 64;
            int o, i, kw;
            if (j.narrow_out) { kw = n / C; o = n - kw * C; i = c; }
            else { o = n; kw = c / C; i = c - kw * C; }
            float v = 0.f;
            if (kw < KW) {
                const int a = j.transposed ? i : o, b = j.transposed ? o : i;
                const int khw = flip_kh ? KH - 1 - kh : kh, kww = flip_kw ? KW - 1 - kw : kw;
                if (a < Co && b < Ci) v = j.w[(((size_t)a * Ci + b) * KH + khw) * KW + kww];
            }
            __nv_bfloat16 h, l;
            split(v, h, l);
            hi[idx] = h;
            if (lo) lo[idx] = l;
        }
        return;
    }
    const int total = Co * Ci * KH * KW;
    for (int idx = first; idx < total; idx += stride) {
        const int kw = idx % KW;
        const int kh = (idx / KW) % KH;
        const int ci = (idx / (KW * KH)) % Ci;
        const int co = idx / (KW * KH * Ci);
        const float v = j.w[idx];
        const int tf = kh * KW + kw, td = (KH - 1 - kh) * KW + (KW - 1 - kw);
        if (j.kind == 1) {
            float* wf = static_cast<float*>(j.o0);
            float* wd = static_cast<float*>(j.o1);
            if (wf) wf[(size_t)(tf * Ci + ci) * Co + co] = v;
            if (wd) wd[(size_t)(td * Co + co) * Ci + ci] = v;
        } else {
            __nv_bfloat16 *f_hi = static_cast<__nv_bfloat16*>(j.o0), *f_lo = static_cast<__nv_bfloat16*>(j.o1);
            __nv_bfloat16 *d_hi = static_cast<__nv_bfloat16*>(j.o2), *d_lo = static_cast<__nv_bfloat16*>(j.o3);
            __nv_bfloat16 h, l;
            split(v, h, l);
            if (f_hi) {
                const size_t o = (size_t)co * (KH * KW * Ci) + (size_t)tf * Ci + ci;
                f_hi[o] = h;
                if (f_lo) f_lo[o] = l;
            }
            if (d_hi) {
                const size_t o = (size_t)ci * (KH * KW * Co) + (size_t)td * Co + co;
                d_hi[o] = h;
                if (d_lo) d_lo[o] = l;
            }
        }
    }
}

}  // namespace

extern "C" {

int tpgsr_pack_job_bytes(void) { return (int)sizeof(PackJob); }

// jobs: device-resident table of n_jobs PackJob records (layout above; built once by the caller — the weights and their packed copies live
// at fixed addresses for the life of a train step).  max_elems = the largest per-job element count (sizes the x dimension of the grid).
int tpgsr_pack_weights_batched(const void* jobs, int n_jobs, int max_elems, cudaStream_t stream) {
    if (!jobs || n_jobs < 1 || n_jobs > 65535 || max_elems < 1) return TPGSR_ERR_BAD_ARG;
    int bx = ceil_div(max_elems, 256 * 4);            // ~4 elements per thread for the largest job
    if (bx > 64) bx = 64;
    if (bx < 1) bx = 1;
    pack_batched_kernel<<<dim3(bx, n_jobs), 256, 0, stream>>>(static_cast<const PackJob*>(jobs));
    return tpgsr_launch_status();
}

}  // extern "C"
