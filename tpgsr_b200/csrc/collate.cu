// Label one-hot collate on the device (second half): the text-label half of the reference's collate functions —
// dataset/dataset.py:1239-1323 (alignCollate_realWTLAMask.__call__; the same loop at :1172-1217 and :1360-1407): every label string becomes a
// [len, 37] one-hot matrix over the alphabet "-0123456789abcdefghijklmnopqrstuvwxyz", the batch is zero-padded to the longest word and permuted
// to [B, 37, 1, max_len].  The reference builds it with a Python loop of torch.zeros / scatter_ / slice assignments per sample.  Here the host
// maps characters to class codes (string handling stays on the host, as in the reference), ONE int32 [B][L] array crosses the bus, and one
// kernel zero-fills and scatters the batch tensor.  Integer / index work: 4 B read per character, 4*37 B written per character slot.
#include "common.cuh"

namespace {

// codes: [B][L] int32, class index in [0, C) or -1 behind the end of a label; out: [B][C][1][L] fp32
__global__ void __launch_bounds__(256)
label_onehot_kernel(const int* __restrict__ codes, int B, int L, int C, float* __restrict__ out) {
    const long total = (long)B * C * L;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int t = (int)(i % L);
        const long q = i / L;
        const int c = (int)(q % C);
        const int b = (int)(q / C);
        out[i] = __ldg(codes + (size_t)b * L + t) == c ? 1.f : 0.f;
    }
}

}  // namespace

extern "C" {

int tpgsr_label_onehot(const int* codes, int B, int L, int C, float* out, cudaStream_t stream) {
    if (!codes || !out || B < 1 || L < 1 || C < 1) return TPGSR_ERR_BAD_ARG;
    const long total = (long)B * C * L;
    const long blocks = (total + 255) / 256;
    label_onehot_kernel<<<(int)(blocks > 148 * 8 ? 148 * 8 : blocks), 256, 0, stream>>>(codes, B, L, C, out);
    return tpgsr_launch_status();
}

}  // extern "C"
