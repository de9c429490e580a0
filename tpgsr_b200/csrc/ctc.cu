// CTC greedy decode on the device: arg-max over the classes of every time step, collapse repeats, drop blanks —
// utils/metrics.py:71-88 (get_string_crnn), which the reference runs per batch in eval() as a Python loop of tensor.item() calls over
// [B][T] elements after copying the logits to the host.  Here one CTA per sample produces the decoded class ids and their count; the host
// fetches B*(T+1) ints in one copy and maps them through the alphabet.
//
// Integer/index work, HBM-bound and tiny (26 x 48 x 37 floats per CRNN batch): phase 1 reads the logits coalesced (one warp per time step,
// lanes across the classes) and leaves the arg-max of each time step in shared memory; phase 2 is a warp scan (ballot + popc compaction).
// Ties resolve to the lowest class index (what the CPU torch.max the fixtures were recorded with returns); NaN logits never win a comparison.
#include "common.cuh"
#include <limits.h>

namespace {

constexpr int CTC_THREADS = 128;

__global__ void __launch_bounds__(CTC_THREADS)
ctc_greedy_kernel(const float* __restrict__ logits, int T, int B, int C, int* __restrict__ ids, int* __restrict__ lens) {
    extern __shared__ int arg[];                  // [T] arg-max per time step
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < T; t += CTC_THREADS / 32) {
        const float* row = logits + ((size_t)t * B + b) * C;
        float bv = -INFINITY;
        int bi = INT_MAX;
        for (int c = lane; c < C; c += 32) {
            const float v = __ldg(row + c);
            if (bi == INT_MAX || v > bv) { bv = v; bi = c; }       // first candidate of the lane, then strictly greater only
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != INT_MAX && (bi == INT_MAX || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) arg[t] = bi;
    }
    __syncthreads();
    if (warp != 0) return;
    int n = 0;           // ids written so far (warp-uniform)
    int carry = 0;       // arg-max of the time step before this chunk ("last" starts out empty, like a blank)
    int* out = ids + (size_t)b * T;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const int cur = t < T ? arg[t] : 0;
        int prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (lane == 0) prev = carry;
        const bool keep = t < T && cur != 0 && cur != prev;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) out[n + __popc(m & ((1u << lane) - 1u))] = cur;
        n += __popc(m);
        const int last = T - 1 - t0 < 31 ? T - 1 - t0 : 31;
        carry = __shfl_sync(0xffffffffu, cur, last);
    }
    for (int i = n + lane; i < T; i += 32) out[i] = -1;
    if (lane == 0) lens[b] = n;
}

}  // namespace

extern "C" {

// logits: [T][B][C] fp32 (the CRNN's output layout, model/crnn/crnn.py:74-90); ids: [B][T] int32, the decoded class indices of sample b in
// ids[b][0 .. lens[b]) and -1 behind them; class 0 is the CTC blank.
int tpgsr_ctc_greedy_decode(const float* logits, int T, int B, int C, int* ids, int* lens, cudaStream_t stream) {
    if (!logits || !ids || !lens || T < 1 || B < 1 || C < 1) return TPGSR_ERR_BAD_ARG;
    if ((size_t)T * sizeof(int) > 48 * 1024) return TPGSR_ERR_UNSUPPORTED;
    ctc_greedy_kernel<<<B, CTC_THREADS, (size_t)T * sizeof(int), stream>>>(logits, T, B, C, ids, lens);
    return tpgsr_launch_status();
}

}  // extern "C"
