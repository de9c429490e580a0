// Micro-benchmark (sm_100a): cost of the building blocks of one GRU step for one warp per scheduler:
// chained / independent mma.sync m16n8k16 bf16, ldmatrix, MUFU gate chains, bar.sync of 128 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void k(long long* out, float* sink, int iters) {
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u};
    uint32_t b0 = 0x3f803f80u, b1 = 0x3f803f80u;
    float d[6][4] = {};
    __shared__ __align__(16) __nv_bfloat16 hb[16][40];
    for (int i = threadIdx.x; i < 16 * 40; i += blockDim.x) hb[i / 40][i % 40] = __float2bfloat16(0.01f * i);
    __syncthreads();
    long long t0, t1;
    // 1. one chain of dependent HMMAs
    t0 = clock64();
    for (int i = 0; i < iters; ++i) mma(d[0], a, b0, b1);
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
    // 2. 6 independent chains x 3 (the step's 18 MMAs)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) mma(d[c], a, b0 + r, b1);
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0) / iters;
    // 3. ldmatrix x4 dependent on address chain
    uint32_t r4[4]; uint32_t addr = (uint32_t)__cvta_generic_to_shared(&hb[threadIdx.x & 15][0]);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r4[0]), "=r"(r4[1]), "=r"(r4[2]), "=r"(r4[3]) : "r"(addr));
        addr += (r4[0] & 0);      // dependency
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0) / iters;
    // 4. gate chain: sigmoid -> tanh -> update
    float x = d[0][0] * 1e-30f + 0.3f, h = 0.1f;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        float r = __fdividef(1.f, 1.f + __expf(-x));
        float n = 1.f - __fdividef(2.f, 1.f + __expf(2.f * (x + r * h)));
        h = (1.f - r) * n + r * h;
        x = h * 0.5f;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0) / iters;
    // 5. bar.sync of the CTA
    t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0) / iters;
    // 6. bf16 split + STS + bar + ldmatrix round trip
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        __nv_bfloat16 hh = __float2bfloat16_rn(h);
        __nv_bfloat16 ll = __float2bfloat16_rn(h - __bfloat162float(hh));
        hb[threadIdx.x & 15][threadIdx.x >> 4] = hh; hb[(threadIdx.x + 1) & 15][8 + (threadIdx.x >> 4)] = ll;
        __syncthreads();
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r4[0]), "=r"(r4[1]), "=r"(r4[2]), "=r"(r4[3]) : "r"(addr));
        h += __uint_as_float(r4[0] & 0);
        __syncthreads();
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[5] = (t1 - t0) / iters;
    sink[threadIdx.x] = d[0][0] + d[1][1] + d[2][2] + d[3][3] + d[4][0] + d[5][1] + h + x + __uint_as_float(r4[1]);
}
int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
    k<<<1, 128>>>(out, sink, 2000);
    long long h[8];
    cudaMemcpy(h, out, 48, cudaMemcpyDeviceToHost);
    printf("cycles per iteration (1 CTA of 4 warps = one warp per scheduler):\n");
    printf("  dependent HMMA m16n8k16 chain:          %lld\n", h[0]);
    printf("  18 HMMAs (6 chains x 3):                %lld\n", h[1]);
    printf("  ldmatrix.x4 (dependent):                %lld\n", h[2]);
    printf("  sigmoid -> tanh -> h update chain:      %lld\n", h[3]);
    printf("  __syncthreads (128 threads):            %lld\n", h[4]);
    printf("  split + STS + bar + ldmatrix + bar:     %lld\n", h[5]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
