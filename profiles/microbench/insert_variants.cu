// profiles/microbench/insert_variants.cu — what bounds "bucket load + RED" on an L2-resident table?
// randacc.cu measured ld.256 alone 280 G/s, RED alone 220 G/s, ld.256 -> RED 77 G/s (24 MB table).
// Variants here separate the candidates: dependence (load -> RED), memory-level parallelism per
// thread, same-sector vs different-sector RED, load width, occupancy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o insert_variants insert_variants.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
__device__ __forceinline__ void ld256(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ uint64_t ld64(const uint64_t *p) {
    uint64_t a; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(a) : "l"(p)); return a;
}
__device__ __forceinline__ void red(uint64_t *p) { atomicAdd((unsigned long long *)p, 1ULL << 42); }

// V 0: ld256 -> RED same bucket (dependent)          1: ld256 ; RED same bucket, independent of the loaded value
// V 2: ld256 -> RED in a DIFFERENT random bucket     3: ld64 -> RED same slot
// V 4: batch of B: B x ld256, then B x RED (dependent on its own load)
// V 5: RED only                                       6: ld256 only
// V 7: two REDs (same bucket, two slots)              8: ld256 -> atomicCAS on the slot (always fails: value != expected)
template <int V, int B>
__global__ void __launch_bounds__(256) k(uint64_t *tab, uint64_t nb, uint64_t n, uint64_t *sink) {
    uint64_t acc = 0;
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < n; i0 += stride * B) {
        uint64_t *p[B]; uint64_t a[B], b[B], c[B], d[B];
#pragma unroll
        for (int j = 0; j < B; j++) {
            const uint64_t i = i0 + j * stride;
            p[j] = tab + 4 * __umul64hi(mix(i), nb);
            a[j] = b[j] = c[j] = d[j] = 0;
            if (V == 3) a[j] = ld64(p[j]);
            else if (V != 5 && V != 7) ld256(p[j], a[j], b[j], c[j], d[j]);
        }
#pragma unroll
        for (int j = 0; j < B; j++) {
            const uint64_t i = i0 + j * stride;
            if (i >= n) continue;
            const uint64_t x = a[j] ^ b[j] ^ c[j] ^ d[j];
            acc ^= x;
            if (V == 0 || V == 4) red(p[j] + (x & 3));
            if (V == 1) red(p[j] + (i & 3));
            if (V == 2) red(tab + 4 * __umul64hi(mix(i + x + 0x9e3779b97f4a7c15ULL), nb) + (x & 3));
            if (V == 3) red(p[j]);
            if (V == 5) red(p[j] + (i & 3));
            if (V == 7) { red(p[j] + (i & 1)); red(p[j] + 2 + (i & 1)); }
            if (V == 8) acc ^= atomicCAS((unsigned long long *)(p[j] + (x & 3)), ~0ULL, i);
        }
    }
    if (acc == 0x1234567) sink[0] = acc;
}
// V 9: the two halves as separate kernels — kA finds the slot of every record (ld256) and writes its index, kB streams the
// indices and does the REDs: no kernel mixes loads and atomics
__global__ void __launch_bounds__(256) kA(const uint64_t *tab, uint64_t nb, uint64_t n, uint32_t *slot) {
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t b = __umul64hi(mix(i), nb);
        uint64_t a, c, d, e;
        ld256(tab + 4 * b, a, c, d, e);
        slot[i] = (uint32_t)(4 * b + ((a ^ c ^ d ^ e) & 3));
    }
}
__global__ void __launch_bounds__(256) kB(uint64_t *tab, uint64_t n, const uint32_t *slot) {
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) red(tab + __ldcs(slot + i));
}
// V 10: phases inside a block: 8 loads per thread, barrier, 8 REDs
__global__ void __launch_bounds__(256) kC(uint64_t *tab, uint64_t nb, uint64_t n, uint64_t *sink) {
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    uint64_t acc = 0;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < n; i0 += stride * 8) {
        uint64_t *p[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint64_t a, c, d, e;
            uint64_t *q = tab + 4 * __umul64hi(mix(i0 + j * stride), nb);
            ld256(q, a, c, d, e);
            p[j] = q + ((a ^ c ^ d ^ e) & 3);
            acc ^= a;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; j++) if (i0 + j * stride < n) red(p[j]);
        __syncthreads();
    }
    if (acc == 0x1234567) sink[0] = acc;
}
template <int V, int B> void run(const char *name, int blocks_per_sm, uint64_t *tab, uint64_t nb, uint64_t n, uint64_t *sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V, B><<<148 * blocks_per_sm, 256>>>(tab, nb, n / 8, sink);
    cudaEventRecord(e0);
    k<V, B><<<148 * blocks_per_sm, 256>>>(tab, nb, n, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-64s blocks/SM %d %8.2f ms  %7.2f G/s\n", name, blocks_per_sm, ms, n / ms / 1e6);
}
int main(int argc, char **argv) {
    double mb = argc > 1 ? atof(argv[1]) : 24.0;
    uint64_t n = argc > 2 ? strtoull(argv[2], 0, 10) : 1000000000ULL;
    uint64_t nb = (uint64_t)(mb * 1048576 / 32);
    uint64_t *tab, *sink;
    uint32_t *slot;
    cudaMalloc(&tab, nb * 32); cudaMalloc(&sink, 8); cudaMalloc(&slot, n * 4);
    cudaMemset(tab, 0, nb * 32);
    printf("table %.0f MB, %llu accesses\n", mb, (unsigned long long)n);
    for (int bps : {8, 4, 2}) {
        run<0, 1>("0 ld256 -> RED same bucket (dependent)", bps, tab, nb, n, sink);
        run<1, 1>("1 ld256 ; RED same bucket (independent)", bps, tab, nb, n, sink);
        run<2, 1>("2 ld256 -> RED other bucket", bps, tab, nb, n, sink);
        run<3, 1>("3 ld64 -> RED same slot", bps, tab, nb, n, sink);
        run<4, 2>("4 batch 2: ld256 x2, RED x2", bps, tab, nb, n, sink);
        run<4, 4>("4 batch 4: ld256 x4, RED x4", bps, tab, nb, n, sink);
        run<4, 8>("4 batch 8: ld256 x8, RED x8", bps, tab, nb, n, sink);
        run<5, 1>("5 RED only", bps, tab, nb, n, sink);
        run<6, 1>("6 ld256 only", bps, tab, nb, n, sink);
        run<6, 4>("6 ld256 only, batch 4", bps, tab, nb, n, sink);
        run<7, 1>("7 RED x2 same bucket", bps, tab, nb, n, sink);
        run<8, 1>("8 ld256 -> CAS (fails)", bps, tab, nb, n, sink);
        {
            cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
            kA<<<148 * bps, 256>>>(tab, nb, n / 8, slot);
            cudaEventRecord(e0);
            kA<<<148 * bps, 256>>>(tab, nb, n, slot);
            cudaEventRecord(e1);
            kB<<<148 * bps, 256>>>(tab, n, slot);
            cudaEventRecord(e2); cudaEventSynchronize(e2);
            float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
            printf("%-64s blocks/SM %d %8.2f ms  %7.2f G/s  (loads %.2f ms + REDs %.2f ms)\n", "9 two kernels: ld256 -> index stream ; index stream -> RED", bps, a + b, n / (a + b) / 1e6, a, b);
            kC<<<148 * bps, 256>>>(tab, nb, n / 8, sink);
            cudaEventRecord(e0);
            kC<<<148 * bps, 256>>>(tab, nb, n, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&a, e0, e1);
            printf("%-64s blocks/SM %d %8.2f ms  %7.2f G/s\n", "10 block phases: 8 x ld256, barrier, 8 x RED", bps, a, n / a / 1e6);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
