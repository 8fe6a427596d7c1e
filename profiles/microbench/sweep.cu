// profiles/microbench/sweep.cu — does a streaming read defeat L2 residency of a small hot table?
// Emulates insert_bins_kernel: each thread streams records (8 B + 4 B) and does one 256-bit bucket
// load + RED.ADD.64 into a table of `mb` megabytes. Variants add L2 eviction-priority hints.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }

// MODE 0: table only   1: + stream (ld.cs)   2: + stream, table ops with evict_last policy
// MODE 3: + stream with evict_first policy (table default)   4: both hints   5: stream via ld.nc.L1::no_allocate
// MODE 6: like 1 but stream read with plain ld
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t *tab, uint64_t nb, const uint64_t *s8, const uint32_t *s4, uint64_t n, uint64_t *sink) {
    uint64_t acc = 0;
    uint64_t pl = 0, pf = 0;
    if (MODE == 2 || MODE == 4) pl = pol_last();
    if (MODE == 3 || MODE == 4) pf = pol_first();
    uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t r = i;
        if (MODE == 1) { r = __ldcs(s8 + i); acc += __ldcs(s4 + i); }
        if (MODE == 2) { r = __ldcs(s8 + i); acc += __ldcs(s4 + i); }
        if (MODE == 3 || MODE == 4) {
            uint32_t t;
            asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(s8 + i), "l"(pf));
            asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(t) : "l"(s4 + i), "l"(pf));
            acc += t;
        }
        if (MODE == 5) { uint32_t t; asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(s8 + i));
                         asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(t) : "l"(s4 + i)); acc += t; }
        if (MODE == 6) { r = s8[i]; acc += s4[i]; }
        uint64_t b = __umul64hi(mix(r + i), nb);
        uint64_t *p = tab + 4 * b;
        uint64_t a, c, d, e;
        if (MODE == 2 || MODE == 4) {
            asm volatile("ld.global.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(c), "=l"(d), "=l"(e) : "l"(p), "l"(pl));
            acc ^= a ^ c ^ d ^ e;
            asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(p + (acc & 3)), "l"(1ULL << 42), "l"(pl) : "memory");
        } else {
            asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(c), "=l"(d), "=l"(e) : "l"(p));
            acc ^= a ^ c ^ d ^ e;
            atomicAdd((unsigned long long *)(p + (acc & 3)), 1ULL << 42);
        }
    }
    if (acc == 0x1234567) sink[0] = acc;
}
template <int MODE> void run(const char *name, uint64_t *tab, uint64_t nb, uint64_t *s8, uint32_t *s4, uint64_t n, uint64_t *sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(tab, nb, s8, s4, n / 8, sink);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(tab, nb, s8, s4, n, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-52s %8.2f ms  %7.2f G rec/s\n", name, ms, n / ms / 1e6);
}
int main(int argc, char **argv) {
    double mb = argc > 1 ? atof(argv[1]) : 24.0;
    uint64_t n = argc > 2 ? strtoull(argv[2], 0, 10) : 400000000ULL;
    uint64_t nb = (uint64_t)(mb * 1048576 / 32);
    uint64_t *tab, *sink, *s8; uint32_t *s4;
    cudaMalloc(&tab, nb * 32); cudaMalloc(&sink, 8); cudaMalloc(&s8, n * 8); cudaMalloc(&s4, n * 4);
    cudaMemset(tab, 0, nb * 32); cudaMemset(s8, 1, n * 8); cudaMemset(s4, 1, n * 4);
    printf("table %.0f MB, %llu records (stream %.1f GB)\n", mb, (unsigned long long)n, n * 12 / 1e9);
    run<0>("0 table only (ld.cg.256 + red)", tab, nb, s8, s4, n, sink);
    run<1>("1 + stream ld.cs", tab, nb, s8, s4, n, sink);
    run<6>("6 + stream plain ld", tab, nb, s8, s4, n, sink);
    run<5>("5 + stream ld.nc.L1::no_allocate", tab, nb, s8, s4, n, sink);
    run<3>("3 + stream L2::evict_first policy", tab, nb, s8, s4, n, sink);
    run<2>("2 + stream ld.cs, table L2::evict_last policy", tab, nb, s8, s4, n, sink);
    run<4>("4 + stream evict_first, table evict_last", tab, nb, s8, s4, n, sink);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
