// profiles/microbench/randacc.cu — what does ONE random table access cost on B200?
// Measures accesses/s for different load widths / L2 hints / atomics on a table far larger than
// L2; run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,...` to get DRAM bytes
// per access. Not part of the product; it only grounds DESIGN.md's random-access roofline.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t *tab, uint64_t nb, int iters, uint64_t *sink) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t acc = 0;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        uint64_t b = __umul64hi(mix(t * 1000003ULL + i), nb);
        uint64_t *p = tab + 4 * b;
        if (MODE == 0) { acc ^= *(volatile uint32_t *)p; }
        else if (MODE == 1) { acc ^= *(volatile uint64_t *)p; }
        else if (MODE == 2) { uint64_t a, c; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(c) : "l"(p)); acc ^= a ^ c; }
        else if (MODE == 3) { uint64_t a, c, d, e; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(c) : "l"(p));
                              asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(d), "=l"(e) : "l"(p + 2)); acc ^= a ^ c ^ d ^ e; }
        else if (MODE == 4) { uint64_t a, c, d, e; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(c), "=l"(d), "=l"(e) : "l"(p)); acc ^= a ^ c ^ d ^ e; }
        else if (MODE == 5) { uint64_t a, c, d, e; asm volatile("ld.global.cg.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(c), "=l"(d), "=l"(e) : "l"(p)); acc ^= a ^ c ^ d ^ e; }
        else if (MODE == 6) { atomicAdd((unsigned long long *)p, 1ULL << 42); }
        else if (MODE == 7) { uint64_t a, c, d, e; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(c), "=l"(d), "=l"(e) : "l"(p));
                              acc ^= a ^ c ^ d ^ e; atomicAdd((unsigned long long *)(p + (acc & 3)), 1ULL << 42); }
        else if (MODE == 8) { uint64_t a, c, d, e; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(c) : "l"(p));
                              asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(d), "=l"(e) : "l"(p + 2));
                              acc ^= a ^ c ^ d ^ e; atomicAdd((unsigned long long *)(p + (acc & 3)), 1ULL << 42); }
        else if (MODE == 9) { acc ^= atomicAdd((unsigned long long *)p, 1ULL << 42); }
        else if (MODE == 10) { uint64_t a, c, d, e; asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(c) : "l"(p));
                              asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(d), "=l"(e) : "l"(p + 2)); acc ^= a ^ c ^ d ^ e; }
    }
    if (acc == 0x1234567) sink[0] = acc;
}
template <int MODE> void run(const char *name, uint64_t *tab, uint64_t nb, uint64_t *sink) {
    int iters = 64, blocks = 148 * 8 * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(tab, nb, 8, sink);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(tab, nb, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * 256 * iters;
    printf("%-34s %8.2f ms  %7.2f G access/s  (x32B = %6.0f GB/s)\n", name, ms, n / ms / 1e6, n * 32 / ms / 1e6);
}
int main(int argc, char **argv) {
    double gb = argc > 1 ? atof(argv[1]) : 8.0;
    uint64_t nb = (uint64_t)(gb * 1e9 / 32);
    uint64_t *tab, *sink;
    cudaMalloc(&tab, nb * 32); cudaMalloc(&sink, 8);
    cudaMemset(tab, 0, nb * 32);
    printf("table %.1f GB, %llu buckets of 32 B\n", gb, (unsigned long long)nb);
    run<0>("0 ld.32", tab, nb, sink);
    run<1>("1 ld.64", tab, nb, sink);
    run<2>("2 ld.cg.128", tab, nb, sink);
    run<3>("3 2x ld.cg.128 (32B)", tab, nb, sink);
    run<4>("4 ld.cg.256", tab, nb, sink);
    run<5>("5 ld.cg.L2::64B.256", tab, nb, sink);
    run<10>("10 2x ld.nc.noalloc.128 (32B)", tab, nb, sink);
    run<6>("6 red.add.64", tab, nb, sink);
    run<9>("9 atom.add.64 (returns)", tab, nb, sink);
    run<7>("7 ld.256 + red.add.64", tab, nb, sink);
    run<8>("8 2x ld.128 + red.add.64", tab, nb, sink);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
