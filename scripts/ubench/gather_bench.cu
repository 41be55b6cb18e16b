// What bounds the membership probe on B200?  Random look-ups shaped like the probe of index_view.cuh: up to three
// dependent table reads (L2-resident tables of 4 / 12 / 31 MB) followed by one random 32-byte suffix window out of
// a 2 GB array, 64-thread CTAs at 24 CTAs/SM like seq_words_kernel MODE 3.  Variants differ only in the NUMBER OF LOAD
// INSTRUCTIONS per look-up (two 16-byte loads vs one 32-byte load per window; two correction bytes vs one aligned
// word; separate vs merged tables), which is what the L1/LSU request rate cares about, not in bytes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu && ./gather_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint4 ld128(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol_first()));
    return v;
}
__device__ __forceinline__ uint32_t ld256x(const uint4* p) {
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p), "l"(pol_first()));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ uint2 ldtab(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol_last()));
    return v;
}
__device__ __forceinline__ int ldb(const int8_t* p) {
    int v;
    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.s8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol_last()));
    return v;
}
__device__ __forceinline__ uint32_t ldw(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol_last()));
    return v;
}
// WIN: 0 none, 1 two 16-byte loads, 2 one 32-byte load;  NWIN: windows per look-up (2 = a second, dependent, adjacent-ish window for half the lanes)
// T1/T2: read the 4 MB / 12 MB tables;  SUBM: 0 none, 1 two byte loads, 2 one aligned word
template <int WIN, int T1, int T2, int SUBM, int SECOND>
__global__ void __launch_bounds__(64, 24) probe_like(const uint4* __restrict__ big, uint64_t n_win, const uint2* __restrict__ t1, uint32_t n1,
                                                     const uint2* __restrict__ t2, uint32_t n2, const int8_t* __restrict__ t3, uint32_t n3,
                                                     uint64_t n, uint32_t* out) {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = mix(i + 12345);
        uint32_t x = (uint32_t)h, y = (uint32_t)(h >> 32);
        if (T1) { const uint2 e = ldtab(t1 + x % n1); y += e.x & 1; x += e.y & 1; }           // values are all ones: the dependency is real, the index stays random
        if (T2) { const uint2 e = ldtab(t2 + y % n2); x += e.x & 1; y ^= e.y & 1; }
        if (SUBM == 1) { const uint32_t j = x % (n3 - 4); const int a = ldb(t3 + j), b = ldb(t3 + j + 1); y += (uint32_t)(a + b) & 1; }
        if (SUBM == 2) { const uint32_t j = x % (n3 - 4); const uint32_t w = ldw(reinterpret_cast<const uint32_t*>(t3) + (j >> 2)); y += w & 1; }
        const uint64_t w = (((uint64_t)x << 32) | y) % n_win;
        const uint4* p = big + w * 2;
        if (WIN == 1) { const uint4 a = ld128(p), b = ld128(p + 1); acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w; }
        if (WIN == 2) acc += ld256x(p);
        if (SECOND && ((acc ^ x) & 1)) {                                                        // half the lanes need the neighbouring window too
            const uint4* q = big + ((w ^ 1) % n_win) * 2;
            if (WIN == 1) { const uint4 a = ld128(q), b = ld128(q + 1); acc += a.x ^ b.w; }
            if (WIN == 2) acc += ld256x(q);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
int main(int argc, char** argv) {
    const uint64_t big_bytes = 2ull << 30, n = 1ull << 29;
    uint4* big; uint2 *t1, *t2; int8_t* t3; uint32_t* out;
    const uint32_t n1 = (4u << 20) / 8, n2 = (12u << 20) / 8, n3 = 31u << 20;
    cudaMalloc(&big, big_bytes); cudaMemset(big, 1, big_bytes);
    cudaMalloc(&t1, n1 * 8ull); cudaMemset(t1, 1, n1 * 8ull);
    cudaMalloc(&t2, n2 * 8ull); cudaMemset(t2, 1, n2 * 8ull);
    cudaMalloc(&t3, n3); cudaMemset(t3, 1, n3);
    cudaMalloc(&out, 4);
    int grid_mult = argc > 1 ? atoi(argv[1]) : 24;
    const int grid = 148 * grid_mult * 8, block = 64;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char* name, auto kern) {
        kern<<<grid, block>>>(big, big_bytes / 32, t1, n1, t2, n2, t3, n3, n, out);
        cudaEventRecord(a);
        kern<<<grid, block>>>(big, big_bytes / 32, t1, n1, t2, n2, t3, n3, n, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-64s %8.3f ms  %6.1f G look-ups/s  (%.2f ms per 1 G)\n", name, ms, n / ms * 1e-6, ms * (double)(1ull << 30) / n * 1000.0 / 1024.0 * 1.024);
    };
    run("window only, 2 x 16 B loads", probe_like<1, 0, 0, 0, 0>);
    run("window only, 1 x 32 B load", probe_like<2, 0, 0, 0, 0>);
    run("window + half a second window, 2 x 16 B", probe_like<1, 0, 0, 0, 1>);
    run("window + half a second window, 1 x 32 B", probe_like<2, 0, 0, 0, 1>);
    run("tables only: t1 -> t2 -> 2 bytes", probe_like<0, 1, 1, 1, 0>);
    run("tables only: t1 -> t2 -> 1 word", probe_like<0, 1, 1, 2, 0>);
    run("tables only: t2 -> 1 word", probe_like<0, 0, 1, 2, 0>);
    run("full: t1 -> t2 -> 2 bytes -> 2 x 16 B (+ half second)  [today]", probe_like<1, 1, 1, 1, 1>);
    run("full: t1 -> t2 -> 1 word -> 1 x 32 B (+ half second)", probe_like<2, 1, 1, 2, 1>);
    run("full: t1 -> t2 -> 1 word -> 1 x 32 B (no second)", probe_like<2, 1, 1, 2, 0>);
    run("full: t2 -> 1 word -> 1 x 32 B (no second)", probe_like<2, 0, 1, 2, 0>);
    run("full: t2 -> 1 x 32 B (no second)", probe_like<2, 0, 1, 0, 0>);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
