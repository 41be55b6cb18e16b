// DRAM fetch granularity and L2 residency on B200: random aligned 32-byte window reads over a large array, optionally
// mixed with random 8-byte table reads over a small table (the probe's access pattern).  Run under ncu:
//   ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum ./dram_gran
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
// WB: bytes per window access (32 or 64); TABLE: also read 8 B at a random table slot (table_bytes) per access, n_tab times
__device__ __forceinline__ uint4 ld64b(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int WB, int NTAB, int LD = 0>
__global__ void probe_like(const uint4* __restrict__ big, uint64_t n_win, const uint2* __restrict__ tab, uint64_t n_tab_entries, uint64_t n, uint32_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = mix(i + 12345);
        uint64_t w = h % n_win;
        const uint4* p = big + w * (WB / 16);
#pragma unroll
        for (int c = 0; c < WB / 16; c++) { uint4 v = LD == 1 ? ld64b(p + c) : __ldg(p + c); acc += v.x ^ v.y ^ v.z ^ v.w; }
#pragma unroll
        for (int t = 0; t < NTAB; t++) { uint2 e = __ldg(tab + mix(h + t) % n_tab_entries); acc += e.x + e.y; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
int main(int argc, char** argv) {
    if (argc > 1) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1])); size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("set fetch granularity %s: %s, now %zu\n", argv[1], cudaGetErrorString(e), g); }
    { size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity %zu\n", g); }
    const uint64_t big_bytes = 4ull << 30, n = 1ull << 28;
    uint4* big; uint2* tab; uint32_t* out;
    cudaMalloc(&big, big_bytes); cudaMemset(big, 1, big_bytes);
    const uint64_t tab_bytes_max = 64ull << 20;
    cudaMalloc(&tab, tab_bytes_max); cudaMemset(tab, 1, tab_bytes_max);
    cudaMalloc(&out, 4);
    const int grid = 148 * 16, block = 256;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char* name, auto kern, uint64_t n_win, uint64_t tab_entries) {
        kern<<<grid, block>>>(big, n_win, tab, tab_entries, n, out);
        cudaEventRecord(a);
        kern<<<grid, block>>>(big, n_win, tab, tab_entries, n, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-40s %.3f ms  %.1f G acc/s\n", name, ms, n / ms * 1e-6);
    };
    run("win32 only", probe_like<32, 0>, big_bytes / 32, 1);
    run("win32 only, ld .L2::64B", probe_like<32, 0, 1>, big_bytes / 32, 1);
    run("win64 only", probe_like<64, 0>, big_bytes / 64, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
