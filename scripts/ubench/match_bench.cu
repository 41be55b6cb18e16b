// micro-benchmark: warp-level peer detection (match.any vs ballot emulation) and shared atomics throughput
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned peers_ballot(uint32_t d) {
    unsigned m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1);
        m &= ((d >> b) & 1) ? v : ~v;
    }
    return m;
}
template <int MODE>
__global__ void k(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t sh[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    uint32_t x = in[blockIdx.x * blockDim.x + threadIdx.x] ^ ((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u); x ^= x >> 13; x *= 0x9E3779B1u; x ^= x >> 16;
    uint32_t acc = 0;
    const int warp = threadIdx.x >> 5;
    for (int it = 0; it < iters; it++) {
        uint32_t d = (x >> 11) & 255u;
        if (MODE == 0) acc += __popc(__match_any_sync(0xffffffffu, d));
        if (MODE == 1) acc += __popc(peers_ballot(d));
        if (MODE == 2) acc += atomicAdd(&sh[warp][d], 1u);
        if (MODE == 3) { unsigned p = peers_ballot(d); if ((p & ((1u << (threadIdx.x & 31)) - 1)) == 0) acc += atomicAdd(&sh[warp][d], __popc(p)); }
        x = x * 1664525u + 1013904223u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    uint32_t *in, *out;
    CK(cudaMalloc(&in, blocks * threads * 4)); CK(cudaMalloc(&out, blocks * threads * 4));
    CK(cudaMemset(in, 0x5a, blocks * threads * 4));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[] = {"match_any", "ballot x8", "smem atomicAdd (ret)", "ballot x8 + leader atomic"};
    for (int mode = 0; mode < 4; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(a);
            if (mode == 0) k<0><<<blocks, threads>>>(in, out, iters);
            if (mode == 1) k<1><<<blocks, threads>>>(in, out, iters);
            if (mode == 2) k<2><<<blocks, threads>>>(in, out, iters);
            if (mode == 3) k<3><<<blocks, threads>>>(in, out, iters);
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            double winst = (double)blocks * threads / 32 * iters;
            if (rep) printf("%-28s %8.3f ms  %.2f G warp-ops/s  (%.1f SM-cycles per warp-op at 1.9 GHz)\n", names[mode], ms, winst / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 / winst);
        }
    }
    return 0;
}
