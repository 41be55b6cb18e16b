// Kernel k3: LSD radix sort of a batch of words (8-bit digits, onesweep style): one histogram pass
// over the keys, then one scatter pass per digit whose tile prefixes come from a decoupled look-back,
// so every pass reads and writes each key exactly once.  Replaces the reference's implicit grouping
// of consecutive equal prefixes (chunk_by, src/wordset/mod.rs:147,172,192,223) with a global order.
//
// Algorithmic traffic: n * sizeof(W) * (2 * passes + 1)   (SURVEY section 8d).
#pragma once
#include "scan.cuh"

namespace cbl {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 16;
constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_INCL = 2u << 30, RS_VAL_MASK = (1u << 30) - 1;
constexpr uint64_t RS_MAX_KEYS = (1ull << 30) - 1;

template <class W> struct RsTile { static constexpr int ITEMS = sizeof(W) == 8 ? 16 : 8; static constexpr int TILE = RS_THREADS * ITEMS; };

template <class W> __device__ __forceinline__ uint32_t digit_of(W key, int shift) { return (uint32_t)(key >> shift) & 255u; }

// digit functors for the scatter pass: a byte of the key (LSD sort) or the owner rank of the word's
// prefix (multi-GPU routing: dest = number of splitters <= prefix, i.e. contiguous prefix ranges)
template <class W> struct ByteDigit {
    int shift;
    __device__ __forceinline__ uint32_t operator()(W key) const { return (uint32_t)(key >> shift) & 255u; }
};
constexpr int ROUTE_MAX_SPLIT = 15;
template <class W> struct DestDigit {
    int suffix_bits;
    uint32_t n_split;
    uint32_t split[ROUTE_MAX_SPLIT];
    __device__ __forceinline__ uint32_t operator()(W key) const {
        const uint32_t prefix = (uint32_t)(key >> suffix_bits);
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < ROUTE_MAX_SPLIT; i++) d += (i < (int)n_split) && (split[i] <= prefix);
        return d;
    }
};

// All digit histograms in one read of the keys.  hist[pass][256] (u64, zeroed by the caller).
template <class W>
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const W* __restrict__ keys, uint64_t n, int n_pass,
                                                                unsigned long long* __restrict__ hist) {
    __shared__ uint32_t sh[RS_MAX_PASSES * 256];
    for (int i = threadIdx.x; i < n_pass * 256; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * RS_THREADS;
    const uint64_t n_round = div_up(n, 32) * 32;
    for (uint64_t i = (uint64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        W key = valid ? keys[i] : (W)0;
        for (int p = 0; p < n_pass; p++) {
            uint32_t d = digit_of<W>(key, 8 * p) | (valid ? 0u : 0x100u);
            unsigned peers = __match_any_sync(0xffffffffu, d);
            if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&sh[p * 256 + (d & 255u)], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pass * 256; i += RS_THREADS)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// hist[pass][256] -> exclusive digit bases, in place (one block per pass)
__global__ void __launch_bounds__(256) radix_scan_hist_kernel(unsigned long long* __restrict__ hist) {
    __shared__ unsigned long long tmp[33];
    unsigned long long* h = hist + (size_t)blockIdx.x * 256;
    unsigned long long v = h[threadIdx.x], total;
    unsigned long long e = block_excl_scan<unsigned long long, 256>(v, tmp, total);
    h[threadIdx.x] = e;
}

// One scatter pass.  status[tile][256] must be zero on entry; tile_counter zero.
// HAS_VAL: a u32 payload travels with every key.  pos_out (may be null): receives, for every input
// slot, the slot its key was moved to (used by the multi-GPU router to bring answers back).
template <class W, bool HAS_VAL, class DigitFn>
__global__ void __launch_bounds__(RS_THREADS) radix_pass_kernel(const W* __restrict__ in, W* __restrict__ out,
                                                                const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                                uint64_t n, DigitFn digit, const unsigned long long* __restrict__ digit_base,
                                                                volatile uint32_t* status, uint32_t* tile_counter,
                                                                uint32_t* __restrict__ pos_out) {
    constexpr int ITEMS = RsTile<W>::ITEMS;
    constexpr int TILE = RsTile<W>::TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* s_keys = reinterpret_cast<W*>(smem_raw);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(W) * TILE);  // only if HAS_VAL
    __shared__ uint32_t s_whist[RS_WARPS][256];
    __shared__ uint32_t s_dstart[256];
    __shared__ long long s_goff[256];
    __shared__ uint32_t s_tmp[33];
    __shared__ uint32_t s_tile;

    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const int tile_n = (int)min((uint64_t)TILE, n - tile_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&s_whist[0][0])[i] = 0;
    __syncthreads();

    // 1. load (warp-striped: order = warp, item, lane) and rank inside the warp (stable)
    W key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rnk[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int local = warp * (32 * ITEMS) + i * 32 + lane;
        const bool valid = local < tile_n;
        key[i] = valid ? in[tile_base + local] : (W)0;
        if (HAS_VAL) val[i] = valid ? vin[tile_base + local] : 0u;
        const uint32_t d = digit(key[i]);
        const unsigned peers = __match_any_sync(0xffffffffu, d | (valid ? 0u : 0x100u));
        const uint32_t lt = __popc(peers & lanemask_lt());
        uint32_t base = valid ? s_whist[warp][d] : 0u;
        __syncwarp();
        if (valid && lt == 0) s_whist[warp][d] = base + __popc(peers);
        __syncwarp();
        rnk[i] = base + lt;
    }
    __syncthreads();

    // 2. per digit (thread t = digit t): exclusive offsets over warps, tile count, look-back
    {
        const int t = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = s_whist[w][t];
            s_whist[w][t] = run;
            run += c;
        }
        if (tile == 0) status[t] = RS_FLAG_INCL | run;
        else status[(size_t)tile * 256 + t] = RS_FLAG_AGG | run;
        uint32_t total;
        uint32_t dstart = block_excl_scan<uint32_t, RS_THREADS>(run, s_tmp, total);
        s_dstart[t] = dstart;
        uint32_t excl = 0;
        if (tile > 0) {
            for (long long prev = (long long)tile - 1; prev >= 0; prev--) {
                uint32_t s;
                do { s = status[(size_t)prev * 256 + t]; } while ((s >> 30) == 0);
                excl += s & RS_VAL_MASK;
                if (s & RS_FLAG_INCL) break;
            }
            status[(size_t)tile * 256 + t] = RS_FLAG_INCL | (excl + run);
        }
        s_goff[t] = (long long)digit_base[t] + (long long)excl - (long long)dstart;
    }
    __syncthreads();

    // 3. local scatter into digit order
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int local = warp * (32 * ITEMS) + i * 32 + lane;
        if (local < tile_n) {
            const uint32_t d = digit(key[i]);
            const uint32_t pos = s_dstart[d] + s_whist[warp][d] + rnk[i];
            s_keys[pos] = key[i];
            if (HAS_VAL) s_vals[pos] = val[i];
            if (pos_out) pos_out[tile_base + local] = (uint32_t)(s_goff[d] + (long long)pos);
        }
    }
    __syncthreads();

    // 4. coalesced global scatter: consecutive smem slots of one digit go to consecutive addresses
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int idx = i * RS_THREADS + threadIdx.x;
        if (idx < tile_n) {
            const W k = s_keys[idx];
            const uint32_t d = digit(k);
            const long long g = s_goff[d] + idx;
            out[g] = k;
            if (HAS_VAL) vout[g] = s_vals[idx];
        }
    }
}

// per-destination counts for the router (<= 16 destinations): registers -> warp reduce -> atomics
template <class W>
__global__ void __launch_bounds__(256) route_hist_kernel(const W* __restrict__ keys, uint64_t n, DestDigit<W> digit,
                                                         unsigned long long* __restrict__ hist /*[256], zeroed*/) {
    uint32_t cnt[ROUTE_MAX_SPLIT + 1];
#pragma unroll
    for (int i = 0; i <= ROUTE_MAX_SPLIT; i++) cnt[i] = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t d = digit(keys[i]);
#pragma unroll
        for (int j = 0; j <= ROUTE_MAX_SPLIT; j++) cnt[j] += (d == (uint32_t)j);
    }
#pragma unroll
    for (int j = 0; j <= ROUTE_MAX_SPLIT; j++) {
        uint32_t c = warp_sum(cnt[j]);
        if (lane_id() == 0 && c) atomicAdd(&hist[j], (unsigned long long)c);
    }
}

__global__ void gather_u8_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ pos, uint64_t n, uint8_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[pos[i]];
}

}  // namespace cbl
