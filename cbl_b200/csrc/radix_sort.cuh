// Kernel k3: LSD radix sort of a batch of words (8-bit digits, onesweep style): one histogram pass
// over the keys, then one scatter pass per digit whose tile prefixes come from a decoupled look-back,
// so every pass reads and writes each key exactly once.  Replaces the reference's implicit grouping
// of consecutive equal prefixes (chunk_by, src/wordset/mod.rs:147,172,192,223) with a global order.
//
// Algorithmic traffic: n * sizeof(W) * (2 * passes + 1)   (SURVEY section 8d).
#pragma once
#include "scan.cuh"

namespace cbl {

#ifndef CBL_RS_MIN_BLOCKS
#define CBL_RS_MIN_BLOCKS 4
#endif
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
    static constexpr bool WANTS_POS = false;   // the sort never asks where a key went
    int shift;
    __device__ __forceinline__ uint32_t operator()(W key) const { return (uint32_t)(key >> shift) & 255u; }
};
constexpr int ROUTE_MAX_SPLIT = 15;
template <class W> struct DestDigit {
    static constexpr bool WANTS_POS = true;
    int suffix_bits;
    uint32_t n_split;
    uint32_t split[ROUTE_MAX_SPLIT];
    __device__ __forceinline__ uint32_t operator()(W key) const {
        const uint32_t prefix = (uint32_t)(key >> suffix_bits);
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < ROUTE_MAX_SPLIT; i++) d += split[i] <= prefix;   // unused splitters hold 0xFFFFFFFF > any prefix (PREFIX_BITS <= 31)
        return d;
    }
};

// routing straight into the owners' receive buffers: one output base pointer per destination (peer
// memory mapped with CUDA IPC; the stores travel over NVLink), see cbl_route_scatter_dev
struct PeerOuts {
    void* p[ROUTE_MAX_SPLIT + 1];
};

// All digit histograms in one read of the keys.  hist[pass][256] (u64, zeroed by the caller).
// Per-warp shared-memory histograms (plain shared atomics: counting needs no order), several keys per
// thread in flight; one global atomic per (block, pass, digit) at the end.
constexpr int RH_THREADS = 512;
constexpr int RH_KEYS = 4;  // keys per thread per round
template <class W>
__global__ void __launch_bounds__(RH_THREADS) radix_hist_kernel(const W* __restrict__ keys, uint64_t n, int n_pass,
                                                                unsigned long long* __restrict__ hist) {
    extern __shared__ uint32_t sh_hist[];  // [n_pass][256]
    for (int i = threadIdx.x; i < n_pass * 256; i += RH_THREADS) sh_hist[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * RH_THREADS * RH_KEYS;
    for (uint64_t base = (uint64_t)blockIdx.x * RH_THREADS * RH_KEYS; base < n; base += stride) {
        W k[RH_KEYS];
        bool ok[RH_KEYS];
#pragma unroll
        for (int j = 0; j < RH_KEYS; j++) {
            const uint64_t i = base + (uint64_t)j * RH_THREADS + threadIdx.x;
            ok[j] = i < n;
            k[j] = ok[j] ? keys[i] : (W)0;
        }
#pragma unroll
        for (int j = 0; j < RH_KEYS; j++) {
            if (ok[j]) {
                W v = k[j];
                for (int p = 0; p < n_pass - 2; p++) {
                    atomicAdd(&sh_hist[p * 256 + ((uint32_t)v & 255u)], 1u);
                    v >>= 8;
                }
            }
        }
        // the two most significant digits are heavily skewed for necklace words (they start with a run
        // of zeros): aggregate equal digits inside the warp first, one shared atomic per distinct digit
#pragma unroll
        for (int j = 0; j < RH_KEYS; j++) {
            for (int p = max(n_pass - 2, 0); p < n_pass; p++) {
                const uint32_t d = digit_of<W>(k[j], 8 * p) | (ok[j] ? 0u : 0x100u);
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                if (ok[j] && (peers & lanemask_lt()) == 0) atomicAdd(&sh_hist[p * 256 + d], (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pass * 256; i += RH_THREADS)
        if (sh_hist[i]) atomicAdd(&hist[i], (unsigned long long)sh_hist[i]);
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
//
// Ranking (stable): all match_any's of a thread are issued back to back, then the highest lane of
// every peer group adds the group size to its warp's digit counter (shared atomic, returns the base)
// and broadcasts it; the atomics of one warp execute in program order, which keeps items ordered.
// Tile prefixes: decoupled look-back, one thread per digit.
//
// MULTI_OUT (router only): digit d's keys go to peers.p[d][digit_base[d] + ...] instead of out[...], and
// pos_out receives local_base[d] + rank of the key among this rank's keys for d (its slot in send order).
template <class W, bool HAS_VAL, class DigitFn, bool MULTI_OUT = false>
__global__ void __launch_bounds__(RS_THREADS, CBL_RS_MIN_BLOCKS) radix_pass_kernel(const W* __restrict__ in, W* __restrict__ out,
                                                                   const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                                   uint64_t n, DigitFn digit, const unsigned long long* __restrict__ digit_base,
                                                                   volatile uint32_t* status, uint32_t* tile_counter,
                                                                   uint32_t* __restrict__ pos_out, PeerOuts peers = PeerOuts(),
                                                                   const unsigned long long* __restrict__ local_base = nullptr) {
    constexpr int ITEMS = RsTile<W>::ITEMS;
    constexpr int TILE = RsTile<W>::TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* s_keys = reinterpret_cast<W*>(smem_raw);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(W) * TILE);  // only if HAS_VAL
    __shared__ uint32_t s_whist[RS_WARPS][256];
    __shared__ long long s_goff[256];
    __shared__ long long s_loff[MULTI_OUT ? ROUTE_MAX_SPLIT + 1 : 1];
    __shared__ W* s_outp[MULTI_OUT ? ROUTE_MAX_SPLIT + 1 : 1];
    __shared__ uint32_t s_tmp[33];
    __shared__ uint32_t s_tile;

    if (MULTI_OUT && threadIdx.x <= ROUTE_MAX_SPLIT) s_outp[threadIdx.x] = reinterpret_cast<W*>(peers.p[threadIdx.x]);
    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const int tile_n = (int)min((uint64_t)TILE, n - tile_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&s_whist[0][0])[i] = 0;
    // peer masks (3 round-robin sets of [warp][256] words) live in the key staging area, which is not
    // needed before step 5
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem_raw);
    for (int i = threadIdx.x; i < 3 * RS_WARPS * 256 / 4; i += RS_THREADS) reinterpret_cast<uint4*>(s_mask)[i] = make_uint4(0, 0, 0, 0);

    // 1. load (warp-striped: order = warp, item, lane)
    W key[ITEMS];
    uint32_t val[ITEMS];
    const bool full = tile_n == TILE;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int local = warp * (32 * ITEMS) + i * 32 + lane;
        const bool valid = full || local < tile_n;
        key[i] = valid ? in[tile_base + local] : (W)0;
        if (HAS_VAL) val[i] = valid ? vin[tile_base + local] : 0u;
    }
    __syncthreads();  // s_whist / s_mask zeroed
    // 2. peer groups of every row (the lanes of the warp holding the same digit), found through shared
    //    memory: every lane ORs its bit into mask[digit], then reads the word back (match.any costs ~60
    //    SM-cycles per warp on B200 when the digits are all different, a shared atomic ~4; measured).
    //    Set i % 3 is cleaned by the leaders two rows later, which needs only one warp barrier per row.
    //    info = rank inside the group (5 bits) | group size << 5 (6 bits) | leader lane << 11 (5 bits);
    //    step 3 adds the group's base << 16 and finally turns info into the rank inside (warp, digit)
    uint32_t info[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int local = warp * (32 * ITEMS) + i * 32 + lane;
        const bool valid = full || local < tile_n;
        uint32_t* M = s_mask + ((i % 3) * RS_WARPS + warp) * 256;
        if (i >= 2) {
            uint32_t* M2 = s_mask + (((i - 2) % 3) * RS_WARPS + warp) * 256;
            if ((int)((info[i - 2] >> 11) & 31u) == lane) M2[digit(key[i - 2])] = 0;
        }
        const uint32_t d = digit(key[i]);
#if CBL_RS_ABLATE == 2
        const unsigned peers = 1u << lane;
#else
        if (valid) atomicOr(&M[d], 1u << lane);
        __syncwarp();
        const unsigned peers = valid ? M[d] : (1u << lane);
#endif
        info[i] = __popc(peers & lanemask_lt()) | (__popc(peers) << 5) | ((31 - __clz(peers)) << 11);
    }
    // 3. leaders reserve their group's slots in the warp histogram (in item order)
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int local = warp * (32 * ITEMS) + i * 32 + lane;
        const bool valid = full || local < tile_n;
        if (valid && (int)((info[i] >> 11) & 31u) == lane) info[i] |= atomicAdd(&s_whist[warp][digit(key[i])], (info[i] >> 5) & 63u) << 16;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) info[i] = __shfl_sync(0xffffffffu, info[i] >> 16, (info[i] >> 11) & 31u) + (info[i] & 31u);
    __syncthreads();

    // 4. per digit (thread t = digit t): exclusive offsets over warps, tile count, look-back
    {
        const int t = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) run += s_whist[w][t];
        if (tile == 0) status[t] = RS_FLAG_INCL | run;
        else status[(size_t)tile * 256 + t] = RS_FLAG_AGG | run;
        uint32_t total;
        const uint32_t dstart = block_excl_scan<uint32_t, RS_THREADS>(run, s_tmp, total);
        uint32_t acc = dstart;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { const uint32_t c = s_whist[w][t]; s_whist[w][t] = acc; acc += c; }   // slot of (warp, digit) in the tile
        uint32_t excl = 0;
#ifndef CBL_RS_ABLATE
#define CBL_RS_ABLATE 0
#endif
        if (tile > 0 && CBL_RS_ABLATE != 1) {
            for (long long prev = (long long)tile - 1; prev >= 0; prev--) {
                uint32_t s;
                do { s = status[(size_t)prev * 256 + t]; } while ((s >> 30) == 0);
                excl += s & RS_VAL_MASK;
                if (s & RS_FLAG_INCL) break;
            }
            status[(size_t)tile * 256 + t] = RS_FLAG_INCL | (excl + run);
        }
        s_goff[t] = (long long)digit_base[t] + (long long)excl - (long long)dstart;
        if (MULTI_OUT && t <= ROUTE_MAX_SPLIT) s_loff[t] = (long long)local_base[t] + (long long)excl - (long long)dstart;
    }
    __syncthreads();

    // 5. local scatter into digit order
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int local = warp * (32 * ITEMS) + i * 32 + lane;
        if (full || local < tile_n) {
            const uint32_t d = digit(key[i]);
            const uint32_t pos = s_whist[warp][d] + info[i];
            s_keys[pos] = key[i];
            if (HAS_VAL) s_vals[pos] = val[i];
            if (DigitFn::WANTS_POS && pos_out != nullptr)
                pos_out[tile_base + local] = (uint32_t)((MULTI_OUT ? s_loff[d] : s_goff[d]) + (long long)pos);
        }
    }
    __syncthreads();

    // 6. coalesced global scatter: consecutive smem slots of one digit go to consecutive addresses
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int idx = i * RS_THREADS + threadIdx.x;
        if (full || idx < tile_n) {
            const W k = s_keys[idx];
            const uint32_t d = digit(k);
            const long long g = s_goff[d] + idx;
            if (MULTI_OUT) s_outp[d][g] = k;
            else out[g] = k;
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
