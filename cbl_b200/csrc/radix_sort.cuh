// Kernel k3: LSD radix sort of a batch of words (8-bit digits, onesweep style): one histogram pass
// over the keys, then one scatter pass per digit whose tile prefixes come from a decoupled look-back,
// so every pass reads and writes each key exactly once.  Replaces the reference's implicit grouping
// of consecutive equal prefixes (chunk_by, src/wordset/mod.rs:147,172,192,223) with a global order.
//
// Algorithmic traffic: n * sizeof(W) * (2 * passes + 1)   (SURVEY section 8d).
#pragma once
#include "scan.cuh"

namespace cbl {

#ifndef CBL_RS_ABLATE
#define CBL_RS_ABLATE 0   // developer ablations: 1 = no look-back, 2 = no peer detection (wrong results, timing only)
#endif
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
                                                                unsigned long long* __restrict__ hist, int first = 0) {
    // rows of hist: digits first, first + 1, .., first + n_pass - 1 of the key (digit d = bits [8d, 8d + 8))
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
            k[j] = ok[j] ? (W)(keys[i] >> (8 * first)) : (W)0;
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
static __global__ void __launch_bounds__(256) radix_scan_hist_kernel(unsigned long long* __restrict__ hist) {
    __shared__ unsigned long long tmp[33];
    unsigned long long* h = hist + (size_t)blockIdx.x * 256;
    unsigned long long v = h[threadIdx.x], total;
    unsigned long long e = block_excl_scan<unsigned long long, 256>(v, tmp, total);
    h[threadIdx.x] = e;
}

// One scatter pass.  status[tile][256] must be zero on entry; tile_counter zero.
// pos_out (may be null): receives, for every input slot, the slot its key was moved to (used by the multi-GPU
// router to bring answers back).  HAS_VAL is kept in the signature for the call sites; no payload variant exists.
//
// Stable ranking of a row (the 32 keys one warp instruction holds) on B200: match.any costs ~60 SM-cycles per warp
// when the digits differ, a shared atomic ~4 (measured, scripts/ubench), so the lanes with equal digits find each
// other through shared memory: every lane ORs its bit into mask[digit] and reads the word back = its peer group.
// The highest lane of a group reserves the group's slots in the warp's digit counter (one shared atomicAdd per
// group, rows in program order => stable) and broadcasts the base.  Three mask sets rotate so that one warp barrier
// per row is enough: every lane clears the word it used two rows ago.
// The body is compiled twice (FULL tile / ragged last tile): the full-tile copy carries no bounds predicates, all
// offsets are 32-bit, and nothing but (rank | digit << 16) is kept per key between the phases (the first version
// executed ~140 SASS instructions per row and spilled; this one ~45).
// Tile prefixes: decoupled look-back, one thread per digit, 8 predecessor tiles per round trip.
//
// MULTI_OUT (router only): digit d's keys go to peers.p[d][digit_base[d] + ...] instead of out[...], and
// pos_out receives local_base[d] + rank of the key among this rank's keys for d (its slot in send order).
template <class W, class DigitFn, bool MULTI_OUT, bool FULL>
__device__ __forceinline__ void radix_tile(const W* __restrict__ in, W* __restrict__ out, uint64_t n, const DigitFn& digit,
                                           const unsigned long long* __restrict__ digit_base, volatile uint32_t* status,
                                           uint32_t* __restrict__ pos_out, const unsigned long long* __restrict__ local_base,
                                           const uint32_t tile, W* s_keys, uint32_t (*s_whist)[256], long long* s_goff64, uint32_t* s_goff32,
                                           long long* s_loff, W** s_outp, uint32_t* s_tmp) {
    constexpr int ITEMS = RsTile<W>::ITEMS;
    constexpr int TILE = RsTile<W>::TILE;
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const int tile_n = FULL ? TILE : (int)(n - tile_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = threadIdx.x;
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_keys);   // the key staging area is free until step 5

    // 0. zero the warp histograms and the mask sets (vector stores)
    {
        uint4* z = reinterpret_cast<uint4*>(&s_whist[0][0]);
#pragma unroll
        for (int i = 0; i < RS_WARPS * 256 / 4 / RS_THREADS; i++) z[i * RS_THREADS + t] = make_uint4(0, 0, 0, 0);
        uint4* zm = reinterpret_cast<uint4*>(s_mask);
#pragma unroll
        for (int i = 0; i < 3 * RS_WARPS * 256 / 4 / RS_THREADS; i++) zm[i * RS_THREADS + t] = make_uint4(0, 0, 0, 0);
    }
    // 1. load (warp-striped: order = warp, item, lane)
    W key[ITEMS];
    const W* src = in + tile_base + warp * (32 * ITEMS) + lane;
    const int local0 = warp * (32 * ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) key[i] = (FULL || local0 + i * 32 < tile_n) ? src[i * 32] : (W)0;
    __syncthreads();  // s_whist / s_mask zeroed

    // 2 + 3. peer group of every row, group leaders reserve slots; rd = rank inside (warp, digit) | digit << 16
    uint32_t rd[ITEMS];
    {
        uint32_t* const Mw = s_mask + warp * 256;
        uint32_t* const Hw = &s_whist[warp][0];
        const uint32_t lanebit = 1u << lane, ltmask = lanebit - 1u;
        uint32_t *a1 = Mw, *a2 = Mw;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const bool valid = FULL || local0 + i * 32 < tile_n;
            const uint32_t d = digit(key[i]);
            uint32_t* const a = Mw + (i % 3) * (RS_WARPS * 256) + d;
#if CBL_RS_ABLATE == 2
            const uint32_t peers = lanebit;
#else
            if (valid) atomicOr(a, lanebit);
            __syncwarp();
            const uint32_t peers = valid ? *reinterpret_cast<volatile uint32_t*>(a) : lanebit;
            if (i >= 2) *reinterpret_cast<volatile uint32_t*>(a2) = 0u;   // every lane of a group clears the same word
#endif
            a2 = a1;
            a1 = a;
            uint32_t base = 0;
            if (valid && (peers >> lane) == 1u) base = atomicAdd(Hw + d, (uint32_t)__popc(peers));   // highest lane of the group
            base = __shfl_sync(0xffffffffu, base, 31 - __clz(peers));
            rd[i] = (base + __popc(peers & ltmask)) | (d << 16);
        }
    }
    __syncthreads();

    // 4a. per digit (thread t = digit t): tile count (published at once as this tile's aggregate), exclusive
    //     offsets over warps
    uint32_t run = 0, dstart;
    {
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) run += s_whist[w][t];
        if (tile == 0) status[t] = RS_FLAG_INCL | run;
        else status[(size_t)tile * 256 + t] = RS_FLAG_AGG | run;
        uint32_t total;
        dstart = block_excl_scan<uint32_t, RS_THREADS>(run, s_tmp, total);
        uint32_t acc = dstart;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { const uint32_t c = s_whist[w][t]; s_whist[w][t] = acc; acc += c; }   // slot of (warp, digit) in the tile
    }
    __syncthreads();

    // 5. local scatter into digit order (before the look-back wait: the keys leave the registers)
    {
        const uint32_t* const Hw = &s_whist[warp][0];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            if (FULL || local0 + i * 32 < tile_n) {
                const uint32_t d = rd[i] >> 16;
                const uint32_t pos = Hw[d] + (rd[i] & 0xFFFFu);
                s_keys[pos] = key[i];
                rd[i] = pos | (d << 16);
            }
        }
    }

    // 4b. decoupled look-back, LB predecessors per round trip (with ~600 resident tiles the nearest tile whose
    //     inclusive prefix is published is typically tens of tiles back; one dependent L2 load per step was slow)
    {
        constexpr int LB = 8;
        uint32_t excl = 0;
        if (tile > 0 && CBL_RS_ABLATE != 1) {
            long long prev = (long long)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t sv[LB];
#pragma unroll
                for (int j = 0; j < LB; j++) {
                    uint32_t v0 = 2u << 30;   // virtual tile before the first: inclusive prefix 0
                    if (prev - j >= 0) v0 = status[(size_t)(prev - j) * 256 + t];
                    sv[j] = v0;
                }
#pragma unroll
                for (int j = 0; j < LB; j++) {
                    if (!done) {
                        uint32_t v = sv[j];
                        while ((v >> 30) == 0) v = status[(size_t)(prev - j) * 256 + t];
                        excl += v & RS_VAL_MASK;
                        done = (v & RS_FLAG_INCL) != 0;
                    }
                }
                prev -= LB;
            }
            status[(size_t)tile * 256 + t] = RS_FLAG_INCL | (excl + run);
        }
        if (MULTI_OUT) {
            s_goff64[t] = (long long)digit_base[t] + (long long)excl - (long long)dstart;
            if (t <= ROUTE_MAX_SPLIT) s_loff[t] = (long long)local_base[t] + (long long)excl - (long long)dstart;
        } else {
            s_goff32[t] = (uint32_t)digit_base[t] + excl - dstart;   // n < 2^30: slots fit 32 bits (wrapping arithmetic)
        }
    }
    __syncthreads();
    if (DigitFn::WANTS_POS && pos_out != nullptr) {
        uint32_t* const po = pos_out + tile_base + local0;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            if (FULL || local0 + i * 32 < tile_n) {
                const uint32_t d = rd[i] >> 16, pos = rd[i] & 0xFFFFu;
                po[i * 32] = MULTI_OUT ? (uint32_t)(s_loff[d] + (long long)pos) : s_goff32[d] + pos;
            }
        }
    }

    // 6. coalesced global scatter: consecutive smem slots of one digit go to consecutive addresses
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int idx = i * RS_THREADS + t;
        if (FULL || idx < tile_n) {
            const W k = s_keys[idx];
            const uint32_t d = digit(k);
            if (MULTI_OUT) s_outp[d][s_goff64[d] + idx] = k;
            else out[s_goff32[d] + (uint32_t)idx] = k;
        }
    }
}

template <class W, bool HAS_VAL, class DigitFn, bool MULTI_OUT = false>
__global__ void __launch_bounds__(RS_THREADS, CBL_RS_MIN_BLOCKS) radix_pass_kernel(const W* __restrict__ in, W* __restrict__ out,
                                                                   const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                                   uint64_t n, DigitFn digit, const unsigned long long* __restrict__ digit_base,
                                                                   volatile uint32_t* status, uint32_t* tile_counter,
                                                                   uint32_t* __restrict__ pos_out, PeerOuts peers = PeerOuts(),
                                                                   const unsigned long long* __restrict__ local_base = nullptr) {
    static_assert(!HAS_VAL, "no payload variant");
    static_assert(RsTile<W>::TILE <= 65536 && sizeof(W) * RsTile<W>::TILE >= 3 * RS_WARPS * 256 * 4, "tile / mask aliasing");
    constexpr int TILE = RsTile<W>::TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* s_keys = reinterpret_cast<W*>(smem_raw);
    __shared__ __align__(16) uint32_t s_whist[RS_WARPS][256];
    __shared__ long long s_goff64[MULTI_OUT ? 256 : 1];
    __shared__ uint32_t s_goff32[MULTI_OUT ? 1 : 256];
    __shared__ long long s_loff[MULTI_OUT ? ROUTE_MAX_SPLIT + 1 : 1];
    __shared__ W* s_outp[MULTI_OUT ? ROUTE_MAX_SPLIT + 1 : 1];
    __shared__ uint32_t s_tmp[33];
    __shared__ uint32_t s_tile;

    if (MULTI_OUT && threadIdx.x <= ROUTE_MAX_SPLIT) s_outp[threadIdx.x] = reinterpret_cast<W*>(peers.p[threadIdx.x]);
    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    if (n - (uint64_t)tile * TILE >= (uint64_t)TILE)
        radix_tile<W, DigitFn, MULTI_OUT, true>(in, out, n, digit, digit_base, status, pos_out, local_base, tile, s_keys, s_whist, s_goff64, s_goff32,
                                                s_loff, s_outp, s_tmp);
    else
        radix_tile<W, DigitFn, MULTI_OUT, false>(in, out, n, digit, digit_base, status, pos_out, local_base, tile, s_keys, s_whist, s_goff64, s_goff32,
                                                 s_loff, s_outp, s_tmp);
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

// answers back into read order.  The byte reads are scattered: fetch 64 B instead of the default 128-byte line per
// L2 miss (see CBL_L2_FETCH_Q in index_view.cuh), four answers per thread so the output store is one word.
static __global__ void gather_u8_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ pos, uint64_t n, uint8_t* __restrict__ out) {
    const uint64_t i4 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    auto ld = [&](uint32_t p) {
        uint32_t v;
        asm("ld.global.nc.L1::no_allocate.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(src + p));
        return v;
    };
    if (i4 + 4 <= n && ((uintptr_t)(out + i4) & 3) == 0 && ((uintptr_t)(pos + i4) & 15) == 0) {
        const uint4 p = *reinterpret_cast<const uint4*>(pos + i4);
        const uint32_t a = ld(p.x), b = ld(p.y), c = ld(p.z), d = ld(p.w);
        *reinterpret_cast<uint32_t*>(out + i4) = a | (b << 8) | (c << 16) | (d << 24);
    } else {
        for (uint64_t i = i4; i < n && i < i4 + 4; i++) out[i] = (uint8_t)ld(pos[i]);
    }
}

}  // namespace cbl
