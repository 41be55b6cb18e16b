// Kernel k3b: finishing step of the batch sort.  After LSD radix passes over the TOP digits only (radix_sort.cuh)
// the words are grouped by their high part (word >> shift) but unordered inside a group ("segment").  Necklace
// words are close to uniformly spread over the low part inside a segment, so each segment is finished by an
// interpolation sort in shared memory instead of ceil(shift / 8) more global scatter passes:
//
//   bin  = segment start + floor(top32(low part) * segment length / 2^32)      (monotone in the key; ~1 key per bin)
//   slot = first slot of the bin (block scan of the bin counters) + arrival rank (shared atomic)
//   final position inside the bin by counting the smaller keys of the same bin (equal keys are ordered by slot, so
//   any multiset is handled).  k-mers that overlap in a read have necklaces with long common heads, so the bins are
//   more crowded than a uniform model predicts (measured: ~10 keys in the fullest bin of 32 neighbouring slots);
//   two bins per key, packed counters and an unrolled counting loop were tried and were slower (more instructions).
//
// Traffic: one read + one write of the batch, whatever `shift` is.  One CTA owns the segments that START inside
// its nominal tile [tile * T, tile * T + T) and reads on to the end of the last of them, so there is no inter-CTA
// communication.  A segment longer than CAP - T keys cannot be staged: the kernel raises *fail and the caller
// re-sorts the batch with the plain LSD passes (exact for any input; only the speed depends on the distribution).
// The caller picks the number of LSD passes from the batch size so that this does not happen for k-mer data
// (Index::sort_keys).  Replaces, like the radix passes, the reference's per-bucket insertion order bookkeeping
// (src/wordset/mod.rs:187-216 groups consecutive equal prefixes; src/trievec/mod.rs:209-220 sorts buckets lazily).
#pragma once
#include "scan.cuh"

namespace cbl {

constexpr int SS_THREADS = 512;
#ifndef CBL_SS_ITEMS8
#define CBL_SS_ITEMS8 11   // keys per thread for 8-byte words (measured: 9 -> 7.01 ms, 11 -> 6.75 ms, 13 -> 7.03 ms per 500 M keys)
#endif
template <class W> struct SsTile {
    static constexpr int ITEMS = sizeof(W) == 8 ? CBL_SS_ITEMS8 : 7;   // odd: blocked shared-memory access without bank conflicts
    static constexpr int CAP = SS_THREADS * ITEMS;          // keys one CTA can stage
    static constexpr int T = CAP / 2;                       // smallest nominal tile (keys a CTA owns); longest segment always handled = CAP - tile
    static constexpr int LIM = CAP + 1;                     // staged slots: slot j holds in[t0 - 1 + j]
    static constexpr size_t OFF_CNT = ((size_t)LIM * sizeof(W) + 15) & ~(size_t)15;
    static constexpr size_t OFF_SEG = OFF_CNT + (((size_t)(CAP + 1) * 4 + 15) & ~(size_t)15);
    static constexpr size_t OFF_BIN = OFF_SEG + (((size_t)(CAP + 2) * 2 + 15) & ~(size_t)15);
    static constexpr size_t SMEM = OFF_BIN + (size_t)CAP * 2;
};

// top 32 bits of the low `shift` bits of a word (monotone in the low part)
template <class W> __device__ __forceinline__ uint32_t low_top32(W key, int shift) {
    const W f = (W)(key & low_mask<W>(shift));
    return shift >= 32 ? (uint32_t)(f >> (shift - 32)) : ((uint32_t)f << (32 - shift));
}

// LOW32: shift <= 32, so two words of one segment compare like their low 32 bits
template <class W, bool LOW32>
__global__ void __launch_bounds__(SS_THREADS, 2) seg_sort_kernel(const W* __restrict__ in, W* __restrict__ out, uint64_t n, int shift,
                                                                 unsigned* __restrict__ fail, const int T) {
    // T = nominal keys per tile (run-time, SsTile<W>::T <= T < CAP): the longest segment always handled is CAP - T, so the
    // caller raises T when it expects short segments and a CTA then sorts close to CAP keys instead of CAP / 2
    constexpr int ITEMS = SsTile<W>::ITEMS, CAP = SsTile<W>::CAP, LIM = SsTile<W>::LIM;
    extern __shared__ __align__(16) unsigned char ss_raw[];
    W* A = reinterpret_cast<W*>(ss_raw);                                                // [LIM] staging, later the binned keys
    uint32_t* cnt = reinterpret_cast<uint32_t*>(ss_raw + SsTile<W>::OFF_CNT);          // [CAP + 1] bin counters -> first slot of every bin
    uint16_t* seg_start = reinterpret_cast<uint16_t*>(ss_raw + SsTile<W>::OFF_SEG);    // [CAP + 2]
    uint16_t* binid = reinterpret_cast<uint16_t*>(ss_raw + SsTile<W>::OFF_BIN);        // [CAP]
    __shared__ uint32_t s_tmp[33];
    __shared__ int s_a0, s_a1;
    const int t = threadIdx.x;
    const W hi_mask = (W)~low_mask<W>(shift);   // two words belong to one segment iff they agree on these bits
    const uint64_t t0 = (uint64_t)blockIdx.x * T;
    if (t0 >= n) return;
    // slot one past the last key of the batch (the end of the data closes the last segment)
    const uint64_t jend64 = n - t0 + 1;
    const int jend = jend64 > (uint64_t)LIM + 1 ? LIM + 1 : (int)jend64;

    for (int i = t; i <= CAP; i += SS_THREADS) cnt[i] = 0;
    if (t == 0) { s_a0 = INT_MAX; s_a1 = INT_MAX; }

    // ---- 1. stage the nominal tile plus one chunk; a0 = first segment start inside the tile, a1 = first one after it
    int loaded = 0;                       // slots [0, loaded) are staged
    auto stage = [&](int upto) {          // upto <= min(LIM, jend)
        const W* const src = in + t0;               // slot j holds src[j - 1]; slot 0 of the first tile holds nothing
        for (int j = loaded + t; j < upto; j += SS_THREADS) A[j] = (j > 0 || t0 > 0) ? src[(long long)j - 1] : (W)0;
    };
    auto find_bounds = [&](int upto) {    // segment starts among slots [loaded, upto); the end of the data counts for a1 only
        for (int j = max(loaded, 1) + t; j < upto; j += SS_THREADS) {
            const bool bnd = (t0 + (uint64_t)j == 1) || ((A[j] ^ A[j - 1]) & hi_mask) != 0;
            if (bnd) atomicMin(j <= T ? &s_a0 : &s_a1, j);
        }
        if (t == 0 && jend <= upto) atomicMin(&s_a1, jend);
    };
    {
        const int upto = min(min(T + 1 + SS_THREADS, LIM), jend);
        stage(upto);
        __syncthreads();
        find_bounds(upto);
        loaded = upto;
        __syncthreads();
    }
    if (s_a0 == INT_MAX) return;          // the tile lies inside a segment that started earlier
    while (s_a1 == INT_MAX) {             // block-uniform: s_a1 only changes between barriers
        const int upto = min(min(loaded + 2 * SS_THREADS, LIM), jend);
        if (upto <= loaded) {             // staging area full: the segment is too long (upto == jend cannot get here, it sets a1)
            if (t == 0) atomicOr(fail, 1u);
            return;
        }
        __syncthreads();                  // everybody has read s_a1
        stage(upto);
        __syncthreads();
        find_bounds(upto);
        loaded = upto;
        __syncthreads();
    }
    const int a0 = s_a0, a1 = s_a1;
    const int m = a1 - a0;                // keys this CTA sorts: slots [a0, a1); m <= CAP because a0 >= 1 and a1 <= LIM
    const uint64_t g0 = t0 - 1 + (uint64_t)a0;   // global index of the first of them

    // ---- 2. blocked pass: keys to registers, segment index of every key, segment starts
    W key[ITEMS];
    uint32_t sidx[ITEMS];
    {
        uint32_t flags = 0;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int e = ITEMS * t + i;
            key[i] = (W)0;
            if (e < m) {
                key[i] = A[a0 + e];
                if (e == 0 || ((key[i] ^ A[a0 + e - 1]) & hi_mask) != 0) flags |= 1u << i;
            }
        }
        uint32_t total;
        const uint32_t excl = block_excl_scan<uint32_t, SS_THREADS>((uint32_t)__popc(flags), s_tmp, total);
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            sidx[i] = excl + __popc(flags & ((2u << i) - 1u)) - 1u;
            if ((flags >> i) & 1u) seg_start[sidx[i]] = (uint16_t)(ITEMS * t + i);
        }
        if (t == 0) seg_start[total] = (uint16_t)m;
    }
    __syncthreads();   // A fully read (its storage becomes the binned array), seg_start complete

    // ---- 3. bin of every key, arrival rank inside the bin
    uint32_t br[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        br[i] = 0;
        if (ITEMS * t + i < m) {
            const uint32_t st = seg_start[sidx[i]], len = seg_start[sidx[i] + 1] - st;
            const uint32_t k32 = LOW32 ? ((uint32_t)key[i] << (32 - shift)) : low_top32<W>(key[i], shift);   // shift in [1, 32] when LOW32
            const uint32_t bin = st + __umulhi(k32, len);
            const uint32_t r = atomicAdd(&cnt[bin], 1u);
            br[i] = bin | (r << 16);
        }
    }
    __syncthreads();

    // ---- 4. bin counters -> first slot of every bin (exclusive scan, blocked), cnt[m] = m
    {
        uint32_t c[ITEMS], sum = 0;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int e = ITEMS * t + i;
            c[i] = e < m ? cnt[e] : 0u;
            sum += c[i];
        }
        uint32_t total;
        uint32_t acc = block_excl_scan<uint32_t, SS_THREADS>(sum, s_tmp, total);
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int e = ITEMS * t + i;
            if (e < m) cnt[e] = acc;
            acc += c[i];
        }
        if (t == 0) cnt[m] = (uint32_t)m;
    }
    __syncthreads();

    // ---- 5. keys into bin order
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        if (ITEMS * t + i < m) {
            const uint32_t bin = br[i] & 0xFFFFu;
            const uint32_t pos = cnt[bin] + (br[i] >> 16);
            A[pos] = key[i];
            binid[pos] = (uint16_t)bin;
        }
    }
    __syncthreads();

    // ---- 6. order inside the bin by counting, coalesced write-out
    for (int q = t; q < m; q += SS_THREADS) {
        const W k = A[q];
        const uint32_t b = binid[q];
        const uint32_t bs = cnt[b], be = cnt[b + 1];
        uint32_t rank = 0;
        if (be - bs > 1) {
            for (uint32_t u = bs; u < be; u++) {
                if (LOW32) {
                    const uint32_t v = (uint32_t)A[u];
                    rank += v < (uint32_t)k || (v == (uint32_t)k && u < (uint32_t)q);
                } else {
                    const W v = A[u];
                    rank += v < k || (v == k && u < (uint32_t)q);
                }
            }
        }
        out[g0 + bs + rank] = k;
    }
}

}  // namespace cbl
