// Device-resident index layout and the probe used by contains / edit generation.
//
// Replaces, on the GPU, the reference's prefix bitvector + rank (src/bitvector/mod.rs:12-62 over
// cxx/rank_bv.h / sux WordDynRankSel), the tiered vector rank->bucket id (cxx/tiered_vec.h) and the
// Vec/Trie suffix buckets (src/trievec, src/trie.rs) with:
//   bitmap[2^P / 64]        u64   prefix presence bits
//   blkrank[2^P / 256]      u32   exclusive count of set bits before each 256-bit block (one 32-byte
//                                 sector holds the 4 words of a block, so rank = 1 directory read +
//                                 1 sector read + <=4 popcounts)
//   bucket_prefix[nb]       u32   prefix of the bucket with rank r (select)
//   bucket_off[nb + 1]      u32   start of bucket r in suf[]  (CSR, indexed by prefix RANK)
//   suf[n]                  Suf   suffixes, ascending inside every bucket => ascending word order overall
#pragma once
#include "scan.cuh"

namespace cbl {

template <class Suf>
struct IndexView {
    const uint64_t* bitmap;
    const uint32_t* blkrank;
    const uint32_t* bucket_prefix;
    const uint32_t* bucket_off;
    const Suf* suf;
    uint32_t nb;
    uint64_t n;
};

__device__ __forceinline__ bool bitmap_test_rank(const uint64_t* __restrict__ bitmap, const uint32_t* __restrict__ blkrank,
                                                 uint32_t prefix, uint32_t& rank) {
    const uint32_t blk = prefix >> 8;
    const uint32_t wi = (prefix >> 6) & 3;
    const uint64_t* w = bitmap + ((size_t)blk << 2);
    uint32_t r = __ldg(blkrank + blk);
    uint64_t w0 = __ldg(w + 0), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
    uint64_t cur = wi == 0 ? w0 : wi == 1 ? w1 : wi == 2 ? w2 : w3;
    if (wi > 0) r += __popcll(w0);
    if (wi > 1) r += __popcll(w1);
    if (wi > 2) r += __popcll(w2);
    const uint32_t b = prefix & 63;
    r += __popcll(cur & ((1ull << b) - 1));
    rank = r;  // exclusive rank (number of occupied prefixes < prefix), as RankBV::rank
    return (cur >> b) & 1;
}

template <class W, class Suf> __device__ __forceinline__ void split_key(W key, const KParams& P, uint32_t& prefix, Suf& suffix) {
    prefix = (uint32_t)(key >> P.suffix_bits);                       // src/wordset/mod.rs:63-71
    suffix = (Suf)(key & low_mask<W>(P.suffix_bits));
}

struct ProbeResult {
    bool found;
    bool prefix_present;
    uint32_t rank;   // exclusive rank of the prefix
    uint64_t pos;    // lower bound position in suf[] (insertion point when !found)
};

// ---- bucket search -----------------------------------------------------------------------------
// Suffixes inside a bucket are sorted and, for k-mer data, close to uniformly spread over the suffix
// space, so the bucket is searched by INTERPOLATION on 32-byte windows (one DRAM sector = 8 x u32 /
// 4 x u64 / 2 x u128): guess the slot from the suffix value, load the aligned window around it, and
// either finish inside the window or tighten both the index range and the value range and guess again
// (secant-like).  After PROBE_MAX_IT windows it falls back to a plain binary search, so any
// distribution is handled exactly; random DNA needs ~1-2 sectors per lookup instead of the
// ~log2(bucket) sectors of a binary search.
constexpr int PROBE_MAX_IT = 6;

template <class Suf> struct Window { static constexpr int N = 32 / (int)sizeof(Suf); };

template <class Suf> __device__ __forceinline__ void load_window(const Suf* __restrict__ p, Suf (&e)[Window<Suf>::N]);
template <> __device__ __forceinline__ void load_window<uint32_t>(const uint32_t* __restrict__ p, uint32_t (&e)[8]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 x = __ldg(q), y = __ldg(q + 1);
    e[0] = x.x; e[1] = x.y; e[2] = x.z; e[3] = x.w; e[4] = y.x; e[5] = y.y; e[6] = y.z; e[7] = y.w;
}
template <> __device__ __forceinline__ void load_window<uint64_t>(const uint64_t* __restrict__ p, uint64_t (&e)[4]) {
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
    ulonglong2 x = __ldg(q), y = __ldg(q + 1);
    e[0] = x.x; e[1] = x.y; e[2] = y.x; e[3] = y.y;
}
template <> __device__ __forceinline__ void load_window<u128>(const u128* __restrict__ p, u128 (&e)[2]) {
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
    ulonglong2 x = __ldg(q), y = __ldg(q + 1);
    e[0] = ((u128)x.y << 64) | x.x;
    e[1] = ((u128)y.y << 64) | y.x;
}

// monotone map of a suffix onto 32 bits (its most significant part) for the interpolation
template <class Suf> __device__ __forceinline__ uint32_t key32(Suf v, int suffix_bits) {
    return suffix_bits >= 32 ? (uint32_t)(v >> (suffix_bits - 32)) : ((uint32_t)v << (32 - suffix_bits));
}

// lower bound of s in suf[lo, hi) and whether it is present.  suf[] is padded so that the aligned
// window around any valid slot is readable.
template <class Suf>
__device__ __forceinline__ void bucket_search(const Suf* __restrict__ suf, uint32_t lo, uint32_t hi, Suf s, int suffix_bits,
                                              bool& found, uint32_t& pos) {
    constexpr int WN = Window<Suf>::N;
    uint32_t L = lo, R = hi;          // invariant: suf[lo, L) < s  and  suf[R, hi) > s
    float fL = 0.f, fR = 4294967296.f;  // key32 bounds of suf[L, R)
    const float fs = (float)key32<Suf>(s, suffix_bits);
    for (int it = 0; it < PROBE_MAX_IT && L < R; it++) {
        const float den = fR - fL;
        float t = den > 0.f ? (fs - fL) / den : 0.5f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        const uint32_t span = R - L;
        uint32_t g = L + min((uint32_t)(t * (float)span), span - 1);
        const uint32_t base = g & ~(uint32_t)(WN - 1);
        Suf e[WN];
        load_window<Suf>(suf + base, e);
        const uint32_t v0 = max(L, base), v1 = min(R, base + WN);  // valid slots of the window
        uint32_t n_lt = 0;
        bool eq = false;
        Suf vmin = e[0], vmax = e[0];
        bool have = false;
#pragma unroll
        for (int i = 0; i < WN; i++) {
            const uint32_t idx = base + i;
            const bool valid = idx >= v0 && idx < v1;
            if (valid) {
                n_lt += e[i] < s;
                eq |= e[i] == s;
                if (!have) { vmin = e[i]; have = true; }
                vmax = e[i];
            }
        }
        const uint32_t n_valid = v1 - v0;
        if (eq || (n_lt > 0 && n_lt < n_valid)) { found = eq; pos = v0 + n_lt; return; }
        if (n_lt == 0) { R = v0; fR = (float)key32<Suf>(vmin, suffix_bits); }   // whole window > s
        else { L = v1; fL = (float)key32<Suf>(vmax, suffix_bits); }             // whole window < s
    }
    while (L < R) {  // exact fallback
        const uint32_t mid = L + ((R - L) >> 1);
        if (suf[mid] < s) L = mid + 1; else R = mid;
    }
    pos = L;
    found = L < hi && suf[L] == s;
}

template <class W, class Suf>
__device__ __forceinline__ ProbeResult probe_key(const IndexView<Suf>& ix, const KParams& P, W key) {
    ProbeResult r;
    uint32_t prefix;
    Suf s;
    split_key<W, Suf>(key, P, prefix, s);
    if (ix.nb == 0) { r.found = false; r.prefix_present = false; r.rank = 0; r.pos = 0; return r; }
    r.prefix_present = bitmap_test_rank(ix.bitmap, ix.blkrank, prefix, r.rank);
    const uint2 off = make_uint2(__ldg(ix.bucket_off + r.rank), r.prefix_present ? __ldg(ix.bucket_off + r.rank + 1) : 0u);
    if (!r.prefix_present) { r.found = false; r.pos = off.x; return r; }
    uint32_t pos;
    bucket_search<Suf>(ix.suf, off.x, off.y, s, P.suffix_bits, r.found, pos);
    r.pos = pos;
    return r;
}

}  // namespace cbl
