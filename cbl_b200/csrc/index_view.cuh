// Device-resident index layout and the probe used by contains / edit generation.
//
// Replaces, on the GPU, the reference's prefix bitvector + rank (src/bitvector/mod.rs:12-62 over
// cxx/rank_bv.h / sux WordDynRankSel), the tiered vector rank->bucket id (cxx/tiered_vec.h) and the
// Vec/Trie suffix buckets (src/trievec, src/trie.rs) with:
//   dir[2^P / 32]           uint2 {bits of 32 prefixes, number of set bits BEFORE this word}: the
//                                 bitvector and its popcount rank directory interleaved, so presence
//                                 test + exclusive rank cost ONE 8-byte load (RankBV::get + ::rank)
//   bucket_prefix[nb]       u32   prefix of the bucket with rank r (select)
//   bucket_off[nb + 1]      u32   start of bucket r in suf[]  (CSR, indexed by prefix RANK)
//   bucket_range[nb]        uint2 {start, end} of bucket r — the same information laid out so a probe
//                                 gets both bounds with one 8-byte load
//   suf[n]                  Suf   suffixes, ascending inside every bucket => ascending word order overall
//   sub[n / 16 + 2]         i8    read-only accelerator for contains: per group of 16 suffixes the deviation
//                                 of the bucket's empirical CDF from a straight line (built lazily)
#pragma once
#include "scan.cuh"

namespace cbl {

// ---- L2 residency hints ----------------------------------------------------------------------------
// The probe's lookup tables (directory, bucket ranges, correction bytes: tens of MB) are re-read all
// the time and should stay in the 126 MB L2, while suffix windows are touched once per lookup and
// would otherwise wash the tables out.  CBL_L2_POLICY=1: tables evict_last, windows evict_first.
#ifndef CBL_L2_POLICY
#define CBL_L2_POLICY 1
#endif
// DRAM fetch size of a probe load.  Measured on B200 (scripts/ubench/dram_gran.cu): a plain ld.global.nc that misses
// L2 pulls the whole 128-byte line from HBM (125 B of DRAM traffic per random 32-byte read), the .L2::64B qualifier
// halves that (63 B); cudaLimitMaxL2FetchGranularity changes nothing.  64B is the smallest size PTX offers.
#ifndef CBL_L2_FETCH_Q
#define CBL_L2_FETCH_Q ".L2::64B"
#endif
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint2 ldg_keep(const uint2* p) {
#if CBL_L2_POLICY
    uint2 v;
    asm("ld.global.nc.L2::cache_hint" CBL_L2_FETCH_Q ".v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(l2_policy_keep()));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ int ldg_keep(const int8_t* p) {
#if CBL_L2_POLICY
    int v;
    asm("ld.global.nc.L2::cache_hint" CBL_L2_FETCH_Q ".s8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_policy_keep()));
    return v;
#else
    return (int)__ldg(p);
#endif
}
__device__ __forceinline__ uint32_t ldg_keep(const uint32_t* p) {
#if CBL_L2_POLICY
    uint32_t v;
    asm("ld.global.nc.L2::cache_hint" CBL_L2_FETCH_Q ".u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_policy_keep()));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
#if CBL_L2_POLICY
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint" CBL_L2_FETCH_Q ".v4.u32 {%0, %1, %2, %3}, [%4], %5;"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(l2_policy_stream()));
    return v;
#else
    return __ldg(p);
#endif
}

template <class Suf>
struct IndexView {
    const uint2* dir;
    const uint32_t* bucket_prefix;
    const uint32_t* bucket_off;
    const uint2* bucket_range;
    const Suf* suf;
    const int8_t* sub;   // interpolation corrections, one per SUB_GROUP suffixes (see "membership probe" below)
    uint32_t nb;
    uint64_t n;
};

// presence bit + exclusive rank (number of occupied prefixes < prefix), as RankBV::get / ::rank
__device__ __forceinline__ bool dir_test_rank(const uint2* __restrict__ dir, uint32_t prefix, uint32_t& rank) {
    const uint2 e = __ldg(dir + (prefix >> 5));
    const uint32_t b = prefix & 31;
    rank = e.y + __popc(e.x & ((1u << b) - 1u));
    return (e.x >> b) & 1u;
}
__device__ __forceinline__ unsigned int* dir_bits_word(uint2* dir, uint32_t prefix) {
    return reinterpret_cast<unsigned int*>(dir + (prefix >> 5));  // .x of the entry
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
// space, so the bucket is searched by INTERPOLATION on aligned windows (32 B = one DRAM sector, or
// 64 B = one L2 fetch granule): guess the slot from the suffix value, load the window around it, and
// either finish inside the window or tighten both the index range and the value range and guess again
// (secant-like).  After PROBE_MAX_IT windows it falls back to a plain binary search, so any
// distribution is handled exactly; random DNA needs ~2-3 windows per lookup instead of the
// ~log2(bucket) sectors of a binary search.
constexpr int PROBE_MAX_IT = 6;

template <class Suf, int WB> struct Window { static constexpr int N = WB / (int)sizeof(Suf); };

template <class Suf> __device__ __forceinline__ void unpack16(const uint4& x, Suf* e);
template <> __device__ __forceinline__ void unpack16<uint32_t>(const uint4& x, uint32_t* e) { e[0] = x.x; e[1] = x.y; e[2] = x.z; e[3] = x.w; }
template <> __device__ __forceinline__ void unpack16<uint64_t>(const uint4& x, uint64_t* e) {
    e[0] = ((uint64_t)x.y << 32) | x.x;
    e[1] = ((uint64_t)x.w << 32) | x.z;
}
template <> __device__ __forceinline__ void unpack16<u128>(const uint4& x, u128* e) {
    e[0] = ((u128)(((uint64_t)x.w << 32) | x.z) << 64) | (u128)(((uint64_t)x.y << 32) | x.x);
}
// one 32-byte load instruction (LDG.256, sm_100): measured on B200 (scripts/ubench/gather_bench.cu) a random 32-byte
// window costs the memory system the same whether it is fetched by one 16-byte request or one 32-byte request, so
// two 16-byte loads per window halve the look-up rate (24 vs 48 G windows/s)
#ifndef CBL_WINDOW_LD256
#define CBL_WINDOW_LD256 1
#endif
__device__ __forceinline__ void ldg_stream256(const void* p, uint4& a, uint4& b) {
#if CBL_L2_POLICY
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint" CBL_L2_FETCH_Q ".v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p), "l"(l2_policy_stream()));
#else
    asm("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
#endif
}
template <class Suf, int WB> __device__ __forceinline__ void load_window(const Suf* __restrict__ p, Suf (&e)[Window<Suf, WB>::N]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    constexpr int EPC = 16 / (int)sizeof(Suf);
    uint4 x[WB / 16];
#if CBL_WINDOW_LD256
    if (WB == 32) ldg_stream256(q, x[0], x[WB / 16 - 1]);
    else
#endif
    {
#pragma unroll
        for (int c = 0; c < WB / 16; c++) x[c] = ldg_stream(q + c);
    }
#pragma unroll
    for (int c = 0; c < WB / 16; c++) unpack16<Suf>(x[c], &e[c * EPC]);
}

// monotone map of a suffix onto 32 bits (its most significant part) for the interpolation
template <class Suf> __device__ __forceinline__ uint32_t key32(Suf v, int suffix_bits) {
    return suffix_bits >= 32 ? (uint32_t)(v >> (suffix_bits - 32)) : ((uint32_t)v << (32 - suffix_bits));
}

// lower bound of s in suf[lo, hi) and whether it is present.  suf[] is padded so that the aligned
// window around any valid slot is readable.
template <class Suf, int WB>
__device__ __forceinline__ void bucket_search(const Suf* __restrict__ suf, uint32_t lo, uint32_t hi, Suf s, int suffix_bits,
                                              bool& found, uint32_t& pos) {
    constexpr int WN = Window<Suf, WB>::N;
    uint32_t L = lo, R = hi;            // invariant: suf[lo, L) < s  and  suf[R, hi) > s
    float fL = 0.f, fR = 4294967296.f;  // key32 bounds of suf[L, R)
    const float fs = (float)key32<Suf>(s, suffix_bits);
    for (int it = 0; it < PROBE_MAX_IT && L < R; it++) {
        const float den = fR - fL;
        float t = den > 0.f ? __fdividef(fs - fL, den) : 0.5f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        const uint32_t span = R - L;
        const uint32_t g = L + min((uint32_t)(t * (float)span), span - 1);
        const uint32_t base = g & ~(uint32_t)(WN - 1);
        Suf e[WN];
        load_window<Suf, WB>(suf + base, e);
        if (base >= L && base + WN <= R) {  // window entirely inside the open range (the common case)
            uint32_t n_lt = 0, n_le = 0;
#pragma unroll
            for (int i = 0; i < WN; i++) { n_lt += e[i] < s; n_le += e[i] <= s; }
            if (n_le != n_lt || (n_lt > 0 && n_lt < (uint32_t)WN)) { found = n_le != n_lt; pos = base + n_lt; return; }
            if (n_lt == 0) { R = base; fR = (float)key32<Suf>(e[0], suffix_bits); }        // whole window > s
            else { L = base + WN; fL = (float)key32<Suf>(e[WN - 1], suffix_bits); }          // whole window < s
        } else {
            const uint32_t v0 = max(L, base), v1 = min(R, base + WN);  // valid slots of the window
            uint32_t n_lt = 0;
            bool eq = false, have = false;
            Suf vmin = e[0], vmax = e[0];
#pragma unroll
            for (int i = 0; i < WN; i++) {
                const uint32_t idx = base + i;
                if (idx >= v0 && idx < v1) {
                    n_lt += e[i] < s;
                    eq |= e[i] == s;
                    if (!have) { vmin = e[i]; have = true; }
                    vmax = e[i];
                }
            }
            const uint32_t n_valid = v1 - v0;
            if (eq || (n_lt > 0 && n_lt < n_valid)) { found = eq; pos = v0 + n_lt; return; }
            if (n_lt == 0) { R = v0; fR = (float)key32<Suf>(vmin, suffix_bits); }
            else { L = v1; fL = (float)key32<Suf>(vmax, suffix_bits); }
        }
    }
    while (L < R) {  // exact fallback
        const uint32_t mid = L + ((R - L) >> 1);
        if (suf[mid] < s) L = mid + 1; else R = mid;
    }
    pos = L;
    found = L < hi && suf[L] == s;
}

template <class W, class Suf, int WB = 32>
__device__ __forceinline__ ProbeResult probe_key(const IndexView<Suf>& ix, const KParams& P, W key) {
    ProbeResult r;
    uint32_t prefix;
    Suf s;
    split_key<W, Suf>(key, P, prefix, s);
    if (ix.nb == 0) { r.found = false; r.prefix_present = false; r.rank = 0; r.pos = 0; return r; }
    r.prefix_present = dir_test_rank(ix.dir, prefix, r.rank);
    if (!r.prefix_present) { r.found = false; r.pos = __ldg(ix.bucket_off + r.rank); return r; }
    const uint2 range = __ldg(ix.bucket_range + r.rank);
    uint32_t pos;
    bucket_search<Suf, WB>(ix.suf, range.x, range.y, s, P.suffix_bits, r.found, pos);
    r.pos = pos;
    return r;
}

// ---- membership probe (contains_seq / contains) --------------------------------------------------
// Only "is it there" is needed, so the search is specialised: the slot of suffix s inside its bucket
// [start, end) is predicted as  start + floor(key32(s) * span / 2^32) + corr  where corr comes from
// `sub`: the bucket owns the slots  first = ceil(start/16) .. end/16  of the sub array; the first
// 2^eb of them (eb = floor(log2(#slots))) hold, for the suffix-space boundaries j * 2^(32-eb), the
// signed difference between the true lower-bound position of that boundary and the straight-line
// prediction (clamped to i8).  Interpolating between two neighbouring corrections leaves an error of
// a couple of slots (binomial noise inside one ~16-element group), so the first aligned 32-byte
// window resolves ~4 of 5 lookups and one neighbouring window nearly all the rest.  Everything is
// exact: a wrong prediction only costs extra windows, and after PROBE_MAX_IT windows the search
// falls back to plain bisection of what is left.
#ifndef CBL_SUB_SHIFT
#define CBL_SUB_SHIFT 4
#endif
#ifndef CBL_SUB_ONE_LOAD
#define CBL_SUB_ONE_LOAD 0
#endif
constexpr int SUB_SHIFT = CBL_SUB_SHIFT;
constexpr int SUB_GROUP = 1 << SUB_SHIFT;

// Corrections are stored in units of 2^sc slots, sc = sub_scale(eb) growing with the bucket (eb = log2 of its number of
// correction slots), so that the signed byte covers a deviation of a fixed fraction of the bucket whatever its size (a
// plain byte saturates at 127 slots: in buckets of thousands of suffixes whose distribution is far from uniform — necklace
// words with few leading zeros — the clamped prediction then misses by hundreds of slots and the lookup pays window
// after window).  The unit depends on the bucket size only: nothing to store, two instructions to compute.
#ifndef CBL_SUB_SCALE_BIAS
#define CBL_SUB_SCALE_BIAS 4     // 99: plain bytes (developer A/B)
#endif
__device__ __forceinline__ int sub_scale(int eb) { return min(max(eb - CBL_SUB_SCALE_BIAS, 0), 8); }
struct SubSlots {
    uint32_t first;  // index of the bucket's first slot in sub[]
    int eb;          // log2 of the number of boundaries kept (0 => no correction for this bucket)
};
__device__ __forceinline__ SubSlots sub_slots(uint32_t start, uint32_t end) {
    SubSlots r;
    r.first = (start + SUB_GROUP - 1) >> SUB_SHIFT;
    const int slots = (int)(end >> SUB_SHIFT) - (int)r.first;
    r.eb = slots >= 2 ? 31 - __clz(slots) : 0;
    return r;
}

// predicted slot (relative to start) of the suffix with monotone 32-bit image k32 inside bucket
// [start, end).  Branch-free (the two byte loads are always in bounds: sub[] has n/16 + 4 entries) so the
// caller can keep several lookups in flight.
__device__ __forceinline__ uint32_t predict_slot(const int8_t* __restrict__ sub, uint32_t start, uint32_t end, uint32_t k32) {
    const uint32_t span = end - start;
    const SubSlots ss = sub_slots(start, end);
    const uint64_t t = (uint64_t)k32 << ss.eb;
    const uint32_t j = (uint32_t)(t >> 32);
    const int frac = (int)((uint32_t)t >> 24);
#if CBL_SUB_ONE_LOAD
    // both correction bytes from ONE aligned 4-byte load (a second one only when they straddle a word: 1 lane in 4)
    const uint32_t bi = ss.first + j;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(sub);
    const uint32_t w0 = ldg_keep(sw + (bi >> 2));
    uint32_t w1 = 0;
    if ((bi & 3u) == 3u) w1 = ldg_keep(sw + (bi >> 2) + 1);
    const uint32_t two = __funnelshift_r(w0, w1, (bi & 3u) * 8);
    const int d0 = (int)(int8_t)(two & 0xFFu);
    int d1 = (int)(int8_t)((two >> 8) & 0xFFu);
#else
    const int d0 = ldg_keep(sub + ss.first + j);
    int d1 = ldg_keep(sub + ss.first + j + 1);
#endif
    d1 = (j + 1 < (1u << ss.eb)) ? d1 : 0;
    const int sc = sub_scale(ss.eb);
    const int v = d0 * 256 + (d1 - d0) * frac;   // interpolated correction in 1/256 of a unit of 2^sc slots
    const int corr = ss.eb > 0 ? (v + (128 >> sc)) >> (8 - sc) : 0;
    const int guess = (int)__umulhi(k32, span) + corr;
    return (uint32_t)min(max(guess, 0), max((int)span - 1, 0));
}

// window evaluation.  Returns 1 found, 0 proven absent, -1 undecided (L/R/g updated for the next window).
template <class Suf, int WB>
__device__ __forceinline__ int eval_window(const Suf (&e)[Window<Suf, WB>::N], uint32_t base, Suf s, uint32_t k32, int suffix_bits,
                                           uint32_t span, uint32_t& L, uint32_t& R, uint32_t& g) {
    constexpr int WN = Window<Suf, WB>::N;
    if (base >= L && base + WN <= R) {  // window entirely inside the open range (the common case); e[] ascending
        if (s < e[0]) {
            R = base;
            if (R <= L) return 0;
            const uint32_t d = __umulhi(key32<Suf>(e[0], suffix_bits) - k32, span);
            g = base - 1 - min(d, base - 1 - L);
            return -1;
        }
        if (s > e[WN - 1]) {
            L = base + WN;
            if (R <= L) return 0;
            const uint32_t d = __umulhi(k32 - key32<Suf>(e[WN - 1], suffix_bits), span);
            g = L + min(d, R - 1 - L);
            return -1;
        }
        bool eq = false;
#pragma unroll
        for (int i = 0; i < WN; i++) eq |= e[i] == s;
        return eq ? 1 : 0;
    }
    // Window sticking out of the open range (first / last window of a bucket, or a range already narrowed): slots
    // [lo, hi) of it are valid and ascending.  p(i) = "slot i is below the range, or valid and < s" is monotone over the
    // window, so its first false slot is found by a branch-free bisection that drags the candidate element along
    // (log2(WN) + 1 compares instead of one masked compare pair per slot).
    const uint32_t lo = L > base ? L - base : 0u, hi = min(R - base, (uint32_t)WN);   // 0 <= lo < hi <= WN
    Suf x[WN];
#pragma unroll
    for (int i = 0; i < WN; i++) x[i] = e[i];
    uint32_t pos = 0;
#pragma unroll
    for (int half = WN / 2; half >= 1; half >>= 1) {
        const uint32_t i = pos + (uint32_t)half - 1u;
        const bool p = i < lo || (i < hi && x[half - 1] < s);
#pragma unroll
        for (int j = 0; j < half; j++) x[j] = p ? x[j + half] : x[j];
        pos += p ? (uint32_t)half : 0u;
    }
    const Suf z = x[0];                                        // = e[pos]
    const bool pz = pos < lo || (pos < hi && z < s);
    const uint32_t n_lt = pos + (pz ? 1u : 0u);                // slots of the window that are "< s" (those below the range included)
    if (!pz && pos < hi && z == s) return 1;
    if (n_lt == lo) { R = base + lo; if (R <= L) return 0; g = R - 1; return -1; }   // every valid slot is > s
    if (n_lt == hi) { L = base + hi; if (R <= L) return 0; g = L; return -1; }       // every valid slot is < s
    return 0;
}

// continue a lookup whose first window did not decide it
template <class Suf, int WB>
__device__ __noinline__ bool contains_slow(const Suf* __restrict__ suf, uint32_t L, uint32_t R, uint32_t g, Suf s, uint32_t k32,
                                           int suffix_bits, uint32_t span, uint32_t end) {
    constexpr int WN = Window<Suf, WB>::N;
    for (int it = 0; it < PROBE_MAX_IT; it++) {
        const uint32_t base = g & ~(uint32_t)(WN - 1);
        Suf e[WN];
        load_window<Suf, WB>(suf + base, e);
        const int r = eval_window<Suf, WB>(e, base, s, k32, suffix_bits, span, L, R, g);
        if (r >= 0) return r != 0;
    }
    while (L < R) {  // exact fallback
        const uint32_t mid = L + ((R - L) >> 1);
        if (suf[mid] < s) L = mid + 1; else R = mid;
    }
    return L < end && suf[L] == s;
}

// one complete lookup (word-level entry points; the sequence kernel stages the same steps by hand)
template <class W, class Suf, int WB = 32>
__device__ __forceinline__ bool contains_key(const IndexView<Suf>& ix, const KParams& P, W key) {
    constexpr int WN = Window<Suf, WB>::N;
    uint32_t prefix, rank;
    Suf s;
    split_key<W, Suf>(key, P, prefix, s);
    if (ix.nb == 0 || !dir_test_rank(ix.dir, prefix, rank)) return false;
    const uint2 range = ldg_keep(ix.bucket_range + rank);
    const uint32_t k32 = key32<Suf>(s, P.suffix_bits), span = range.y - range.x;
    uint32_t g = range.x + predict_slot(ix.sub, range.x, range.y, k32);
    uint32_t L = range.x, R = range.y;
    const uint32_t base = g & ~(uint32_t)(WN - 1);
    Suf e[WN];
    load_window<Suf, WB>(ix.suf + base, e);
    const int r = eval_window<Suf, WB>(e, base, s, k32, P.suffix_bits, span, L, R, g);
    if (r >= 0) return r != 0;
    return contains_slow<Suf, WB>(ix.suf, L, R, g, s, k32, P.suffix_bits, span, range.y);
}

}  // namespace cbl
