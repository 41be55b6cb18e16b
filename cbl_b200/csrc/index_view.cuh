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

template <class W, class Suf>
__device__ __forceinline__ ProbeResult probe_key(const IndexView<Suf>& ix, const KParams& P, W key) {
    ProbeResult r;
    uint32_t prefix;
    Suf s;
    split_key<W, Suf>(key, P, prefix, s);
    if (ix.nb == 0) { r.found = false; r.prefix_present = false; r.rank = 0; r.pos = 0; return r; }
    r.prefix_present = bitmap_test_rank(ix.bitmap, ix.blkrank, prefix, r.rank);
    uint32_t lo = __ldg(ix.bucket_off + r.rank);
    if (!r.prefix_present) { r.found = false; r.pos = lo; return r; }
    uint32_t hi = __ldg(ix.bucket_off + r.rank + 1);
    while (lo < hi) {  // lower bound of s in suf[lo, hi)
        uint32_t mid = lo + ((hi - lo) >> 1);
        Suf v = ix.suf[mid];
        if (v < s) lo = mid + 1; else hi = mid;
    }
    r.pos = lo;
    uint32_t end = __ldg(ix.bucket_off + r.rank + 1);
    r.found = lo < end && ix.suf[lo] == s;
    return r;
}

}  // namespace cbl
