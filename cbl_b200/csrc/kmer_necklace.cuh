// k-mer arithmetic shared by every kernel: 2-bit packing, parity-canonical form, reverse complement
// and the necklace transform.  All functions are __host__ __device__ so the exact code the GPU runs
// can also be exercised on the CPU by tests/host_check (no GPU in the build container).
//
// Semantics follow the reference (imartayan/CBL @ e6ca8a4):
//   base code A=0 C=1 T=2 G=3, complement = ^2 ............ src/kmer.rs:11-24,218-220
//   canonical = even popcount, else reverse complement ...... src/kmer.rs:93-106
//   necklace = min over bit rotations, ties -> smallest pos . src/necklace/mod.rs:13-25
//   word = (necklace << POS_BITS) | pos ..................... src/cbl.rs:181-184
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CBL_HD __host__ __device__ __forceinline__
#else
#define CBL_HD inline
#endif

namespace cbl {

typedef unsigned __int128 u128;

struct KParams {
    int k;            // bases per k-mer
    int bits;         // 2k
    int pos_bits;     // ceil(log2(2k))            src/cbl.rs:66
    int prefix_bits;  // PREFIX_BITS
    int suffix_bits;  // bits + pos_bits - prefix_bits   src/cbl.rs:29-32
    int canonical;
};

template <class W> struct WordTraits;
template <> struct WordTraits<uint64_t> { static constexpr int BITS = 64; };
template <> struct WordTraits<u128> { static constexpr int BITS = 128; };

template <class W> CBL_HD W low_mask(int bits) { return (W)(((W)1 << bits) - 1); }  // bits < width of W

CBL_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
CBL_HD int clz64(uint64_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}
CBL_HD int popc_w(uint64_t x) { return popc64(x); }
CBL_HD int popc_w(u128 x) { return popc64((uint64_t)x) + popc64((uint64_t)(x >> 64)); }
CBL_HD int top_bit(uint64_t x) { return 63 - clz64(x); }  // x != 0
CBL_HD int top_bit(u128 x) {
    uint64_t hi = (uint64_t)(x >> 64);
    return hi ? 127 - clz64(hi) : 63 - clz64((uint64_t)x);
}

// reverse the 32 two-bit groups of a 64-bit word
CBL_HD uint64_t rev2_64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    x = __brevll(x);  // full bit reversal, then un-swap the two bits of every group
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
#else
    x = __builtin_bswap64(x);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    return x;
#endif
}

// reverse complement of a k-mer held in the low 2k bits (src/kmer.rs:327-348)
CBL_HD uint64_t revcomp(uint64_t x, int k) { return (rev2_64(x) ^ 0xAAAAAAAAAAAAAAAAull) >> (64 - 2 * k); }
CBL_HD u128 revcomp(u128 x, int k) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    u128 r = ((u128)(rev2_64(lo) ^ 0xAAAAAAAAAAAAAAAAull) << 64) | (u128)(rev2_64(hi) ^ 0xAAAAAAAAAAAAAAAAull);
    return r >> (128 - 2 * k);
}

template <class W> CBL_HD W rotl_ring(W w, int p, int bits, W mask) {
    // p in [0, bits); bits < width of W so (w >> bits) == 0 covers p == 0  (src/necklace/queue.rs:48-50)
    return (W)(((w << p) & mask) | (w >> (bits - p)));
}

// Normative brute force: min over p of (rotl(w, p), p).  Every rotation is formed directly from w
// (an incremental "rotate by one" loop gave wrong 128-bit results on sm_100a with nvcc 12.9 although
// the same source is right on the host — first GPU run, 2K=118; the direct form below is the one the
// fast path also uses and is bit-exact against the oracle).
template <class W> CBL_HD void necklace_brute(W w, int bits, W& neck, int& pos) {
    const W mask = low_mask<W>(bits);
    W best = w;
    int bp = 0;
    for (int p = 1; p < bits; p++) {
        W rot = rotl_ring<W>(w, p, bits, mask);
        if (rot < best) { best = rot; bp = p; }
    }
    neck = best;
    pos = bp;
}

#ifndef CBL_NECKLACE_ALGO
#define CBL_NECKLACE_ALGO 1
#endif
// Exact fast path, variant 1 (default): lexicographic elimination.  Bit i of `cand` stands for the rotation that
// brings bit i of w to the top (p = bits-1-i); its k-th bit from the top is bit i of rotl(w, k).  Step k keeps the
// candidates whose k-th bit is 0 if there are any (otherwise all of them have a 1 there and all stay), so after
// step k the survivors are exactly the rotations with the smallest (k+1)-bit head.  The loop ends when one candidate
// is left or bits steps are done (the survivors are then equal rotations of a periodic word and the highest bit =
// smallest p wins, src/necklace/mod.rs:17-23).  The zero-run search of variant 0 is the first phase of the same
// process; the tie-break between equally long runs costs one shift+mask per step instead of a rotate+compare per
// candidate.  Steps are taken two at a time (an extra step on a single survivor changes nothing).
template <class W> CBL_HD void necklace_elim(W w, int bits, W& neck, int& pos) {
    const W mask = low_mask<W>(bits);
    W cand = (W)(~w & mask);
    if (cand == 0 || w == 0) { neck = w; pos = 0; return; }  // all ones / all zeros: every rotation equal
    W wk = w;
    for (int k = 1; k < bits && (cand & (W)(cand - 1)) != 0; k += 2) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int u = 0; u < 2; u++) {
            wk = (W)(((wk << 1) & mask) | (wk >> (bits - 1)));
            const W n = (W)(cand & ~wk);
            if (n != 0) cand = n;
        }
    }
    const int p = bits - 1 - top_bit(cand);
    neck = rotl_ring<W>(w, p, bits, mask);
    pos = p;
}

// The same elimination for 64-bit words of 33..63 bits with a cheaper step (the integer ALU pipe is what bounds the
// kernels; measured per step 15.5 -> 12.5 instructions).  wk holds a 64-bit window of the PERIODIC extension of the ring
// (bit j = ring bit j mod bits): rotating the ring by one is then  (wk << 1) | (wk >> (bits - 1))  with no masking — the
// bits shifted in at the bottom are the periodic continuation, and for bits >= 33 they fit the low 32-bit half, so the
// high half is a plain shift.  cand only ever holds bits below `bits`, so the extension above never shows in cand & ~wk.
// "More than one candidate left" is tested with population counts (their pipe is idle) instead of cand & (cand - 1).
#ifndef CBL_NECKLACE_PERIODIC
#define CBL_NECKLACE_PERIODIC 1
#endif
CBL_HD void necklace_elim_periodic(uint64_t w, int bits, uint64_t& neck, int& pos) {
    const uint64_t mask = low_mask<uint64_t>(bits);
    uint64_t cand = ~w & mask;
    if (cand == 0 || w == 0) { neck = w; pos = 0; return; }
    uint64_t wk = w | (w << bits);   // bits >= 33: two copies cover the 64-bit window
    const int sh = bits - 1;
    for (int k = 1; k < bits && popc64(cand) > 1; k += 2) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int u = 0; u < 2; u++) {
            wk = (wk << 1) | (uint64_t)(uint32_t)(wk >> sh);   // 64 - sh <= 32 significant bits come down
            const uint64_t n = cand & ~wk;
            if (n != 0) cand = n;
        }
    }
    const int p = bits - 1 - top_bit(cand);
    neck = rotl_ring<uint64_t>(w, p, bits, mask);
    pos = p;
}

template <class W> CBL_HD void necklace_runs(W w, int bits, W& neck, int& pos);
// 64-bit words take the elimination loop; 128-bit words keep variant 0: the elimination loop compiled for u128 gives
// wrong words on sm_100a with nvcc 12.9 (GPU parity tests, 2K = 118) although the same source is right on the host —
// the same class of miscompile as the incremental rotation noted at necklace_brute.
template <class W> CBL_HD void necklace_fast(W w, int bits, W& neck, int& pos) {
#if CBL_NECKLACE_ALGO == 1
    if (sizeof(W) == 8) {
#if CBL_NECKLACE_PERIODIC
        if (bits >= 33) {
            uint64_t n64;
            necklace_elim_periodic((uint64_t)w, bits, n64, pos);
            neck = (W)n64;
            return;
        }
#endif
        necklace_elim<W>(w, bits, neck, pos);
    } else necklace_runs<W>(w, bits, neck, pos);
#else
    necklace_runs<W>(w, bits, neck, pos);
#endif
}

// Exact fast path, variant 0 (CBL_NECKLACE_ALGO=0, kept for A/B runs).  The minimal rotation starts with the longest circular run of zero bits, so only
// the starts of maximal-length zero runs are candidates.  y_k has bit i set iff bits i, i-1, .., i-k+1
// (circularly) are all zero; y_{k+1} = z & rotl1(y_k).  The last non-empty y marks the candidates.
// Candidates are visited by ascending rotation amount and replaced only on strict '<', which keeps
// the reference's tie rule (smallest pos; src/necklace/mod.rs:17-23, pinned by :83-98).
template <class W> CBL_HD void necklace_runs(W w, int bits, W& neck, int& pos) {
    const W mask = low_mask<W>(bits);
    const W z = (W)(~w & mask);
    if (z == 0 || w == 0) { neck = w; pos = 0; return; }  // all ones / all zeros: every rotation equal
    W y = z, cand;
    do {
        cand = y;
        y = (W)(z & (((y << 1) & mask) | (y >> (bits - 1))));
    } while (y != 0);
    W best = ~(W)0;
    int bp = 0;
    while (cand != 0) {
        int i = top_bit(cand);
        cand = (W)(cand & ~((W)1 << i));
        int p = bits - 1 - i;
        W r = rotl_ring<W>(w, p, bits, mask);
        if (r < best) { best = r; bp = p; }
    }
    neck = best;
    pos = bp;
}

// k-mer integer -> word  (src/cbl.rs:199-206 with the canonical switch)
template <class W> CBL_HD W kmer_to_word(W x, const KParams& P, bool brute = false) {
    if (P.canonical && (popc_w(x) & 1)) x = revcomp(x, P.k);
    W neck;
    int pos;
    if (brute) necklace_brute<W>(x, P.bits, neck, pos);
    else necklace_fast<W>(x, P.bits, neck, pos);
    return (W)((neck << P.pos_bits) | (W)pos);
}

// word -> k-mer integer (src/cbl.rs:210-215, src/necklace/mod.rs:29-31)
template <class W> CBL_HD W word_to_kmer(W word, const KParams& P) {
    W neck = (W)(word >> P.pos_bits);
    int pos = (int)(word & low_mask<W>(P.pos_bits));
    const W mask = low_mask<W>(P.bits);
    return (W)(((neck << (P.bits - pos)) & mask) | (neck >> pos));
}

// ASCII -> 2-bit code; valid only for ACGTacgt where code == (c >> 1) & 3  
CBL_HD uint32_t pack4(uint32_t ascii4) {
    // four bytes (first base in the lowest byte) -> 8 bits, first base most significant
    uint32_t c = (ascii4 >> 1) & 0x03030303u;
    return (c * 0x40100401u) >> 24;
}
CBL_HD bool is_acgt(uint8_t c) {
    c &= 0xDF;
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

}  // namespace cbl
