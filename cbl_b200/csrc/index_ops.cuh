// Kernels k8 and helpers: CSR -> words expansion (iter / export), the probe's interpolation corrections, bucket sizes.
// (The first-generation mutation path — unique + probe/edit lists + directory rebuild — was removed in round 2: every
// mutation is the streaming merge of merge_ops.cuh.)
//
// Replaces WordSet::iter (src/wordset/mod.rs:298-362) + CBL::recover_kmer (src/cbl.rs:210-215) and the
// bucket-size statistics (src/wordset/mod.rs:258-263).
#pragma once
#include "index_view.cuh"

namespace cbl {

constexpr int OP_THREADS = 256;
constexpr int OP_ITEMS = 8;
constexpr int OP_TILE = OP_THREADS * OP_ITEMS;

enum : int { EDIT_INS = 1, EDIT_DEL = 2 };   // mutation kinds of the sequence / word entry points

// ---------------------------------------------------------------------------------------------
// CSR -> words (ascending) for elements [e0, e0 + count); optionally rotated back into k-mers
// (WordSet::iter + CBL::recover_kmer, src/wordset/mod.rs:349-361, src/cbl.rs:210-215).
// ---------------------------------------------------------------------------------------------
template <class W, class Suf>
__global__ void __launch_bounds__(OP_THREADS) expand_kernel(IndexView<Suf> ix, KParams P, uint64_t e0, uint64_t count,
                                                            int to_kmers, W* __restrict__ out) {
    __shared__ uint32_t s_tmp[33];
    __shared__ uint32_t s_r[2];
    __shared__ uint8_t s_start[OP_TILE];
    const uint64_t t0 = e0 + (uint64_t)blockIdx.x * OP_TILE;
    const uint64_t t1 = min(e0 + count, t0 + (uint64_t)OP_TILE);
    if (threadIdx.x == 0) {
        // bucket containing t0: last r with off[r] <= t0 ; first bucket starting at or after t1
        s_r[0] = (uint32_t)(upper_bound_dev<uint32_t>(ix.bucket_off, (uint64_t)ix.nb + 1, (uint32_t)t0) - 1);
        s_r[1] = (uint32_t)lower_bound_dev<uint32_t>(ix.bucket_off, (uint64_t)ix.nb + 1, (uint32_t)t1);
    }
    for (int i = threadIdx.x; i < OP_TILE; i += OP_THREADS) s_start[i] = 0;
    __syncthreads();
    const uint32_t r_lo = s_r[0], r_hi = s_r[1];
    for (uint32_t r = r_lo + 1 + threadIdx.x; r < r_hi; r += OP_THREADS) s_start[ix.bucket_off[r] - t0] = 1;
    __syncthreads();
    const int s0 = threadIdx.x * OP_ITEMS;
    uint32_t c = 0;
#pragma unroll
    for (int e = 0; e < OP_ITEMS; e++) c += s_start[s0 + e];
    uint32_t total;
    uint32_t before = block_excl_scan<uint32_t, OP_THREADS>(c, s_tmp, total);
    uint32_t rank = r_lo + before;
#pragma unroll
    for (int e = 0; e < OP_ITEMS; e++) {
        const uint64_t idx = t0 + s0 + e;
        rank += s_start[s0 + e];
        if (idx < t1) {
            W key = (W)(((W)ix.bucket_prefix[rank] << P.suffix_bits) | (W)ix.suf[idx]);
            out[idx - e0] = to_kmers ? word_to_kmer<W>(key, P) : key;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Interpolation corrections for the membership probe (index_view.cuh): one signed byte per group of
// SUB_GROUP suffixes.  Slot q belongs to the bucket that holds element q * SUB_GROUP; if it is one of
// the bucket's first 2^eb slots it stores  lower_bound(boundary j * 2^(32-eb)) - straight-line prediction.
// ---------------------------------------------------------------------------------------------
template <class Suf>
__global__ void __launch_bounds__(256) build_sub_kernel(IndexView<Suf> ix, int suffix_bits, int8_t* __restrict__ sub, uint64_t n_slots) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_slots) return;
    int dv = 0;
    const uint64_t pos = q << SUB_SHIFT;
    if (pos < ix.n && ix.nb) {
        const uint32_t r = (uint32_t)(upper_bound_dev<uint32_t>(ix.bucket_off, (uint64_t)ix.nb + 1, (uint32_t)pos) - 1);
        const uint32_t start = ix.bucket_off[r], end = ix.bucket_off[r + 1];
        const SubSlots ss = sub_slots(start, end);
        const uint32_t j = (uint32_t)q - ss.first;
        if (ss.eb > 0 && j > 0 && j < (1u << ss.eb)) {
            const uint32_t kb = j << (32 - ss.eb);
            uint32_t lo = start, hi = end;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (key32<Suf>(ix.suf[mid], suffix_bits) < kb) lo = mid + 1; else hi = mid;
            }
            dv = (int)(lo - start) - (int)__umulhi(kb, end - start);
            const int sc = sub_scale(ss.eb);   // stored in units of 2^sc slots (rounded to nearest)
            dv = (dv + ((1 << sc) >> 1)) >> sc;
            dv = min(max(dv, -127), 127);
        }
    }
    sub[q] = (int8_t)dv;
}

// bucket sizes (prefix, size) — a by-product of the CSR offsets (src/wordset/mod.rs:258-263)
static __global__ void bucket_sizes_kernel(const uint32_t* __restrict__ bucket_off, uint32_t nb, uint32_t* __restrict__ sizes) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nb) sizes[r] = bucket_off[r + 1] - bucket_off[r];
}

}  // namespace cbl
