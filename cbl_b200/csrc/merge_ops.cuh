// Kernels k4-k7 (second generation): every mutation of a shard — insert_seq, remove_seq, | & - ^ — is ONE
// streaming merge of two ascending word sequences:
//     A = the resident index (CSR: prefix of the bucket, suffix of the element), never materialised as words
//     B = the batch (sorted words, duplicates allowed: a repeated word counts once, so the sort needs no
//         separate unique pass) or, for set operations, the OTHER index, read as CSR too (BCSR: no expansion
//         of the operand into words, traffic N_b * S instead of N_b * (S + 2 W))
// Replaces WordSet::insert_batch / remove_batch (src/wordset/mod.rs:187-237) and the binary set
// operations (src/wordset/set_ops.rs:78-410, src/trievec/set_ops.rs:5-257, src/bitvector/set_ops.rs:4-106).
//
//   merge_partition_kernel  merge-path split points: tile t covers merged positions [t*TILE, (t+1)*TILE)
//   merge_apply_kernel      per tile: stage A and B words in shared memory, per-thread merge-path split,
//                           8-step serial merge with duplicate detection, emit by set-op rule, write new
//                           suffixes (tile offsets by decoupled look-back) and add run lengths to the
//                           dense per-prefix counters (one atomic per (tile, prefix) run)
//   dir_bits_kernel         per-prefix counters -> bitvector + popcount rank directory (warp ballots,
//                           block scan, look-back) and the element offset of every directory word
//   dir_fill_kernel         bucket_prefix / bucket_off / bucket_range from the counters
//
// Traffic per mutation: N*S (old suffixes) + nB*W (batch) + N'*S (new suffixes) + 3 * 2^P * 4 (counters).
#pragma once
#include "index_view.cuh"

namespace cbl {

constexpr int MG_THREADS = 256;
#ifndef CBL_MG_ITEMS
#define CBL_MG_ITEMS 12   // 8-byte words, measured on B200 (500 M batch words): 8 -> 3.22 ms, 10 -> 2.97, 11 -> 3.30, 12 -> 2.88, 16 -> 3.66
#endif
#ifndef CBL_MG_ITEMS16
#define CBL_MG_ITEMS16 12  // 16-byte words: the tile's staging area is twice as large (occupancy), see DESIGN.md section 4
#endif
// elements per thread / per tile / staged words, by word width
template <class W> struct MgCfg {
    static constexpr int ITEMS = sizeof(W) == 8 ? CBL_MG_ITEMS : CBL_MG_ITEMS16;
    static constexpr int TILE = MG_THREADS * ITEMS;
    static constexpr int SMEM_ELEMS = TILE + 2 + TILE / 32 + 2;   // staging words (also holds the padded output)
};

enum : int { MERGE_OR = 0, MERGE_AND = 1, MERGE_SUB = 2, MERGE_XOR = 3 };  // same numbering as SetOp

// A[i] <= key ?   (A = the index seen as an ascending word sequence; i < ix.n, ix.nb > 0)
template <class W, class Suf>
__device__ __forceinline__ bool index_elem_le(const IndexView<Suf>& ix, const KParams& P, uint32_t i, W key) {
    uint32_t prefix, rank;
    Suf s;
    split_key<W, Suf>(key, P, prefix, s);
    const bool present = dir_test_rank(ix.dir, prefix, rank);
    const uint32_t start = __ldg(ix.bucket_off + rank);  // first element whose prefix is >= key's prefix
    if (i < start) return true;
    if (!present) return false;
    const uint32_t end = __ldg(ix.bucket_off + rank + 1);
    if (i >= end) return false;
    return ix.suf[i] <= s;
}

// The B side of a merge: a sorted word array (batch) or a second index in CSR form (set operations).
template <class W, class Suf, bool BCSR> struct MergeB;
template <class W, class Suf> struct MergeB<W, Suf, false> {
    const W* words;
    uint64_t n;
    __device__ __forceinline__ uint64_t size() const { return n; }
    __device__ __forceinline__ W at(uint64_t j, const KParams&) const { return words[j]; }
    __device__ __forceinline__ uint32_t rank_of(uint64_t) const { return 0; }
};
template <class W, class Suf> struct MergeB<W, Suf, true> {
    IndexView<Suf> ix;
    __device__ __forceinline__ uint64_t size() const { return ix.n; }
    __device__ __forceinline__ uint32_t rank_of(uint64_t j) const {   // bucket holding element j (nb when j == n)
        return j < ix.n ? (uint32_t)(upper_bound_dev<uint32_t>(ix.bucket_off, (uint64_t)ix.nb + 1, (uint32_t)j) - 1) : ix.nb;
    }
    __device__ __forceinline__ W at(uint64_t j, const KParams& P) const {
        const uint32_t r = rank_of(j);
        return (W)(((W)ix.bucket_prefix[r] << P.suffix_bits) | (W)ix.suf[j]);
    }
};

// part_i[t] = number of A elements among the first min(t * TILE, nA + nB) merged elements (A first on
// ties); part_r[t] = rank of the bucket holding A[part_i[t]] (nb when part_i[t] == nA); part_rb[t] (BCSR only) = the
// same for B[t * TILE - part_i[t]].  t in [0, tiles].
template <class W, class Suf, bool BCSR>
__global__ void merge_partition_kernel(IndexView<Suf> ix, KParams P, MergeB<W, Suf, BCSR> B, uint64_t tiles,
                                       uint32_t* __restrict__ part_i, uint32_t* __restrict__ part_r, uint32_t* __restrict__ part_rb,
                                       const unsigned long long* __restrict__ skip) {
    // skip: device flag raised by the batch sort when B could NOT be sorted (seg_sort.cuh); merging an unsorted B would
    // index shared memory out of bounds, and the host re-sorts and re-merges anyway once it has read the flag
    if (skip && *skip) return;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > tiles) return;
    const uint64_t nA = ix.n, nB = B.size();
    const uint64_t D = min(t * (uint64_t)MgCfg<W>::TILE, nA + nB);
    uint64_t lo = D > nB ? D - nB : 0, hi = min(D, nA);
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (index_elem_le<W, Suf>(ix, P, (uint32_t)mid, B.at(D - 1 - mid, P))) lo = mid + 1; else hi = mid;
    }
    part_i[t] = (uint32_t)lo;
    part_r[t] = lo < nA ? (uint32_t)(upper_bound_dev<uint32_t>(ix.bucket_off, (uint64_t)ix.nb + 1, (uint32_t)lo) - 1) : ix.nb;
    if (BCSR) part_rb[t] = B.rank_of(D - lo);
}

// stage elements [i0, i0 + n) of a CSR index as words into dst[0, n): r0 / r1 = bucket ranks of elements i0 and i0 + n
// (nb past the end).  Uses s_pos / s_pfx as scratch (bucket starts inside the range); block-wide, ends with a barrier.
template <class W, class Suf>
__device__ __forceinline__ void stage_csr_words(const IndexView<Suf>& ix, const KParams& P, uint32_t i0, int n, uint32_t r0, uint32_t r1, W* dst,
                                                uint16_t* s_pos, uint32_t* s_pfx) {
    int n_bk = 0;
    if (n > 0) {
        const uint32_t r_hi = min(r1, ix.nb - 1);             // last candidate rank
        n_bk = (int)(r_hi - r0) + 1;
        if (r_hi > r0 && ix.bucket_off[r_hi] >= i0 + (uint32_t)n) n_bk--;   // the bucket of element i0 + n starts exactly there
        for (int j = threadIdx.x; j < n_bk; j += MG_THREADS) {
            s_pos[j] = j == 0 ? 0 : (uint16_t)(ix.bucket_off[r0 + j] - i0);
            s_pfx[j] = ix.bucket_prefix[r0 + j];
        }
    }
    __syncthreads();
    // Blocked: thread t stages elements [t * per, t * per + per).  All of its suffix loads are issued before anything
    // depends on them (the strided, one-load-per-iteration version stalled on every load: 26 % of the kernel's stall
    // samples), its first element's bucket is found by ONE bisection of the bucket starts and the following elements walk
    // on from there (consecutive elements rarely cross more than one bucket boundary).
    constexpr int ITEMS = MgCfg<W>::ITEMS;
    const int per = (n + MG_THREADS - 1) / MG_THREADS;       // <= ITEMS
    const int s0 = (int)threadIdx.x * per, s1 = min(n, s0 + per);
    Suf v[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
        if (s0 + i < s1) v[i] = ix.suf[i0 + s0 + i];
    int lo = 0;
    if (s0 < s1) {
        int hi = n_bk;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((int)s_pos[mid] <= s0) lo = mid; else hi = mid;
        }
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int e = s0 + i;
        if (e < s1) {
            while (lo + 1 < n_bk && (int)s_pos[lo + 1] <= e) lo++;
            dst[e] = (W)(((W)s_pfx[lo] << P.suffix_bits) | (W)v[i]);
        }
    }
    __syncthreads();
}

// shared-memory slot of output element k: one pad word per 32 elements makes the "8 consecutive elements
// per thread" write pattern bank-conflict free while keeping the sequential read-out conflict free
__device__ __forceinline__ uint32_t mg_pad(uint32_t k) { return k + (k >> 5); }

template <class W, class Suf, int OP, bool BCSR>
__global__ void __launch_bounds__(MG_THREADS) merge_apply_kernel(IndexView<Suf> ix, KParams P, MergeB<W, Suf, BCSR> B,
                                                                 const uint32_t* __restrict__ part_i, const uint32_t* __restrict__ part_r,
                                                                 const uint32_t* __restrict__ part_rb,
                                                                 Suf* __restrict__ suf_out, uint32_t* __restrict__ prefix_cnt,
                                                                 volatile uint64_t* status, uint32_t* tile_counter,
                                                                 unsigned long long* __restrict__ n_out,
                                                                 const unsigned long long* __restrict__ skip) {
    if (skip && *skip) return;   // see merge_partition_kernel (block-uniform: every tile leaves, nobody waits in the look-back)
    extern __shared__ __align__(16) unsigned char mg_smem[];
    W* sK = reinterpret_cast<W*>(mg_smem);                 // [0] A halo, [1, na] A, (na, na + nb] B, [na + nb + 1] B halo
    Suf* s_out = reinterpret_cast<Suf*>(mg_smem);          // after the merge: emitted suffixes (padded slots)
    __shared__ uint16_t s_pos[MgCfg<W>::TILE + 1];                // first: bucket starts inside the tile; later: run heads
    __shared__ uint32_t s_pfx[MgCfg<W>::TILE + 1];                // first: prefix of those buckets;       later: prefix of the heads
    __shared__ uint32_t s_last[MG_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;
    __shared__ uint32_t s_tmp[33];
    __shared__ W s_bprev;                                   // B[j0 - 1]: a B element equal to its predecessor is a repeat
    // Exhausted sides compare as SENTINEL.  No k-mer word is all ones (the position field of an all-ones necklace is
    // 0); the word-level entry points that accept foreign words reject it (Index::check_foreign_words).
    const W SENTINEL = ~(W)0;
    constexpr uint32_t NONE = 0xFFFFFFFFu;

    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const uint64_t nA = ix.n, nB = B.size();
    const uint64_t D0 = min((uint64_t)tile * MgCfg<W>::TILE, nA + nB), D1 = min((uint64_t)(tile + 1) * MgCfg<W>::TILE, nA + nB);
    const uint32_t i0 = part_i[tile], i1 = part_i[tile + 1];
    const uint64_t j0 = D0 - i0, j1 = D1 - i1;
    const int na = (int)(i1 - i0), nb = (int)(j1 - j0);
    const uint32_t r0 = part_r[tile], r1 = part_r[tile + 1];

    // ---- stage A words and B words (coalesced global reads, consecutive shared slots) ----
    stage_csr_words<W, Suf>(ix, P, i0, na, r0, r1, sK + 1, s_pos, s_pfx);
    if constexpr (BCSR) {
        stage_csr_words<W, Suf>(B.ix, P, (uint32_t)j0, nb, part_rb[tile], part_rb[tile + 1], sK + 1 + na, s_pos, s_pfx);
    } else {
        for (int s = threadIdx.x; s < nb; s += MG_THREADS) sK[1 + na + s] = B.words[j0 + s];
    }
    if (threadIdx.x == 0) {
        W h = SENTINEL;
        if (i0 > 0) {  // A[i0 - 1]: in bucket r0 unless that bucket starts exactly at i0
            const uint32_t rp = (r0 < ix.nb && ix.bucket_off[r0] < i0) ? r0 : r0 - 1;
            h = (W)(((W)ix.bucket_prefix[rp] << P.suffix_bits) | (W)ix.suf[i0 - 1]);
        }
        sK[0] = h;
        W nxt = SENTINEL;
        if (j1 < nB) {
            // B[j1]: for a CSR operand its bucket is the one the partition kernel already found for this split point (a
            // bisection of the bucket table here, by one thread with the whole CTA waiting, cost 21 dependent loads per tile)
            if constexpr (BCSR) nxt = (W)(((W)B.ix.bucket_prefix[part_rb[tile + 1]] << P.suffix_bits) | (W)B.ix.suf[j1]);
            else nxt = B.at(j1, P);
        }
        sK[1 + na + nb] = nxt;
        s_bprev = (!BCSR && j0 > 0) ? B.at(j0 - 1, P) : SENTINEL;   // an index holds no repeats
    }
    __syncthreads();

    // ---- per-thread merge-path split inside the tile, then MgCfg<W>::ITEMS serial steps ----
    const W* sA = sK + 1;
    const W* sB = sK + 1 + na;
    const int d = min((int)threadIdx.x * MgCfg<W>::ITEMS, na + nb);
    int lo = max(0, d - nb), hi = min(d, na);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sA[mid] <= sB[d - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    int ia = lo, ib = d - lo;
    W out[MgCfg<W>::ITEMS];
    uint32_t emit_mask = 0;
    {
        // the current and the previous element of both sides live in registers: one shared load per step
        W a = ia < na ? sA[ia] : SENTINEL;
        W b = sB[ib];                                  // ib == nb reads the B halo
        W pa = sA[ia - 1];                             // the largest A element taken so far (A halo when ia == 0)
        W pb = ib > 0 ? sB[ib - 1] : s_bprev;          // a B element equal to its predecessor is a repeat
#pragma unroll
        for (int e = 0; e < MgCfg<W>::ITEMS; e++) {
            out[e] = 0;
            if (ia + ib < na + nb) {
                const bool take_a = ia < na && (ib >= nb || a <= b);
                bool emit;
                if (take_a) {
                    const bool eq_b = b == a;          // b is the smallest B element >= a
                    emit = OP == MERGE_OR ? true : OP == MERGE_AND ? eq_b : !eq_b;
                    out[e] = a;
                    pa = a;
                    ia++;
                    a = ia < na ? sA[ia] : SENTINEL;
                } else {
                    emit = (OP == MERGE_OR || OP == MERGE_XOR) && pa != b && pb != b;
                    out[e] = b;
                    pb = b;
                    ib++;
                    b = sB[ib];
                }
                emit_mask |= (emit ? 1u : 0u) << e;
            }
        }
    }
    // ---- prefix runs: a head is an emitted element whose prefix differs from the previously emitted one ----
    uint32_t last = NONE;
#pragma unroll
    for (int e = 0; e < MgCfg<W>::ITEMS; e++)
        if ((emit_mask >> e) & 1u) last = (uint32_t)(out[e] >> P.suffix_bits);
    // prev = prefix of the element emitted last before this thread's range ("last defined value" scan)
    uint32_t incl = last;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, dd);
        if ((int)lane_id() >= dd && incl == NONE) incl = t;
    }
    uint32_t prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane_id() == 0) prev = NONE;
    if (lane_id() == 31) s_last[threadIdx.x >> 5] = incl;
    __syncthreads();   // also: every thread is done reading sK
    for (int w = (int)(threadIdx.x >> 5) - 1; w >= 0 && prev == NONE; w--) prev = s_last[w];
    uint32_t hm = 0;
#pragma unroll
    for (int e = 0; e < MgCfg<W>::ITEMS; e++)
        if ((emit_mask >> e) & 1u) {
            const uint32_t p = (uint32_t)(out[e] >> P.suffix_bits);
            if (p != prev) hm |= 1u << e;
            prev = p;
        }
    uint32_t packed_tot;
    const uint32_t packed = block_excl_scan<uint32_t, MG_THREADS>(__popc(emit_mask) | (__popc(hm) << 16), s_tmp, packed_tot);
    const uint32_t off = packed & 0xFFFFu, tile_emitted = packed_tot & 0xFFFFu, n_heads = packed_tot >> 16;
    uint32_t hoff = packed >> 16;
    // publish the tile's count at once, stage the output while the predecessors finish, resolve the prefix last
    if (threadIdx.x == 0) lookback_publish(status, tile, tile_emitted);
    {
        uint32_t k = off;
#pragma unroll
        for (int e = 0; e < MgCfg<W>::ITEMS; e++)
            if ((emit_mask >> e) & 1u) {
                s_out[mg_pad(k)] = (Suf)(out[e] & low_mask<W>(P.suffix_bits));
                if ((hm >> e) & 1u) { s_pos[hoff] = (uint16_t)k; s_pfx[hoff] = (uint32_t)(out[e] >> P.suffix_bits); hoff++; }
                k++;
            }
    }
    if (threadIdx.x < 32) {
        const uint64_t e = lookback_resolve(status, tile, tile_emitted);
        if (threadIdx.x == 0) s_excl = e;
    }
    __syncthreads();
    const uint64_t excl = s_excl;
    for (uint32_t k = threadIdx.x; k < tile_emitted; k += MG_THREADS) suf_out[excl + k] = s_out[mg_pad(k)];
    for (uint32_t h = threadIdx.x; h < n_heads; h += MG_THREADS) {
        const uint32_t p0 = s_pos[h], p1 = h + 1 < n_heads ? s_pos[h + 1] : tile_emitted;
        atomicAdd(prefix_cnt + s_pfx[h], p1 - p0);
    }
    if (threadIdx.x == 0 && D1 == nA + nB) *n_out = excl + tile_emitted;
}

// Dense per-prefix counters -> directory.  One warp handles 32 directory words (1024 prefixes):
// coalesced counter rows, ballots give the bit words, warp sums the element counts.
// dir[w] = {bits, rank}; word_off[w] = number of elements in prefixes below 32 * w; totals[0] = nb, totals[1] = n
// (the caller passes its totals array + 1: slot 0 of that array is the merge kernel's own element count).
static __global__ void __launch_bounds__(256) dir_bits_kernel(const uint32_t* __restrict__ prefix_cnt, uint64_t n_words, uint2* __restrict__ dir,
                                                       uint32_t* __restrict__ word_off, volatile uint64_t* status_rank,
                                                       volatile uint64_t* status_off, uint32_t* tile_counter,
                                                       unsigned long long* __restrict__ totals) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;
    __shared__ uint32_t s_tmp[33];
    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const int lane = threadIdx.x & 31;
    const uint64_t w = (uint64_t)tile * 256 + threadIdx.x;                   // this thread's directory word
    const uint64_t row0 = ((uint64_t)tile * 256 + (threadIdx.x & ~31)) * 32;  // first prefix of the warp's 32 words
    uint32_t bits = 0, sum = 0;
    for (int r = 0; r < 32; r++) {
        const uint64_t wr = (uint64_t)tile * 256 + (threadIdx.x & ~31) + r;
        const uint32_t v = wr < n_words ? prefix_cnt[row0 + (uint64_t)r * 32 + lane] : 0u;
        const uint32_t bal = __ballot_sync(0xffffffffu, v != 0);
        const uint32_t sm = warp_sum(v);
        if (lane == r) { bits = bal; sum = sm; }
    }
    uint32_t tot_bits, tot_sum;
    const uint32_t rk = block_excl_scan<uint32_t, 256>(__popc(bits), s_tmp, tot_bits);
    const uint32_t of = block_excl_scan<uint32_t, 256>(sum, s_tmp, tot_sum);
    const uint64_t ex_rank = block_lookback(status_rank, tile, tot_bits, &s_excl);
    const uint64_t ex_off = block_lookback(status_off, tile, tot_sum, &s_excl);
    if (w < n_words) {
        dir[w] = make_uint2(bits, (uint32_t)ex_rank + rk);
        word_off[w] = (uint32_t)ex_off + of;
    }
    if (threadIdx.x == 0 && (uint64_t)(tile + 1) * 256 >= n_words && (uint64_t)tile * 256 < n_words) {
        totals[0] = ex_rank + tot_bits;
        totals[1] = ex_off + tot_sum;
    }
}

// bucket_prefix / bucket_off / bucket_range of every occupied prefix (thread = directory word).  The bucket and
// element totals are read from the device (totals[1] = buckets, totals[2] = elements, written by dir_bits_kernel), so
// the host never waits for them in the middle of a mutation; totals[3] receives the prefix of the last bucket.
static __global__ void dir_fill_kernel(const uint32_t* __restrict__ prefix_cnt, const uint2* __restrict__ dir, const uint32_t* __restrict__ word_off,
                                uint64_t n_words, uint32_t* __restrict__ bucket_prefix, uint32_t* __restrict__ bucket_off,
                                uint2* __restrict__ bucket_range, unsigned long long* __restrict__ totals) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nb = (uint32_t)totals[1];
    if (w == 0) bucket_off[nb] = (uint32_t)totals[2];
    if (w >= n_words) return;
    const uint2 e = dir[w];
    uint32_t bits = e.x, r = e.y, o = word_off[w];
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const uint32_t p = (uint32_t)(w * 32 + b), c = prefix_cnt[p];
        bucket_prefix[r] = p;
        bucket_off[r] = o;
        bucket_range[r] = make_uint2(o, o + c);
        if (r + 1 == nb) totals[3] = p;
        o += c;
        r++;
    }
}

}  // namespace cbl
