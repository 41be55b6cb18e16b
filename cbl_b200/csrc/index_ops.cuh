// Kernels k3b-k8: unique, edit generation (probe), bucket directory rebuild (bitvector + popcount rank
// directory + CSR offsets), the streaming merge / anti-merge that rewrites the suffix array, and the
// CSR -> words expansion used by iter/export and the set operations.
//
// Together they replace WordSet::insert_batch / remove_batch / contains_batch
// (src/wordset/mod.rs:163-237), the bucket containers (src/trievec/mod.rs, src/trie.rs), the dynamic
// rank bitvector (cxx/rank_bv.h) and the binary set operations (src/wordset/set_ops.rs:78-410,
// src/trievec/set_ops.rs:5-257, src/bitvector/set_ops.rs:4-106).
#pragma once
#include "index_view.cuh"

namespace cbl {

constexpr int OP_THREADS = 256;
constexpr int OP_ITEMS = 8;
constexpr int OP_TILE = OP_THREADS * OP_ITEMS;

// ---------------------------------------------------------------------------------------------
// unique: sorted keys -> distinct keys (single pass, look-back).  *n_out receives the count.
// ---------------------------------------------------------------------------------------------
template <class W>
__global__ void __launch_bounds__(OP_THREADS) unique_kernel(const W* __restrict__ in, uint64_t n, W* __restrict__ out,
                                                            volatile uint64_t* status, uint32_t* tile_counter,
                                                            unsigned long long* __restrict__ n_out) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;
    __shared__ uint32_t s_tmp[33];
    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const uint64_t base = (uint64_t)tile * OP_TILE + (uint64_t)threadIdx.x * OP_ITEMS;
    W k[OP_ITEMS];
    bool head[OP_ITEMS];
    uint32_t cnt = 0;
    W prev = (base > 0 && base <= n) ? in[base - 1] : (W)0;
#pragma unroll
    for (int i = 0; i < OP_ITEMS; i++) {
        const uint64_t idx = base + i;
        const bool valid = idx < n;
        k[i] = valid ? in[idx] : (W)0;
        head[i] = valid && (idx == 0 || k[i] != prev);
        prev = k[i];
        cnt += head[i];
    }
    uint32_t total;
    uint32_t off = block_excl_scan<uint32_t, OP_THREADS>(cnt, s_tmp, total);
    const uint64_t excl = block_lookback(status, tile, total, &s_excl);
    uint64_t o = excl + off;
#pragma unroll
    for (int i = 0; i < OP_ITEMS; i++)
        if (head[i]) out[o++] = k[i];
    const uint64_t tile_end = (uint64_t)(tile + 1) * OP_TILE;
    if (threadIdx.x == 0 && tile_end >= n && (uint64_t)tile * OP_TILE < n) *n_out = excl + total;
}

// ---------------------------------------------------------------------------------------------
// Edit generation.  For every (sorted, distinct) probe key decide what happens to the TARGET index:
//   EDIT_INS      key absent from target      -> insert it            (insert_seq, |=, ^=)
//   EDIT_DEL      key present in target       -> delete that element  (remove_seq, -=, ^=)
//   EDIT_KEEP_ONLY (self mode) the probe keys ARE the target's own elements (element j <-> key j);
//                 they are looked up in `ix` (the OTHER index) and deleted from the target when
//                 absent there                                         (&=)
// Outputs: ins_key / ins_vpos (virtual position = insertion point + number of earlier inserts),
// del_idx (ascending target positions), per-target-bucket size deltas, and the bits of brand-new
// prefixes OR-ed into new_dir (a copy of the target's directory; ranks are recomputed afterwards).
// ---------------------------------------------------------------------------------------------
enum : int { EDIT_INS = 1, EDIT_DEL = 2, EDIT_KEEP_ONLY = 4 };

template <class W, class Suf>
__global__ void __launch_bounds__(OP_THREADS) probe_edits_kernel(
    const W* __restrict__ keys, uint64_t n, IndexView<Suf> ix, IndexView<Suf> target, KParams P, int mode,
    W* __restrict__ ins_key, uint64_t* __restrict__ ins_vpos, uint64_t* __restrict__ del_idx, int* __restrict__ delta,
    uint2* __restrict__ new_dir, volatile uint64_t* status_ins, volatile uint64_t* status_del,
    uint32_t* tile_counter, unsigned long long* __restrict__ counts /* [0]=n_ins [1]=n_del */) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;
    __shared__ uint32_t s_tmp[33];
    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const uint64_t base = (uint64_t)tile * OP_TILE + (uint64_t)threadIdx.x * OP_ITEMS;
    W k[OP_ITEMS];
    uint64_t apos[OP_ITEMS];
    bool ins[OP_ITEMS], del[OP_ITEMS];
    uint32_t n_ins = 0, n_del = 0;
#pragma unroll
    for (int i = 0; i < OP_ITEMS; i++) {
        const uint64_t idx = base + i;
        ins[i] = del[i] = false;
        apos[i] = 0;
        if (idx < n) {
            k[i] = keys[idx];
            ProbeResult r = probe_key<W, Suf>(ix, P, k[i]);
            if (mode & EDIT_KEEP_ONLY) {
                if (!r.found) {
                    del[i] = true;
                    apos[i] = idx;
                    uint32_t prefix = (uint32_t)(k[i] >> P.suffix_bits), trank;
                    dir_test_rank(target.dir, prefix, trank);
                    atomicAdd(delta + trank, -1);
                }
            } else {
                apos[i] = r.pos;
                if (!r.found && (mode & EDIT_INS)) {
                    ins[i] = true;
                    if (r.prefix_present) atomicAdd(delta + r.rank, 1);
                    else {
                        uint32_t prefix = (uint32_t)(k[i] >> P.suffix_bits);
                        atomicOr(dir_bits_word(new_dir, prefix), 1u << (prefix & 31));
                    }
                }
                if (r.found && (mode & EDIT_DEL)) {
                    del[i] = true;
                    atomicAdd(delta + r.rank, -1);
                }
            }
        }
        n_ins += ins[i];
        n_del += del[i];
    }
    uint32_t tot_ins, tot_del;
    uint32_t off_ins = block_excl_scan<uint32_t, OP_THREADS>(n_ins, s_tmp, tot_ins);
    uint32_t off_del = block_excl_scan<uint32_t, OP_THREADS>(n_del, s_tmp, tot_del);
    const uint64_t ex_ins = block_lookback(status_ins, tile, tot_ins, &s_excl);
    const uint64_t ex_del = block_lookback(status_del, tile, tot_del, &s_excl);
    uint64_t oi = ex_ins + off_ins, od = ex_del + off_del;
#pragma unroll
    for (int i = 0; i < OP_ITEMS; i++) {
        if (ins[i]) { ins_key[oi] = k[i]; ins_vpos[oi] = apos[i] + oi; oi++; }
        if (del[i]) { del_idx[od] = apos[i]; od++; }
    }
    if (threadIdx.x == 0 && (uint64_t)(tile + 1) * OP_TILE >= n && (uint64_t)tile * OP_TILE < n) {
        counts[0] = ex_ins + tot_ins;
        counts[1] = ex_del + tot_del;
    }
}

// ---------------------------------------------------------------------------------------------
// Bucket directory rebuild
// ---------------------------------------------------------------------------------------------
// (a) clear the bits of buckets that end up empty
static __global__ void clear_emptied_kernel(const uint32_t* __restrict__ bucket_prefix, const uint32_t* __restrict__ bucket_off,
                                     const int* __restrict__ delta, uint32_t nb, uint2* __restrict__ new_dir) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    int sz = (int)(bucket_off[r + 1] - bucket_off[r]) + delta[r];
    if (sz == 0) {
        uint32_t p = bucket_prefix[r];
        atomicAnd(dir_bits_word(new_dir, p), ~(1u << (p & 31)));
    }
}

// (b) rank directory: dir[i].y = number of set bits before word i; *nb_out = total set bits.
// Single pass: per-thread popcounts, warp-shuffle scan, block scan, decoupled look-back across tiles.
static __global__ void __launch_bounds__(OP_THREADS) rank_directory_kernel(uint2* __restrict__ dir, uint64_t n_words, volatile uint64_t* status,
                                                                    uint32_t* tile_counter, unsigned long long* __restrict__ nb_out) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;
    __shared__ uint32_t s_tmp[33];
    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const uint64_t base = (uint64_t)tile * OP_TILE + (uint64_t)threadIdx.x * OP_ITEMS;
    uint32_t c[OP_ITEMS], cnt = 0;
#pragma unroll
    for (int i = 0; i < OP_ITEMS; i++) {
        const uint64_t w = base + i;
        c[i] = w < n_words ? __popc(dir[w].x) : 0;
        cnt += c[i];
    }
    uint32_t total;
    uint32_t off = block_excl_scan<uint32_t, OP_THREADS>(cnt, s_tmp, total);
    const uint64_t excl = block_lookback(status, tile, total, &s_excl);
    uint32_t run = (uint32_t)excl + off;
#pragma unroll
    for (int i = 0; i < OP_ITEMS; i++) {
        const uint64_t w = base + i;
        if (w < n_words) dir[w].y = run;
        run += c[i];
    }
    if (threadIdx.x == 0 && (uint64_t)(tile + 1) * OP_TILE >= n_words && (uint64_t)tile * OP_TILE < n_words) *nb_out = excl + total;
}

// (c) surviving old buckets -> their new rank
static __global__ void fill_sizes_old_kernel(const uint32_t* __restrict__ bucket_prefix, const uint32_t* __restrict__ bucket_off,
                                      const int* __restrict__ delta, uint32_t nb, const uint2* __restrict__ new_dir,
                                      uint32_t* __restrict__ size_new, uint32_t* __restrict__ prefix_new) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb) return;
    int sz = (int)(bucket_off[r + 1] - bucket_off[r]) + delta[r];
    if (sz > 0) {
        uint32_t p = bucket_prefix[r], nr;
        dir_test_rank(new_dir, p, nr);
        size_new[nr] = (uint32_t)sz;
        prefix_new[nr] = p;
    }
}

// (d) inserted keys whose prefix did not exist before -> count them into their new bucket
template <class W>
__global__ void fill_sizes_ins_kernel(const W* __restrict__ ins_key, uint64_t ni, KParams P, const uint2* __restrict__ old_dir,
                                      const uint2* __restrict__ new_dir, uint32_t* __restrict__ size_new,
                                      uint32_t* __restrict__ prefix_new) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    uint32_t p = (uint32_t)(ins_key[i] >> P.suffix_bits);
    const bool was = (__ldg(old_dir + (p >> 5)).x >> (p & 31)) & 1u;
    if (!was) {
        uint32_t nr;
        dir_test_rank(new_dir, p, nr);
        atomicAdd(size_new + nr, 1u);
        prefix_new[nr] = p;
    }
}

// (e) exclusive scan of u32 sizes -> u32 offsets (n entries in, n+1 out: out[n] = total)
static __global__ void __launch_bounds__(OP_THREADS) scan_sizes_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ out,
                                                                uint2* __restrict__ range, volatile uint64_t* status,
                                                                uint32_t* tile_counter) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;
    __shared__ uint32_t s_tmp[33];
    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const uint64_t base = (uint64_t)tile * OP_TILE + (uint64_t)threadIdx.x * OP_ITEMS;
    uint32_t c[OP_ITEMS], cnt = 0;
#pragma unroll
    for (int i = 0; i < OP_ITEMS; i++) {
        c[i] = (base + i < n) ? in[base + i] : 0u;
        cnt += c[i];
    }
    uint32_t total;
    uint32_t off = block_excl_scan<uint32_t, OP_THREADS>(cnt, s_tmp, total);
    const uint64_t excl = block_lookback(status, tile, total, &s_excl);
    uint32_t run = (uint32_t)excl + off;
#pragma unroll
    for (int i = 0; i < OP_ITEMS; i++) {
        if (base + i <= n) out[base + i] = run;  // includes the terminal entry out[n]
        if (base + i < n) range[base + i] = make_uint2(run, run + c[i]);
        run += c[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming merge / anti-merge: new suffix array = old one with `ins` spliced in and `del` dropped.
// Tiles cover the virtual stream (old elements interleaved with the inserted ones); every old
// suffix is read once and every surviving suffix written once.
// ---------------------------------------------------------------------------------------------
template <class W, class Suf>
__global__ void __launch_bounds__(OP_THREADS) apply_edits_kernel(const Suf* __restrict__ suf_old, uint64_t n_old,
                                                                 const uint64_t* __restrict__ ins_vpos, const W* __restrict__ ins_key,
                                                                 uint64_t ni, const uint64_t* __restrict__ del_idx, uint64_t nd,
                                                                 Suf* __restrict__ out, KParams P, volatile uint64_t* status,
                                                                 uint32_t* tile_counter) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;
    __shared__ uint32_t s_tmp[33];
    __shared__ uint64_t s_bounds[4];
    __shared__ uint8_t s_ins[OP_TILE];
    __shared__ uint8_t s_dead[OP_TILE];
    const uint64_t V = n_old + ni;
    const uint32_t tile = block_ticket(tile_counter, &s_tile);
    const uint64_t v0 = (uint64_t)tile * OP_TILE;
    const uint64_t v1 = min(V, v0 + (uint64_t)OP_TILE);
    if (threadIdx.x == 0) {
        uint64_t j_lo = lower_bound_dev<uint64_t>(ins_vpos, ni, v0);
        uint64_t j_hi = lower_bound_dev<uint64_t>(ins_vpos, ni, v1);
        s_bounds[0] = j_lo;
        s_bounds[1] = j_hi;
        uint64_t a_lo = v0 - j_lo, a_hi = v1 - j_hi;
        s_bounds[2] = lower_bound_dev<uint64_t>(del_idx, nd, a_lo);
        s_bounds[3] = lower_bound_dev<uint64_t>(del_idx, nd, a_hi);
    }
    for (int i = threadIdx.x; i < OP_TILE; i += OP_THREADS) { s_ins[i] = 0; s_dead[i] = 0; }
    __syncthreads();
    const uint64_t j_lo = s_bounds[0], j_hi = s_bounds[1], d_lo = s_bounds[2], d_hi = s_bounds[3];
    const uint64_t a_lo = v0 - j_lo;
    for (uint64_t j = j_lo + threadIdx.x; j < j_hi; j += OP_THREADS) s_ins[ins_vpos[j] - v0] = 1;
    for (uint64_t d = d_lo + threadIdx.x; d < d_hi; d += OP_THREADS) s_dead[del_idx[d] - a_lo] = 1;
    __syncthreads();

    const int s0 = threadIdx.x * OP_ITEMS;
    uint32_t nins = 0;
#pragma unroll
    for (int e = 0; e < OP_ITEMS; e++) nins += s_ins[s0 + e];
    uint32_t tot_ins;
    uint32_t ins_before = block_excl_scan<uint32_t, OP_THREADS>(nins, s_tmp, tot_ins);
    Suf val[OP_ITEMS];
    bool live[OP_ITEMS];
    uint32_t nlive = 0;
#pragma unroll
    for (int e = 0; e < OP_ITEMS; e++) {
        const int s = s0 + e;
        live[e] = false;
        if (v0 + s < v1) {
            if (s_ins[s]) {
                W key = ins_key[j_lo + ins_before];
                val[e] = (Suf)(key & low_mask<W>(P.suffix_bits));
                live[e] = true;
                ins_before++;
            } else {
                const uint32_t arel = (uint32_t)s - ins_before;
                if (!s_dead[arel]) {
                    val[e] = suf_old[a_lo + arel];
                    live[e] = true;
                }
            }
        }
        nlive += live[e];
    }
    uint32_t tot_live;
    uint32_t off = block_excl_scan<uint32_t, OP_THREADS>(nlive, s_tmp, tot_live);
    const uint64_t excl = block_lookback(status, tile, tot_live, &s_excl);
    uint64_t o = excl + off;
#pragma unroll
    for (int e = 0; e < OP_ITEMS; e++)
        if (live[e]) out[o++] = val[e];
}

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

// words -> membership flags (single-k-mer API and the sharded path after the all-to-all)
template <class W, class Suf>
__global__ void probe_words_kernel(const W* __restrict__ words, uint64_t n, IndexView<Suf> ix, KParams P, uint8_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = contains_key<W, Suf>(ix, P, words[i]) ? 1 : 0;
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
