// Sharded contains_seq as ONE kernel per GPU: every warp alternates between PRODUCING (2-bit encode + necklace of a
// slice of this rank's reads, every word stored straight into the receive region of the rank that owns its prefix, over
// NVLink peer memory) and CONSUMING (probing a block of words some rank stored into THIS rank's receive buffer, answers
// stored straight back into the asking rank's answer buffer).  Replaces, for the prefix-sharded set, the per-chunk loop
// of CBL::contains_seq (src/cbl.rs:311-324): get_seq_words (src/cbl.rs:247-289) on the source rank, WordSet::contains_batch
// (src/wordset/mod.rs:160-185) on the owner.
//
// Why one kernel.  On one GPU the integer-bound necklace hides under the memory stalls of the probe because every warp
// does both (seq_words_kernel MODE 1: 10.2 + 18.3 -> 20.1 ms per 1 G k-mers).  Sharded, a word is made on one GPU and
// probed on another, and with two kernels (route, then probe; or producer and consumer co-resident with capped grids)
// the two halves add up: the probe needs ~48 warps per SM to cover its latency, the producer needs the same issue
// slots, and the register file cannot hold both populations.  Here every resident warp is both: it takes a produce
// task, then looks (without blocking) whether the block whose ticket it holds has been completed by the peers; the SM
// sees the same statistical mix of integer work and outstanding loads as in the single-GPU kernel.
//
// Protocol.  The receive buffers hold the SENTINEL (all ones: never a valid word, see sq_sentinel) wherever no word
// has arrived, so a word is its own "I have landed" flag and no fence, counter or release/acquire pair is needed on the
// data path (an aligned 8- or 16-byte store by one thread arrives whole):
//   source s, per staged unit of <= SQ_UNIT words: owner and rank-inside-the-owner's-run of every word in (warp-private)
//   shared memory, space in region (s -> d) reserved with ONE device-local atomic per (unit, owner), the words stored
//   straight to the owner.  When all words of s are reserved, s publishes final(s -> d) = epoch << 48 | (words sent + 1)
//   (low bits all ones if the region ran out of capacity: nothing of the overflow was written); the epoch tag tells the
//   owner that the value belongs to this call, so the slots need no zeroing (and no barrier) between calls.
//   owner d: tickets t -> (block t / g of source t % g); a block is taken up when its LAST word (of SQ_BLOCK or, once
//   final is known, of what is left of the region) differs from the sentinel: runs are reserved and stored in order, so
//   the rest has landed or is in flight, and the consumer waits for the odd straggler as it goes through the block.  It
//   puts the sentinel back as it reads, so the buffer is clean for the next call.  A warp never waits for a block while produce tasks remain, and
//   producers never wait for consumers, so the kernels of the g ranks cannot deadlock; only after a rank's own
//   production is finished do its warps spin (with a time-out that raises an error flag) on blocks that other ranks
//   are still filling.
#pragma once
#include "seq_words.cuh"

namespace cbl {

constexpr int SQ_THREADS = 64;   // two independent warps per CTA (no block-level barrier after the prologue)
constexpr int SQ_UNIT = 256;     // words staged per flush (8 rounds of 32 k-mers)
constexpr int SQ_BLOCK = 1024;   // words per consumer block (one ticket)
constexpr int SQ_HALF = 1024;    // k-mers per produce task: half a reference chunk, lane l packs bases [32 l, 32 l + 32)
constexpr int SQ_MAX_RANKS = ROUTE_MAX_SPLIT + 1;
#ifndef CBL_SQ_OPAQUE_IDS
#define CBL_SQ_OPAQUE_IDS 1
#endif
#ifndef CBL_SQ_MIN_BLOCKS
#define CBL_SQ_MIN_BLOCKS 24     // same residency as the single-GPU fused probe
#endif
constexpr unsigned long long SQ_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;   // a peer that never shows up
constexpr unsigned long long SQ_FINAL_OVERFLOW = (1ull << 48) - 1;          // low 48 bits of a final: the region overflowed
constexpr uint32_t SQ_SPIN_MAX = 50u * 1000 * 1000;                        // ... or a straggler word that never lands (>= 10 s)

template <class W> struct ShardQueryArgs {
    // ---- producing side (this rank as a source)
    uint32_t split[ROUTE_MAX_SPLIT];                  // splitters; unused ones 0xFFFFFFFF
    int suffix_bits;
    W* peer[SQ_MAX_RANKS];                            // my region inside owner d's receive buffer
    unsigned long long* peer_final[SQ_MAX_RANKS];     // my final-count slot at owner d
    unsigned long long* cnt;                          // [16] words reserved per owner (local, zero on entry); [16] their sum
    uint32_t* pos;                                    // [k-mers] d * cap + index inside region d: where the answer will be
    unsigned long long cap;                           // words per region
    unsigned long long n_kmers;                       // words this rank produces
    unsigned* prod_next;                              // local: next produce task (zero on entry)
    uint32_t n_tasks;                                 // 2 * chunks
    // ---- consuming side (this rank as an owner)
    W* seg_words[SQ_MAX_RANKS];                       // region of my receive buffer written by source s (sentinel-filled)
    uint8_t* seg_out[SQ_MAX_RANKS];                   // my region inside source s's answer buffer
    const unsigned long long* final_[SQ_MAX_RANKS];   // local slot where source s publishes epoch << 48 | (words sent + 1)
    unsigned long long epoch;                         // 1 .. 65535, the same on every rank, different from the previous call's
    unsigned* ticket;                                 // local, zero on entry
    uint32_t max_blocks;                              // cap / SQ_BLOCK
    int g;                                            // ranks
    unsigned long long* err;                          // [0] smallest offending byte offset (ULLONG_MAX: none), [1] time-out flag,
                                                      // [2] / [3] globaltimer at kernel start / when the production was complete (trace)
    int dev_flags;                                    // developer knobs (CBL_SQ_FLAGS): 1 = produce only (no answers), 2 = consume only after the production
};

// all ones is never a word: a word has 2K + POS_BITS bits with pos < 2K in its low POS_BITS bits, so either the type has
// spare zero bits on top or (K = 29, 64 bits used) pos = 58..63 does not occur
template <class W> __device__ __forceinline__ W sq_sentinel() { return ~(W)0; }
__device__ __forceinline__ void st_word_plain(uint64_t* p, uint64_t v) { *p = v; }
__device__ __forceinline__ void st_word_plain(u128* p, u128 v) {
    *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2((unsigned long long)v, (unsigned long long)(v >> 64));   // one 16-byte store
}

// has the word of this slot landed?  An aligned 8-byte store arrives whole.  A 16-byte store is one transaction too, but the
// memory model does not promise it: the high half of a 128-bit word always has zero top bits (2K + POS_BITS <= 125), so
// "high half all ones" means not landed for sure, and "low half all ones under a landed high half" is either a torn pair or
// (once in 2^64 words) a real value — wait a little, then take it
__device__ __forceinline__ bool sq_missing(uint64_t w, uint32_t) { return w == ~0ull; }
__device__ __forceinline__ bool sq_missing(u128 w, uint32_t spins) {
    const uint64_t hi = (uint64_t)(w >> 64), lo = (uint64_t)w;
    return hi == ~0ull || (lo == ~0ull && spins < 2000u);
}

__device__ __forceinline__ unsigned long long sq_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long sq_ld_sys(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}

// my production is complete (every word has its place reserved): tell every owner how many words I sent it.  No ordering
// with the words themselves is needed: the owner waits until that many words differ from the sentinel.
template <class W> __device__ __forceinline__ void publish_finals_dev(const ShardQueryArgs<W>& a, int lane) {
    if (lane < a.g) {
        const unsigned long long c = atomicAdd(a.cnt + lane, 0ull);
        const unsigned long long v = (a.epoch << 48) | (c > a.cap ? SQ_FINAL_OVERFLOW : c + 1);
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.peer_final[lane]), "l"(v) : "memory");
    }
    if (lane == 0) a.err[3] = sq_now_ns();
}

// everything a warp keeps in shared memory, in one block so that one base register addresses all of it.  A warp is
// either staging a unit or probing a block: the two layouts share the bytes.
template <class W, class Suf> struct SqWarpMem {
    static constexpr int QN = PENDING_CAP;
    struct Stage {
        W word[SQ_UNIT];
        uint16_t dr[SQ_UNIT];          // owner << 12 | rank inside the owner's run
    };
    struct Probe {
        uint4 q_a[QN];                 // {L, R, g, span}
        PendingKey<Suf> q_b[QN];       // {suffix, slot | rounds << 16}
        uint8_t flags[SQ_BLOCK];
    };
    union alignas(16) {
        Stage st;
        Probe pr;
    } u;
    W* dst[16];                        // unit: where the run of owner d starts (null: the region is full)
    uint32_t pos0[16];                 // unit: d * cap + start of the run
    uint32_t cnt[16];
};

// CANON: the set is canonical (compile-time, so that the plain instantiation carries none of the F6 bookkeeping in registers)
template <class W, class Suf, int WB, bool CANON>
__global__ void __launch_bounds__(SQ_THREADS, CBL_SQ_MIN_BLOCKS) shard_query_kernel(SeqBatch b, KParams P, IndexView<Suf> ix, ShardQueryArgs<W> a) {
    constexpr int WN = Window<Suf, WB>::N;
    constexpr int QN = PENDING_CAP;
    __shared__ SqWarpMem<W, Suf> s_w[SQ_THREADS / 32];
    __shared__ uint32_t s_split[16];
    // lane and warp index as opaque register values: derived from threadIdx.x the compiler re-reads the special register and
    // redoes the shift / multiply at nearly every shared-memory access under this kernel's register pressure (measured: 5 %
    // of all instructions); an asm volatile result cannot be rematerialised
    int lane;
    uint32_t widx;
#if CBL_SQ_OPAQUE_IDS
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("{\n .reg .u32 t;\n mov.u32 t, %%tid.x;\n shr.u32 %0, t, 5;\n}" : "=r"(widx));
#else
    lane = threadIdx.x & 31;
    widx = threadIdx.x >> 5;
#endif
    if (threadIdx.x < 16) {
        s_split[threadIdx.x] = threadIdx.x < ROUTE_MAX_SPLIT ? a.split[threadIdx.x] : 0xFFFFFFFFu;
        s_w[0].cnt[threadIdx.x] = 0;
        s_w[1].cnt[threadIdx.x] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.err[2] = sq_now_ns();
    __syncthreads();   // the only block-level barrier: from here on the two warps are independent
    SqWarpMem<W, Suf>& wm = s_w[widx];

    // ---- produce task t: half `t & 1` of chunk `t >> 1` ---------------------------------------------------------
    auto produce = [&](uint32_t task) {
        const uint64_t chunk = task >> 1;
        const int half = (int)(task & 1u);
        uint32_t piece = 0;
        if (lane == 0) piece = (uint32_t)(upper_bound_dev<uint64_t>(b.piece_chunk0, (uint64_t)b.n_pieces + 1, chunk) - 1);
        piece = __shfl_sync(0xffffffffu, piece, 0);
        const uint64_t ci = chunk - b.piece_chunk0[piece];
        const uint32_t pk = b.piece_kmers[piece];
        const uint32_t ks = (uint32_t)(ci * CHUNK_KMERS);
        const int m = (int)min((uint32_t)CHUNK_KMERS, pk - ks);   // k-mers in the chunk
        const int kbase = SQ_HALF * half;
        if (kbase >= m) return;                                   // warp-uniform
        const uint8_t* cbase = b.seq + b.piece_byte[piece] + ks;
        const int nbytes = m + P.k - 1;
        uint32_t* const opos = a.pos + b.piece_out[piece] + ks;
        bool bad = false;
        uint64_t Wd, H = 0;
        {
            const int off = kbase + 32 * lane;
            Wd = load_pack32(cbase + off, b.seq_end, min(32, nbytes - off), bad);
            if (lane < 2) {
                const int hoff = kbase + SQ_HALF + 32 * lane;
                H = load_pack32(cbase + hoff, b.seq_end, min(32, nbytes - hoff), bad);
            }
            if (bad) atomicMin(a.err, (unsigned long long)(b.piece_byte[piece] + ks + off));
        }
        const uint64_t Wd1 = __shfl_sync(0xffffffffu, lane == 0 ? H : Wd, (lane + 1) & 31);
        const uint64_t Wd2 = sizeof(W) == 16 ? __shfl_sync(0xffffffffu, lane < 2 ? H : Wd, (lane + 2) & 31) : 0ull;   // a 64-bit window spans two packed words
        // canonical mode: answers of a chunk come "forward k-mers first, then the reverse-complemented ones";
        // lane j keeps the parity ballot of round j of this half, the other half contributes its forward count only
        uint32_t my_bal = 0, fwd_before = 0, nfwd_total = 0;
        if (CANON) {
            for (int j = 0; j < 32; j++) {
                const uint64_t A = __shfl_sync(0xffffffffu, Wd, j), B = __shfl_sync(0xffffffffu, Wd1, j), C = __shfl_sync(0xffffffffu, Wd2, j);
                const W x = cut_window<W>(A, B, C, 2 * lane, P.bits);
                const uint32_t bal = __ballot_sync(0xffffffffu, kbase + 32 * j + lane < m && !(popc_w(x) & 1));
                if (lane == j) my_bal = bal;
            }
            const uint32_t c_mine = warp_sum((uint32_t)__popc(my_bal));
            uint32_t c_other = 0;
            const int obase = SQ_HALF * (1 - half);
            if (obase < m) {
                bool obad = false;   // the other half's warp reports its own bytes
                const int off = obase + 32 * lane;
                const uint64_t oW = load_pack32(cbase + off, b.seq_end, min(32, nbytes - off), obad);
                uint64_t oH = 0;
                if (lane < 2) {
                    const int hoff = obase + SQ_HALF + 32 * lane;
                    oH = load_pack32(cbase + hoff, b.seq_end, min(32, nbytes - hoff), obad);
                }
                const uint64_t oW1 = __shfl_sync(0xffffffffu, lane == 0 ? oH : oW, (lane + 1) & 31);
                const uint64_t oW2 = __shfl_sync(0xffffffffu, lane < 2 ? oH : oW, (lane + 2) & 31);
                for (int j = 0; j < 32; j++) {
                    const uint64_t A = __shfl_sync(0xffffffffu, oW, j), B = __shfl_sync(0xffffffffu, oW1, j), C = __shfl_sync(0xffffffffu, oW2, j);
                    const W x = cut_window<W>(A, B, C, 2 * lane, P.bits);
                    c_other += (obase + 32 * j + lane < m && !(popc_w(x) & 1)) ? 1u : 0u;
                }
                c_other = warp_sum(c_other);
            }
            nfwd_total = c_mine + c_other;
            fwd_before = half == 0 ? 0u : c_other;
        }
        for (int j0 = 0; j0 < 32; j0 += SQ_UNIT / 32) {
            if (kbase + 32 * j0 >= m) break;   // warp-uniform
            const int m_sub = min(SQ_UNIT, m - kbase - 32 * j0);
            // pass 1: words of the unit, owner and rank inside the owner's run
#pragma unroll 1
            for (int jj = 0; jj < SQ_UNIT / 32; jj++) {
                const int j = j0 + jj;
                const uint64_t A = __shfl_sync(0xffffffffu, Wd, j), B = __shfl_sync(0xffffffffu, Wd1, j);
                const uint64_t C = sizeof(W) == 16 ? __shfl_sync(0xffffffffu, Wd2, j) : 0ull;
                const W x = cut_window<W>(A, B, C, 2 * lane, P.bits);
                const int li = jj * 32 + lane;
                if (li < m_sub) {
                    const W word = kmer_to_word<W>(x, P, false);
                    const uint32_t pfx = (uint32_t)(word >> a.suffix_bits);
                    uint32_t d = s_split[7] <= pfx ? 8u : 0u;   // owner = number of splitters <= prefix
                    d += s_split[d + 3] <= pfx ? 4u : 0u;
                    d += s_split[d + 1] <= pfx ? 2u : 0u;
                    d += s_split[d] <= pfx ? 1u : 0u;
                    const uint32_t r = atomicAdd(&wm.cnt[d], 1u);
                    wm.u.st.word[li] = word;
                    wm.u.st.dr[li] = (uint16_t)((d << 12) | r);
                }
            }
            __syncwarp();
            if (lane < 16) {
                const uint32_t c_run = wm.cnt[lane];
                unsigned long long base = 0;
                if (c_run) base = atomicAdd(a.cnt + lane, (unsigned long long)c_run);   // reserve the run in my region at owner `lane`
                wm.dst[lane] = (base + c_run <= a.cap) ? a.peer[lane] + base : nullptr;
                wm.pos0[lane] = (uint32_t)((unsigned long long)lane * a.cap + base);
                wm.cnt[lane] = 0;
            }
            __syncwarp();
            // pass 2: every word straight to its place in its owner's region (the lanes of one owner hold consecutive ranks:
            // contiguous pieces); where its answer will come back
#pragma unroll 1
            for (int jj = 0; jj < SQ_UNIT / 32; jj++) {
                const int j = j0 + jj;
                const int li = jj * 32 + lane;
                uint32_t bal = 0;
                if (CANON) bal = __shfl_sync(0xffffffffu, my_bal, j);
                if (li < m_sub) {
                    const uint32_t dr = wm.u.st.dr[li], d = dr >> 12, r = dr & 4095u;
                    W* const dst = wm.dst[d];
                    if (dst) st_word_plain(dst + r, wm.u.st.word[li]);
                    const uint32_t kidx = (uint32_t)(kbase + 32 * j + lane);
                    uint32_t slot = kidx;
                    if (CANON) {
                        const uint32_t fb = fwd_before + __popc(bal & ((1u << lane) - 1u));
                        slot = ((bal >> lane) & 1u) ? fb : nfwd_total + (kidx - fb);
                    }
                    opos[slot] = wm.pos0[d] + r;
                }
                if (CANON) fwd_before += __popc(bal);
            }
            __syncwarp();   // the staging area is reused by the next unit
        }
        // all of my words reserved?  then tell every owner how many I sent it
        unsigned long long before = 0;
        if (lane == 0) before = atomicAdd(a.cnt + 16, (unsigned long long)(m - kbase < SQ_HALF ? m - kbase : SQ_HALF));
        before = __shfl_sync(0xffffffffu, before, 0);
        if (before + (unsigned long long)min(m - kbase, SQ_HALF) == a.n_kmers) publish_finals_dev(a, lane);
    };
    // ---- consume: probe block `blk` (m words) of source `src` ---------------------------------------------------
    uint32_t q_head = 0, q_count = 0;   // warp-uniform
    auto q_push = [&](bool und, Suf s, uint32_t L, uint32_t R, uint32_t g, uint32_t span, uint32_t slot_it) {
        const uint32_t bal = __ballot_sync(0xffffffffu, und);
        if (und) {
            const uint32_t i = (q_head + q_count + __popc(bal & ((1u << lane) - 1u))) & (QN - 1);
            wm.u.pr.q_a[i] = make_uint4(L, R, g, span);
            PendingKey<Suf> pk;
            pk.s = s;
            pk.slot_it = slot_it;
            wm.u.pr.q_b[i] = pk;
        }
        q_count += __popc(bal);
        __syncwarp();
    };
    auto q_drain = [&]() {
        const uint32_t take = min(q_count, 32u);
        const bool mine = (uint32_t)lane < take;
        const uint32_t i = (q_head + lane) & (QN - 1);
        const uint4 qa = wm.u.pr.q_a[i];
        const PendingKey<Suf> pk = wm.u.pr.q_b[i];
        __syncwarp();
        q_head = (q_head + take) & (QN - 1);
        q_count -= take;
        const Suf s = pk.s;
        uint32_t L = qa.x, R = qa.y, g = qa.z, slot_it = pk.slot_it;
        bool und = false;
        if (mine) {
            const uint32_t base = g & ~(uint32_t)(WN - 1);
            Suf e[WN];
            load_window<Suf, WB>(ix.suf + base, e);
            int r = eval_window<Suf, WB>(e, base, s, key32<Suf>(s, P.suffix_bits), P.suffix_bits, qa.w, L, R, g);
            slot_it += 1u << 16;
            if (r < 0 && (slot_it >> 16) >= (uint32_t)PROBE_MAX_IT) {   // exact fallback: bisect what is left
                const uint32_t R0 = R;
                while (L < R) {
                    const uint32_t mid = L + ((R - L) >> 1);
                    if (ix.suf[mid] < s) L = mid + 1; else R = mid;
                }
                r = (L < R0 && ix.suf[L] == s) ? 1 : 0;
            }
            if (r >= 0) wm.u.pr.flags[slot_it & 0xFFFFu] = (uint8_t)r;
            else und = true;
        }
        q_push(und, s, L, R, g, qa.w, slot_it);
    };
    auto consume = [&](uint32_t src, uint32_t blk, int m) {
        W* in_words = a.seg_words[src] + (uint64_t)blk * SQ_BLOCK;
        uint8_t* of = a.seg_out[src] + (uint64_t)blk * SQ_BLOCK;
        q_head = 0;
        q_count = 0;
        // the word of round j + 1 is fetched while round j is looked up (one more load in flight at the head of the chain
        // word -> directory -> range -> corrections -> window)
        W next = lane < m ? ld_cg_word(in_words + lane) : (W)0;
#pragma unroll 1
        for (int j = 0; 32 * j < m; j++) {
            const int idx = 32 * j + lane;
            const bool active = idx < m;
            W word = next;
            next = idx + 32 < m ? ld_cg_word(in_words + idx + 32) : (W)0;
            if (active) {
                // stored by a peer moments ago (read past L1).  The block's last word has landed, so a word still missing here
                // is in flight (reserved together with words that have arrived): wait for it
                for (uint32_t spin = 0; sq_missing(word, spin); spin++) {
                    if (spin > SQ_SPIN_MAX) { atomicExch(a.err + 1, 2ull); word = (W)0; break; }
                    __nanosleep(200);
                    word = ld_cg_word(in_words + idx);
                }
                st_word_plain(in_words + idx, sq_sentinel<W>());     // leave the slot clean for the next call
            }
            uint32_t prefix;
            Suf s;
            split_key<W, Suf>(word, P, prefix, s);
            const uint32_t k32 = key32<Suf>(s, P.suffix_bits);
            bool present = active && ix.nb != 0 && (prefix >> P.prefix_bits) == 0;   // anything that is not a word of this set (a corrupted buffer) is absent, not a wild read
            const uint2 de = ldg_keep(ix.dir + (present ? (prefix >> 5) : 0u));
            const uint32_t bit = prefix & 31;
            const uint32_t rank = de.y + __popc(de.x & ((1u << bit) - 1u));
            present = present && ((de.x >> bit) & 1u);
            const uint2 range = ldg_keep(ix.bucket_range + (present ? rank : 0u));
            const uint32_t lo = present ? range.x : 0u, hi = present ? range.y : 0u;
            uint32_t gg = lo + predict_slot(ix.sub, lo, hi, k32);
            Suf e[WN];
            load_window<Suf, WB>(ix.suf + (gg & ~(uint32_t)(WN - 1)), e);
            int r = 0;
            uint32_t L = lo, R = hi;
            if (present) r = eval_window<Suf, WB>(e, gg & ~(uint32_t)(WN - 1), s, k32, P.suffix_bits, R - L, L, R, gg);
            if (active && r >= 0) wm.u.pr.flags[idx] = (uint8_t)r;
            q_push(r < 0, s, L, R, gg, hi - lo, (uint32_t)idx);
            while (q_count >= 32) q_drain();
        }
        while (q_count > 0) q_drain();
        __syncwarp();
        // 16 answers per store (the regions are 16-byte aligned: capacity is a multiple of SQ_BLOCK)
        const int m16 = m >> 4;
        for (int i = lane; i < m16; i += 32) reinterpret_cast<uint4*>(of)[i] = reinterpret_cast<const uint4*>(wm.u.pr.flags)[i];
        for (int i = (m16 << 4) + lane; i < m; i += 32) of[i] = wm.u.pr.flags[i];
        __syncwarp();   // the flags area is reused
    };
    // state of the block behind ticket (src, blk): > 0 complete with that many words, 0 not yet, -1 does not exist
    auto block_state = [&](uint32_t src, uint32_t blk) -> int {
        unsigned long long f = 0;
        if (lane == 0) f = sq_ld_sys(a.final_[src]);
        f = __shfl_sync(0xffffffffu, f, 0);
        f = (f >> 48) == a.epoch ? (f & SQ_FINAL_OVERFLOW) : 0ull;   // a value of an earlier call: not yet known
        if (f == SQ_FINAL_OVERFLOW) return -1;
        int m = SQ_BLOCK;
        if (f != 0) {
            const unsigned long long total = f - 1, first = (unsigned long long)blk * SQ_BLOCK;
            if (total <= first) return -1;
            m = (int)min((unsigned long long)SQ_BLOCK, total - first);
        }
        const W* in_words = a.seg_words[src] + (uint64_t)blk * SQ_BLOCK;
        // runs are reserved and stored in order of arrival: when the last word of the block is there, the rest has landed or
        // is in flight (consume waits for a straggler)
        W last = sq_sentinel<W>();
        if (lane == 0) last = ld_cg_word(in_words + (m - 1));
        if (__shfl_sync(0xffffffffu, (int)sq_missing(last, 0u), 0)) return 0;
        return m;
    };

    if (a.n_kmers == 0 && blockIdx.x == 0 && threadIdx.x < 32) publish_finals_dev(a, lane);   // a rank without reads still answers its peers

    bool have_ticket = false, tickets_left = !(a.dev_flags & 1), prod_left = a.n_tasks != 0;
    uint32_t t_src = 0, t_blk = 0;
    unsigned long long wait_since = 0;
    for (;;) {
        if (!have_ticket && tickets_left && !((a.dev_flags & 2) && prod_left)) {
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(a.ticket, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            t_src = t % (uint32_t)a.g;
            t_blk = t / (uint32_t)a.g;
            if (t_blk >= a.max_blocks) tickets_left = false;
            else have_ticket = true;
        }
        if (have_ticket) {
            const int st = block_state(t_src, t_blk);
            if (st != 0) {
                if (st > 0) consume(t_src, t_blk, st);
                have_ticket = false;
                wait_since = 0;
                continue;
            }
        }
        if (prod_left) {
            uint32_t task = 0;
            if (lane == 0) task = atomicAdd(a.prod_next, 1u);
            task = __shfl_sync(0xffffffffu, task, 0);
            if (task >= a.n_tasks) { prod_left = false; continue; }
            produce(task);
            continue;
        }
        if (!have_ticket) break;   // nothing left to produce, no ticket left to take
        // my production is done and the block I hold is still being filled by another rank
        if (wait_since == 0) wait_since = sq_now_ns();
        else if (sq_now_ns() - wait_since > SQ_TIMEOUT_NS) {
            if (lane == 0) atomicExch(a.err + 1, 1ull);
            break;
        }
        __nanosleep(500);
    }
}

}  // namespace cbl
