// Kernel k1+k2: reads -> 2-bit packed windows -> (canonical) -> necklace -> word, fused (optionally)
// with the membership probe.  Replaces src/kmer.rs (from_nucs/append/canonical) + src/necklace/queue.rs
// + CBL::get_seq_words / get_seq_chunks (src/cbl.rs:239-289) for a whole batch of records.
//
// Work decomposition mirrors the reference's own chunking so that results come out in ITS order:
// one CTA (2 warps) per 2048-k-mer chunk of a "piece" (a record, or a 2048-aligned slice of a long
// record).  Inside a warp every lane packs 32 bases (two/three aligned 16-byte loads + funnel shifts)
// into one 64-bit word; iteration j broadcasts words j, j+1(, j+2) by warp shuffle and lane l cuts
// the window starting at base 32j + l out of them, so global stores are fully coalesced (lane l
// writes k-mer 32j + l).  Canonical mode reproduces the per-chunk "forward words first, then the
// reverse-complemented ones" order of src/cbl.rs:248-275 with ballot prefix counts.
#pragma once
#include "index_view.cuh"
#include "radix_sort.cuh"   // DestDigit (owner rank of a word)

namespace cbl {

// Arguments of the sharded modes of seq_words_kernel (multi-GPU path, cbl_b200/sharded.py).
// MODE 2 (source side): every word is stored DIRECTLY into the receive buffer of the rank that owns its prefix
// (peer memory mapped with CUDA IPC, the stores travel over NVLink).  peer[d] already points at THIS rank's
// region inside owner d's buffer; a region holds `cap` words.  Space is reserved chunk by chunk with one atomic
// per (chunk, owner) on the device-local counters cnt[16], so no counting pass and no send buffer exist.
// pos[i] = d * cap + (index of word i inside the region) tells where the answer of k-mer i will come back.
// A region that would overflow is not written; cnt[d] > cap afterwards tells the host to retry with a larger cap.
// MODE 3 (owner side): membership of in_words[0, n_in), one answer byte per word to out_flags, which may itself
// be peer memory (the answer region inside the source rank's buffer).
constexpr int PROBE_MAX_SEG = 16;
template <class W> struct ShardArgs {
    // MODE 3: the words come as up to PROBE_MAX_SEG segments (the per-source regions of a sharded receive buffer), all
    // probed by ONE launch: segment s holds seg_n[s] words, its answers go to seg_out[s] (one byte per word; may be
    // peer memory); seg_chunk0[s] = number of 2048-word chunks in the segments before s
    const W* seg_words[PROBE_MAX_SEG];
    uint8_t* seg_out[PROBE_MAX_SEG];
    unsigned long long seg_n[PROBE_MAX_SEG];
    unsigned long long seg_chunk0[PROBE_MAX_SEG + 1];
    int n_seg = 0;
    // MODE 0: digit histograms of the words for the LSD passes of the batch sort, folded into the kernel that makes
    // the words (saves the separate read of the batch by radix_hist_kernel): hist[p][256] (u64, zeroed by the caller)
    // counts digit hist_first + p, p < hist_np <= SW_HIST_MAX; null = no histogram
    unsigned long long* hist = nullptr;
    int hist_first = 0, hist_np = 0;
    DestDigit<W> dest;                    // MODE 2: owner rank of a word
    W* peer[ROUTE_MAX_SPLIT + 1];         // MODE 2
    unsigned long long* cnt = nullptr;    // MODE 2: [16] words reserved per owner (zeroed by the caller)
    uint32_t* pos = nullptr;              // MODE 2: may be null (mutations need no answers)
    unsigned long long cap = 0;           // MODE 2: words per region
};

constexpr unsigned long long STREAM_OVERFLOW = ~0ull;   // final count of a region whose producer ran out of capacity
constexpr int SW_HIST_MAX = 4;      // digit histograms the words kernel can carry (the hybrid sort uses <= 4 LSD passes)
constexpr int CHUNK_KMERS = 2048;  // src/cbl.rs:67 CHUNK_SIZE
constexpr int SW_THREADS = 64;
constexpr int PENDING_CAP = 64;   // per-warp queue of undecided lookups (power of two, >= 63)
template <class Suf> struct alignas(sizeof(Suf) == 4 ? 8 : 16) PendingKey {
    Suf s;
    uint32_t slot_it;
};

// a word stored by a peer GPU moments ago: read it past the (non-coherent) L1
__device__ __forceinline__ uint64_t ld_cg_word(const uint64_t* p) { return __ldcg(reinterpret_cast<const unsigned long long*>(p)); }
__device__ __forceinline__ u128 ld_cg_word(const u128* p) {
    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(p));
    return ((u128)v.y << 64) | (u128)v.x;
}

struct SeqBatch {
    const uint8_t* seq;          // concatenated record bytes (device)
    const uint8_t* seq_end;      // one past the last readable byte
    const uint64_t* piece_byte;  // [n_pieces]   first byte of the piece
    const uint64_t* piece_out;   // [n_pieces]   output slot of the piece's first k-mer
    const uint32_t* piece_kmers; // [n_pieces]   number of k-mers in the piece
    const uint64_t* piece_chunk0;// [n_pieces+1] prefix sum of chunks per piece
    uint32_t n_pieces;
    uint64_t n_chunks;
};

// 32 bases starting at p (any alignment) -> one 64-bit word, first base most significant.
// Only bytes below `nvalid` are validated (and must be readable below `end`); the rest read as 'A'.
__device__ __forceinline__ uint64_t load_pack32(const uint8_t* p, const uint8_t* end, int nvalid, bool& bad) {
    if (nvalid <= 0) return 0;
    const uintptr_t a = (uintptr_t)p;
    const uint4* a0 = (const uint4*)(a & ~(uintptr_t)15);
    const int sh = (int)(a & 15);
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0;
    v0 = __ldg(a0);  // contains p itself, which is readable since nvalid > 0
    if ((const uint8_t*)(a0 + 1) < end) v1 = __ldg(a0 + 1);
    if (sh != 0 && (const uint8_t*)(a0 + 2) < end) v2 = __ldg(a0 + 2);
    uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
    const int wo = sh >> 2, bs = (sh & 3) * 8;
    uint32_t x[8];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (wo == c) {
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = __funnelshift_r(w[c + j], w[c + j + 1], bs);
        }
    }
    uint64_t packed = 0;
    uint32_t badbits = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int nb = nvalid - 4 * j;
        uint32_t rmask = nb >= 4 ? 0xFFFFFFFFu : nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u);
        uint32_t u = x[j] & 0xDFDFDFDFu;
        uint32_t ok = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
        badbits |= ~ok & rmask;
        uint32_t xv = (x[j] & rmask) | (0x41414141u & ~rmask);  // out-of-range bytes read as 'A'
        packed |= (uint64_t)pack4(xv) << (56 - 8 * j);
    }
    if (badbits) bad = true;
    return packed;
}

template <class W> __device__ __forceinline__ W cut_window(uint64_t A, uint64_t B, uint64_t C, int sh2, int bits);
template <> __device__ __forceinline__ uint64_t cut_window<uint64_t>(uint64_t A, uint64_t B, uint64_t, int sh2, int bits) {
    uint64_t v = sh2 ? ((A << sh2) | (B >> (64 - sh2))) : A;
    return v >> (64 - bits);
}
template <> __device__ __forceinline__ u128 cut_window<u128>(uint64_t A, uint64_t B, uint64_t C, int sh2, int bits) {
    uint64_t hi = sh2 ? ((A << sh2) | (B >> (64 - sh2))) : A;
    uint64_t lo = sh2 ? ((B << sh2) | (C >> (64 - sh2))) : B;
    u128 v = ((u128)hi << 64) | lo;
    return v >> (128 - bits);
}

// MODE 0: write words (W) to out_words.   MODE 1: probe the index, write one byte per k-mer.
// MODE 2: fused route (sharded path, source side): the words go straight to their owner ranks, see ShardArgs.
// MODE 3: the words are read from sa.seg_words instead of being computed (sharded path, owner side, and the
//         word-level contains of the C ABI); answers to out_flags.
// BRUTE: use the normative brute-force necklace instead of the fast one (debug / cross-check).
// U: k-mers per lane processed together.  In MODE 1 the U lookups advance in lock step (directory
// words, bucket ranges, correction bytes, suffix windows), each stage issuing U independent loads.
// Measured on B200: U = 1 with high occupancy (40 registers) beats U = 2 / 4 (80 / 120 registers);
// only U = 1 is instantiated.
#ifndef CBL_SW_OPAQUE_IDS
#define CBL_SW_OPAQUE_IDS 1
#endif
#ifndef CBL_SW_MIN_BLOCKS_WORDS
#define CBL_SW_MIN_BLOCKS_WORDS 24   // word-level probe (MODE 3): 32 blocks (32 registers, no spills) measured 24.4 vs 19.3 ms: more warps thrash L1/L2
#endif
#ifndef CBL_SW_MIN_BLOCKS
#define CBL_SW_MIN_BLOCKS 24   // fused probe: latency bound, occupancy beats a few spilled registers (measured)
#endif
template <class W, class Suf, int MODE, bool BRUTE, int WB, int U>
__global__ void __launch_bounds__(SW_THREADS, (MODE == 3 && U == 1) ? CBL_SW_MIN_BLOCKS_WORDS : ((MODE == 1 && U == 1) ? CBL_SW_MIN_BLOCKS : 1)) seq_words_kernel(SeqBatch b, KParams P, W* __restrict__ out_words,
                                                               uint8_t* __restrict__ out_flags, IndexView<Suf> ix,
                                                               unsigned long long* __restrict__ err_pos, ShardArgs<W> sa) {
    static_assert(32 % U == 0, "U must divide 32");
    constexpr int WN = Window<Suf, WB>::N;
    constexpr bool WORDS_IN = MODE == 3;   // the words are read, not computed
    constexpr bool PROBE = MODE == 1 || WORDS_IN;
    constexpr int QN = PROBE ? PENDING_CAP : 1;
    // MODE 2: the chunk's words (by output slot), owner | rank-inside-owner of every slot, slots in owner order
    constexpr int RN = MODE == 2 ? CHUNK_KMERS : 1;
    __shared__ W s_stage[RN];
    __shared__ uint16_t s_dr[RN], s_inv[RN];
    __shared__ uint32_t s_cnt[16], s_off[17], s_base[16];
    __shared__ W* s_peer[16];
    __shared__ uint32_t s_split[16];   // MODE 2: the splitters (unused ones 0xFFFFFFFF), searched by bisection
    if (MODE == 2) {
        if (threadIdx.x < 16) {
            s_cnt[threadIdx.x] = 0;
            s_peer[threadIdx.x] = sa.peer[threadIdx.x];
            s_split[threadIdx.x] = threadIdx.x < ROUTE_MAX_SPLIT ? sa.dest.split[threadIdx.x] : 0xFFFFFFFFu;
        }
        __syncthreads();
    }
    __shared__ uint32_t s_fwd[2][32];
    __shared__ uint32_t s_piece;
    __shared__ uint32_t s_hist[MODE == 0 ? SW_HIST_MAX * 256 : 1];
    const bool do_hist = MODE == 0 && sa.hist != nullptr;
    uint32_t hc_d0 = 0, hc_n0 = 0, hc_d1 = 0, hc_n1 = 0;   // per-lane run counters of the two top digits (MODE 0 histogram)
    // MODE 1: answers of the chunk (written out coalesced at the end) and, per warp, the queue of
    // lookups their first window did not decide.  Those are not finished on the spot (that would
    // stall the other lanes of the warp, 4 of 5 of which are already done) but collected and worked
    // off 32 at a time, one window per lane per round.
    __shared__ __align__(16) uint8_t s_flags[PROBE ? CHUNK_KMERS : 16];
    __shared__ uint4 q_a[2][QN];                 // {L, R, g, span}
    __shared__ PendingKey<Suf> q_b[2][QN];       // {suffix, slot | rounds << 16}
    // lane / warp index as opaque register values in the probing modes: derived from threadIdx.x they are rematerialised
    // (special-register read + shift / mask) at nearly every shared-memory access under the 40-register budget
    int lane, warp;
#if CBL_SW_OPAQUE_IDS
    if (PROBE) {
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
        asm volatile("{\n .reg .u32 t;\n mov.u32 t, %%tid.x;\n shr.u32 %0, t, 5;\n}" : "=r"(warp));
    } else
#endif
    {
        lane = threadIdx.x & 31;
        warp = threadIdx.x >> 5;
    }
    uint32_t q_head = 0, q_count = 0;  // warp-uniform

    // enqueue the lanes with `und` set (warp-collective)
    auto q_push = [&](bool und, Suf s, uint32_t L, uint32_t R, uint32_t g, uint32_t span, uint32_t slot_it) {
        const uint32_t bal = __ballot_sync(0xffffffffu, und);
        if (und) {
            const uint32_t i = (q_head + q_count + __popc(bal & ((1u << lane) - 1u))) & (QN - 1);
            q_a[warp][i] = make_uint4(L, R, g, span);
            PendingKey<Suf> pk;
            pk.s = s;
            pk.slot_it = slot_it;
            q_b[warp][i] = pk;
        }
        q_count += __popc(bal);
        __syncwarp();
    };
    // one round: the first min(count, 32) queued lookups each examine one more window (warp-collective)
    auto q_drain = [&]() {
        const uint32_t take = min(q_count, 32u);
        const bool mine = (uint32_t)lane < take;
        const uint32_t i = (q_head + lane) & (QN - 1);
        const uint4 a = q_a[warp][i];
        const PendingKey<Suf> pk = q_b[warp][i];
        __syncwarp();
        q_head = (q_head + take) & (QN - 1);
        q_count -= take;
        const Suf s = pk.s;
        uint32_t L = a.x, R = a.y, g = a.z, slot_it = pk.slot_it;
        bool und = false;
        if (mine) {
            const uint32_t base = g & ~(uint32_t)(WN - 1);
            Suf e[WN];
            load_window<Suf, WB>(ix.suf + base, e);
            int r = eval_window<Suf, WB>(e, base, s, key32<Suf>(s, P.suffix_bits), P.suffix_bits, a.w, L, R, g);
            slot_it += 1u << 16;
            if (r < 0 && (slot_it >> 16) >= (uint32_t)PROBE_MAX_IT) {  // exact fallback: bisect what is left
                const uint32_t R0 = R;  // everything at or beyond R0 is known to be > s
                while (L < R) {
                    const uint32_t mid = L + ((R - L) >> 1);
                    if (ix.suf[mid] < s) L = mid + 1; else R = mid;
                }
                r = (L < R0 && ix.suf[L] == s) ? 1 : 0;
            }
            if (r >= 0) s_flags[slot_it & 0xFFFFu] = (uint8_t)r;
            else und = true;
        }
        q_push(und, s, L, R, g, a.w, slot_it);
    };
    const uint64_t n_chunks = MODE == 3 ? sa.seg_chunk0[sa.n_seg] : b.n_chunks;
    for (uint64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        int m;
        W* ow = nullptr;
        uint8_t* of = nullptr;
        uint32_t* opos = nullptr;
        uint64_t Wd = 0, H = 0;
        const W* in_words = nullptr;   // MODE 3 / 4: first word of this chunk
        if (MODE == 3) {
            int sg = 0;
#pragma unroll
            for (int t = 1; t < PROBE_MAX_SEG; t++) sg += (t < sa.n_seg && sa.seg_chunk0[t] <= chunk) ? 1 : 0;
            const uint64_t w0 = (chunk - sa.seg_chunk0[sg]) * CHUNK_KMERS;
            m = (int)min((unsigned long long)CHUNK_KMERS, sa.seg_n[sg] - w0);
            of = sa.seg_out[sg] + w0;
            in_words = sa.seg_words[sg] + w0;
        } else {
            if (threadIdx.x == 0) s_piece = (uint32_t)(upper_bound_dev<uint64_t>(b.piece_chunk0, (uint64_t)b.n_pieces + 1, chunk) - 1);
            if (do_hist)
                for (int i = threadIdx.x; i < sa.hist_np * 256; i += SW_THREADS) s_hist[i] = 0;
            __syncthreads();
            const uint32_t piece = s_piece;
            const uint64_t ci = chunk - b.piece_chunk0[piece];
            const uint32_t pk = b.piece_kmers[piece];
            const uint32_t ks = (uint32_t)(ci * CHUNK_KMERS);
            m = (int)min((uint32_t)CHUNK_KMERS, pk - ks);                     // k-mers in this chunk
            const uint8_t* cbase = b.seq + b.piece_byte[piece] + ks;          // first byte of the chunk
            const int nbytes = m + P.k - 1;                                   // bytes the chunk may read
            ow = MODE == 0 ? out_words + b.piece_out[piece] + ks : nullptr;
            of = MODE == 1 ? out_flags + b.piece_out[piece] + ks : nullptr;
            opos = (MODE == 2 && sa.pos) ? sa.pos + b.piece_out[piece] + ks : nullptr;

            // pack: lane l holds bases [1024*warp + 32*l, +32); lanes 0,1 also hold the two halo words
            bool bad = false;
            const int off = 1024 * warp + 32 * lane;
            Wd = load_pack32(cbase + off, b.seq_end, min(32, nbytes - off), bad);
            if (lane < 2) {
                const int hoff = 1024 * warp + 1024 + 32 * lane;
                H = load_pack32(cbase + hoff, b.seq_end, min(32, nbytes - hoff), bad);
            }
            if (bad) atomicMin(err_pos, (unsigned long long)(b.piece_byte[piece] + ks + off));
        }

        // Wd1 / Wd2: the packed words one and two lanes further on (wrapping into the halo words held
        // by lanes 0 and 1), so the window loops need plain broadcasts only
        const uint64_t Wd1 = __shfl_sync(0xffffffffu, lane == 0 ? H : Wd, (lane + 1) & 31);
        const uint64_t Wd2 = __shfl_sync(0xffffffffu, lane < 2 ? H : Wd, (lane + 2) & 31);
        const int kbase = 1024 * warp;
        uint32_t fwd_before = 0, nfwd_total = 0;
        if (!WORDS_IN && P.canonical) {
            // pass 1: parity of every window -> ballots, so output slots are known up front
            for (int j = 0; j < 32; j++) {
                uint64_t A = __shfl_sync(0xffffffffu, Wd, j);
                uint64_t B = __shfl_sync(0xffffffffu, Wd1, j);
                uint64_t C = __shfl_sync(0xffffffffu, Wd2, j);
                W x = cut_window<W>(A, B, C, 2 * lane, P.bits);
                bool active = kbase + 32 * j + lane < m;
                uint32_t bal = __ballot_sync(0xffffffffu, active && !(popc_w(x) & 1));
                if (lane == 0) s_fwd[warp][j] = bal;
            }
            __syncthreads();
            uint32_t c0 = __popc(s_fwd[0][lane]), c1 = __popc(s_fwd[1][lane]);
            c0 = warp_sum(c0);
            c1 = warp_sum(c1);
            nfwd_total = c0 + c1;
            fwd_before = warp == 0 ? 0 : c0;
        }
        for (int j0 = 0; j0 < 32; j0 += U) {
            if (kbase + 32 * j0 >= m) break;  // warp-uniform
            W word[U];
            uint32_t slot[U];
            bool active[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = j0 + u;
                uint64_t A = __shfl_sync(0xffffffffu, Wd, j);
                uint64_t B = __shfl_sync(0xffffffffu, Wd1, j);
                uint64_t C = __shfl_sync(0xffffffffu, Wd2, j);
                const int kidx = kbase + 32 * j + lane;
                active[u] = kidx < m;
                W x = cut_window<W>(A, B, C, 2 * lane, P.bits);
                slot[u] = (uint32_t)kidx;
                if (WORDS_IN) {
                    word[u] = active[u] ? in_words[kidx] : (W)0;
                    continue;
                }
                if (P.canonical) {
                    const uint32_t bal = s_fwd[warp][j];
                    const uint32_t fb = fwd_before + __popc(bal & ((1u << lane) - 1u));
                    const bool is_fwd = (bal >> lane) & 1;
                    slot[u] = is_fwd ? fb : nfwd_total + ((uint32_t)kidx - fb);
                    fwd_before += __popc(bal);
                }
                word[u] = kmer_to_word<W>(x, P, BRUTE);
            }
            if (MODE == 0) {
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (active[u]) ow[slot[u]] = word[u];
                    if (do_hist && active[u]) {
                        const W v = (W)(word[u] >> (8 * sa.hist_first));
                        const int np = sa.hist_np;
                        for (int p = 0; p < np - 2; p++) atomicAdd(&s_hist[p * 256 + ((uint32_t)(v >> (8 * p)) & 255u)], 1u);
                        // the two most significant digits are heavily skewed for necklace words (they start with a run of
                        // zeros): every lane counts runs of equal digits in registers and touches shared memory only when the
                        // digit changes (match.any aggregation cost as much ALU time as the separate histogram pass it replaced)
                        if (np >= 2) {
                            const uint32_t d = (uint32_t)(v >> (8 * (np - 2))) & 255u;
                            if (d != hc_d0) { if (hc_n0) atomicAdd(&s_hist[(np - 2) * 256 + hc_d0], hc_n0); hc_d0 = d; hc_n0 = 0; }
                            hc_n0++;
                        }
                        if (np >= 1) {
                            const uint32_t d = (uint32_t)(v >> (8 * (np - 1))) & 255u;
                            if (d != hc_d1) { if (hc_n1) atomicAdd(&s_hist[(np - 1) * 256 + hc_d1], hc_n1); hc_d1 = d; hc_n1 = 0; }
                            hc_n1++;
                        }
                    }
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (active[u]) {
                        // owner = number of splitters <= prefix: 4 bisection steps over the 15 (padded) splitters in shared
                        // memory instead of 15 compares against kernel parameters
                        const uint32_t pfx = (uint32_t)(word[u] >> sa.dest.suffix_bits);
                        uint32_t d = s_split[7] <= pfx ? 8u : 0u;
                        d += s_split[d + 3] <= pfx ? 4u : 0u;
                        d += s_split[d + 1] <= pfx ? 2u : 0u;
                        d += s_split[d] <= pfx ? 1u : 0u;
                        const uint32_t r = atomicAdd(&s_cnt[d], 1u);   // rank among the chunk's words for owner d
                        s_stage[slot[u]] = word[u];
                        s_dr[slot[u]] = (uint16_t)((d << 11) | r);
                    }
            } else {
                // staged membership probe (index_view.cuh): every stage issues U independent loads
                Suf s[U];
                uint32_t k32[U], lo[U], hi[U], g[U];
                bool present[U];
                uint2 de[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    uint32_t prefix;
                    split_key<W, Suf>(word[u], P, prefix, s[u]);
                    k32[u] = key32<Suf>(s[u], P.suffix_bits);
                    present[u] = active[u] && ix.nb != 0 && (!WORDS_IN || (prefix >> P.prefix_bits) == 0);   // words handed in from outside may be anything
                    lo[u] = prefix;
                    de[u] = ldg_keep(ix.dir + (present[u] ? (prefix >> 5) : 0u));
                }
                uint2 range[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t bit = lo[u] & 31;
                    const uint32_t rank = de[u].y + __popc(de[u].x & ((1u << bit) - 1u));
                    present[u] = present[u] && ((de[u].x >> bit) & 1u);
                    range[u] = ldg_keep(ix.bucket_range + (present[u] ? rank : 0u));
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    lo[u] = present[u] ? range[u].x : 0u;
                    hi[u] = present[u] ? range[u].y : 0u;
                    g[u] = lo[u] + predict_slot(ix.sub, lo[u], hi[u], k32[u]);
                }
                Suf e[U][WN];
#pragma unroll
                for (int u = 0; u < U; u++) {
#if defined(CBL_ABLATE_WINDOWS)   // developer ablation (wrong answers): no suffix-window traffic, tables only
                    for (int i = 0; i < WN; i++) e[u][i] = s[u];
#else
                    load_window<Suf, WB>(ix.suf + (g[u] & ~(uint32_t)(WN - 1)), e[u]);
#endif
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    int r = 0;
                    uint32_t L = lo[u], R = hi[u], gg = g[u];
                    if (present[u]) r = eval_window<Suf, WB>(e[u], gg & ~(uint32_t)(WN - 1), s[u], k32[u], P.suffix_bits, R - L, L, R, gg);
                    if (active[u] && r >= 0) s_flags[slot[u]] = (uint8_t)r;
                    q_push(r < 0, s[u], L, R, gg, hi[u] - lo[u], slot[u]);
                    while (q_count >= 32) q_drain();  // keeps the queue below 32 + 32 entries
                }
            }
        }
        if (PROBE) {
            while (q_count > 0) q_drain();
            __syncthreads();
            if (((uintptr_t)of & 15) == 0) {   // 16 answers per store (always the case for MODE 3: NVLink-friendly)
                const int m16 = m >> 4;
                for (int i = threadIdx.x; i < m16; i += SW_THREADS) reinterpret_cast<uint4*>(of)[i] = reinterpret_cast<const uint4*>(s_flags)[i];
                for (int i = (m16 << 4) + threadIdx.x; i < m; i += SW_THREADS) of[i] = s_flags[i];
            } else {
                for (int i = threadIdx.x; i < m; i += SW_THREADS) of[i] = s_flags[i];
            }
        }
        if (MODE == 2) {
            __syncthreads();  // s_stage / s_dr / s_cnt of the chunk are complete
            if (threadIdx.x < 16) {
                const uint32_t c = s_cnt[threadIdx.x];
                uint32_t inc = c;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                s_off[threadIdx.x] = inc - c;
                if (threadIdx.x == 15) s_off[16] = inc;
                unsigned long long base = 0;
                if (c) base = atomicAdd(sa.cnt + threadIdx.x, (unsigned long long)c);   // reserve c slots in my region at owner d
                s_base[threadIdx.x] = (base + c <= sa.cap) ? (uint32_t)base : 0xFFFFFFFFu;
                s_cnt[threadIdx.x] = 0;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < m; i += SW_THREADS) {
                const uint32_t dr = s_dr[i], d = dr >> 11, r = dr & 2047u;
                s_inv[s_off[d] + r] = (uint16_t)i;
                if (opos) opos[i] = (uint32_t)(d * sa.cap) + s_base[d] + r;
            }
            __syncthreads();
            // owner order: consecutive q of one owner -> consecutive addresses in its region (coalesced NVLink stores)
            for (int q = threadIdx.x; q < m; q += SW_THREADS) {
                uint32_t d = s_off[8] <= (uint32_t)q ? 8u : 0u;   // owner of output slot q: largest d with s_off[d] <= q (bisection)
                d += s_off[d + 4] <= (uint32_t)q ? 4u : 0u;
                d += s_off[d + 2] <= (uint32_t)q ? 2u : 0u;
                d += s_off[d + 1] <= (uint32_t)q ? 1u : 0u;
                const uint32_t bd = s_base[d];
                if (bd != 0xFFFFFFFFu) s_peer[d][(size_t)bd + ((uint32_t)q - s_off[d])] = s_stage[s_inv[q]];
            }
        }
        if (do_hist) {   // the chunk's counters: lane run counters -> shared, then one global RED per digit value that occurred
            if (hc_n0) atomicAdd(&s_hist[(sa.hist_np - 2) * 256 + hc_d0], hc_n0);
            if (hc_n1) atomicAdd(&s_hist[(sa.hist_np - 1) * 256 + hc_d1], hc_n1);
            hc_n0 = hc_n1 = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < sa.hist_np * 256; i += SW_THREADS)
                if (s_hist[i]) atomicAdd(&sa.hist[i], (unsigned long long)s_hist[i]);
        }
        __syncthreads();  // s_piece / s_fwd / s_flags / s_hist reuse in the next grid-stride iteration
    }
}

// k-mer integers (lo/hi arrays) -> words; single-k-mer API (src/cbl.rs:199-235)
template <class W>
__global__ void kmers_to_words_kernel(const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi, uint64_t n, KParams P,
                                      W* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    W x;
    if (sizeof(W) == 8) x = (W)lo[i];
    else x = (W)(((u128)(hi ? hi[i] : 0) << 64) | lo[i]);
    x &= low_mask<W>(P.bits);
    out[i] = kmer_to_word<W>(x, P);
}

}  // namespace cbl
