// One device-resident CBL shard: the class template behind IIndex, instantiated once per (word, suffix) type in
// its own translation unit (inst_*.cu) so the four instantiations compile in parallel.
// Host-side orchestration: batches on a CUDA stream, sort -> merge -> directory rebuild, set operations, export.
// B200 counterpart of src/cbl.rs + src/wordset/mod.rs (see DESIGN.md for the kernel map).
#pragma once
#include "cbl_index.cuh"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "index_ops.cuh"
#include "merge_ops.cuh"
#include "radix_sort.cuh"
#include "seg_sort.cuh"
#ifndef CBL_PROBE_WB
#define CBL_PROBE_WB 32   // bytes per suffix window of the membership probe
#endif
#include "sanitize.cuh"
#include "seq_words.cuh"

namespace cbl {

inline int pos_bits_for(int kmer_bits) {  // src/cbl.rs:66
    int p = 0;
    while ((1 << p) < kmer_bits) p++;
    return p;
}

// CBL_TRACE=1: host-side timeline of a mutation on stderr (developer aid)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    cudaStream_t s;
    explicit Trace(cudaStream_t st) : on(getenv("CBL_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), s(st) {}
    void mark(const char* what, bool sync = false) {
        if (!on) return;
        if (sync) cudaStreamSynchronize(s);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[cbl trace] %-28s %9.3f ms%s\n", what, std::chrono::duration<double, std::milli>(t - t0).count(), sync ? " (synced)" : "");
        t0 = std::chrono::steady_clock::now();
    }
};

inline uint64_t env_u64(const char* name, uint64_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return strtoull(v, nullptr, 10);
}

// One zeroed look-back workspace per kernel launch (status words + ticket counter).
struct Lookback {
    DevBuf<uint64_t> status;
    DevBuf<uint32_t> counter;
    Lookback(uint64_t tiles, cudaStream_t s) : status(tiles ? tiles : 1, s), counter(1, s) {
        status.zero();
        counter.zero();
    }
};

template <class W, class Suf>
class Index final : public IIndex {
    Config cfg_;
    KParams P_;
    cudaStream_t st_ = nullptr;
    cudaStream_t side_[2] = {nullptr, nullptr};
    DevBuf<uint2> dir_, bucket_range_;
    DevBuf<uint32_t> bucket_prefix_, bucket_off_;
    DevBuf<Suf> suf_;
    bool use_merge_ = true;   // CBL_MUTATE=edits selects the first-generation probe/edit-list path (kept for A/B runs)
    DevBuf<int8_t> sub_;      // interpolation corrections for the membership probe, rebuilt lazily
    bool sub_valid_ = false;
    uint32_t nb_ = 0;
    uint64_t n_ = 0;
    uint32_t last_prefix_ = 0;  // prefix of the last bucket (for the reference's is_empty quirk)
    uint64_t n_dir_ = 0;  // directory words: 32 prefixes each
    uint64_t batch_kmers_;
    static constexpr uint64_t SUF_PAD = 16;  // suffix arrays are over-allocated: probe windows are 32-byte aligned loads

public:
    explicit Index(const Config& cfg) : cfg_(cfg) {
        P_.k = cfg.k;
        P_.bits = 2 * cfg.k;
        P_.pos_bits = pos_bits_for(2 * cfg.k);
        P_.prefix_bits = cfg.prefix_bits;
        P_.suffix_bits = P_.bits + P_.pos_bits - cfg.prefix_bits;
        P_.canonical = cfg.canonical;
        CUDA_CHECK(cudaSetDevice(cfg.device));
        if (env_u64("CBL_STREAM_HIGH_PRIORITY", 0)) {   // the sharded pipeline's router handle: its CTAs are placed first
            int lo_p = 0, hi_p = 0;
            CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
            CUDA_CHECK(cudaStreamCreateWithPriority(&st_, cudaStreamNonBlocking, hi_p));
        } else {
            CUDA_CHECK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
        }
        for (auto& s : side_) CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, cfg.device));
        uint64_t thr = UINT64_MAX;
        CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        uint64_t bits = 1ull << cfg.prefix_bits;
        if (bits < 32) bits = 32;
        n_dir_ = bits / 32;
        dir_.alloc(n_dir_, st_);
        dir_.zero();
        bucket_range_.alloc(1, st_);
        bucket_prefix_.alloc(1, st_);
        bucket_off_.alloc(1, st_);
        bucket_off_.zero();
        suf_.alloc(SUF_PAD, st_);
        suf_.zero();
        sub_.alloc(4, st_);
        sub_.zero();
        { const char* m = getenv("CBL_MUTATE"); use_merge_ = !(m && std::string(m) == "edits"); }
        batch_kmers_ = env_u64("CBL_BATCH_KMERS", sizeof(W) == 8 ? (1ull << 29) : (1ull << 28));   // sort buffers: 2 x 4.3 GB of the 180 GB
        if (batch_kmers_ > RS_MAX_KEYS) batch_kmers_ = RS_MAX_KEYS;
        if (batch_kmers_ < CHUNK_KMERS) batch_kmers_ = CHUNK_KMERS;
        if (uint64_t g = env_u64("CBL_L2_FETCH", 0)) CUDA_CHECK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g));
        // L2::evict_last lines live in the persisting set-aside; without one the hint is a no-op
        if (uint64_t mb = env_u64("CBL_L2_PERSIST_MB", 0)) {
            int max_persist = 0;
            CUDA_CHECK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, cfg.device));
            CUDA_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)mb << 20, (size_t)max_persist)));
        }
        static bool attr_done = false;
        if (!attr_done) {
            CUDA_CHECK(cudaFuncSetAttribute((radix_pass_kernel<W, false, ByteDigit<W>>), cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            CUDA_CHECK(cudaFuncSetAttribute((radix_pass_kernel<W, false, DestDigit<W>>), cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            CUDA_CHECK(cudaFuncSetAttribute((radix_pass_kernel<W, false, ByteDigit<W>>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CUDA_CHECK(cudaFuncSetAttribute((seg_sort_kernel<W, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SsTile<W>::SMEM));
            CUDA_CHECK(cudaFuncSetAttribute((seg_sort_kernel<W, true>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CUDA_CHECK(cudaFuncSetAttribute((seg_sort_kernel<W, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SsTile<W>::SMEM));
            CUDA_CHECK(cudaFuncSetAttribute((seg_sort_kernel<W, false>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CUDA_CHECK(cudaFuncSetAttribute((merge_apply_kernel<W, Suf, MERGE_OR>), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CUDA_CHECK(cudaFuncSetAttribute((merge_apply_kernel<W, Suf, MERGE_AND>), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CUDA_CHECK(cudaFuncSetAttribute((merge_apply_kernel<W, Suf, MERGE_SUB>), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CUDA_CHECK(cudaFuncSetAttribute((merge_apply_kernel<W, Suf, MERGE_XOR>), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CUDA_CHECK(cudaFuncSetAttribute((radix_pass_kernel<W, false, DestDigit<W>>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CUDA_CHECK(cudaFuncSetAttribute((radix_pass_kernel<W, false, DestDigit<W>, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            CUDA_CHECK(cudaFuncSetAttribute((radix_pass_kernel<W, false, DestDigit<W>, true>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            attr_done = true;
        }
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }
    ~Index() override {
        cudaSetDevice(cfg_.device);
        dir_.release(); bucket_range_.release(); bucket_prefix_.release(); bucket_off_.release(); suf_.release(); sub_.release();
        if (st_) { cudaStreamSynchronize(st_); arena::retire_stream(st_); cudaStreamDestroy(st_); }
        for (auto& s : side_) if (s) { cudaStreamSynchronize(s); arena::retire_stream(s); cudaStreamDestroy(s); }
    }

    const Config& config() const override { return cfg_; }
    const KParams& params() const override { return P_; }
    cudaStream_t stream() const override { return st_; }
    uint64_t count() const override { return n_; }
    uint32_t n_buckets() const override { return nb_; }
    // WordSet::is_empty is prefixes.count() == 0 and RankBV::count_ones() ignores the last bit
    // (src/wordset/mod.rs:57-60, cxx/rank_bv.h:34): a set holding only words with the
    // all-ones prefix reports empty.  Reproduced for drop-in behaviour.
    bool is_empty_reference_semantics() const override {
        if (nb_ == 0) return true;
        return nb_ == 1 && last_prefix_ == (uint32_t)((1ull << cfg_.prefix_bits) - 1);
    }
    void sync() override { CUDA_CHECK(cudaSetDevice(cfg_.device)); CUDA_CHECK(cudaStreamSynchronize(st_)); }

    IndexView<Suf> view() const {
        IndexView<Suf> v;
        v.dir = dir_.get(); v.bucket_prefix = bucket_prefix_.get(); v.bucket_off = bucket_off_.get();
        v.bucket_range = bucket_range_.get(); v.suf = suf_.get(); v.sub = sub_.get(); v.nb = nb_; v.n = n_;
        return v;
    }
    // (re)build the probe's correction bytes if the set changed since they were last computed
    void ensure_sub() {
        if (sub_valid_) return;
        const uint64_t n_slots = (n_ >> SUB_SHIFT) + 4;
        sub_.alloc(n_slots, st_);
        CBL_LAUNCH((build_sub_kernel<Suf>), (unsigned)div_up(n_slots, 256), 256, 0, st_, view(), P_.suffix_bits, sub_.get(), n_slots);
        sub_valid_ = true;
    }

    IIndex* clone() override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        std::unique_ptr<Index> c(new Index(cfg_));
        sync();
        c->copy_state_from(*this);
        c->sync();
        return c.release();
    }
    void copy_state_from(const Index& o) {
        nb_ = o.nb_; n_ = o.n_; last_prefix_ = o.last_prefix_;
        sub_valid_ = false;
        CUDA_CHECK(cudaMemcpyAsync(dir_.get(), o.dir_.get(), n_dir_ * sizeof(uint2), cudaMemcpyDeviceToDevice, st_));
        bucket_range_.alloc(nb_ ? nb_ : 1, st_);
        if (nb_) CUDA_CHECK(cudaMemcpyAsync(bucket_range_.get(), o.bucket_range_.get(), (size_t)nb_ * sizeof(uint2), cudaMemcpyDeviceToDevice, st_));
        bucket_prefix_.alloc(nb_ ? nb_ : 1, st_);
        bucket_off_.alloc((uint64_t)nb_ + 1, st_);
        suf_.alloc(n_ + SUF_PAD, st_);
        if (nb_) CUDA_CHECK(cudaMemcpyAsync(bucket_prefix_.get(), o.bucket_prefix_.get(), (size_t)nb_ * 4, cudaMemcpyDeviceToDevice, st_));
        CUDA_CHECK(cudaMemcpyAsync(bucket_off_.get(), o.bucket_off_.get(), ((size_t)nb_ + 1) * 4, cudaMemcpyDeviceToDevice, st_));
        if (n_) CUDA_CHECK(cudaMemcpyAsync(suf_.get(), o.suf_.get(), n_ * sizeof(Suf), cudaMemcpyDeviceToDevice, st_));
    }

    // ------------------------------------------------------------------------------------------
    // records -> pieces (2048-aligned slices so chunking is identical to src/cbl.rs:239-243)
    // ------------------------------------------------------------------------------------------
    static constexpr uint32_t PIECE_KMERS = CHUNK_KMERS * 8192;  // 16.7M k-mers per piece

    void check_records(const uint64_t* offsets, size_t n_seqs) const {
        for (size_t i = 0; i < n_seqs; i++) {
            if (offsets[i + 1] < offsets[i]) throw Error(CBL_EINVAL, "record offsets must be non-decreasing");
            uint64_t len = offsets[i + 1] - offsets[i];
            if (len < (uint64_t)cfg_.k)  // src/cbl.rs:294-299,329-334
                throw Error(CBL_EINVAL, "Sequence size (" + std::to_string(len) + ") is smaller than K (" + std::to_string(cfg_.k) + ")");
        }
    }
    // pieces of records [r0, r1) — out offsets are k-mer ranks counted from record r0
    void build_pieces(const uint64_t* offsets, size_t r0, size_t r1, PieceList& pl, uint32_t piece_kmers = PIECE_KMERS) const {
        pl = PieceList();
        pl.chunk0.push_back(0);
        for (size_t r = r0; r < r1; r++) {
            uint64_t nk = offsets[r + 1] - offsets[r] - (uint64_t)cfg_.k + 1;
            for (uint64_t s = 0; s < nk; s += piece_kmers) {
                uint32_t m = (uint32_t)std::min<uint64_t>(piece_kmers, nk - s);
                pl.byte_off.push_back(offsets[r] + s);
                pl.out_off.push_back(pl.n_kmers + s);
                pl.kmers.push_back(m);
                pl.n_chunks += div_up(m, CHUNK_KMERS);
                pl.chunk0.push_back(pl.n_chunks);
            }
            pl.n_kmers += nk;
        }
    }

    struct DevPieces {
        DevBuf<uint64_t> byte_off, out_off, chunk0;
        DevBuf<uint32_t> kmers;
        SeqBatch batch;
    };
    // upload pieces [p0, p1) with out offsets rebased by out_base
    void upload_pieces(const PieceList& pl, size_t p0, size_t p1, uint64_t out_base, const uint8_t* d_seq, uint64_t n_bytes,
                       DevPieces& dp, cudaStream_t s) const {
        size_t np = p1 - p0;
        std::vector<uint64_t> out(np), ch(np + 1);
        for (size_t i = 0; i < np; i++) { out[i] = pl.out_off[p0 + i] - out_base; ch[i] = pl.chunk0[p0 + i] - pl.chunk0[p0]; }
        ch[np] = pl.chunk0[p1] - pl.chunk0[p0];
        dp.byte_off.alloc(np, s); dp.out_off.alloc(np, s); dp.chunk0.alloc(np + 1, s); dp.kmers.alloc(np, s);
        CUDA_CHECK(cudaMemcpyAsync(dp.byte_off.get(), pl.byte_off.data() + p0, np * 8, cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(dp.out_off.get(), out.data(), np * 8, cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(dp.chunk0.get(), ch.data(), (np + 1) * 8, cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(dp.kmers.get(), pl.kmers.data() + p0, np * 4, cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaStreamSynchronize(s));  // `out`/`ch` are stack temporaries
        dp.batch.seq = d_seq; dp.batch.seq_end = d_seq + n_bytes;
        dp.batch.piece_byte = dp.byte_off.get(); dp.batch.piece_out = dp.out_off.get();
        dp.batch.piece_kmers = dp.kmers.get(); dp.batch.piece_chunk0 = dp.chunk0.get();
        dp.batch.n_pieces = (uint32_t)np; dp.batch.n_chunks = ch[np];
    }

    // launches the fused encode + necklace (+ probe) kernel (no synchronisation); *err must hold
    // ULLONG_MAX on entry and receives the smallest offending byte offset if a non-ACGT byte is seen
    void launch_seq_words(const SeqBatch& b, int mode, bool brute, W* d_words, uint8_t* d_flags, unsigned long long* err, cudaStream_t s) {
        if (b.n_chunks == 0) return;
        unsigned grid = (unsigned)std::min<uint64_t>(b.n_chunks, 1u << 30);
        IndexView<Suf> v = view();
        const ShardArgs<W> sa{};
        if (mode == 0) {
            if (brute) CBL_LAUNCH((seq_words_kernel<W, Suf, 0, true, 32, 1>), grid, SW_THREADS, 0, s, b, P_, d_words, d_flags, v, err, sa);
            else CBL_LAUNCH((seq_words_kernel<W, Suf, 0, false, 32, 1>), grid, SW_THREADS, 0, s, b, P_, d_words, d_flags, v, err, sa);
        } else {
            CBL_LAUNCH((seq_words_kernel<W, Suf, 1, false, CBL_PROBE_WB, 1>), grid, SW_THREADS, 0, s, b, P_, d_words, d_flags, v, err, sa);
        }
    }
    // membership of n words (MODE 3 of the fused kernel: same staged probe + deferred queue, no sequence front end);
    // d_flags may be peer memory
    void launch_probe_words(const W* d_words, uint64_t n, uint8_t* d_flags, cudaStream_t s) {
        if (n == 0) return;
        ShardArgs<W> sa{};
        sa.in_words = d_words;
        sa.n_in = n;
        // CBL_WORDS_GRID / CBL_ROUTE_GRID cap the grids (the kernels stride over the chunks) so that the owner-side probe and
        // the next sub-batch's route kernel can be co-resident on every SM (sharded pipeline, cbl_b200/sharded.py)
        const unsigned grid = (unsigned)std::min<uint64_t>(div_up(n, CHUNK_KMERS), env_u64("CBL_WORDS_GRID", 1u << 30));
        CBL_LAUNCH((seq_words_kernel<W, Suf, 3, false, CBL_PROBE_WB, 1>), grid, SW_THREADS, 0, s, SeqBatch{}, P_, (W*)nullptr, d_flags, view(),
                   (unsigned long long*)nullptr, sa);
    }
    static void throw_bad_byte(unsigned long long e) {
        throw Error(CBL_EINVAL, "non-ACGT byte in sequence near byte offset " + std::to_string(e) +
                                    " (not supported on this entry point: the fused multi-GPU route; see DESIGN.md)");
    }
    // Reads with non-nucleotide bytes (sanitize.cuh): every reference chunk of `in` becomes one clean
    // single-chunk piece of K + b bytes that yields 1 + b words.  Returns the number of words the batch produces.
    struct Sanitized {
        DevBuf<uint8_t> clean;
        DevPieces dp;
        uint64_t n_kmers = 0;
    };
    void sanitize_batch(const SeqBatch& in, Sanitized& out, cudaStream_t s) {
        const uint64_t nc = in.n_chunks;
        DevBuf<uint32_t> counts(nc, s);
        const unsigned grid = (unsigned)std::min<uint64_t>(nc, 1u << 30);
        CBL_LAUNCH(sanitize_chunks_kernel, grid, SAN_THREADS, 0, s, in, cfg_.k, counts.get(), (const uint64_t*)nullptr, (uint8_t*)nullptr);
        std::vector<uint32_t> b(nc);
        CUDA_CHECK(cudaMemcpyAsync(b.data(), counts.get(), nc * 4, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        PieceList pl;
        pl.chunk0.push_back(0);
        uint64_t bytes = 0;
        for (uint64_t c = 0; c < nc; c++) {
            pl.byte_off.push_back(bytes);
            pl.out_off.push_back(pl.n_kmers);
            pl.kmers.push_back(1 + b[c]);
            pl.n_kmers += 1 + b[c];
            pl.chunk0.push_back(c + 1);
            bytes += (uint64_t)cfg_.k + b[c];
        }
        pl.n_chunks = nc;
        out.clean.alloc(bytes + 64, s);
        upload_pieces(pl, 0, nc, 0, out.clean.get(), bytes, out.dp, s);
        CBL_LAUNCH(sanitize_chunks_kernel, grid, SAN_THREADS, 0, s, in, cfg_.k, (uint32_t*)nullptr, (const uint64_t*)out.dp.byte_off.get(), out.clean.get());
        out.n_kmers = pl.n_kmers;
    }
    // Synchronous.  Returns the number of words / answers written (at the front of d_words / d_flags): the batch's
    // k-mer count, or fewer when the reads held non-ACGT bytes and the reference's behaviour (F8) was reproduced.
    uint64_t run_seq_words(const SeqBatch& b, uint64_t n_kmers, int mode, bool brute, W* d_words, uint8_t* d_flags, cudaStream_t s) {
        if (b.n_chunks == 0) return 0;
        DevBuf<unsigned long long> err(1, s);
        CUDA_CHECK(cudaMemsetAsync(err.get(), 0xFF, 8, s));
        launch_seq_words(b, mode, brute, d_words, d_flags, err.get(), s);
        unsigned long long e = 0;
        CUDA_CHECK(cudaMemcpyAsync(&e, err.get(), 8, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        if (e == ULLONG_MAX) return n_kmers;
        Sanitized sn;
        sanitize_batch(b, sn, s);
        CUDA_CHECK(cudaMemsetAsync(err.get(), 0xFF, 8, s));
        launch_seq_words(sn.dp.batch, mode, brute, d_words, d_flags, err.get(), s);
        CUDA_CHECK(cudaMemcpyAsync(&e, err.get(), 8, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        if (e != ULLONG_MAX) throw Error(CBL_ECUDA, "internal: sanitised reads still hold a non-ACGT byte");
        return sn.n_kmers;
    }

    // ------------------------------------------------------------------------------------------
    // sort / unique
    // ------------------------------------------------------------------------------------------
    // sorts n keys; returns the buffer (a or b) that holds the result
    // Hybrid (default): LSD passes over the top digits only, then the in-shared-memory segment sort (seg_sort.cuh).
    // The number of passes is chosen so that the largest group of words sharing their sorted top bits fits a segment
    // tile: for k-mer data the most frequent b-bit head of a necklace word has mass ~ 2K / 2^b.  Any other
    // distribution is still sorted exactly: a segment that does not fit raises the fail flag and the batch is re-sorted
    // by the plain LSD passes.  CBL_SORT=lsd selects the plain passes outright.
    W* sort_keys(W* a, W* b, uint64_t n) {
        if (n <= 1) return a;
        if (n > RS_MAX_KEYS) throw Error(CBL_EINVAL, "internal: sort batch too large");
        const int key_bits = P_.bits + P_.pos_bits;
        const int n_digits = (key_bits + 7) / 8;
        int n_pass = n_digits;
        const char* sort_env = getenv("CBL_SORT");
        const bool hybrid = !(sort_env && std::string(sort_env) == "lsd");
        if (hybrid) {
            for (int c = 1; c < n_digits - 1; c++) {   // c LSD passes leave 8 * (n_digits - c) low bits to the segment sort
                const int b = key_bits - 8 * (n_digits - c);
                const double est = (double)n * (double)P_.bits / std::ldexp(1.0, b);
                if (est * 1.5 <= (double)SsTile<W>::T) { n_pass = c; break; }
            }
        }
        W* res = lsd_passes(a, b, n, n_digits - n_pass, n_pass);
        if (n_pass == n_digits) return res;
        W* other = res == a ? b : a;
        DevBuf<unsigned> fail(1, st_);
        fail.zero();
        const int shift = 8 * (n_digits - n_pass);
        // nominal tile: leave room for segments 3x longer than the longest one expected, at least 512 keys
        const double est_seg = (double)n * (double)P_.bits / std::ldexp(1.0, key_bits - shift);
        const int room = (int)std::min<double>(SsTile<W>::CAP / 2, std::max<double>(512.0, std::ceil(3.0 * est_seg / 256.0) * 256.0));
        const int tile = SsTile<W>::CAP - room;
        if (shift <= 32)
            CBL_LAUNCH((seg_sort_kernel<W, true>), (unsigned)div_up(n, (uint64_t)tile), SS_THREADS, SsTile<W>::SMEM, st_, res, other, n, shift, fail.get(), tile);
        else
            CBL_LAUNCH((seg_sort_kernel<W, false>), (unsigned)div_up(n, (uint64_t)tile), SS_THREADS, SsTile<W>::SMEM, st_, res, other, n, shift, fail.get(), tile);
        unsigned h_fail = 0;
        CUDA_CHECK(cudaMemcpyAsync(&h_fail, fail.get(), sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        if (!h_fail) return other;
        g_sort_fallbacks.fetch_add(1, std::memory_order_relaxed);
        return lsd_passes(res, other, n, 0, n_digits);   // `res` still holds every word (grouped by its top digits)
    }
    // LSD passes over digits [first, first + n_pass); returns the buffer that holds the result
    W* lsd_passes(W* a, W* b, uint64_t n, int first, int n_pass) {
        DevBuf<unsigned long long> hist((size_t)n_pass * 256, st_);
        hist.zero();
        unsigned hgrid = (unsigned)std::min<uint64_t>(div_up(n, RH_THREADS * RH_KEYS), 148 * 4);
        CBL_LAUNCH((radix_hist_kernel<W>), hgrid, RH_THREADS, (size_t)n_pass * 256 * sizeof(uint32_t), st_, a, n, n_pass, hist.get(), first);
        CBL_LAUNCH(radix_scan_hist_kernel, n_pass, 256, 0, st_, hist.get());
        const uint64_t tiles = div_up(n, RsTile<W>::TILE);
        DevBuf<uint32_t> status(tiles * 256, st_), counter(1, st_);
        const size_t smem = sizeof(W) * RsTile<W>::TILE;
        W *src = a, *dst = b;
        for (int p = 0; p < n_pass; p++) {
            status.zero();
            counter.zero();
            CBL_LAUNCH((radix_pass_kernel<W, false, ByteDigit<W>>), (unsigned)tiles, RS_THREADS, smem, st_, src, dst, nullptr, nullptr, n,
                       ByteDigit<W>{8 * (first + p)}, hist.get() + (size_t)p * 256, status.get(), counter.get(), (uint32_t*)nullptr);
            std::swap(src, dst);
        }
        return src;
    }
    uint64_t read_u64(const unsigned long long* d) {
        unsigned long long v = 0;
        CUDA_CHECK(cudaMemcpyAsync(&v, d, 8, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        return v;
    }
    uint64_t unique_keys(const W* sorted, uint64_t n, W* out) {
        if (n == 0) return 0;
        const uint64_t tiles = div_up(n, OP_TILE);
        Lookback lb(tiles, st_);
        DevBuf<unsigned long long> cnt(1, st_);
        cnt.zero();
        CBL_LAUNCH((unique_kernel<W>), (unsigned)tiles, OP_THREADS, 0, st_, sorted, n, out, lb.status.get(), lb.counter.get(), cnt.get());
        return read_u64(cnt.get());
    }

    // ------------------------------------------------------------------------------------------
    // mutation: sorted distinct probe keys -> edits -> new directory -> new suffix array
    // ------------------------------------------------------------------------------------------
    struct NewState {
        DevBuf<uint2> dir, bucket_range;
        DevBuf<uint32_t> bucket_prefix, bucket_off;
        DevBuf<Suf> suf;
        uint32_t nb = 0;
        uint64_t n = 0;
        uint32_t last_prefix = 0;
        bool changed = false;
    };
    void adopt(NewState& ns) {
        dir_.swap(ns.dir); bucket_range_.swap(ns.bucket_range); bucket_prefix_.swap(ns.bucket_prefix);
        bucket_off_.swap(ns.bucket_off); suf_.swap(ns.suf);
        nb_ = ns.nb; n_ = ns.n; last_prefix_ = ns.last_prefix;
        sub_valid_ = false;
        dir_.rebind(st_); bucket_range_.rebind(st_); bucket_prefix_.rebind(st_); bucket_off_.rebind(st_); suf_.rebind(st_);
    }

    // keys: sorted distinct words.  probe_ix: index they are looked up in (own view unless KEEP_ONLY).
    // scratch_ins: optional buffer (>= nk words) reused for the insert list.
    void compute_new_state(const W* keys, uint64_t nk, int mode, const IndexView<Suf>& probe_ix, W* scratch_ins, NewState& ns) {
        ns.changed = false;
        if (nk == 0) return;
        const IndexView<Suf> self = view();
        const uint64_t tiles = div_up(nk, OP_TILE);
        DevBuf<W> ins_own;
        W* ins_key = scratch_ins;
        const bool want_ins = (mode & EDIT_INS) != 0;
        const bool want_del = (mode & (EDIT_DEL | EDIT_KEEP_ONLY)) != 0;
        if (want_ins && !ins_key) { ins_own.alloc(nk, st_); ins_key = ins_own.get(); }
        DevBuf<uint64_t> ins_vpos(want_ins ? nk : 1, st_), del_idx(want_del ? nk : 1, st_);
        DevBuf<int> delta(nb_ ? nb_ : 1, st_);
        delta.zero();
        ns.dir.alloc(n_dir_, st_);
        CUDA_CHECK(cudaMemcpyAsync(ns.dir.get(), dir_.get(), n_dir_ * sizeof(uint2), cudaMemcpyDeviceToDevice, st_));
        Lookback lb_ins(tiles, st_), lb_del(tiles, st_);
        DevBuf<unsigned long long> counts(4, st_);
        counts.zero();
        CBL_LAUNCH((probe_edits_kernel<W, Suf>), (unsigned)tiles, OP_THREADS, 0, st_, keys, nk, probe_ix, self, P_, mode, ins_key,
                   ins_vpos.get(), del_idx.get(), delta.get(), ns.dir.get(), lb_ins.status.get(),
                   lb_del.status.get(), lb_ins.counter.get(), counts.get());
        unsigned long long h_counts[2] = {0, 0};
        CUDA_CHECK(cudaMemcpyAsync(h_counts, counts.get(), 16, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        const uint64_t ni = h_counts[0], nd = h_counts[1];
        if (ni == 0 && nd == 0) return;
        const uint64_t n_new = n_ + ni - nd;
        if (n_new >= (1ull << 32))
            throw Error(CBL_EINVAL, "shard would hold >= 2^32 k-mers; bucket offsets are 32-bit — shard the index over more GPUs");

        // directory
        if (nb_) CBL_LAUNCH(clear_emptied_kernel, (unsigned)div_up(nb_, 256), 256, 0, st_, bucket_prefix_.get(), bucket_off_.get(),
                            delta.get(), nb_, ns.dir.get());
        {
            const uint64_t t = div_up(n_dir_, OP_TILE);
            Lookback lb(t, st_);
            CBL_LAUNCH(rank_directory_kernel, (unsigned)t, OP_THREADS, 0, st_, ns.dir.get(), n_dir_, lb.status.get(), lb.counter.get(),
                       counts.get() + 2);
        }
        const uint64_t nb_new = read_u64(counts.get() + 2);
        DevBuf<uint32_t> size_new(nb_new + 1, st_);
        size_new.zero();
        ns.bucket_prefix.alloc(nb_new ? nb_new : 1, st_);
        if (nb_) CBL_LAUNCH(fill_sizes_old_kernel, (unsigned)div_up(nb_, 256), 256, 0, st_, bucket_prefix_.get(), bucket_off_.get(),
                            delta.get(), nb_, ns.dir.get(), size_new.get(), ns.bucket_prefix.get());
        if (ni) CBL_LAUNCH((fill_sizes_ins_kernel<W>), (unsigned)div_up(ni, 256), 256, 0, st_, ins_key, ni, P_, dir_.get(), ns.dir.get(),
                           size_new.get(), ns.bucket_prefix.get());
        ns.bucket_off.alloc(nb_new + 1, st_);
        ns.bucket_range.alloc(nb_new ? nb_new : 1, st_);
        {
            const uint64_t t = div_up(nb_new + 1, OP_TILE);
            Lookback lb(t, st_);
            CBL_LAUNCH(scan_sizes_kernel, (unsigned)t, OP_THREADS, 0, st_, size_new.get(), nb_new, ns.bucket_off.get(), ns.bucket_range.get(), lb.status.get(),
                       lb.counter.get());
        }
        // suffixes
        ns.suf.alloc(n_new + SUF_PAD, st_);
        {
            const uint64_t V = n_ + ni;
            const uint64_t t = div_up(V, OP_TILE);
            Lookback lb(t, st_);
            if (t) CBL_LAUNCH((apply_edits_kernel<W, Suf>), (unsigned)t, OP_THREADS, 0, st_, suf_.get(), n_, ins_vpos.get(), ins_key, ni,
                              del_idx.get(), nd, ns.suf.get(), P_, lb.status.get(), lb.counter.get());
        }
        uint32_t tail[2] = {0, 0};  // total from the offsets scan, prefix of the last bucket
        CUDA_CHECK(cudaMemcpyAsync(&tail[0], ns.bucket_off.get() + nb_new, 4, cudaMemcpyDeviceToHost, st_));
        if (nb_new) CUDA_CHECK(cudaMemcpyAsync(&tail[1], ns.bucket_prefix.get() + (nb_new - 1), 4, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        if ((uint64_t)tail[0] != n_new)
            throw Error(CBL_ECUDA, "internal: directory total " + std::to_string(tail[0]) + " != element count " + std::to_string(n_new));
        ns.nb = (uint32_t)nb_new;
        ns.n = n_new;
        ns.last_prefix = tail[1];
        ns.changed = true;
    }

    // ------------------------------------------------------------------------------------------
    // mutation by one streaming merge (merge_ops.cuh): keys = sorted distinct words, op = MERGE_*
    // ------------------------------------------------------------------------------------------
    static uint64_t round_capacity(uint64_t n) {
        if (n < (1u << 20)) return n;
        int top = 63 - __builtin_clzll(n);
        const uint64_t step = 1ull << (top - 2);
        return (n + step - 1) / step * step;
    }
    template <int OP> void launch_merge_apply(unsigned tiles, const IndexView<Suf>& v, const W* keys, uint64_t nk, const uint32_t* part_i,
                                              const uint32_t* part_r, Suf* suf_out, uint32_t* cnt, uint64_t* status, uint32_t* counter,
                                              unsigned long long* n_out) {
        const size_t smem = (size_t)MG_SMEM_ELEMS * sizeof(W);
        CBL_LAUNCH((merge_apply_kernel<W, Suf, OP>), tiles, MG_THREADS, smem, st_, v, P_, keys, nk, part_i, part_r, suf_out, cnt, status,
                   counter, n_out);
    }
    void merge_new_state(const W* keys, uint64_t nk, int op, NewState& ns) {
        ns.changed = false;
        const uint64_t V = n_ + nk;
        if (nk == 0 || (n_ == 0 && (op == MERGE_AND || op == MERGE_SUB))) {
            if (op == MERGE_AND && n_ != 0) {  // A & {} = {}
                ns.dir.alloc(n_dir_, st_); ns.dir.zero();
                ns.bucket_range.alloc(1, st_); ns.bucket_prefix.alloc(1, st_); ns.bucket_off.alloc(1, st_); ns.bucket_off.zero();
                ns.suf.alloc(SUF_PAD, st_);
                ns.nb = 0; ns.n = 0; ns.last_prefix = 0; ns.changed = true;
            }
            return;
        }
        const IndexView<Suf> v = view();
        const uint64_t tiles = div_up(V, MG_TILE);
        DevBuf<uint32_t> part_i(tiles + 1, st_), part_r(tiles + 1, st_);
        CBL_LAUNCH((merge_partition_kernel<W, Suf>), (unsigned)div_up(tiles + 1, 128), 128, 0, st_, v, P_, keys, nk, tiles, part_i.get(), part_r.get());
        const uint64_t n_prefix = n_dir_ * 32;
        DevBuf<uint32_t> cnt(n_prefix, st_);
        cnt.zero();
        // capacity is rounded up to 4 steps per octave so that the blocks freed by earlier, smaller states of a
        // growing index can be reused by the memory pool instead of a fresh driver allocation per batch
        const uint64_t cap = round_capacity((op == MERGE_AND || op == MERGE_SUB) ? n_ : V);
        ns.suf.alloc(cap + SUF_PAD, st_);
        DevBuf<unsigned long long> totals(4, st_);
        totals.zero();
        {
            Lookback lb(tiles, st_);
            switch (op) {
                case MERGE_OR: launch_merge_apply<MERGE_OR>((unsigned)tiles, v, keys, nk, part_i.get(), part_r.get(), ns.suf.get(), cnt.get(), lb.status.get(), lb.counter.get(), totals.get()); break;
                case MERGE_AND: launch_merge_apply<MERGE_AND>((unsigned)tiles, v, keys, nk, part_i.get(), part_r.get(), ns.suf.get(), cnt.get(), lb.status.get(), lb.counter.get(), totals.get()); break;
                case MERGE_SUB: launch_merge_apply<MERGE_SUB>((unsigned)tiles, v, keys, nk, part_i.get(), part_r.get(), ns.suf.get(), cnt.get(), lb.status.get(), lb.counter.get(), totals.get()); break;
                default: launch_merge_apply<MERGE_XOR>((unsigned)tiles, v, keys, nk, part_i.get(), part_r.get(), ns.suf.get(), cnt.get(), lb.status.get(), lb.counter.get(), totals.get()); break;
            }
        }
        ns.dir.alloc(n_dir_, st_);
        DevBuf<uint32_t> word_off(n_dir_, st_);
        {
            const uint64_t t = div_up(n_dir_, 256);
            Lookback lb_rank(t, st_), lb_off(t, st_);
            CBL_LAUNCH(dir_bits_kernel, (unsigned)t, 256, 0, st_, cnt.get(), n_dir_, ns.dir.get(), word_off.get(), lb_rank.status.get(),
                       lb_off.status.get(), lb_rank.counter.get(), totals.get() + 1);
        }
        unsigned long long h[3] = {0, 0, 0};  // merged element count, buckets, elements by the directory
        CUDA_CHECK(cudaMemcpyAsync(h, totals.get(), sizeof h, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        const uint64_t n_new = h[0], nb_new = h[1];
        if (h[2] != n_new)
            throw Error(CBL_ECUDA, "internal: directory total " + std::to_string(h[2]) + " != element count " + std::to_string(n_new));
        if (n_new >= (1ull << 32))
            throw Error(CBL_EINVAL, "shard would hold >= 2^32 k-mers; bucket offsets are 32-bit — shard the index over more GPUs");
        ns.bucket_prefix.alloc(nb_new ? nb_new : 1, st_);
        ns.bucket_off.alloc(nb_new + 1, st_);
        ns.bucket_range.alloc(nb_new ? nb_new : 1, st_);
        CBL_LAUNCH(dir_fill_kernel, (unsigned)div_up(n_dir_, 256), 256, 0, st_, cnt.get(), ns.dir.get(), word_off.get(), n_dir_,
                   ns.bucket_prefix.get(), ns.bucket_off.get(), ns.bucket_range.get(), (uint32_t)nb_new, (uint32_t)n_new);
        uint32_t last = 0;
        if (nb_new) CUDA_CHECK(cudaMemcpyAsync(&last, ns.bucket_prefix.get() + (nb_new - 1), 4, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        ns.nb = (uint32_t)nb_new;
        ns.n = n_new;
        ns.last_prefix = last;
        ns.changed = true;
    }

    // unsorted words in `a` (n of them, `b` same-size scratch) -> applied to this index
    void mutate_with_words(W* a, W* b, uint64_t n, int mode) {
        if (n == 0) return;
        Trace tr(st_);
        W* sorted = sort_keys(a, b, n);
        tr.mark("  sort enqueue");
        NewState ns;
        if (use_merge_) {
            // the merge treats a repeated batch word as one (merge_ops.cuh): no unique pass, no count read-back
            merge_new_state(sorted, n, mode == EDIT_INS ? MERGE_OR : MERGE_SUB, ns);
        } else {
            W* other = sorted == a ? b : a;
            uint64_t nu = unique_keys(sorted, n, other);
            tr.mark("  unique (sync inside)");
            compute_new_state(other, nu, mode, view(), sorted, ns);
        }
        tr.mark("  new state (syncs inside)");
        if (ns.changed) adopt(ns);
        tr.mark("  adopt");
    }

    // ------------------------------------------------------------------------------------------
    // sequence front ends
    // ------------------------------------------------------------------------------------------
    void mutate_seqs_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, int mode) {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        PieceList pl;
        build_pieces(offsets, 0, n_seqs, pl);
        last_produced = 0;
        size_t p = 0;
        const size_t np = pl.kmers.size();
        while (p < np) {
            size_t q = p;
            uint64_t nk = 0;
            while (q < np && (q == p || nk + pl.kmers[q] <= batch_kmers_)) { nk += pl.kmers[q]; q++; }
            Trace tr(st_);
            DevPieces dp;
            upload_pieces(pl, p, q, pl.out_off[p], d_seq, n_bytes, dp, st_);
            tr.mark("upload pieces");
            DevBuf<W> a(nk, st_), b(nk, st_);
            tr.mark("alloc a,b");
            const uint64_t produced = run_seq_words(dp.batch, nk, 0, false, a.get(), nullptr, st_);
            last_produced += produced;
            tr.mark("seq_words (sync inside)");
            mutate_with_words(a.get(), b.get(), produced, mode);
            tr.mark("mutate_with_words", true);
            p = q;
        }
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }
    void insert_seqs_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs) override {
        mutate_seqs_dev(d_seq, n_bytes, offsets, n_seqs, EDIT_INS);
    }
    void remove_seqs_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs) override {
        mutate_seqs_dev(d_seq, n_bytes, offsets, n_seqs, EDIT_DEL);
    }
    void contains_seqs_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, uint8_t* d_out) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        PieceList pl;
        build_pieces(offsets, 0, n_seqs, pl);
        last_produced = 0;
        if (pl.kmers.empty()) return;
        ensure_sub();
        DevPieces dp;
        upload_pieces(pl, 0, pl.kmers.size(), 0, d_seq, n_bytes, dp, st_);
        last_produced = run_seq_words(dp.batch, pl.n_kmers, 1, false, nullptr, d_out, st_);
    }
    void seq_words_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, void* d_words, bool brute) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        PieceList pl;
        build_pieces(offsets, 0, n_seqs, pl);
        last_produced = 0;
        if (pl.kmers.empty()) return;
        DevPieces dp;
        upload_pieces(pl, 0, pl.kmers.size(), 0, d_seq, n_bytes, dp, st_);
        last_produced = run_seq_words(dp.batch, pl.n_kmers, 0, brute, (W*)d_words, nullptr, st_);
    }

    // host buffers: records are grouped (~CBL_GROUP_BYTES each) and streamed through the device
    template <class F> void for_each_group(const uint64_t* offsets, size_t n_seqs, F&& f) const {
        const uint64_t group_bytes = env_u64("CBL_GROUP_BYTES", 256ull << 20);
        size_t r = 0;
        while (r < n_seqs) {
            size_t q = r;
            while (q < n_seqs && (q == r || offsets[q + 1] - offsets[r] <= group_bytes)) q++;
            f(r, q);
            r = q;
        }
    }
    void insert_seqs(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, bool remove) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        for_each_group(offsets, n_seqs, [&](size_t r, size_t q) {
            const uint64_t b0 = offsets[r], nbytes = offsets[q] - b0;
            DevBuf<uint8_t> d(nbytes + 64, st_);
            CUDA_CHECK(cudaMemcpyAsync(d.get(), seq + b0, nbytes, cudaMemcpyHostToDevice, st_));
            std::vector<uint64_t> off(q - r + 1);
            for (size_t i = r; i <= q; i++) off[i - r] = offsets[i] - b0;
            mutate_seqs_dev(d.get(), nbytes, off.data(), q - r, remove ? EDIT_DEL : EDIT_INS);
        });
    }
    // Host buffers -> answers on the host, software-pipelined over N_SLOTS streams: while one group's
    // kernel runs, the next group's reads are on their way in and the previous group's answers on
    // their way out (H2D and D2H use separate copy engines).  Nothing blocks the host until the end;
    // with pinned host memory the copies are true DMA, with pageable memory the driver stages them.
    static constexpr int N_SLOTS = 3;
    struct Slot {
        cudaStream_t s = nullptr;
        cudaEvent_t done = nullptr;
        DevBuf<uint8_t> seq, flags;
        DevBuf<uint64_t> byte_off, out_off, chunk0;
        DevBuf<uint32_t> kmers;
        DevBuf<unsigned long long> err;
        uint64_t *h_byte = nullptr, *h_out = nullptr, *h_chunk0 = nullptr;  // pinned staging for the piece arrays
        uint32_t* h_kmers = nullptr;
        size_t cap_pieces = 0;
        bool used = false;
        void free_host() {
            if (h_byte) cudaFreeHost(h_byte);
            if (h_out) cudaFreeHost(h_out);
            if (h_chunk0) cudaFreeHost(h_chunk0);
            if (h_kmers) cudaFreeHost(h_kmers);
            h_byte = h_out = h_chunk0 = nullptr;
            h_kmers = nullptr;
        }
    };
    void contains_seqs(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, uint8_t* out) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        ensure_sub();
        CUDA_CHECK(cudaStreamSynchronize(st_));  // index state is final before the pipeline streams read it
        // cut all records into pieces of <= piece_kmers k-mers, then group consecutive pieces
        const uint64_t group_kmers = std::max<uint64_t>(CHUNK_KMERS, env_u64("CBL_GROUP_BYTES", 32ull << 20));
        const uint32_t piece_kmers = (uint32_t)std::min<uint64_t>(PIECE_KMERS, std::max<uint64_t>(CHUNK_KMERS, (group_kmers / 4) / CHUNK_KMERS * CHUNK_KMERS));
        PieceList pl;
        build_pieces(offsets, 0, n_seqs, pl, piece_kmers);
        const size_t np = pl.kmers.size();
        if (np == 0) return;
        // groups [g0, g1) of pieces
        std::vector<std::pair<size_t, size_t>> groups;
        size_t max_pieces = 0;
        uint64_t max_bytes = 0, max_kmers = 0;
        for (size_t p = 0; p < np;) {
            size_t q = p;
            uint64_t nk = 0;
            while (q < np && (q == p || nk + pl.kmers[q] <= group_kmers)) { nk += pl.kmers[q]; q++; }
            groups.push_back({p, q});
            max_pieces = std::max(max_pieces, q - p);
            max_kmers = std::max(max_kmers, nk);
            max_bytes = std::max<uint64_t>(max_bytes, pl.byte_off[q - 1] + pl.kmers[q - 1] + cfg_.k - 1 - pl.byte_off[p]);
            p = q;
        }
        Slot slots[N_SLOTS];
        struct Cleanup {
            Slot* sl;
            ~Cleanup() {
                for (int i = 0; i < N_SLOTS; i++) {
                    if (sl[i].s) cudaStreamSynchronize(sl[i].s);
                    sl[i].seq.release(); sl[i].flags.release(); sl[i].byte_off.release(); sl[i].out_off.release();
                    sl[i].chunk0.release(); sl[i].kmers.release(); sl[i].err.release();
                    if (sl[i].s) cudaStreamSynchronize(sl[i].s);
                    sl[i].free_host();
                    if (sl[i].done) cudaEventDestroy(sl[i].done);
                    if (sl[i].s) { arena::retire_stream(sl[i].s); cudaStreamDestroy(sl[i].s); }
                }
            }
        } cleanup{slots};
        const int n_slots = (int)std::min<size_t>(N_SLOTS, groups.size());
        for (int i = 0; i < n_slots; i++) {
            Slot& sl = slots[i];
            CUDA_CHECK(cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking));
            CUDA_CHECK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
            sl.seq.alloc(max_bytes + 64, sl.s);
            sl.flags.alloc(max_kmers, sl.s);
            sl.byte_off.alloc(max_pieces, sl.s); sl.out_off.alloc(max_pieces, sl.s); sl.chunk0.alloc(max_pieces + 1, sl.s);
            sl.kmers.alloc(max_pieces, sl.s);
            sl.err.alloc(1, sl.s);
            CUDA_CHECK(cudaMemsetAsync(sl.err.get(), 0xFF, 8, sl.s));
            CUDA_CHECK(cudaMallocHost((void**)&sl.h_byte, max_pieces * 8));
            CUDA_CHECK(cudaMallocHost((void**)&sl.h_out, max_pieces * 8));
            CUDA_CHECK(cudaMallocHost((void**)&sl.h_chunk0, (max_pieces + 1) * 8));
            CUDA_CHECK(cudaMallocHost((void**)&sl.h_kmers, max_pieces * 4));
        }
        for (size_t g = 0; g < groups.size(); g++) {
            Slot& sl = slots[g % n_slots];
            if (sl.used) CUDA_CHECK(cudaEventSynchronize(sl.done));  // staging buffers of this slot are free again
            const size_t p0 = groups[g].first, p1 = groups[g].second, m = p1 - p0;
            const uint64_t b0 = pl.byte_off[p0];
            const uint64_t nbytes = pl.byte_off[p1 - 1] + pl.kmers[p1 - 1] + cfg_.k - 1 - b0;
            const uint64_t k0 = pl.out_off[p0];
            uint64_t nk = 0;
            for (size_t i = 0; i < m; i++) {
                sl.h_byte[i] = pl.byte_off[p0 + i] - b0;
                sl.h_out[i] = pl.out_off[p0 + i] - k0;
                sl.h_kmers[i] = pl.kmers[p0 + i];
                sl.h_chunk0[i] = pl.chunk0[p0 + i] - pl.chunk0[p0];
                nk += pl.kmers[p0 + i];
            }
            sl.h_chunk0[m] = pl.chunk0[p1] - pl.chunk0[p0];
            CUDA_CHECK(cudaMemcpyAsync(sl.seq.get(), seq + b0, nbytes, cudaMemcpyHostToDevice, sl.s));
            CUDA_CHECK(cudaMemcpyAsync(sl.byte_off.get(), sl.h_byte, m * 8, cudaMemcpyHostToDevice, sl.s));
            CUDA_CHECK(cudaMemcpyAsync(sl.out_off.get(), sl.h_out, m * 8, cudaMemcpyHostToDevice, sl.s));
            CUDA_CHECK(cudaMemcpyAsync(sl.chunk0.get(), sl.h_chunk0, (m + 1) * 8, cudaMemcpyHostToDevice, sl.s));
            CUDA_CHECK(cudaMemcpyAsync(sl.kmers.get(), sl.h_kmers, m * 4, cudaMemcpyHostToDevice, sl.s));
            SeqBatch b;
            b.seq = sl.seq.get(); b.seq_end = sl.seq.get() + nbytes;
            b.piece_byte = sl.byte_off.get(); b.piece_out = sl.out_off.get(); b.piece_kmers = sl.kmers.get(); b.piece_chunk0 = sl.chunk0.get();
            b.n_pieces = (uint32_t)m; b.n_chunks = sl.h_chunk0[m];
            launch_seq_words(b, 1, false, nullptr, sl.flags.get(), sl.err.get(), sl.s);
            CUDA_CHECK(cudaMemcpyAsync(out + k0, sl.flags.get(), nk, cudaMemcpyDeviceToHost, sl.s));
            CUDA_CHECK(cudaEventRecord(sl.done, sl.s));
            sl.used = true;
        }
        unsigned long long errs[N_SLOTS];
        for (int i = 0; i < n_slots; i++) {
            errs[i] = ULLONG_MAX;
            CUDA_CHECK(cudaMemcpyAsync(&errs[i], slots[i].err.get(), 8, cudaMemcpyDeviceToHost, slots[i].s));
        }
        for (int i = 0; i < n_slots; i++) CUDA_CHECK(cudaStreamSynchronize(slots[i].s));
        last_produced = pl.n_kmers;
        bool bad = false;
        for (int i = 0; i < n_slots; i++) bad = bad || errs[i] != ULLONG_MAX;
        if (bad) contains_seqs_unpipelined(seq, offsets, n_seqs, out);   // non-ACGT bytes: the reference's behaviour (F8), slow path
    }
    // group after group through the device path; answers are compacted in reference order (record after record,
    // chunk after chunk), last_produced = how many there are
    void contains_seqs_unpipelined(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, uint8_t* out) {
        uint64_t total = 0;
        for_each_group(offsets, n_seqs, [&](size_t r, size_t q) {
            const uint64_t b0 = offsets[r], nbytes = offsets[q] - b0;
            DevBuf<uint8_t> d(nbytes + 64, st_);
            CUDA_CHECK(cudaMemcpyAsync(d.get(), seq + b0, nbytes, cudaMemcpyHostToDevice, st_));
            std::vector<uint64_t> off(q - r + 1);
            uint64_t nk = 0;
            for (size_t i = r; i <= q; i++) off[i - r] = offsets[i] - b0;
            for (size_t i = r; i < q; i++) nk += offsets[i + 1] - offsets[i] - (uint64_t)cfg_.k + 1;
            DevBuf<uint8_t> flags(nk, st_);
            contains_seqs_dev(d.get(), nbytes, off.data(), q - r, flags.get());
            if (last_produced) CUDA_CHECK(cudaMemcpyAsync(out + total, flags.get(), last_produced, cudaMemcpyDeviceToHost, st_));
            CUDA_CHECK(cudaStreamSynchronize(st_));
            total += last_produced;
        });
        last_produced = total;
    }
    void seq_words(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, uint64_t* lo, uint64_t* hi, bool brute) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        const uint64_t nbytes = offsets[n_seqs] - offsets[0];
        std::vector<uint64_t> off(n_seqs + 1);
        for (size_t i = 0; i <= n_seqs; i++) off[i] = offsets[i] - offsets[0];
        uint64_t nk = 0;
        for (size_t i = 0; i < n_seqs; i++) nk += off[i + 1] - off[i] - cfg_.k + 1;
        DevBuf<uint8_t> d(nbytes + 64, st_);
        CUDA_CHECK(cudaMemcpyAsync(d.get(), seq + offsets[0], nbytes, cudaMemcpyHostToDevice, st_));
        DevBuf<W> w(nk, st_);
        seq_words_dev(d.get(), nbytes, off.data(), n_seqs, w.get(), brute);
        download_words(w.get(), last_produced, lo, hi);
    }
    void download_words(const W* d, uint64_t n, uint64_t* lo, uint64_t* hi) {
        if (n == 0) return;
        std::vector<W> h(n);
        CUDA_CHECK(cudaMemcpyAsync(h.data(), d, n * sizeof(W), cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        for (uint64_t i = 0; i < n; i++) {
            lo[i] = (uint64_t)h[i];
            if (hi) hi[i] = sizeof(W) == 16 ? (uint64_t)((u128)h[i] >> 64) : 0;
        }
    }

    // ------------------------------------------------------------------------------------------
    // k-mer / word level operations
    // ------------------------------------------------------------------------------------------
    void words_op_dev(int op, const void* d_words, uint64_t n, uint8_t* d_out) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        if (n == 0) return;
        if (d_out) ensure_sub();
        if (d_out) launch_probe_words((const W*)d_words, n, d_out, st_);
        if (op == 0) return;
        uint64_t done = 0;
        while (done < n) {
            uint64_t m = std::min<uint64_t>(batch_kmers_, n - done);
            DevBuf<W> a(m, st_), b(m, st_);
            CUDA_CHECK(cudaMemcpyAsync(a.get(), (const W*)d_words + done, m * sizeof(W), cudaMemcpyDeviceToDevice, st_));
            mutate_with_words(a.get(), b.get(), m, op == 1 ? EDIT_INS : EDIT_DEL);
            done += m;
        }
    }
    // insert / remove the words of several device segments (the per-source regions of a sharded receive buffer) as ONE
    // batch: the segments are gathered into the sort buffer, so the shard is rewritten once, not once per segment
    void words_op_segments_dev(int op, const void* const* seg, const uint64_t* seg_n, uint32_t n_seg) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        uint64_t remaining = 0;
        for (uint32_t i = 0; i < n_seg; i++) remaining += seg_n[i];
        uint32_t si = 0;
        uint64_t so = 0;
        while (remaining) {
            const uint64_t m = std::min<uint64_t>(batch_kmers_, remaining);
            DevBuf<W> a(m, st_), b(m, st_);
            uint64_t filled = 0;
            while (filled < m) {
                const uint64_t take = std::min<uint64_t>(seg_n[si] - so, m - filled);
                if (take) CUDA_CHECK(cudaMemcpyAsync(a.get() + filled, (const W*)seg[si] + so, take * sizeof(W), cudaMemcpyDeviceToDevice, st_));
                filled += take;
                so += take;
                if (so == seg_n[si]) { si++; so = 0; }
            }
            mutate_with_words(a.get(), b.get(), m, op == 1 ? EDIT_INS : EDIT_DEL);
            remaining -= m;
        }
    }
    void kmers_op(int op, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        if (n == 0) return;
        DevBuf<uint64_t> dlo(n, st_), dhi(hi ? n : 1, st_);
        CUDA_CHECK(cudaMemcpyAsync(dlo.get(), lo, n * 8, cudaMemcpyHostToDevice, st_));
        if (hi) CUDA_CHECK(cudaMemcpyAsync(dhi.get(), hi, n * 8, cudaMemcpyHostToDevice, st_));
        DevBuf<W> w(n, st_);
        CBL_LAUNCH((kmers_to_words_kernel<W>), (unsigned)div_up(n, 256), 256, 0, st_, dlo.get(), hi ? dhi.get() : nullptr, (uint64_t)n, P_, w.get());
        DevBuf<uint8_t> flags(n, st_);
        words_op_dev(op, w.get(), n, out ? flags.get() : nullptr);
        if (out) CUDA_CHECK(cudaMemcpyAsync(out, flags.get(), n, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }
    void load_sorted_words(const uint64_t* lo, const uint64_t* hi, uint64_t n) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        if (n == 0) return;
        std::vector<W> h(n);
        for (uint64_t i = 0; i < n; i++) h[i] = sizeof(W) == 16 ? (W)(((u128)(hi ? hi[i] : 0) << 64) | lo[i]) : (W)lo[i];
        DevBuf<W> d(n, st_);
        CUDA_CHECK(cudaMemcpyAsync(d.get(), h.data(), n * sizeof(W), cudaMemcpyHostToDevice, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        words_op_dev(1, d.get(), n, nullptr);
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }

    // ------------------------------------------------------------------------------------------
    // multi-GPU routing building blocks: stable partition of words by owner rank + answer gather
    // ------------------------------------------------------------------------------------------
    void route_words_dev(const void* d_words, uint64_t n, const uint32_t* splitters, uint32_t n_split, void* d_send, uint32_t* d_pos,
                         uint64_t* counts) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        const DestDigit<W> dg = make_dest_digit(splitters, n_split);
        for (uint32_t i = 0; i <= n_split; i++) counts[i] = 0;
        if (n == 0) return;
        if (n > RS_MAX_KEYS) throw Error(CBL_EINVAL, "route batch too large (max 2^30 - 1 words per call)");
        DevBuf<unsigned long long> hist(256, st_);
        hist.zero();
        unsigned hgrid = (unsigned)std::min<uint64_t>(div_up(n, 256 * 8), 148 * 16);
        CBL_LAUNCH((route_hist_kernel<W>), hgrid, 256, 0, st_, (const W*)d_words, n, dg, hist.get());
        unsigned long long h[ROUTE_MAX_SPLIT + 1];
        CUDA_CHECK(cudaMemcpyAsync(h, hist.get(), sizeof(h), cudaMemcpyDeviceToHost, st_));
        CBL_LAUNCH(radix_scan_hist_kernel, 1, 256, 0, st_, hist.get());
        const uint64_t tiles = div_up(n, RsTile<W>::TILE);
        DevBuf<uint32_t> status(tiles * 256, st_), counter(1, st_);
        status.zero();
        counter.zero();
        CBL_LAUNCH((radix_pass_kernel<W, false, DestDigit<W>>), (unsigned)tiles, RS_THREADS, sizeof(W) * RsTile<W>::TILE, st_, (const W*)d_words,
                   (W*)d_send, nullptr, nullptr, n, dg, hist.get(), status.get(), counter.get(), d_pos);
        CUDA_CHECK(cudaStreamSynchronize(st_));
        for (uint32_t i = 0; i <= n_split; i++) counts[i] = h[i];
    }
    DestDigit<W> make_dest_digit(const uint32_t* splitters, uint32_t n_split) const {
        if (n_split > ROUTE_MAX_SPLIT) throw Error(CBL_EINVAL, "too many splitters");
        DestDigit<W> dg;
        dg.suffix_bits = P_.suffix_bits;
        dg.n_split = n_split;
        for (int i = 0; i < ROUTE_MAX_SPLIT; i++) dg.split[i] = i < (int)n_split ? splitters[i] : 0xFFFFFFFFu;
        return dg;
    }
    void route_counts_dev(const void* d_words, uint64_t n, const uint32_t* splitters, uint32_t n_split, uint64_t* counts) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        const DestDigit<W> dg = make_dest_digit(splitters, n_split);
        for (uint32_t i = 0; i <= n_split; i++) counts[i] = 0;
        if (n == 0) return;
        DevBuf<unsigned long long> hist(256, st_);
        hist.zero();
        unsigned hgrid = (unsigned)std::min<uint64_t>(div_up(n, 256 * 8), 148 * 16);
        CBL_LAUNCH((route_hist_kernel<W>), hgrid, 256, 0, st_, (const W*)d_words, n, dg, hist.get());
        unsigned long long h[ROUTE_MAX_SPLIT + 1];
        CUDA_CHECK(cudaMemcpyAsync(h, hist.get(), sizeof(h), cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        for (uint32_t i = 0; i <= n_split; i++) counts[i] = h[i];
    }
    // every word goes straight to peer_recv[dest][recv_offset[dest] + (rank among this rank's words for dest)];
    // counts = this rank's per-destination counts (from route_counts_dev); d_pos[i] = slot of word i in send order
    void route_scatter_dev(const void* d_words, uint64_t n, const uint32_t* splitters, uint32_t n_split, void* const* peer_recv,
                           const uint64_t* recv_offset, const uint64_t* counts, uint32_t* d_pos) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        const DestDigit<W> dg = make_dest_digit(splitters, n_split);
        if (n == 0) return;
        if (n > RS_MAX_KEYS) throw Error(CBL_EINVAL, "route batch too large (max 2^30 - 1 words per call)");
        unsigned long long h_base[512];
        PeerOuts po;
        unsigned long long run = 0;
        for (int d = 0; d < 256; d++) {
            h_base[d] = d <= (int)n_split ? recv_offset[d] : 0;
            h_base[256 + d] = run;
            if (d <= (int)n_split) run += counts[d];
        }
        for (int d = 0; d <= ROUTE_MAX_SPLIT; d++) po.p[d] = d <= (int)n_split ? peer_recv[d] : nullptr;
        if (run != n) throw Error(CBL_EINVAL, "route_scatter: counts do not add up to n");
        DevBuf<unsigned long long> base(512, st_);
        CUDA_CHECK(cudaMemcpyAsync(base.get(), h_base, sizeof h_base, cudaMemcpyHostToDevice, st_));
        const uint64_t tiles = div_up(n, RsTile<W>::TILE);
        DevBuf<uint32_t> status(tiles * 256, st_), counter(1, st_);
        status.zero();
        counter.zero();
        CBL_LAUNCH((radix_pass_kernel<W, false, DestDigit<W>, true>), (unsigned)tiles, RS_THREADS, sizeof(W) * RsTile<W>::TILE, st_,
                   (const W*)d_words, (W*)nullptr, nullptr, nullptr, n, dg, base.get(), status.get(), counter.get(), d_pos, po, base.get() + 256);
        CUDA_CHECK(cudaStreamSynchronize(st_));  // h_base is a stack array; the stores to the peers are complete
    }
    // Fused encode + necklace + route (seq_words_kernel MODE 2): every word of the records goes straight into this rank's
    // region (cap words) of its owner's receive buffer; counts[d] = words reserved for owner d (> cap: nothing of the
    // overflow was written, the caller retries with a larger cap); d_pos[i] = d * cap + index inside the region.
    void seq_route_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, const uint32_t* splitters,
                       uint32_t n_split, void* const* peer_region, uint64_t cap, uint32_t* d_pos, uint64_t* counts) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        ShardArgs<W> sa{};
        sa.dest = make_dest_digit(splitters, n_split);
        for (uint32_t i = 0; i <= ROUTE_MAX_SPLIT; i++) sa.peer[i] = i <= n_split ? (W*)peer_region[i] : nullptr;
        for (uint32_t i = 0; i <= n_split; i++) counts[i] = 0;
        if ((uint64_t)(n_split + 1) * cap >= (1ull << 32)) throw Error(CBL_EINVAL, "seq_route: (ranks x region capacity) must stay below 2^32 words");
        PieceList pl;
        build_pieces(offsets, 0, n_seqs, pl);
        if (pl.kmers.empty()) return;
        DevPieces dp;
        upload_pieces(pl, 0, pl.kmers.size(), 0, d_seq, n_bytes, dp, st_);
        DevBuf<unsigned long long> cnt(17, st_);   // [16] per-owner counters, [16] = error offset
        CUDA_CHECK(cudaMemsetAsync(cnt.get(), 0, 16 * 8, st_));
        CUDA_CHECK(cudaMemsetAsync(cnt.get() + 16, 0xFF, 8, st_));
        sa.cnt = cnt.get();
        sa.pos = d_pos;
        sa.cap = cap;
        const unsigned grid = (unsigned)std::min<uint64_t>(dp.batch.n_chunks, env_u64("CBL_ROUTE_GRID", 1u << 30));
        CBL_LAUNCH((seq_words_kernel<W, Suf, 2, false, 32, 1>), grid, SW_THREADS, 0, st_, dp.batch, P_, (W*)nullptr, (uint8_t*)nullptr, view(),
                   cnt.get() + 16, sa);
        unsigned long long h[17];
        CUDA_CHECK(cudaMemcpyAsync(h, cnt.get(), sizeof h, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));   // the stores to the peers are complete
        if (h[16] != ULLONG_MAX) throw_bad_byte(h[16]);
        for (uint32_t i = 0; i <= n_split; i++) counts[i] = h[i];
    }
    void gather_u8_dev(const uint8_t* d_src, const uint32_t* d_pos, uint64_t n, uint8_t* d_out) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        if (n == 0) return;
        CBL_LAUNCH(gather_u8_kernel, (unsigned)div_up(div_up(n, (uint64_t)4), (uint64_t)256), 256, 0, st_, d_src, d_pos, n, d_out);
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }

    // ------------------------------------------------------------------------------------------
    // export / iteration (ascending word order == ascending prefix then suffix)
    // ------------------------------------------------------------------------------------------
    void export_words_dev(uint64_t start, uint64_t count, int to_kmers, void* d_out) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        if (count == 0) return;
        if (start + count > n_) throw Error(CBL_EINVAL, "export range out of bounds");
        CBL_LAUNCH((expand_kernel<W, Suf>), (unsigned)div_up(count, OP_TILE), OP_THREADS, 0, st_, view(), P_, start, count, to_kmers, (W*)d_out);
    }
    void export_words(uint64_t start, uint64_t cap, int to_kmers, uint64_t* lo, uint64_t* hi, uint64_t* n_out) override {
        uint64_t cnt = start >= n_ ? 0 : std::min<uint64_t>(cap, n_ - start);
        *n_out = cnt;
        if (!cnt) return;
        DevBuf<W> d(cnt, st_);
        export_words_dev(start, cnt, to_kmers, d.get());
        download_words(d.get(), cnt, lo, hi);
    }
    void bucket_sizes(uint32_t* prefixes, uint32_t* sizes, uint64_t cap, uint64_t* n_out) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        *n_out = nb_;
        if (!prefixes || !nb_) return;
        if (cap < nb_) throw Error(CBL_EINVAL, "output buffer too small");
        DevBuf<uint32_t> sz(nb_, st_);
        CBL_LAUNCH(bucket_sizes_kernel, (unsigned)div_up(nb_, 256), 256, 0, st_, bucket_off_.get(), nb_, sz.get());
        CUDA_CHECK(cudaMemcpyAsync(prefixes, bucket_prefix_.get(), (size_t)nb_ * 4, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaMemcpyAsync(sizes, sz.get(), (size_t)nb_ * 4, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }

    // ------------------------------------------------------------------------------------------
    // set operations (src/cbl.rs:411-569): both operands must agree on every parameter
    // ------------------------------------------------------------------------------------------
    Index* check_other(IIndex* o) {
        auto* p = dynamic_cast<Index*>(o);
        const Config& c = o->config();
        if (!p || c.k != cfg_.k || c.prefix_bits != cfg_.prefix_bits || c.word_bits != cfg_.word_bits)
            throw Error(CBL_EINVAL, "set operation between indexes with different K / T / PREFIX_BITS");
        if (c.canonical != cfg_.canonical) throw Error(CBL_EINVAL, "One of the index is canonical while the other isn't");  // cbl.rs:422-425
        if (c.device != cfg_.device) throw Error(CBL_EINVAL, "set operation between indexes on different devices");
        return p;
    }
    void setop_new_state(int op, Index* o, NewState& ns) {
        o->sync();
        if (use_merge_) {
            if (o->n_ == 0 && op != SETOP_AND) { ns.changed = false; return; }
            DevBuf<W> theirs(o->n_ ? o->n_ : 1, st_);
            if (o->n_) CBL_LAUNCH((expand_kernel<W, Suf>), (unsigned)div_up(o->n_, OP_TILE), OP_THREADS, 0, st_, o->view(), P_, (uint64_t)0, o->n_, 0, theirs.get());
            merge_new_state(theirs.get(), o->n_, op, ns);
            return;
        }
        if (op == SETOP_AND) {
            if (n_ == 0) { ns.changed = false; return; }
            DevBuf<W> mine(n_, st_);
            export_words_dev(0, n_, 0, mine.get());
            compute_new_state(mine.get(), n_, EDIT_KEEP_ONLY, o->view(), nullptr, ns);
        } else {
            if (o->n_ == 0) { ns.changed = false; return; }
            DevBuf<W> theirs(o->n_, st_);
            // expand the other operand on OUR stream (its state is final after o->sync())
            CBL_LAUNCH((expand_kernel<W, Suf>), (unsigned)div_up(o->n_, OP_TILE), OP_THREADS, 0, st_, o->view(), P_, (uint64_t)0, o->n_, 0, theirs.get());
            const int mode = op == SETOP_OR ? EDIT_INS : op == SETOP_SUB ? EDIT_DEL : (EDIT_INS | EDIT_DEL);
            compute_new_state(theirs.get(), o->n_, mode, view(), nullptr, ns);
        }
    }
    void setop_assign(int op, IIndex* other) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        Index* o = check_other(other);
        if (o == this) {  // x op x
            if (op == SETOP_SUB || op == SETOP_XOR) clear();
            return;
        }
        NewState ns;
        setop_new_state(op, o, ns);
        if (ns.changed) adopt(ns);
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }
    IIndex* setop(int op, IIndex* other) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        Index* o = check_other(other);
        std::unique_ptr<Index> res(new Index(cfg_));
        if (o == this) {
            if (op == SETOP_OR || op == SETOP_AND) { sync(); res->copy_state_from(*this); res->sync(); }
            return res.release();
        }
        NewState ns;
        setop_new_state(op, o, ns);
        CUDA_CHECK(cudaStreamSynchronize(st_));
        if (ns.changed) res->adopt(ns);   // buffers were allocated on our stream; work on them is complete
        else { res->copy_state_from(*this); res->sync(); }
        return res.release();
    }
    void clear() {
        sub_valid_ = false;
        dir_.zero();
        bucket_off_.alloc(1, st_);
        bucket_off_.zero();
        nb_ = 0; n_ = 0; last_prefix_ = 0;
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }
};

#define CBL_INSTANTIATE_INDEX(NAME, W, SUF) \
    IIndex* NAME(const Config& cfg) { return new Index<W, SUF>(cfg); }

}  // namespace cbl
