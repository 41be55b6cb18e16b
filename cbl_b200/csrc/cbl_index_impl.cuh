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
#include <mutex>
#include <set>

#include "index_ops.cuh"
#include "merge_ops.cuh"
#include "radix_sort.cuh"
#include "seg_sort.cuh"
#ifndef CBL_PROBE_WB
#define CBL_PROBE_WB 32   // bytes per suffix window of the membership probe
#endif
#include "sanitize.cuh"
#include "seq_words.cuh"
#include "shard_query.cuh"

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
    DevBuf<int8_t> sub_;      // interpolation corrections for the membership probe, rebuilt lazily
    bool sub_valid_ = false;
    uint32_t nb_ = 0;
    uint64_t n_ = 0;
    uint32_t last_prefix_ = 0;  // prefix of the last bucket (for the reference's is_empty quirk)
    uint64_t n_dir_ = 0;  // directory words: 32 prefixes each
    uint64_t batch_kmers_;
    double sort_conc_ = 1.0;    // the batches cover ~1 / sort_conc_ of the prefix mass (set_sort_concentration; grows after a fallback)
    bool sort_hybrid_ = true;   // CBL_SORT=lsd (read once, here): plain LSD passes instead of top passes + segment sort
    static constexpr uint64_t SUF_PAD = 16;  // suffix arrays are over-allocated: probe windows are 32-byte aligned loads
    // pinned host staging of this handle: piece tables on their way in, status words on their way out.  A mutation
    // enqueues ALL of its kernels and reads this block back with ONE synchronisation at its end.
    struct HostStage {
        uint8_t* p = nullptr;
        size_t cap = 0;
        void reserve(size_t bytes) {
            if (bytes <= cap) return;
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = std::max<size_t>(bytes + bytes / 2, 1 << 16);
            CUDA_CHECK(cudaMallocHost((void**)&p, cap));
        }
        ~HostStage() { if (p) cudaFreeHost(p); }
    } stage_;
    cudaEvent_t stage_ev_ = nullptr;           // the last copy out of stage_ has completed
    unsigned long long* h_status_ = nullptr;   // pinned (32 words): [0] merged elements, [1] buckets, [2] elements by the directory,
                                               // [3] last prefix, [4] segment-sort fail flag, [5] non-ACGT byte offset

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
        if (env_u64("CBL_ARENA", 1) == 0) {   // cudaMallocAsync mode only: keep freed blocks in the driver's pool
            cudaMemPool_t pool;
            CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, cfg.device));
            uint64_t thr = UINT64_MAX;
            CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        }
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
        sub_.alloc(8, st_);
        sub_.zero();
        { const char* m = getenv("CBL_SORT"); sort_hybrid_ = !(m && std::string(m) == "lsd"); }
        CUDA_CHECK(cudaMallocHost((void**)&h_status_, 32 * sizeof(unsigned long long)));
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
        // opt-ins for > 48 KB of dynamic shared memory are per device (context): done once per (instantiation, device),
        // under a lock so that handles may be created from several threads / on several GPUs of one process
        {
            static std::mutex mu;
            static std::set<int> done;
            std::lock_guard<std::mutex> lk(mu);
            if (!done.count(cfg.device)) {
                auto opt_in = [](auto kernel, size_t bytes) {
                    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                };
                constexpr size_t MG_DYN = (size_t)MgCfg<W>::SMEM_ELEMS * sizeof(W) + 1024;
                opt_in(radix_pass_kernel<W, false, ByteDigit<W>>, 96 * 1024);
                opt_in(radix_pass_kernel<W, false, DestDigit<W>>, 96 * 1024);
                opt_in(radix_pass_kernel<W, false, DestDigit<W>, true>, 96 * 1024);
                opt_in(seg_sort_kernel<W, true>, SsTile<W>::SMEM);
                opt_in(seg_sort_kernel<W, false>, SsTile<W>::SMEM);
                opt_in(merge_apply_kernel<W, Suf, MERGE_OR, false>, MG_DYN);
                opt_in(merge_apply_kernel<W, Suf, MERGE_SUB, false>, MG_DYN);
                opt_in(merge_apply_kernel<W, Suf, MERGE_OR, true>, MG_DYN);
                opt_in(merge_apply_kernel<W, Suf, MERGE_AND, true>, MG_DYN);
                opt_in(merge_apply_kernel<W, Suf, MERGE_SUB, true>, MG_DYN);
                opt_in(merge_apply_kernel<W, Suf, MERGE_XOR, true>, MG_DYN);
                done.insert(cfg.device);
            }
        }
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }
    ~Index() override {
        cudaSetDevice(cfg_.device);
        dir_.release(); bucket_range_.release(); bucket_prefix_.release(); bucket_off_.release(); suf_.release(); sub_.release();
        if (st_) { cudaStreamSynchronize(st_); arena::retire_stream(st_); cudaStreamDestroy(st_); }
        for (auto& s : side_) if (s) { cudaStreamSynchronize(s); arena::retire_stream(s); cudaStreamDestroy(s); }
        if (h_status_) cudaFreeHost(h_status_);
        if (stage_ev_) cudaEventDestroy(stage_ev_);
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
    void set_sort_concentration(double factor) override { sort_conc_ = factor >= 1.0 ? factor : 1.0; }

    IndexView<Suf> view() const {
        IndexView<Suf> v;
        v.dir = dir_.get(); v.bucket_prefix = bucket_prefix_.get(); v.bucket_off = bucket_off_.get();
        v.bucket_range = bucket_range_.get(); v.suf = suf_.get(); v.sub = sub_.get(); v.nb = nb_; v.n = n_;
        return v;
    }
    // (re)build the probe's correction bytes if the set changed since they were last computed
    void ensure_sub() {
        if (sub_valid_) return;
        const uint64_t n_slots = (n_ >> SUB_SHIFT) + 8;
        sub_.alloc(n_slots, st_);
        CBL_LAUNCH((build_sub_kernel<Suf>), (unsigned)div_up(n_slots, 256), 256, 0, st_, view(), P_.suffix_bits, sub_.get(), n_slots);
        sub_valid_ = true;
    }

    IIndex* clone() override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        std::unique_ptr<Index> c(new Index(cfg_));
        c->sort_conc_ = sort_conc_;
        sync();
        c->copy_state_from(*this);
        c->sync();
        return c.release();
    }
    IIndex* new_empty(int canonical = -1) override {
        Config c = cfg_;
        if (canonical >= 0) c.canonical = canonical;
        Index* e = new Index(c);
        e->sort_conc_ = sort_conc_;
        return e;
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
        DevBuf<uint8_t> blob;   // [byte_off | out_off | chunk0 | kmers] of the pieces
        SeqBatch batch;
    };
    // upload pieces [p0, p1) with out offsets rebased by out_base.  The tables go through the handle's pinned staging
    // block, so the copies are true asynchronous DMA and the host does not wait for them.
    void upload_pieces(const PieceList& pl, size_t p0, size_t p1, uint64_t out_base, const uint8_t* d_seq, uint64_t n_bytes,
                       DevPieces& dp, cudaStream_t s) {
        const size_t np = p1 - p0;
        if (stage_ev_) CUDA_CHECK(cudaEventSynchronize(stage_ev_));   // the previous user of the staging block is done (normally long ago)
        else CUDA_CHECK(cudaEventCreateWithFlags(&stage_ev_, cudaEventDisableTiming));
        const size_t o_byte = 0, o_out = np * 8, o_ch = 2 * np * 8, o_km = (3 * np + 1) * 8;
        stage_.reserve(o_km + np * 4);
        uint64_t* h_byte = reinterpret_cast<uint64_t*>(stage_.p + o_byte);
        uint64_t* h_out = reinterpret_cast<uint64_t*>(stage_.p + o_out);
        uint64_t* h_ch = reinterpret_cast<uint64_t*>(stage_.p + o_ch);
        uint32_t* h_km = reinterpret_cast<uint32_t*>(stage_.p + o_km);
        for (size_t i = 0; i < np; i++) {
            h_byte[i] = pl.byte_off[p0 + i];
            h_out[i] = pl.out_off[p0 + i] - out_base;
            h_ch[i] = pl.chunk0[p0 + i] - pl.chunk0[p0];
            h_km[i] = pl.kmers[p0 + i];
        }
        h_ch[np] = pl.chunk0[p1] - pl.chunk0[p0];
        // one device block, one copy: [byte_off | out_off | chunk0 | kmers]
        dp.blob.alloc(o_km + np * 4, s);
        CUDA_CHECK(cudaMemcpyAsync(dp.blob.get(), stage_.p, o_km + np * 4, cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaEventRecord(stage_ev_, s));
        dp.batch.seq = d_seq; dp.batch.seq_end = d_seq + n_bytes;
        dp.batch.piece_byte = reinterpret_cast<const uint64_t*>(dp.blob.get() + o_byte);
        dp.batch.piece_out = reinterpret_cast<const uint64_t*>(dp.blob.get() + o_out);
        dp.batch.piece_chunk0 = reinterpret_cast<const uint64_t*>(dp.blob.get() + o_ch);
        dp.batch.piece_kmers = reinterpret_cast<const uint32_t*>(dp.blob.get() + o_km);
        dp.batch.n_pieces = (uint32_t)np; dp.batch.n_chunks = h_ch[np];
    }

    // plan of the LSD part of the batch sort, known before the words exist (it only depends on their number)
    struct SortPlan {
        int n_digits = 0, n_pass = 0, first = 0;   // LSD passes over digits [first, first + n_pass); n_pass == n_digits: plain LSD sort
    };
    SortPlan plan_sort(uint64_t n) const {
        SortPlan sp;
        const int key_bits = P_.bits + P_.pos_bits;
        sp.n_digits = (key_bits + 7) / 8;
        sp.n_pass = sp.n_digits;
        // Hybrid (default): LSD passes over the top digits only, then the in-shared-memory segment sort (seg_sort.cuh).
        // The number of passes is chosen so that the largest group of words sharing their sorted top bits fits a segment
        // tile: for k-mer data the most frequent b-bit head of a necklace word has mass ~ 2K / 2^b.  Any other
        // distribution is still sorted exactly: a segment that does not fit raises the fail flag and the batch is
        // re-sorted by the plain LSD passes.  A shard of a prefix-sharded set sees only 1 / g of the prefix mass, i.e. heads g
        // times as frequent: sort_conc_ = g (set by the sharded hosts; it also grows by itself after a fallback).
        if (sort_hybrid_) {
            for (int c = 1; c < sp.n_digits - 1; c++) {   // c LSD passes leave 8 * (n_digits - c) low bits to the segment sort
                const int b = key_bits - 8 * (sp.n_digits - c);
                const double est = (double)n * sort_conc_ * (double)P_.bits / std::ldexp(1.0, b);
                if (est * 1.5 <= (double)SsTile<W>::T) { sp.n_pass = c; break; }
            }
        }
        sp.first = sp.n_digits - sp.n_pass;
        return sp;
    }
    // launches the fused encode + necklace (+ probe) kernel (no synchronisation); *err must hold
    // ULLONG_MAX on entry and receives the smallest offending byte offset if a non-ACGT byte is seen.
    // hist != null (mode 0): the kernel also counts the digits of plan's LSD passes (hist zeroed by the caller).
    void launch_seq_words(const SeqBatch& b, int mode, bool brute, W* d_words, uint8_t* d_flags, unsigned long long* err, cudaStream_t s,
                          unsigned long long* hist = nullptr, const SortPlan* plan = nullptr) {
        if (b.n_chunks == 0) return;
        unsigned grid = (unsigned)std::min<uint64_t>(b.n_chunks, 1u << 30);
        IndexView<Suf> v = view();
        ShardArgs<W> sa{};
        if (mode == 0) {
            const uint64_t persist = env_u64("CBL_WORDS_PERSIST", 0);   // developer knob: persistent grid 0 never (measured: one CTA per chunk is 0.8 ms faster per 500 M k-mers), 1 always, 2 with the histogram
            if ((hist && plan && plan->n_pass <= SW_HIST_MAX && !brute) || persist == 1) {
                // persistent grid: every CTA keeps its digit counters in shared memory over all of its chunks
                static int occ = 0;
                if (!occ) CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, seq_words_kernel<W, Suf, 0, false, 32, 1>, SW_THREADS, 0));
                int sms = 0;
                CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg_.device));
                if (persist) grid = (unsigned)std::min<uint64_t>(b.n_chunks, (uint64_t)sms * (uint64_t)std::max(occ, 1));
                if (hist && plan && plan->n_pass <= SW_HIST_MAX && !brute) { sa.hist = hist; sa.hist_first = plan->first; sa.hist_np = plan->n_pass; }
            }
            if (brute) CBL_LAUNCH((seq_words_kernel<W, Suf, 0, true, 32, 1>), grid, SW_THREADS, 0, s, b, P_, d_words, d_flags, v, err, sa);
            else CBL_LAUNCH((seq_words_kernel<W, Suf, 0, false, 32, 1>), grid, SW_THREADS, 0, s, b, P_, d_words, d_flags, v, err, sa);
        } else {
            CBL_LAUNCH((seq_words_kernel<W, Suf, 1, false, CBL_PROBE_WB, 1>), grid, SW_THREADS, 0, s, b, P_, d_words, d_flags, v, err, sa);
        }
    }
    // membership of the words of up to PROBE_MAX_SEG device segments in ONE launch (MODE 3 of the fused kernel: same
    // staged probe + deferred queue, no sequence front end); the answer pointers may be peer memory
    void launch_probe_segments(const void* const* seg, const uint64_t* seg_n, uint8_t* const* seg_out, uint32_t n_seg, cudaStream_t s) {
        uint32_t done = 0;
        while (done < n_seg) {
            ShardArgs<W> sa{};
            uint64_t chunks = 0;
            int k = 0;
            for (; done < n_seg && k < PROBE_MAX_SEG; done++) {
                if (seg_n[done] == 0) continue;
                sa.seg_words[k] = (const W*)seg[done];
                sa.seg_out[k] = seg_out[done];
                sa.seg_n[k] = seg_n[done];
                sa.seg_chunk0[k] = chunks;
                chunks += div_up(seg_n[done], CHUNK_KMERS);
                k++;
            }
            for (int t = k; t <= PROBE_MAX_SEG; t++) sa.seg_chunk0[t] = chunks;
            sa.n_seg = k;
            if (k == 0) continue;
            // CBL_WORDS_GRID caps the grid (the kernel strides over the chunks)
            const unsigned grid = (unsigned)std::min<uint64_t>(chunks, env_u64("CBL_WORDS_GRID", 1u << 30));
            CBL_LAUNCH((seq_words_kernel<W, Suf, 3, false, CBL_PROBE_WB, 1>), grid, SW_THREADS, 0, s, SeqBatch{}, P_, (W*)nullptr, (uint8_t*)nullptr,
                       view(), (unsigned long long*)nullptr, sa);
        }
    }
    void launch_probe_words(const W* d_words, uint64_t n, uint8_t* d_flags, cudaStream_t s) {
        const void* seg[1] = {d_words};
        uint8_t* out[1] = {d_flags};
        launch_probe_segments(seg, &n, out, 1, s);
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
        CBL_LAUNCH(sanitize_chunks_kernel, grid, SAN_THREADS, 0, s, in, cfg_.k, (uint32_t*)nullptr, out.dp.batch.piece_byte, out.clean.get());
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
    // batch sort (no host synchronisation inside: the fail flag is read with the mutation's status block)
    // ------------------------------------------------------------------------------------------
    // device status block of one mutation, read back once at its end (h_status_ mirrors it)
    enum { ST_MERGED = 0, ST_BUCKETS = 1, ST_DIR_ELEMS = 2, ST_LAST_PREFIX = 3, ST_SORT_FAIL = 4, ST_BAD_BYTE = 5, ST_WORDS = 8 };
    void init_status(DevBuf<unsigned long long>& st) {
        st.alloc(ST_WORDS, st_);
        CUDA_CHECK(cudaMemsetAsync(st.get(), 0, ST_WORDS * 8, st_));
        CUDA_CHECK(cudaMemsetAsync(st.get() + ST_BAD_BYTE, 0xFF, 8, st_));
    }
    void read_status(const DevBuf<unsigned long long>& st) {   // the ONE synchronisation of a mutation
        CUDA_CHECK(cudaMemcpyAsync(h_status_, st.get(), ST_WORDS * 8, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }
    struct Sorted {
        W* out;        // the sorted words
        W* grouped;    // hybrid sort: the buffer that still holds every word grouped by its top digits (input of the segment sort)
    };
    // sorts n keys held in a (b: same-size scratch).  hist: digit counters of plan's passes already filled by the words
    // kernel, or null (counted here).  fail: device flag the segment sort raises when a group does not fit its tile.
    Sorted sort_keys(W* a, W* b, uint64_t n, const SortPlan& sp, unsigned long long* hist, unsigned long long* fail) {
        if (n <= 1) return {a, a};
        if (n > RS_MAX_KEYS) throw Error(CBL_EINVAL, "internal: sort batch too large");
        W* res = lsd_passes(a, b, n, sp.first, sp.n_pass, hist);
        if (sp.n_pass == sp.n_digits) return {res, res};
        W* other = res == a ? b : a;
        const int key_bits = P_.bits + P_.pos_bits;
        const int shift = 8 * (sp.n_digits - sp.n_pass);
        // nominal tile: leave room for segments 3x longer than the longest one expected, at least 512 keys
        const double est_seg = (double)n * (double)P_.bits / std::ldexp(1.0, key_bits - shift);
        const int room = (int)std::min<double>(SsTile<W>::CAP / 2, std::max<double>(512.0, std::ceil(3.0 * est_seg / 256.0) * 256.0));
        const int tile = SsTile<W>::CAP - room;
        unsigned* f = reinterpret_cast<unsigned*>(fail);
        if (shift <= 32)
            CBL_LAUNCH((seg_sort_kernel<W, true>), (unsigned)div_up(n, (uint64_t)tile), SS_THREADS, SsTile<W>::SMEM, st_, res, other, n, shift, f, tile);
        else
            CBL_LAUNCH((seg_sort_kernel<W, false>), (unsigned)div_up(n, (uint64_t)tile), SS_THREADS, SsTile<W>::SMEM, st_, res, other, n, shift, f, tile);
        return {other, res};
    }
    // LSD passes over digits [first, first + n_pass); returns the buffer that holds the result.  ext_hist: counters of
    // exactly these digits, already filled (not yet scanned); null: counted here with one more read of the keys.
    W* lsd_passes(W* a, W* b, uint64_t n, int first, int n_pass, unsigned long long* ext_hist = nullptr) {
        DevBuf<unsigned long long> own;
        unsigned long long* hist = ext_hist;
        if (!hist) {
            own.alloc((size_t)n_pass * 256, st_);
            own.zero();
            hist = own.get();
            unsigned hgrid = (unsigned)std::min<uint64_t>(div_up(n, RH_THREADS * RH_KEYS), 148 * 4);
            CBL_LAUNCH((radix_hist_kernel<W>), hgrid, RH_THREADS, (size_t)n_pass * 256 * sizeof(uint32_t), st_, a, n, n_pass, hist, first);
        }
        CBL_LAUNCH(radix_scan_hist_kernel, n_pass, 256, 0, st_, hist);
        const uint64_t tiles = div_up(n, RsTile<W>::TILE);
        DevBuf<uint32_t> status(tiles * 256, st_), counter(1, st_);
        const size_t smem = sizeof(W) * RsTile<W>::TILE;
        W *src = a, *dst = b;
        for (int p = 0; p < n_pass; p++) {
            status.zero();
            counter.zero();
            CBL_LAUNCH((radix_pass_kernel<W, false, ByteDigit<W>>), (unsigned)tiles, RS_THREADS, smem, st_, src, dst, nullptr, nullptr, n,
                       ByteDigit<W>{8 * (first + p)}, hist + (size_t)p * 256, status.get(), counter.get(), (uint32_t*)nullptr);
            std::swap(src, dst);
        }
        return src;
    }

    // ------------------------------------------------------------------------------------------
    // new state of the shard after a mutation (built out of place, then swapped in)
    // ------------------------------------------------------------------------------------------
    struct NewState {
        DevBuf<uint2> dir, bucket_range;
        DevBuf<uint32_t> bucket_prefix, bucket_off;
        DevBuf<Suf> suf;
        uint32_t nb = 0;
        uint64_t n = 0;
        uint32_t last_prefix = 0;
        bool changed = false;   // a new state was built (its totals still have to be read: finish_new_state)
        bool pending = false;   // kernels enqueued, totals not read yet
    };
    void adopt(NewState& ns) {
        dir_.swap(ns.dir); bucket_range_.swap(ns.bucket_range); bucket_prefix_.swap(ns.bucket_prefix);
        bucket_off_.swap(ns.bucket_off); suf_.swap(ns.suf);
        nb_ = ns.nb; n_ = ns.n; last_prefix_ = ns.last_prefix;
        sub_valid_ = false;
        dir_.rebind(st_); bucket_range_.rebind(st_); bucket_prefix_.rebind(st_); bucket_off_.rebind(st_); suf_.rebind(st_);
    }

    // ------------------------------------------------------------------------------------------
    // mutation by one streaming merge (merge_ops.cuh): B = sorted words (repeats allowed) or another index (CSR)
    // ------------------------------------------------------------------------------------------
    static uint64_t round_capacity(uint64_t n) {
        if (n < (1u << 20)) return n;
        int top = 63 - __builtin_clzll(n);
        const uint64_t step = 1ull << (top - 2);
        return (n + step - 1) / step * step;
    }
    template <bool BCSR> void launch_merge_apply(int op, unsigned tiles, const IndexView<Suf>& v, const MergeB<W, Suf, BCSR>& B, const uint32_t* part_i,
                                                 const uint32_t* part_r, const uint32_t* part_rb, Suf* suf_out, uint32_t* cnt, uint64_t* status,
                                                 uint32_t* counter, unsigned long long* n_out, const unsigned long long* skip) {
        const size_t smem = (size_t)MgCfg<W>::SMEM_ELEMS * sizeof(W);
#define CBL_MERGE_CASE(OP) \
    CBL_LAUNCH((merge_apply_kernel<W, Suf, OP, BCSR>), tiles, MG_THREADS, smem, st_, v, P_, B, part_i, part_r, part_rb, suf_out, cnt, status, counter, n_out, skip)
        if constexpr (BCSR) {
            switch (op) {
                case MERGE_OR: CBL_MERGE_CASE(MERGE_OR); break;
                case MERGE_AND: CBL_MERGE_CASE(MERGE_AND); break;
                case MERGE_SUB: CBL_MERGE_CASE(MERGE_SUB); break;
                default: CBL_MERGE_CASE(MERGE_XOR); break;
            }
        } else {   // batches only insert or remove
            if (op == MERGE_OR) CBL_MERGE_CASE(MERGE_OR);
            else if (op == MERGE_SUB) CBL_MERGE_CASE(MERGE_SUB);
            else throw Error(CBL_EINVAL, "internal: word batches support insert / remove only");
        }
#undef CBL_MERGE_CASE
    }
    // Enqueues the whole mutation (partition, merge, directory, bucket tables) on the handle's stream and returns without
    // waiting: the totals land in dstat (ST_MERGED .. ST_LAST_PREFIX).  finish_new_state() turns them into ns.n / ns.nb
    // after the caller's read_status().
    template <bool BCSR> void merge_new_state(const MergeB<W, Suf, BCSR>& B, uint64_t nB, int op, NewState& ns, unsigned long long* dstat) {
        ns.changed = false;
        ns.pending = false;
        const uint64_t V = n_ + nB;
        if (nB == 0 || (n_ == 0 && (op == MERGE_AND || op == MERGE_SUB))) {
            if (op == MERGE_AND && n_ != 0) {  // A & {} = {}
                ns.dir.alloc(n_dir_, st_); ns.dir.zero();
                ns.bucket_range.alloc(1, st_); ns.bucket_prefix.alloc(1, st_); ns.bucket_off.alloc(1, st_); ns.bucket_off.zero();
                ns.suf.alloc(SUF_PAD, st_);
                ns.nb = 0; ns.n = 0; ns.last_prefix = 0; ns.changed = true;
            }
            return;
        }
        const IndexView<Suf> v = view();
        const uint64_t tiles = div_up(V, MgCfg<W>::TILE);
        DevBuf<uint32_t> part_i(tiles + 1, st_), part_r(tiles + 1, st_), part_rb(BCSR ? tiles + 1 : 1, st_);
        CBL_LAUNCH((merge_partition_kernel<W, Suf, BCSR>), (unsigned)div_up(tiles + 1, 128), 128, 0, st_, v, P_, B, tiles, part_i.get(), part_r.get(), part_rb.get(),
                   (const unsigned long long*)(dstat + ST_SORT_FAIL));
        const uint64_t n_prefix = n_dir_ * 32;
        DevBuf<uint32_t> cnt(n_prefix, st_);
        cnt.zero();
        // capacity is rounded up to 4 steps per octave so that the blocks freed by earlier, smaller states of a
        // growing index can be reused by the arena instead of a fresh driver allocation per batch
        const uint64_t out_max = (op == MERGE_AND || op == MERGE_SUB) ? n_ : V;
        ns.suf.alloc(round_capacity(out_max) + SUF_PAD, st_);
        {
            Lookback lb(tiles, st_);
            launch_merge_apply<BCSR>(op, (unsigned)tiles, v, B, part_i.get(), part_r.get(), part_rb.get(), ns.suf.get(), cnt.get(), lb.status.get(),
                                     lb.counter.get(), dstat + ST_MERGED, dstat + ST_SORT_FAIL);
        }
        ns.dir.alloc(n_dir_, st_);
        DevBuf<uint32_t> word_off(n_dir_, st_);
        {
            const uint64_t t = div_up(n_dir_, 256);
            Lookback lb_rank(t, st_), lb_off(t, st_);
            CBL_LAUNCH(dir_bits_kernel, (unsigned)t, 256, 0, st_, cnt.get(), n_dir_, ns.dir.get(), word_off.get(), lb_rank.status.get(),
                       lb_off.status.get(), lb_rank.counter.get(), dstat + ST_BUCKETS);
        }
        // the bucket tables are sized by an upper bound (the bucket count is still on the device)
        const uint64_t nb_max = std::max<uint64_t>(1, std::min<uint64_t>(n_prefix, out_max));
        ns.bucket_prefix.alloc(nb_max, st_);
        ns.bucket_off.alloc(nb_max + 1, st_);
        ns.bucket_range.alloc(nb_max, st_);
        CBL_LAUNCH(dir_fill_kernel, (unsigned)div_up(n_dir_, 256), 256, 0, st_, cnt.get(), ns.dir.get(), word_off.get(), n_dir_,
                   ns.bucket_prefix.get(), ns.bucket_off.get(), ns.bucket_range.get(), dstat);
        ns.changed = true;
        ns.pending = true;
    }
    void finish_new_state(NewState& ns) {   // after read_status()
        if (!ns.pending) return;
        const uint64_t n_new = h_status_[ST_MERGED], nb_new = h_status_[ST_BUCKETS];
        if (h_status_[ST_DIR_ELEMS] != n_new)
            throw Error(CBL_ECUDA, "internal: directory total " + std::to_string(h_status_[ST_DIR_ELEMS]) + " != element count " + std::to_string(n_new));
        if (n_new >= (1ull << 32))
            throw Error(CBL_EINVAL, "shard would hold >= 2^32 k-mers; bucket offsets are 32-bit — shard the index over more GPUs");
        ns.nb = (uint32_t)nb_new;
        ns.n = n_new;
        ns.last_prefix = nb_new ? (uint32_t)h_status_[ST_LAST_PREFIX] : 0;
        ns.pending = false;
    }

    // unsorted words in `a` (n of them, `b` same-size scratch) -> applied to this index.  hist (optional): digit counters
    // of plan_sort(n)'s passes already filled by the kernel that produced the words; dstat (optional): the caller's status
    // block (then the caller has initialised it and ST_BAD_BYTE is meaningful).  Returns false when the words kernel
    // reported a non-ACGT byte (nothing was applied; the caller re-runs the batch through the sanitiser).
    bool mutate_with_words(W* a, W* b, uint64_t n, int mode, unsigned long long* hist = nullptr, DevBuf<unsigned long long>* ext_stat = nullptr) {
        if (n == 0) return true;
        Trace tr(st_);
        const int op = mode == EDIT_INS ? MERGE_OR : MERGE_SUB;
        const SortPlan sp = plan_sort(n);
        DevBuf<unsigned long long> own_stat;
        if (!ext_stat) init_status(own_stat);
        DevBuf<unsigned long long>& stat = ext_stat ? *ext_stat : own_stat;
        Sorted so = sort_keys(a, b, n, sp, hist, stat.get() + ST_SORT_FAIL);
        NewState ns;
        // the merge treats a repeated batch word as one (merge_ops.cuh): no unique pass, no count read-back
        merge_new_state<false>(MergeB<W, Suf, false>{so.out, n}, n, op, ns, stat.get());
        tr.mark("  sort + merge enqueued");
        read_status(stat);
        tr.mark("  status read (the one sync)");
        if (h_status_[ST_BAD_BYTE] != ULLONG_MAX) return false;
        if (h_status_[ST_SORT_FAIL]) {   // a group of equal top digits did not fit a segment tile: plain LSD passes (exact for any input)
            g_sort_fallbacks.fetch_add(1, std::memory_order_relaxed);
            sort_conc_ = std::max(sort_conc_, 1.0) * 256.0;   // the words of this handle are denser than planned: one more LSD pass from now on
            W* other = so.grouped == a ? b : a;
            W* sorted = lsd_passes(so.grouped, other, n, 0, sp.n_digits);
            init_status(stat);
            ns = NewState();
            merge_new_state<false>(MergeB<W, Suf, false>{sorted, n}, n, op, ns, stat.get());
            read_status(stat);
        }
        finish_new_state(ns);
        if (ns.changed) adopt(ns);
        tr.mark("  adopt");
        return true;
    }

    // ------------------------------------------------------------------------------------------
    // sequence front ends
    // ------------------------------------------------------------------------------------------
    // One batch = words kernel (with the sort's digit histograms folded in) -> LSD passes -> segment sort -> merge ->
    // directory, all enqueued back to back; the host waits once, at the end, for the status block.
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
            DevBuf<W> a(nk, st_), b(nk, st_);
            tr.mark("upload pieces, alloc a,b");
            const SortPlan sp = plan_sort(nk);
            DevBuf<unsigned long long> stat, hist;
            init_status(stat);
            const bool fused_hist = sp.n_pass <= SW_HIST_MAX && nk > 1 && env_u64("CBL_HIST_FUSED", 1) != 0;
            if (fused_hist) { hist.alloc((size_t)sp.n_pass * 256, st_); hist.zero(); }
            launch_seq_words(dp.batch, 0, false, a.get(), nullptr, stat.get() + ST_BAD_BYTE, st_, fused_hist ? hist.get() : nullptr, &sp);
            uint64_t produced = nk;
            if (!mutate_with_words(a.get(), b.get(), nk, mode, fused_hist ? hist.get() : nullptr, &stat)) {
                // non-ACGT bytes: nothing was applied; the sanitised reads take the plain path
                Sanitized sn;
                sanitize_batch(dp.batch, sn, st_);
                produced = run_seq_words(sn.dp.batch, sn.n_kmers, 0, false, a.get(), nullptr, st_);
                mutate_with_words(a.get(), b.get(), produced, mode);
            }
            last_produced += produced;
            tr.mark("batch", true);
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
    // Host buffers, software-pipelined (src/cbl.rs:328-339 for a whole record loop, examples/cbl.rs:160-163): a batch of
    // <= batch_kmers_ k-mers is cut into groups of ~CBL_INS_GROUP_BYTES; while the words kernel of group g runs, the reads
    // of group g + 1 are on their way in (copy stream, two staging slots), all groups write into the batch's one sort
    // buffer and fold their digits into its one histogram; then ONE sort + ONE merge per batch.
    struct InSlot {
        DevBuf<uint8_t> seq, blob;
        HostStage stage;
        cudaEvent_t copied = nullptr, consumed = nullptr;
        bool used = false;
        ~InSlot() {
            if (copied) cudaEventDestroy(copied);
            if (consumed) cudaEventDestroy(consumed);
        }
    };
    void insert_seqs(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, bool remove) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        const int mode = remove ? EDIT_DEL : EDIT_INS;
        last_produced = 0;
        const uint64_t group_bytes = std::max<uint64_t>(1 << 20, env_u64("CBL_INS_GROUP_BYTES", 64ull << 20));
        const uint32_t piece_kmers = (uint32_t)std::min<uint64_t>(PIECE_KMERS, std::max<uint64_t>(CHUNK_KMERS, (group_bytes / 4) / CHUNK_KMERS * CHUNK_KMERS));
        PieceList pl;
        build_pieces(offsets, 0, n_seqs, pl, piece_kmers);
        const size_t np = pl.kmers.size();
        cudaStream_t cs = side_[0];
        auto piece_end = [&](size_t i) { return pl.byte_off[i] + pl.kmers[i] + (uint64_t)cfg_.k - 1; };
        size_t p = 0;
        while (p < np) {
            size_t q = p;
            uint64_t nk = 0;
            while (q < np && (q == p || nk + pl.kmers[q] <= batch_kmers_)) { nk += pl.kmers[q]; q++; }
            DevBuf<W> a(nk, st_), b(nk, st_);
            const SortPlan sp = plan_sort(nk);
            DevBuf<unsigned long long> stat, hist;
            init_status(stat);
            const bool fused_hist = sp.n_pass <= SW_HIST_MAX && nk > 1 && env_u64("CBL_HIST_FUSED", 1) != 0;
            if (fused_hist) { hist.alloc((size_t)sp.n_pass * 256, st_); hist.zero(); }
            InSlot slots[2];
            for (auto& sl : slots) {
                CUDA_CHECK(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
                CUDA_CHECK(cudaEventCreateWithFlags(&sl.consumed, cudaEventDisableTiming));
            }
            size_t g0 = p;
            int gi = 0;
            while (g0 < q) {
                size_t g1 = g0;
                while (g1 < q && (g1 == g0 || piece_end(g1) - pl.byte_off[g0] <= group_bytes)) g1++;
                InSlot& sl = slots[gi & 1];
                if (sl.used) CUDA_CHECK(cudaEventSynchronize(sl.consumed));   // the words kernel that read this slot is done
                const uint64_t b0 = pl.byte_off[g0], nbytes = piece_end(g1 - 1) - b0;
                const size_t m = g1 - g0;
                const size_t o_out = m * 8, o_ch = 2 * m * 8, o_km = (3 * m + 1) * 8, blob_bytes = o_km + m * 4;
                sl.stage.reserve(blob_bytes);
                uint64_t* h_byte = reinterpret_cast<uint64_t*>(sl.stage.p);
                uint64_t* h_out = reinterpret_cast<uint64_t*>(sl.stage.p + o_out);
                uint64_t* h_ch = reinterpret_cast<uint64_t*>(sl.stage.p + o_ch);
                uint32_t* h_km = reinterpret_cast<uint32_t*>(sl.stage.p + o_km);
                for (size_t i = 0; i < m; i++) {
                    h_byte[i] = pl.byte_off[g0 + i] - b0;
                    h_out[i] = pl.out_off[g0 + i] - pl.out_off[p];
                    h_ch[i] = pl.chunk0[g0 + i] - pl.chunk0[g0];
                    h_km[i] = pl.kmers[g0 + i];
                }
                h_ch[m] = pl.chunk0[g1] - pl.chunk0[g0];
                if (!sl.used || sl.seq.size() < nbytes + 64) sl.seq.alloc(std::max<uint64_t>(nbytes, group_bytes) + 64, cs);
                sl.blob.alloc(blob_bytes, cs);
                CUDA_CHECK(cudaMemcpyAsync(sl.seq.get(), seq + b0, nbytes, cudaMemcpyHostToDevice, cs));
                CUDA_CHECK(cudaMemcpyAsync(sl.blob.get(), sl.stage.p, blob_bytes, cudaMemcpyHostToDevice, cs));
                CUDA_CHECK(cudaEventRecord(sl.copied, cs));
                SeqBatch sb;
                sb.seq = sl.seq.get(); sb.seq_end = sl.seq.get() + nbytes;
                sb.piece_byte = reinterpret_cast<const uint64_t*>(sl.blob.get());
                sb.piece_out = reinterpret_cast<const uint64_t*>(sl.blob.get() + o_out);
                sb.piece_chunk0 = reinterpret_cast<const uint64_t*>(sl.blob.get() + o_ch);
                sb.piece_kmers = reinterpret_cast<const uint32_t*>(sl.blob.get() + o_km);
                sb.n_pieces = (uint32_t)m; sb.n_chunks = h_ch[m];
                CUDA_CHECK(cudaStreamWaitEvent(st_, sl.copied, 0));
                launch_seq_words(sb, 0, false, a.get(), nullptr, stat.get() + ST_BAD_BYTE, st_, fused_hist ? hist.get() : nullptr, &sp);
                CUDA_CHECK(cudaEventRecord(sl.consumed, st_));
                CUDA_CHECK(cudaStreamWaitEvent(cs, sl.consumed, 0));          // the slot's buffers go back to the copy stream's arena list
                sl.used = true;
                g0 = g1;
                gi++;
            }
            uint64_t produced = nk;
            const bool ok = mutate_with_words(a.get(), b.get(), nk, mode, fused_hist ? hist.get() : nullptr, &stat);
            CUDA_CHECK(cudaStreamSynchronize(cs));
            if (!ok) {
                // non-ACGT bytes: nothing was applied; the batch's bytes go in at once and take the sanitiser
                const uint64_t b0 = pl.byte_off[p], nbytes = piece_end(q - 1) - b0;
                DevBuf<uint8_t> d(nbytes + 64, st_);
                CUDA_CHECK(cudaMemcpyAsync(d.get(), seq + b0, nbytes, cudaMemcpyHostToDevice, st_));
                PieceList sub;
                sub.chunk0.push_back(0);
                for (size_t i = p; i < q; i++) {
                    sub.byte_off.push_back(pl.byte_off[i] - b0);
                    sub.out_off.push_back(pl.out_off[i] - pl.out_off[p]);
                    sub.kmers.push_back(pl.kmers[i]);
                    sub.n_chunks += div_up(pl.kmers[i], CHUNK_KMERS);
                    sub.chunk0.push_back(sub.n_chunks);
                }
                sub.n_kmers = nk;
                DevPieces dp;
                upload_pieces(sub, 0, q - p, 0, d.get(), nbytes, dp, st_);
                Sanitized sn;
                sanitize_batch(dp.batch, sn, st_);
                produced = run_seq_words(sn.dp.batch, sn.n_kmers, 0, false, a.get(), nullptr, st_);
                mutate_with_words(a.get(), b.get(), produced, mode);
            }
            last_produced += produced;
            p = q;
        }
        CUDA_CHECK(cudaStreamSynchronize(st_));
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
    // membership of the words of n_seg device segments, answers to seg_out[i] (may be peer memory): one launch
    void words_contains_segments_dev(const void* const* seg, const uint64_t* seg_n, uint8_t* const* seg_out, uint32_t n_seg) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        ensure_sub();
        launch_probe_segments(seg, seg_n, seg_out, n_seg, st_);
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
    void words_op(int op, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        if (n == 0) return;
        std::vector<W> h(n);
        const int key_bits = P_.bits + P_.pos_bits;
        for (size_t i = 0; i < n; i++) {
            h[i] = sizeof(W) == 16 ? (W)(((u128)(hi ? hi[i] : 0) << 64) | lo[i]) : (W)lo[i];
            // a mutation must not be fed anything that is not the word of a k-mer: the merge uses the all-ones word as its
            // exhausted-side sentinel and the directory is sized for PREFIX_BITS-bit prefixes (a membership test just answers no)
            const bool too_wide = key_bits < (int)(8 * sizeof(W)) && (h[i] >> key_bits) != 0;
            if (op != 0 && (too_wide || h[i] == ~(W)0))
                throw Error(CBL_EINVAL, "word " + std::to_string(i) + " is not the word of a k-mer (wider than 2K + POS_BITS bits, or all ones)");
        }
        DevBuf<W> d(n, st_);
        DevBuf<uint8_t> flags(n, st_);
        CUDA_CHECK(cudaMemcpyAsync(d.get(), h.data(), n * sizeof(W), cudaMemcpyHostToDevice, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        words_op_dev(op, d.get(), n, out ? flags.get() : nullptr);
        if (out) CUDA_CHECK(cudaMemcpyAsync(out, flags.get(), n, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
    }
    void load_sorted_words(const uint64_t* lo, const uint64_t* hi, uint64_t n) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        if (n == 0) return;
        std::vector<W> h(n);
        const int key_bits = P_.bits + P_.pos_bits;
        for (uint64_t i = 0; i < n; i++) {
            h[i] = sizeof(W) == 16 ? (W)(((u128)(hi ? hi[i] : 0) << 64) | lo[i]) : (W)lo[i];
            // foreign words (a file): a word wider than the key, or the all-ones word (the merge's exhausted-side sentinel;
            // no k-mer maps to it: the position of an all-ones necklace is 0), cannot come from a k-mer
            const bool too_wide = key_bits < (int)(8 * sizeof(W)) && (h[i] >> key_bits) != 0;
            if (too_wide || h[i] == ~(W)0) throw Error(CBL_EINVAL, "word " + std::to_string(i) + " is not the word of a k-mer (wider than 2K + POS_BITS bits, or all ones)");
        }
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
    // The fused sharded query of one rank (shard_query.cuh): ONE kernel whose warps alternate between producing (encode +
    // necklace + route of this rank's reads, words stored straight into their owners' receive regions) and consuming
    // (probing the blocks of words the peers have completed in this rank's receive buffer, answers stored straight into
    // the asking rank's answer buffer).  Returns when the kernel is done, i.e. when this rank has answered every block sent
    // to it; counts[d] = words sent to owner d (> cap: overflow, the caller retries with larger regions).
    void seq_contains_fused_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, const FusedQuery& q,
                                uint64_t* counts) override {
        CUDA_CHECK(cudaSetDevice(cfg_.device));
        check_records(offsets, n_seqs);
        const uint32_t g = q.n_split + 1;
        if (g > (uint32_t)SQ_MAX_RANKS) throw Error(CBL_EINVAL, "too many ranks");
        if ((uint64_t)g * q.cap >= (1ull << 32)) throw Error(CBL_EINVAL, "fused query: (ranks x region capacity) must stay below 2^32 words");
        if (q.cap % SQ_BLOCK) throw Error(CBL_EINVAL, "fused query: region capacity must be a multiple of " + std::to_string(SQ_BLOCK) + " words");
        if (q.epoch == 0 || q.epoch > 65535) throw Error(CBL_EINVAL, "fused query: epoch must be in 1 .. 65535");
        for (uint32_t i = 0; i < g; i++) counts[i] = 0;
        ensure_sub();
        PieceList pl;
        build_pieces(offsets, 0, n_seqs, pl);
        if (2 * pl.n_chunks >= (1ull << 32)) throw Error(CBL_EINVAL, "fused query: too many chunks in one call");
        DevPieces dp;
        if (!pl.kmers.empty()) upload_pieces(pl, 0, pl.kmers.size(), 0, d_seq, n_bytes, dp, st_);
        // [0, 16) words reserved per owner, [16] their sum, [17] smallest offending byte offset, [18] time-out flag, [19] / [20] globaltimer at kernel start / end of the production, [21] task counter | block ticket (2 x u32)
        DevBuf<unsigned long long> ctl(22, st_);
        CUDA_CHECK(cudaMemsetAsync(ctl.get(), 0, 22 * 8, st_));
        CUDA_CHECK(cudaMemsetAsync(ctl.get() + 17, 0xFF, 8, st_));
        ShardQueryArgs<W> a{};
        for (int i = 0; i < ROUTE_MAX_SPLIT; i++) a.split[i] = (uint32_t)i < q.n_split ? q.splitters[i] : 0xFFFFFFFFu;
        a.suffix_bits = P_.suffix_bits;
        for (uint32_t i = 0; i < (uint32_t)SQ_MAX_RANKS; i++) {
            a.peer[i] = i < g ? (W*)q.peer_region[i] : nullptr;
            a.peer_final[i] = i < g ? q.peer_final[i] : nullptr;
            a.seg_words[i] = i < g ? (W*)q.recv_region[i] : nullptr;
            a.seg_out[i] = i < g ? q.answer_region[i] : nullptr;
            a.final_[i] = i < g ? q.final_[i] : nullptr;
        }
        a.cnt = ctl.get();
        a.err = ctl.get() + 17;
        a.prod_next = reinterpret_cast<unsigned*>(ctl.get() + 21);
        a.ticket = a.prod_next + 1;
        a.epoch = q.epoch;
        a.n_kmers = pl.n_kmers;
        a.pos = q.d_pos;
        a.cap = q.cap;
        a.n_tasks = (uint32_t)(2 * pl.n_chunks);
        a.max_blocks = (uint32_t)(q.cap / SQ_BLOCK);
        a.g = (int)g;
        a.dev_flags = (int)env_u64("CBL_SQ_FLAGS", 0);
        static int occ = 0;
        if (!occ) CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, shard_query_kernel<W, Suf, CBL_PROBE_WB, false>, SQ_THREADS, 0));
        int sms = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg_.device));
        // persistent grid: every resident warp keeps taking produce tasks and block tickets until both run out
        // (ranks that share a device split its CTA slots: the kernels wait for each other's words, so all must be resident)
        const uint64_t per_sm = std::min<uint64_t>(env_u64("CBL_SQ_CTAS", (uint64_t)std::max(occ, 1)), 32);
        const unsigned grid = (unsigned)std::max<uint64_t>((uint64_t)sms * per_sm / std::max<uint32_t>(q.grid_share, 1), 1);
        SeqBatch sb = pl.kmers.empty() ? SeqBatch{} : dp.batch;
        if (P_.canonical) CBL_LAUNCH((shard_query_kernel<W, Suf, CBL_PROBE_WB, true>), grid, SQ_THREADS, 0, st_, sb, P_, view(), a);
        else CBL_LAUNCH((shard_query_kernel<W, Suf, CBL_PROBE_WB, false>), grid, SQ_THREADS, 0, st_, sb, P_, view(), a);
        const auto t_host = std::chrono::steady_clock::now();
        CUDA_CHECK(cudaMemcpyAsync(h_status_, ctl.get(), 21 * 8, cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        if (getenv("CBL_SHARD_TRACE") != nullptr) {
            unsigned long long sent = 0;
            for (uint32_t i = 0; i < g; i++) sent += h_status_[i];
            fprintf(stderr, "[fused query] device %d: production of %llu words complete %.2f ms after the kernel started; launch -> done %.2f ms (host)\n",
                    cfg_.device, sent, (double)(h_status_[20] - h_status_[19]) * 1e-6,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host).count());
        }
        if (h_status_[18] != 0) throw Error(CBL_ECUDA, "fused query: timed out waiting for the words of another rank (is every rank of the group in the call?)");
        if (h_status_[17] != ULLONG_MAX) throw_bad_byte(h_status_[17]);
        for (uint32_t i = 0; i < g; i++) counts[i] = h_status_[i];
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
    // A op= B as ONE streaming merge of the two CSR indexes (operand B is never expanded into words); synchronous.
    void setop_new_state(int op, Index* o, NewState& ns) {
        o->sync();
        ns.changed = false;
        if (o->n_ == 0 && op != SETOP_AND) return;
        DevBuf<unsigned long long> stat;
        init_status(stat);
        merge_new_state<true>(MergeB<W, Suf, true>{o->view()}, o->n_, op, ns, stat.get());
        if (ns.pending) { read_status(stat); finish_new_state(ns); }
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
