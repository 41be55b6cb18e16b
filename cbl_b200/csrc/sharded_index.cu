// One CBL set prefix-sharded over several GPUs of ONE process, behind the same IIndex interface (and therefore the same
// C ABI) as a single-GPU set: cbl_create_sharded(k, word_bits, prefix_bits, canonical, n_gpus, devices, &h) and every
// host-buffer entry point of include/cbl_gpu.h works on the handle; the *_dev entry points (raw device pointers of one
// GPU) do not apply and return CBL_EINVAL.
//
// Design (north star item 4):
//   * the 2^PREFIX_BITS prefix space is cut into n_gpus contiguous ranges by sample-based, cost-weighted splitters
//     (necklace prefixes are extremely skewed); shard d = a complete single-GPU Index on device d;
//   * a batch of records is cut into n_gpus contiguous record groups; one host thread per device copies its group in
//     and runs the fused encode + necklace + route kernel, which stores every word STRAIGHT INTO THE OWNER'S HBM
//     (cudaDeviceEnablePeerAccess: the stores travel over NVLink; no send buffer, no collective library);
//   * the owner merges (insert / remove) or probes (contains) what landed in its receive buffer, answers are stored
//     straight back into the asking GPU's answer buffer and gathered into read order there;
//   * set operations run shard by shard (both operands must share devices and splitters); iteration / export /
//     serialisation concatenate the shards in device order, which IS ascending word order (the ranges are ascending).
// The process-per-GPU variant of the same design (torch.distributed, CUDA IPC mappings) is cbl_b200/sharded.py.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "cbl_index.cuh"
#include "splitters.hpp"

namespace cbl {

namespace {

// run f(d) for d in [0, g) on g host threads (one per device); the first exception is rethrown on the caller
template <class F> void parallel(int g, F&& f) {
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err(g);
    for (int d = 0; d < g; d++)
        th.emplace_back([&, d] {
            try { f(d); } catch (...) { err[d] = std::current_exception(); }
        });
    for (auto& t : th) t.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
}

uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

class ShardedIndex final : public IIndex {
    Config cfg_;
    KParams P_;
    std::vector<int> dev_;
    std::vector<std::unique_ptr<IIndex>> sh_;
    std::vector<uint32_t> split_;
    int g_ = 0, wb_ = 8;
    // exchange buffers of device d: recv = g regions of cap words (region s written by device s), back = g regions of cap
    // answer bytes (region o written by owner o)
    std::vector<void*> recv_;
    std::vector<uint8_t*> back_;
    uint64_t cap_ = 0;
    // fused query (shard_query.cuh): final-count slots of device d (slot s written by device s), call counter, and whether
    // the receive buffers may hold anything else than the 0xFF "no word here" pattern
    std::vector<unsigned long long*> fin_;
    uint32_t epoch_ = 0;
    bool recv_dirty_ = true;

    static void not_for_sharded(const char* what) {
        throw Error(CBL_EINVAL, std::string(what) + " takes device pointers of ONE GPU and is not available on a sharded handle (use the host-buffer entry points)");
    }
    uint32_t owner_of(uint32_t prefix) const { return (uint32_t)(std::upper_bound(split_.begin(), split_.end(), prefix) - split_.begin()); }
    uint32_t prefix_of(uint64_t lo, uint64_t hi) const {
        const u128 w = ((u128)hi << 64) | lo;
        return (uint32_t)(w >> P_.suffix_bits);
    }
    void free_exchange() {
        for (int d = 0; d < g_; d++) {
            cudaSetDevice(dev_[d]);
            if (d < (int)recv_.size() && recv_[d]) cudaFree(recv_[d]);
            if (d < (int)back_.size() && back_[d]) cudaFree(back_[d]);
            if (d < (int)fin_.size() && fin_[d]) cudaFree(fin_[d]);
        }
        recv_.assign(g_, nullptr);
        back_.assign(g_, nullptr);
        fin_.assign(g_, nullptr);
        cap_ = 0;
        recv_dirty_ = true;
    }
    void ensure_cap(uint64_t cap_words) {
        if (cap_words <= cap_) return;
        for (auto& s : sh_) s->sync();
        free_exchange();
        const uint64_t cap = (cap_words + 2047) / 2048 * 2048;
        if ((uint64_t)g_ * cap >= (1ull << 32)) throw Error(CBL_EINVAL, "batch too large for one exchange: pass the records in smaller batches");
        for (int d = 0; d < g_; d++) {
            CUDA_CHECK(cudaSetDevice(dev_[d]));
            CUDA_CHECK(cudaMalloc(&recv_[d], (size_t)g_ * cap * wb_));
            CUDA_CHECK(cudaMalloc((void**)&back_[d], (size_t)g_ * cap));
            CUDA_CHECK(cudaMalloc((void**)&fin_[d], 16 * sizeof(unsigned long long)));
            CUDA_CHECK(cudaMemset(fin_[d], 0, 16 * sizeof(unsigned long long)));
        }
        cap_ = cap;
        recv_dirty_ = true;
    }
    // the fused query wants 0xFF wherever the receive buffers hold no word; it leaves them that way itself
    void clean_recv() {
        if (!recv_dirty_) return;
        parallel(g_, [&](int d) {
            CUDA_CHECK(cudaSetDevice(dev_[d]));
            CUDA_CHECK(cudaMemsetAsync(recv_[d], 0xFF, (size_t)g_ * cap_ * wb_, sh_[d]->stream()));
            CUDA_CHECK(cudaStreamSynchronize(sh_[d]->stream()));
        });
        recv_dirty_ = false;
    }
    void* recv_region(int owner, int src) const { return (uint8_t*)recv_[owner] + ((size_t)src * cap_) * wb_; }
    uint8_t* back_region(int src, int owner) const { return back_[src] + (size_t)owner * cap_; }

    // contiguous record groups with about the same number of bytes: group d = records [cut[d], cut[d + 1])
    std::vector<size_t> cut_records(const uint64_t* offsets, size_t n_seqs) const {
        std::vector<size_t> cut(g_ + 1, n_seqs);
        cut[0] = 0;
        const uint64_t total = offsets[n_seqs] - offsets[0];
        size_t r = 0;
        for (int d = 1; d < g_; d++) {
            const uint64_t target = offsets[0] + total * d / g_;
            while (r < n_seqs && offsets[r] < target) r++;
            cut[d] = r;
        }
        for (int d = 1; d <= g_; d++) cut[d] = std::max(cut[d], cut[d - 1]);
        return cut;
    }
    void check_records(const uint64_t* offsets, size_t n_seqs) const {
        for (size_t i = 0; i < n_seqs; i++) {
            if (offsets[i + 1] < offsets[i]) throw Error(CBL_EINVAL, "record offsets must be non-decreasing");
            const uint64_t len = offsets[i + 1] - offsets[i];
            if (len < (uint64_t)cfg_.k)  // src/cbl.rs:294-299,329-334
                throw Error(CBL_EINVAL, "Sequence size (" + std::to_string(len) + ") is smaller than K (" + std::to_string(cfg_.k) + ")");
        }
    }
    struct Group {   // one device's share of a batch, resident on that device
        void* d_seq = nullptr;
        uint64_t n_bytes = 0, n_kmers = 0, kmer0 = 0;
        std::vector<uint64_t> off;
        uint32_t* d_pos = nullptr;
        std::vector<uint64_t> counts;
    };
    static bool is_bad_byte(const Error& e) { return e.code == CBL_EINVAL && std::string(e.what()).find("non-ACGT") != std::string::npos; }

    // cut the batch into one group of records per device and copy every group to its device
    void prepare_groups(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, bool want_pos, std::vector<Group>& gr) {
        const std::vector<size_t> cut = cut_records(offsets, n_seqs);
        gr.assign(g_, Group());
        uint64_t k0 = 0, n_max = 0;
        for (int d = 0; d < g_; d++) {
            Group& G = gr[d];
            G.off.resize(cut[d + 1] - cut[d] + 1);
            for (size_t i = cut[d]; i <= cut[d + 1]; i++) G.off[i - cut[d]] = offsets[i] - offsets[cut[d]];
            G.n_bytes = G.off.back();
            for (size_t i = cut[d]; i < cut[d + 1]; i++) G.n_kmers += offsets[i + 1] - offsets[i] - (uint64_t)cfg_.k + 1;
            G.kmer0 = k0;
            k0 += G.n_kmers;
            n_max = std::max(n_max, G.n_kmers);
            G.counts.assign(g_, 0);
        }
        ensure_cap((uint64_t)((double)n_max / g_ * 1.3) + 4096);
        parallel(g_, [&](int d) {   // copy in (once; the buffers survive a retry with larger regions)
            Group& G = gr[d];
            CUDA_CHECK(cudaSetDevice(dev_[d]));
            cudaStream_t st = sh_[d]->stream();
            G.d_seq = arena::alloc(G.n_bytes + 64, st);
            if (want_pos) G.d_pos = (uint32_t*)arena::alloc(std::max<uint64_t>(G.n_kmers, 1) * 4, st);
            if (G.n_bytes) CUDA_CHECK(cudaMemcpyAsync(G.d_seq, seq + offsets[cut[d]], G.n_bytes, cudaMemcpyHostToDevice, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
        });
    }
    // route every group's words to their owners; returns false if the reads hold non-ACGT bytes (nothing useful was routed)
    bool route_all(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, bool want_pos, std::vector<Group>& gr) {
        prepare_groups(seq, offsets, n_seqs, want_pos, gr);
        std::atomic<bool> bad{false};
        for (;;) {
            recv_dirty_ = true;
            parallel(g_, [&](int d) {
                Group& G = gr[d];
                if (G.off.size() <= 1) return;
                std::vector<void*> mine(g_);
                for (int o = 0; o < g_; o++) mine[o] = recv_region(o, d);
                try {
                    sh_[d]->seq_route_dev((const uint8_t*)G.d_seq, G.n_bytes, G.off.data(), G.off.size() - 1, split_.data(), (uint32_t)split_.size(), mine.data(),
                                          cap_, G.d_pos, G.counts.data());
                } catch (const Error& e) {
                    if (!is_bad_byte(e)) throw;
                    bad = true;
                }
            });
            if (bad) return false;
            uint64_t mx = 0;
            for (auto& G : gr) for (uint64_t c : G.counts) mx = std::max(mx, c);
            if (mx <= cap_) return true;
            ensure_cap((uint64_t)(mx * 1.1) + 4096);   // a region overflowed (nothing past cap was written): everybody again
        }
    }
    // contains_seq of every group as the fused query (one kernel per device, running together: the words travel to their
    // owners and the answers back over NVLink while the kernels run); returns false if the reads hold non-ACGT bytes
    bool fused_query_all(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, std::vector<Group>& gr) {
        prepare_groups(seq, offsets, n_seqs, true, gr);
        std::atomic<bool> bad{false};
        for (;;) {
            clean_recv();
            epoch_ = epoch_ % 65535 + 1;
            recv_dirty_ = true;   // until the call has come back clean
            parallel(g_, [&](int d) {
                Group& G = gr[d];
                std::vector<void*> mine(g_), theirs(g_);
                std::vector<unsigned long long*> my_final(g_);
                std::vector<const unsigned long long*> their_final(g_);
                std::vector<uint8_t*> ans(g_);
                for (int o = 0; o < g_; o++) {
                    mine[o] = recv_region(o, d);          // my region at owner o
                    my_final[o] = fin_[o] + d;            // my slot at owner o
                    theirs[o] = recv_region(d, o);        // the region source o writes here
                    their_final[o] = fin_[d] + o;
                    ans[o] = back_region(o, d);           // my region in source o's answer buffer
                }
                FusedQuery q;
                q.splitters = split_.data(); q.n_split = (uint32_t)split_.size();
                q.peer_region = mine.data(); q.peer_final = my_final.data();
                q.cap = cap_; q.d_pos = G.d_pos;
                q.recv_region = theirs.data(); q.answer_region = ans.data(); q.final_ = their_final.data();
                q.epoch = epoch_;
                q.grid_share = (uint32_t)std::count(dev_.begin(), dev_.end(), dev_[d]);
                try {
                    sh_[d]->seq_contains_fused_dev((const uint8_t*)G.d_seq, G.n_bytes, G.off.data(), G.off.size() - 1, q, G.counts.data());
                } catch (const Error& e) {
                    if (!is_bad_byte(e)) throw;
                    bad = true;   // the kernel has run to its end all the same: the other devices are not left waiting
                }
            });
            if (bad) return false;
            uint64_t mx = 0;
            for (auto& G : gr) for (uint64_t c : G.counts) mx = std::max(mx, c);
            if (mx <= cap_) { recv_dirty_ = false; return true; }
            ensure_cap((uint64_t)(mx * 1.1) + 4096);   // a region overflowed (nothing past cap was written): everybody again
        }
    }
    void free_groups(std::vector<Group>& gr) {
        for (int d = 0; d < (int)gr.size(); d++) {
            cudaSetDevice(dev_[d]);
            if (gr[d].d_seq) arena::release(gr[d].d_seq, sh_[d]->stream());
            if (gr[d].d_pos) arena::release(gr[d].d_pos, sh_[d]->stream());
            gr[d].d_seq = nullptr;
            gr[d].d_pos = nullptr;
        }
    }
    // Reads with non-nucleotide bytes: the words come from the single-GPU path of shard 0, which reproduces the
    // reference's dropping behaviour, and are handed to their owners through the host (slow path, exact).
    uint64_t words_via_host(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, std::vector<uint64_t>& lo, std::vector<uint64_t>& hi) {
        uint64_t nk = 0;
        for (size_t i = 0; i < n_seqs; i++) nk += offsets[i + 1] - offsets[i] - (uint64_t)cfg_.k + 1;
        lo.assign(nk ? nk : 1, 0);
        hi.assign(nk ? nk : 1, 0);
        sh_[0]->seq_words(seq, offsets, n_seqs, lo.data(), hi.data(), false);
        const uint64_t produced = sh_[0]->last_produced;
        lo.resize(produced);
        hi.resize(produced);
        return produced;
    }

public:
    ShardedIndex(const Config& cfg, const int* devices, int n_gpus, const uint32_t* splitters) : cfg_(cfg) {
        if (n_gpus < 1 || n_gpus > 16) throw Error(CBL_EINVAL, "n_gpus must be in 1..=16");
        if (!devices) throw Error(CBL_EINVAL, "null pointer: devices");
        g_ = n_gpus;
        dev_.assign(devices, devices + n_gpus);
        cfg_.device = dev_[0];
        for (int d = 0; d < g_; d++) {
            Config c = cfg;
            c.device = dev_[d];
            sh_.emplace_back(make_index(c));
            sh_.back()->set_sort_concentration((double)g_);   // a shard's words cover 1 / g of the prefix mass
        }
        P_ = sh_[0]->params();
        wb_ = (P_.bits + P_.pos_bits) <= 64 ? 8 : 16;
        recv_.assign(g_, nullptr);
        back_.assign(g_, nullptr);
        // every device maps every other device's memory (the route / probe kernels store into it over NVLink)
        for (int a = 0; a < g_; a++)
            for (int b = 0; b < g_; b++) {
                if (dev_[a] == dev_[b]) continue;
                int can = 0;
                CUDA_CHECK(cudaDeviceCanAccessPeer(&can, dev_[a], dev_[b]));
                if (!can) throw Error(CBL_ECUDA, "GPUs " + std::to_string(dev_[a]) + " and " + std::to_string(dev_[b]) + " cannot access each other's memory");
                CUDA_CHECK(cudaSetDevice(dev_[a]));
                cudaError_t e = cudaDeviceEnablePeerAccess(dev_[b], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else CUDA_CHECK(e);
            }
        if (splitters) split_.assign(splitters, splitters + (g_ - 1));
        else if (g_ > 1) {
            // a fixed-seed sample of uniform random DNA pushed through the real necklace kernel (shard 0)
            const size_t nb = 2'000'000;
            std::vector<uint8_t> s(nb);
            uint64_t st = 20240229;
            for (size_t i = 0; i < nb; i += 32) {
                uint64_t r = splitmix(st);
                for (size_t j = i; j < std::min(nb, i + 32); j++, r >>= 2) s[j] = "ACTG"[r & 3];
            }
            const uint64_t off[2] = {0, nb};
            const size_t nk = nb - cfg.k + 1;
            std::vector<uint64_t> lo(nk), hi(nk);
            sh_[0]->seq_words(s.data(), off, 1, lo.data(), hi.data(), false);
            std::vector<uint32_t> pre(nk);
            for (size_t i = 0; i < nk; i++) pre[i] = prefix_of(lo[i], hi[i]);
            split_ = equal_cost_splitters(std::move(pre), g_);
        }
        for (size_t i = 1; i < split_.size(); i++)
            if (split_[i] <= split_[i - 1]) throw Error(CBL_EINVAL, "splitters must be strictly increasing");
    }
    ~ShardedIndex() override {
        for (auto& s : sh_) if (s) { try { s->sync(); } catch (...) {} }
        free_exchange();
    }
    const std::vector<uint32_t>& splitters() const { return split_; }
    const std::vector<int>& devices() const { return dev_; }

    const Config& config() const override { return cfg_; }
    const KParams& params() const override { return P_; }
    cudaStream_t stream() const override { return sh_[0]->stream(); }
    uint64_t count() const override {
        uint64_t n = 0;
        for (auto& s : sh_) n += s->count();
        return n;
    }
    uint32_t n_buckets() const override {
        uint64_t n = 0;
        for (auto& s : sh_) n += s->n_buckets();
        return (uint32_t)n;
    }
    // prefixes.count() == 0 with the last-bit quirk of RankBV::count_ones: the all-ones prefix lives in the last shard
    bool is_empty_reference_semantics() const override {
        for (int d = 0; d + 1 < g_; d++) if (sh_[d]->n_buckets() != 0) return false;
        return sh_[g_ - 1]->is_empty_reference_semantics();
    }
    void sync() override { for (auto& s : sh_) s->sync(); }
    void set_sort_concentration(double factor) override { for (auto& s : sh_) s->set_sort_concentration(factor * g_); }
    IIndex* new_empty(int canonical = -1) override {
        Config c = cfg_;
        if (canonical >= 0) c.canonical = canonical;
        return new ShardedIndex(c, dev_.data(), g_, g_ > 1 ? split_.data() : nullptr);
    }
    IIndex* clone() override {
        std::unique_ptr<ShardedIndex> c(static_cast<ShardedIndex*>(new_empty()));
        parallel(g_, [&](int d) { c->sh_[d].reset(sh_[d]->clone()); });
        return c.release();
    }

    // ---- sequences (host buffers) ------------------------------------------------------------------------------------
    void insert_seqs(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, bool remove) override {
        check_records(offsets, n_seqs);
        last_produced = 0;
        if (n_seqs == 0) return;
        std::vector<Group> gr;
        struct Free { ShardedIndex* s; std::vector<Group>* g; ~Free() { s->free_groups(*g); } } fr{this, &gr};
        const int op = remove ? 2 : 1;
        if (route_all(seq, offsets, n_seqs, false, gr)) {
            parallel(g_, [&](int o) {   // owner o: ONE batch out of the g regions of its receive buffer
                std::vector<const void*> seg(g_);
                std::vector<uint64_t> n(g_);
                uint64_t tot = 0;
                for (int s = 0; s < g_; s++) { seg[s] = recv_region(o, s); n[s] = gr[s].counts[o]; tot += n[s]; }
                if (tot) sh_[o]->words_op_segments_dev(op, seg.data(), n.data(), (uint32_t)g_);
                sh_[o]->sync();
            });
            for (auto& G : gr) last_produced += G.n_kmers;
            return;
        }
        std::vector<uint64_t> lo, hi;
        last_produced = words_via_host(seq, offsets, n_seqs, lo, hi);
        words_op(op, lo.data(), hi.data(), lo.size(), nullptr);
    }
    void contains_seqs(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, uint8_t* out) override {
        check_records(offsets, n_seqs);
        last_produced = 0;
        if (n_seqs == 0) return;
        std::vector<Group> gr;
        struct Free { ShardedIndex* s; std::vector<Group>* g; ~Free() { s->free_groups(*g); } } fr{this, &gr};
        if (fused_query_all(seq, offsets, n_seqs, gr)) {
            parallel(g_, [&](int d) {   // every answer has landed: into read order, out to the host
                Group& G = gr[d];
                if (!G.n_kmers) return;
                CUDA_CHECK(cudaSetDevice(dev_[d]));
                cudaStream_t st = sh_[d]->stream();
                uint8_t* d_ans = (uint8_t*)arena::alloc(G.n_kmers, st);
                struct F { uint8_t* p; cudaStream_t s; ~F() { arena::release(p, s); } } f{d_ans, st};
                sh_[d]->gather_u8_dev(back_[d], G.d_pos, G.n_kmers, d_ans);
                CUDA_CHECK(cudaMemcpyAsync(out + G.kmer0, d_ans, G.n_kmers, cudaMemcpyDeviceToHost, st));
                CUDA_CHECK(cudaStreamSynchronize(st));
            });
            for (auto& G : gr) last_produced += G.n_kmers;
            return;
        }
        std::vector<uint64_t> lo, hi;
        last_produced = words_via_host(seq, offsets, n_seqs, lo, hi);
        words_op(0, lo.data(), hi.data(), lo.size(), out);
    }
    void seq_words(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, uint64_t* lo, uint64_t* hi, bool brute) override {
        sh_[0]->seq_words(seq, offsets, n_seqs, lo, hi, brute);
        last_produced = sh_[0]->last_produced;
    }

    // ---- words / k-mers on the host: handed to their owners -----------------------------------------------------------
    void words_op(int op, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) override {
        std::vector<std::vector<uint64_t>> plo(g_), phi(g_);
        std::vector<std::vector<size_t>> idx(g_);
        for (size_t i = 0; i < n; i++) {
            const uint32_t o = owner_of(prefix_of(lo[i], hi ? hi[i] : 0));
            plo[o].push_back(lo[i]);
            phi[o].push_back(hi ? hi[i] : 0);
            idx[o].push_back(i);
        }
        parallel(g_, [&](int o) {
            if (plo[o].empty()) return;
            std::vector<uint8_t> f(plo[o].size());
            sh_[o]->words_op(op, plo[o].data(), phi[o].data(), plo[o].size(), out ? f.data() : nullptr);
            if (out) for (size_t j = 0; j < f.size(); j++) out[idx[o][j]] = f[j];
        });
    }
    void kmers_op(int op, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) override {
        std::vector<uint64_t> wlo(n), whi(n);
        const u128 mask = low_mask<u128>(P_.bits);
        for (size_t i = 0; i < n; i++) {   // src/cbl.rs:199-206 on the host (the same __host__ __device__ code the kernels run)
            const u128 x = ((((u128)(hi ? hi[i] : 0)) << 64) | lo[i]) & mask;
            const u128 w = kmer_to_word<u128>(x, P_);
            wlo[i] = (uint64_t)w;
            whi[i] = (uint64_t)(w >> 64);
        }
        words_op(op, wlo.data(), whi.data(), n, out);
    }
    void load_sorted_words(const uint64_t* lo, const uint64_t* hi, uint64_t n) override {
        std::vector<std::vector<uint64_t>> plo(g_), phi(g_);
        for (uint64_t i = 0; i < n; i++) {
            const uint32_t o = owner_of(prefix_of(lo[i], hi ? hi[i] : 0));
            plo[o].push_back(lo[i]);
            phi[o].push_back(hi ? hi[i] : 0);
        }
        parallel(g_, [&](int o) { if (!plo[o].empty()) sh_[o]->load_sorted_words(plo[o].data(), phi[o].data(), plo[o].size()); });
    }

    // ---- set operations: shard by shard -------------------------------------------------------------------------------
    ShardedIndex* check_other(IIndex* other) const {
        auto* o = dynamic_cast<ShardedIndex*>(other);
        if (!o) throw Error(CBL_EINVAL, "set operation between a sharded and an unsharded index");
        const Config& c = o->cfg_;
        if (c.k != cfg_.k || c.prefix_bits != cfg_.prefix_bits || c.word_bits != cfg_.word_bits)
            throw Error(CBL_EINVAL, "set operation between indexes with different K / T / PREFIX_BITS");
        if (c.canonical != cfg_.canonical) throw Error(CBL_EINVAL, "One of the index is canonical while the other isn't");  // cbl.rs:422-425
        if (o->dev_ != dev_ || o->split_ != split_) throw Error(CBL_EINVAL, "set operation between indexes sharded differently (devices / splitters)");
        return o;
    }
    IIndex* setop(int op, IIndex* other) override {
        ShardedIndex* o = check_other(other);
        std::unique_ptr<ShardedIndex> r(static_cast<ShardedIndex*>(new_empty()));
        parallel(g_, [&](int d) { r->sh_[d].reset(sh_[d]->setop(op, o->sh_[d].get())); });
        return r.release();
    }
    void setop_assign(int op, IIndex* other) override {
        ShardedIndex* o = check_other(other);
        parallel(g_, [&](int d) { sh_[d]->setop_assign(op, o->sh_[d].get()); });
    }

    // ---- export / stats: shards in device order == ascending word order ----------------------------------------------------
    void export_words(uint64_t start, uint64_t cap, int to_kmers, uint64_t* lo, uint64_t* hi, uint64_t* n_out) override {
        uint64_t done = 0, base = 0;
        for (int d = 0; d < g_ && done < cap; d++) {
            const uint64_t nd = sh_[d]->count();
            if (start + done < base + nd) {
                uint64_t got = 0;
                sh_[d]->export_words(start + done - base, cap - done, to_kmers, lo + done, hi ? hi + done : nullptr, &got);
                done += got;
            }
            base += nd;
        }
        *n_out = done;
    }
    void bucket_sizes(uint32_t* prefixes, uint32_t* sizes, uint64_t cap, uint64_t* n_out) override {
        uint64_t nb = 0;
        for (auto& s : sh_) nb += s->n_buckets();
        *n_out = nb;
        if (!prefixes || !nb) return;
        if (cap < nb) throw Error(CBL_EINVAL, "output buffer too small");
        uint64_t at = 0;
        for (auto& s : sh_) {
            uint64_t got = 0;
            if (s->n_buckets()) s->bucket_sizes(prefixes + at, sizes + at, cap - at, &got);
            at += got;
        }
    }

    // ---- single-GPU device-pointer entry points ------------------------------------------------------------------------------
    void insert_seqs_dev(const uint8_t*, uint64_t, const uint64_t*, size_t) override { not_for_sharded("cbl_insert_seqs_dev"); }
    void remove_seqs_dev(const uint8_t*, uint64_t, const uint64_t*, size_t) override { not_for_sharded("cbl_remove_seqs_dev"); }
    void contains_seqs_dev(const uint8_t*, uint64_t, const uint64_t*, size_t, uint8_t*) override { not_for_sharded("cbl_contains_seqs_dev"); }
    void seq_words_dev(const uint8_t*, uint64_t, const uint64_t*, size_t, void*, bool) override { not_for_sharded("cbl_seq_words_dev"); }
    void words_op_dev(int, const void*, uint64_t, uint8_t*) override { not_for_sharded("cbl_words_op_dev"); }
    void words_op_segments_dev(int, const void* const*, const uint64_t*, uint32_t) override { not_for_sharded("cbl_words_op_segments_dev"); }
    void words_contains_segments_dev(const void* const*, const uint64_t*, uint8_t* const*, uint32_t) override { not_for_sharded("cbl_words_contains_segments_dev"); }
    void route_words_dev(const void*, uint64_t, const uint32_t*, uint32_t, void*, uint32_t*, uint64_t*) override { not_for_sharded("cbl_route_words_dev"); }
    void gather_u8_dev(const uint8_t*, const uint32_t*, uint64_t, uint8_t*) override { not_for_sharded("cbl_gather_u8_dev"); }
    void seq_contains_fused_dev(const uint8_t*, uint64_t, const uint64_t*, size_t, const FusedQuery&, uint64_t*) override { not_for_sharded("cbl_seq_contains_fused_dev"); }
    void route_counts_dev(const void*, uint64_t, const uint32_t*, uint32_t, uint64_t*) override { not_for_sharded("cbl_route_counts_dev"); }
    void route_scatter_dev(const void*, uint64_t, const uint32_t*, uint32_t, void* const*, const uint64_t*, const uint64_t*, uint32_t*) override { not_for_sharded("cbl_route_scatter_dev"); }
    void seq_route_dev(const uint8_t*, uint64_t, const uint64_t*, size_t, const uint32_t*, uint32_t, void* const*, uint64_t, uint32_t*, uint64_t*) override { not_for_sharded("cbl_seq_route_dev"); }
    void export_words_dev(uint64_t, uint64_t, int, void*) override { not_for_sharded("cbl_export_words_dev"); }
};

IIndex* make_sharded_index(const Config& cfg, const int* devices, int n_gpus, const uint32_t* splitters) {
    return new ShardedIndex(cfg, devices, n_gpus, splitters);
}
bool sharded_splitters(IIndex* ix, std::vector<uint32_t>& out) {
    auto* s = dynamic_cast<ShardedIndex*>(ix);
    if (!s) return false;
    out = s->splitters();
    return true;
}

}  // namespace cbl
