// CPU ORACLE — TEST INFRASTRUCTURE ONLY (see cbl_oracle.hpp).
// extern "C" surface over the restatement so tests/ and bench.py's cpu_baseline leg can drive it
// with ctypes.  Words and k-mers cross the boundary as (lo, hi) pairs of u64.
#include "cbl_oracle.hpp"

#include <chrono>
#include <cstdio>

using namespace orc;

namespace {

thread_local std::string g_err;

struct ICbl {
    virtual ~ICbl() {}
    virtual ICbl* clone() const = 0;
    virtual int k() const = 0;
    virtual int prefix_bits() const = 0;
    virtual bool canonical() const = 0;
    virtual size_t count() const = 0;
    virtual bool is_empty() const = 0;
    virtual void seq_words(const uint8_t* s, size_t n, std::vector<u128>& out) = 0;
    virtual void insert_seq(const uint8_t* s, size_t n) = 0;
    virtual void remove_seq(const uint8_t* s, size_t n) = 0;
    virtual void contains_seq(const uint8_t* s, size_t n, std::vector<uint8_t>& out) = 0;
    virtual bool contains_all(const uint8_t* s, size_t n) = 0;
    virtual bool insert(u128 kmer) = 0;
    virtual bool remove(u128 kmer) = 0;
    virtual bool contains(u128 kmer) const = 0;
    virtual u128 get_word(u128 kmer) const = 0;
    virtual u128 recover_kmer(u128 word) const = 0;
    virtual void iter_words(std::vector<u128>& out) const = 0;
    virtual ICbl* binary_op(int op, ICbl* other) = 0;
    virtual void assign_op(int op, ICbl* other) = 0;
    virtual ICbl* merge_many(std::vector<ICbl*>& v, bool intersect) = 0;
    virtual void serialize(std::vector<uint8_t>& out) const = 0;
    virtual ICbl* deserialize(const uint8_t* p, size_t n) const = 0;
    virtual void bucket_sizes(std::vector<std::pair<size_t, size_t>>& out) const = 0;
};

template <class T, int BYTES>
struct CblImpl final : ICbl {
    CBL<T, BYTES> c;
    CblImpl(int k, int p, bool canon) : c(k, p, canon) {}
    explicit CblImpl(CBL<T, BYTES>&& o) : c(std::move(o)) {}
    explicit CblImpl(const CBL<T, BYTES>& o) : c(o) {}
    ICbl* clone() const override { return new CblImpl(c); }
    int k() const override { return c.K; }
    int prefix_bits() const override { return c.PREFIX_BITS; }
    bool canonical() const override { return c.canonical; }
    size_t count() const override { return c.count(); }
    bool is_empty() const override { return c.is_empty(); }
    void seq_words(const uint8_t* s, size_t n, std::vector<u128>& out) override { c.seq_words(s, n, out); }
    void insert_seq(const uint8_t* s, size_t n) override { c.insert_seq(s, n); }
    void remove_seq(const uint8_t* s, size_t n) override { c.remove_seq(s, n); }
    void contains_seq(const uint8_t* s, size_t n, std::vector<uint8_t>& out) override { c.contains_seq(s, n, out); }
    bool contains_all(const uint8_t* s, size_t n) override { return c.contains_all(s, n); }
    bool insert(u128 kmer) override { return c.insert((T)kmer); }
    bool remove(u128 kmer) override { return c.remove((T)kmer); }
    bool contains(u128 kmer) const override { return c.contains((T)kmer); }
    u128 get_word(u128 kmer) const override { return (u128)c.get_word((T)kmer); }
    u128 recover_kmer(u128 word) const override { return (u128)c.recover_kmer((T)word); }
    void iter_words(std::vector<u128>& out) const override { c.iter_words(out); }
    static CblImpl* cast(ICbl* o) {
        auto* p = dynamic_cast<CblImpl*>(o);
        if (!p) throw std::invalid_argument("set operation between indexes of different types");
        return p;
    }
    void check_canon(ICbl* o) const {  // cbl.rs:422-425
        if (o->canonical() != c.canonical) throw std::invalid_argument("One of the index is canonical while the other isn't");
    }
    ICbl* binary_op(int op, ICbl* other) override {
        auto* o = cast(other);
        check_canon(o);
        auto ws = WordSet<BYTES>::binary_op((typename WordSet<BYTES>::Op)op, c.wordset, o->c.wordset);
        return new CblImpl(CBL<T, BYTES>(c.K, c.PREFIX_BITS, c.canonical, std::move(ws)));
    }
    void assign_op(int op, ICbl* other) override {
        auto* o = cast(other);
        check_canon(o);
        switch (op) {
            case 0: c.wordset.or_assign(o->c.wordset); break;
            case 1: c.wordset.and_assign(o->c.wordset); break;
            case 2: c.wordset.sub_assign(o->c.wordset); break;
            default: c.wordset.xor_assign(o->c.wordset); break;
        }
    }
    ICbl* merge_many(std::vector<ICbl*>& v, bool intersect) override {  // cbl.rs:108-124
        std::vector<WordSet<BYTES>*> ws;
        for (auto* x : v) { auto* o = cast(x); check_canon(o); ws.push_back(&o->c.wordset); }
        auto r = intersect ? WordSet<BYTES>::intersect(ws) : WordSet<BYTES>::merge(ws);
        return new CblImpl(CBL<T, BYTES>(c.K, c.PREFIX_BITS, c.canonical, std::move(r)));
    }
    void serialize(std::vector<uint8_t>& out) const override { ByteWriter w; c.serialize(w); out.swap(w.out); }
    ICbl* deserialize(const uint8_t* p, size_t n) const override {
        ByteReader r(p, n);
        return new CblImpl(CBL<T, BYTES>::deserialize(c.K, c.PREFIX_BITS, r));
    }
    void bucket_sizes(std::vector<std::pair<size_t, size_t>>& out) const override { out = c.wordset.buckets_sizes(); }
};

template <class T, int B> ICbl* make_b(int k, int p, bool canon, int bytes) {
    if constexpr (B > 16) { (void)k; (void)p; (void)canon; (void)bytes; throw std::invalid_argument("unsupported suffix width"); }
    else {
        if (bytes == B) return new CblImpl<T, B>(k, p, canon);
        return make_b<T, B + 1>(k, p, canon, bytes);
    }
}

ICbl* make(int k, int tbits, int p, bool canon) {
    if (p > 32) throw std::invalid_argument("PREFIX_BITS should be <= 32");
    int sb = 2 * k + pos_bits_for(2 * k) - p;
    if (sb <= 0) throw std::invalid_argument("SUFFIX_BITS should be != 0");
    int bytes = (sb + 7) / 8;
    switch (tbits) {
        case 32: return make_b<uint32_t, 1>(k, p, canon, bytes);
        case 64: return make_b<uint64_t, 1>(k, p, canon, bytes);
        case 128: return make_b<u128, 1>(k, p, canon, bytes);
        default: throw std::invalid_argument("T must be u32, u64 or u128");
    }
}

inline u128 mk(uint64_t lo, uint64_t hi) { return ((u128)hi << 64) | lo; }

template <class F> int guard(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
    catch (...) { g_err = "unknown error"; return 1; }
}

// generic small objects for the KAT replays
struct QueueBox {
    int tbits; bool rev;
    std::unique_ptr<NecklaceQueue<uint64_t, false>> f64;
    std::unique_ptr<NecklaceQueue<uint64_t, true>> r64;
    std::unique_ptr<NecklaceQueue<u128, false>> f128;
    std::unique_ptr<NecklaceQueue<u128, true>> r128;
    std::unique_ptr<NecklaceQueue<uint32_t, false>> f32;
    std::unique_ptr<NecklaceQueue<uint32_t, true>> r32;
};

}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int orc_uses_reference_cxx() {
#ifdef ORACLE_USE_REFERENCE_CXX
    return 1;
#else
    return 0;
#endif
}

// ---- CBL ----
int orc_cbl_create(int k, int tbits, int prefix_bits, int canonical, void** out) {
    return guard([&] { *out = make(k, tbits, prefix_bits, canonical != 0); });
}
void orc_cbl_destroy(void* h) { delete (ICbl*)h; }
int orc_cbl_clone(void* h, void** out) { return guard([&] { *out = ((ICbl*)h)->clone(); }); }
uint64_t orc_cbl_count(void* h) { return ((ICbl*)h)->count(); }
int orc_cbl_is_empty(void* h) { return ((ICbl*)h)->is_empty() ? 1 : 0; }
int orc_cbl_is_canonical(void* h) { return ((ICbl*)h)->canonical() ? 1 : 0; }

// number of words a sequence yields (valid for ACGT-only input): len - K + 1
int orc_cbl_seq_words(void* h, const uint8_t* seq, size_t len, uint64_t* lo, uint64_t* hi, size_t cap, size_t* n_out) {
    return guard([&] {
        std::vector<u128> w;
        ((ICbl*)h)->seq_words(seq, len, w);
        *n_out = w.size();
        if (w.size() > cap) throw std::invalid_argument("output buffer too small");
        for (size_t i = 0; i < w.size(); i++) { lo[i] = (uint64_t)w[i]; if (hi) hi[i] = (uint64_t)(w[i] >> 64); }
    });
}
int orc_cbl_insert_seq(void* h, const uint8_t* seq, size_t len) { return guard([&] { ((ICbl*)h)->insert_seq(seq, len); }); }
int orc_cbl_remove_seq(void* h, const uint8_t* seq, size_t len) { return guard([&] { ((ICbl*)h)->remove_seq(seq, len); }); }
int orc_cbl_contains_seq(void* h, const uint8_t* seq, size_t len, uint8_t* out, size_t cap, size_t* n_out) {
    return guard([&] {
        std::vector<uint8_t> r;
        ((ICbl*)h)->contains_seq(seq, len, r);
        *n_out = r.size();
        if (r.size() > cap) throw std::invalid_argument("output buffer too small");
        if (!r.empty()) memcpy(out, r.data(), r.size());
    });
}
int orc_cbl_contains_all(void* h, const uint8_t* seq, size_t len, int* out) {
    return guard([&] { *out = ((ICbl*)h)->contains_all(seq, len) ? 1 : 0; });
}
int orc_cbl_insert(void* h, uint64_t lo, uint64_t hi) { return ((ICbl*)h)->insert(mk(lo, hi)) ? 1 : 0; }
int orc_cbl_remove(void* h, uint64_t lo, uint64_t hi) { return ((ICbl*)h)->remove(mk(lo, hi)) ? 1 : 0; }
int orc_cbl_contains(void* h, uint64_t lo, uint64_t hi) { return ((ICbl*)h)->contains(mk(lo, hi)) ? 1 : 0; }
void orc_cbl_get_word(void* h, uint64_t lo, uint64_t hi, uint64_t* wlo, uint64_t* whi) {
    u128 w = ((ICbl*)h)->get_word(mk(lo, hi));
    *wlo = (uint64_t)w; *whi = (uint64_t)(w >> 64);
}
void orc_cbl_recover_kmer(void* h, uint64_t lo, uint64_t hi, uint64_t* klo, uint64_t* khi) {
    u128 w = ((ICbl*)h)->recover_kmer(mk(lo, hi));
    *klo = (uint64_t)w; *khi = (uint64_t)(w >> 64);
}
// words in the reference's iteration order (sorted != 0: ascending word order, the F5 definition)
int orc_cbl_iter_words(void* h, int sorted, uint64_t* lo, uint64_t* hi, size_t cap, size_t* n_out) {
    return guard([&] {
        std::vector<u128> w;
        ((ICbl*)h)->iter_words(w);
        if (sorted) std::sort(w.begin(), w.end());
        *n_out = w.size();
        if (w.size() > cap) throw std::invalid_argument("output buffer too small");
        for (size_t i = 0; i < w.size(); i++) { lo[i] = (uint64_t)w[i]; if (hi) hi[i] = (uint64_t)(w[i] >> 64); }
    });
}
int orc_cbl_binary_op(int op, void* a, void* b, void** out) { return guard([&] { *out = ((ICbl*)a)->binary_op(op, (ICbl*)b); }); }
int orc_cbl_assign_op(int op, void* a, void* b) { return guard([&] { ((ICbl*)a)->assign_op(op, (ICbl*)b); }); }
int orc_cbl_merge_many(void** hs, size_t n, int intersect, void** out) {
    return guard([&] {
        if (n == 0) throw std::invalid_argument("empty list");
        std::vector<ICbl*> v;
        for (size_t i = 0; i < n; i++) v.push_back((ICbl*)hs[i]);
        *out = v[0]->merge_many(v, intersect != 0);
    });
}
int orc_cbl_serialize(void* h, uint8_t* out, size_t cap, size_t* n_out) {
    return guard([&] {
        std::vector<uint8_t> b;
        ((ICbl*)h)->serialize(b);
        *n_out = b.size();
        if (out) { if (b.size() > cap) throw std::invalid_argument("output buffer too small"); memcpy(out, b.data(), b.size()); }
    });
}
int orc_cbl_deserialize(void* proto, const uint8_t* p, size_t n, void** out) {
    return guard([&] { *out = ((ICbl*)proto)->deserialize(p, n); });
}
int orc_cbl_bucket_sizes(void* h, uint64_t* prefixes, uint64_t* sizes, size_t cap, size_t* n_out) {
    return guard([&] {
        std::vector<std::pair<size_t, size_t>> v;
        ((ICbl*)h)->bucket_sizes(v);
        *n_out = v.size();
        if (prefixes) {
            if (v.size() > cap) throw std::invalid_argument("output buffer too small");
            for (size_t i = 0; i < v.size(); i++) { prefixes[i] = v[i].first; sizes[i] = v[i].second; }
        }
    });
}

// timed CPU baseline: 1 core, returns seconds for insert_seq / contains_seq over n_seqs records
double orc_cbl_time_insert_seqs(void* h, const uint8_t* buf, const uint64_t* offs, size_t n_seqs) {
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < n_seqs; i++) ((ICbl*)h)->insert_seq(buf + offs[i], (size_t)(offs[i + 1] - offs[i]));
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
double orc_cbl_time_contains_seqs(void* h, const uint8_t* buf, const uint64_t* offs, size_t n_seqs, uint64_t* n_pos) {
    auto t0 = std::chrono::steady_clock::now();
    uint64_t pos = 0;
    std::vector<uint8_t> r;
    for (size_t i = 0; i < n_seqs; i++) {
        r.clear();
        ((ICbl*)h)->contains_seq(buf + offs[i], (size_t)(offs[i + 1] - offs[i]), r);
        for (uint8_t b : r) pos += b;
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (n_pos) *n_pos = pos;
    return dt;
}

// ---- kmer / necklace primitives for the KAT replays ----
// revcomp through nucleotides (src/kmer.rs:355-378)
int orc_revcomp_nucs(int k, int tbits, const uint8_t* in, uint8_t* out) {
    return guard([&] {
        if (tbits == 32) { KmerOps<uint32_t> o(k); o.to_nucs(o.rev_comp(o.from_nucs(in, (size_t)k)), out); }
        else if (tbits == 64) { KmerOps<uint64_t> o(k); o.to_nucs(o.rev_comp(o.from_nucs(in, (size_t)k)), out); }
        else if (tbits == 128) { KmerOps<u128> o(k); o.to_nucs(o.rev_comp(o.from_nucs(in, (size_t)k)), out); }
        else throw std::invalid_argument("tbits");
    });
}
void orc_kmer_from_nucs(int k, const uint8_t* in, size_t n, uint64_t* lo, uint64_t* hi) {
    KmerOps<u128> o(k);
    u128 x = o.from_nucs(in, n);
    *lo = (uint64_t)x; *hi = (uint64_t)(x >> 64);
}
void orc_kmer_revcomp(int k, int tbits, uint64_t lo, uint64_t hi, uint64_t* olo, uint64_t* ohi) {
    u128 r;
    if (tbits == 32) r = KmerOps<uint32_t>(k).rev_comp((uint32_t)lo);
    else if (tbits == 64) r = KmerOps<uint64_t>(k).rev_comp(lo);
    else r = KmerOps<u128>(k).rev_comp(mk(lo, hi));
    *olo = (uint64_t)r; *ohi = (uint64_t)(r >> 64);
}
// brute-force necklace (src/necklace/mod.rs:13-25)
void orc_necklace_pos(int bits, uint64_t lo, uint64_t hi, uint64_t* nlo, uint64_t* nhi, uint64_t* pos) {
    auto r = necklace_pos<u128>(mk(lo, hi), bits);
    *nlo = (uint64_t)r.first; *nhi = (uint64_t)(r.first >> 64); *pos = r.second;
}
void orc_revert_necklace_pos(int bits, uint64_t lo, uint64_t hi, uint64_t pos, uint64_t* wlo, uint64_t* whi) {
    u128 w = revert_necklace_pos<u128>(mk(lo, hi), (size_t)pos, bits);
    *wlo = (uint64_t)w; *whi = (uint64_t)(w >> 64);
}
// batch: brute force vs streaming queue on an array of words; returns number of mismatches
uint64_t orc_necklace_queue_vs_brute(int bits, int width, int reverse, const uint64_t* lo, const uint64_t* hi, size_t n) {
    uint64_t bad = 0;
    for (size_t i = 0; i < n; i++) {
        u128 w = mk(lo[i], hi ? hi[i] : 0);
        auto b = necklace_pos<u128>(w, bits);
        std::pair<u128, size_t> q;
        if (reverse) { NecklaceQueue<u128, true> nq(bits, (size_t)width); nq.insert_full(w); q = nq.get_necklace_pos(); }
        else { NecklaceQueue<u128, false> nq(bits, (size_t)width); nq.insert_full(w); q = nq.get_necklace_pos(); }
        if (b != q) bad++;
    }
    return bad;
}

// streaming queue object (u64 / u128 words)
void* orc_queue_new(int bits, int width, int reverse, int tbits) {
    auto* q = new QueueBox();
    q->tbits = tbits; q->rev = reverse != 0;
    if (tbits == 32) { if (q->rev) q->r32.reset(new NecklaceQueue<uint32_t, true>(bits, (size_t)width)); else q->f32.reset(new NecklaceQueue<uint32_t, false>(bits, (size_t)width)); }
    else if (tbits == 64) { if (q->rev) q->r64.reset(new NecklaceQueue<uint64_t, true>(bits, (size_t)width)); else q->f64.reset(new NecklaceQueue<uint64_t, false>(bits, (size_t)width)); }
    else { if (q->rev) q->r128.reset(new NecklaceQueue<u128, true>(bits, (size_t)width)); else q->f128.reset(new NecklaceQueue<u128, false>(bits, (size_t)width)); }
    return q;
}
void orc_queue_free(void* p) { delete (QueueBox*)p; }
#define QDISPATCH(q, call)                                   \
    do {                                                     \
        if ((q)->f32) (q)->f32->call;                        \
        else if ((q)->r32) (q)->r32->call;                   \
        else if ((q)->f64) (q)->f64->call;                   \
        else if ((q)->r64) (q)->r64->call;                   \
        else if ((q)->f128) (q)->f128->call;                 \
        else (q)->r128->call;                                \
    } while (0)
void orc_queue_insert_full(void* p, uint64_t lo, uint64_t hi) {
    auto* q = (QueueBox*)p;
    if (q->f32) q->f32->insert_full((uint32_t)lo); else if (q->r32) q->r32->insert_full((uint32_t)lo);
    else if (q->f64) q->f64->insert_full(lo); else if (q->r64) q->r64->insert_full(lo);
    else if (q->f128) q->f128->insert_full(mk(lo, hi)); else q->r128->insert_full(mk(lo, hi));
}
void orc_queue_insert(void* p, uint64_t bit) { auto* q = (QueueBox*)p; QDISPATCH(q, insert(bit)); }
void orc_queue_insert2(void* p, uint64_t two) { auto* q = (QueueBox*)p; QDISPATCH(q, insert2(two)); }
void orc_queue_get(void* p, uint64_t* nlo, uint64_t* nhi, uint64_t* pos) {
    auto* q = (QueueBox*)p;
    u128 n; size_t ps;
    if (q->f32) { auto r = q->f32->get_necklace_pos(); n = r.first; ps = r.second; }
    else if (q->r32) { auto r = q->r32->get_necklace_pos(); n = r.first; ps = r.second; }
    else if (q->f64) { auto r = q->f64->get_necklace_pos(); n = r.first; ps = r.second; }
    else if (q->r64) { auto r = q->r64->get_necklace_pos(); n = r.first; ps = r.second; }
    else if (q->f128) { auto r = q->f128->get_necklace_pos(); n = r.first; ps = r.second; }
    else { auto r = q->r128->get_necklace_pos(); n = r.first; ps = r.second; }
    *nlo = (uint64_t)n; *nhi = (uint64_t)(n >> 64); *pos = ps;
}

// LexMinQueue (src/necklace/minimizer.rs:110-166)
void* orc_lexmin_new(int width) { return new LexMinQueue<uint64_t>((size_t)width); }
void orc_lexmin_free(void* p) { delete (LexMinQueue<uint64_t>*)p; }
void orc_lexmin_insert_full(void* p, const uint64_t* vals, size_t n) {
    ((LexMinQueue<uint64_t>*)p)->insert_full(std::vector<uint64_t>(vals, vals + n));
}
void orc_lexmin_insert(void* p, uint64_t u) { ((LexMinQueue<uint64_t>*)p)->insert(u); }
void orc_lexmin_insert2(void* p, uint64_t u, uint64_t v) { ((LexMinQueue<uint64_t>*)p)->insert2(u, v); }
size_t orc_lexmin_min_pos(void* p, uint64_t* out, size_t cap) {
    auto v = ((LexMinQueue<uint64_t>*)p)->iter_min_pos();
    for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
    return v.size();
}

// Bitvector (src/bitvector/mod.rs:149-187)
void* orc_bv_new(int bitlength) { return new Bitvector(bitlength); }
void orc_bv_free(void* p) { delete (Bitvector*)p; }
int orc_bv_insert(void* p, uint64_t i) { return ((Bitvector*)p)->insert((size_t)i) ? 1 : 0; }
int orc_bv_remove(void* p, uint64_t i) { return ((Bitvector*)p)->remove((size_t)i) ? 1 : 0; }
int orc_bv_contains(void* p, uint64_t i) { return ((Bitvector*)p)->contains((size_t)i) ? 1 : 0; }
uint64_t orc_bv_rank(void* p, uint64_t i) { return ((Bitvector*)p)->rank((size_t)i); }
uint64_t orc_bv_count(void* p) { return ((Bitvector*)p)->count(); }
size_t orc_bv_iter(void* p, uint64_t* out, size_t cap) {
    auto v = ((Bitvector*)p)->indices();
    for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
    return v.size();
}
int orc_bv_assign_op(void* a, int op, void* b) { ((Bitvector*)a)->assign_op((Bitvector::Op)op, *(Bitvector*)b); return 0; }

// Tiered vector (src/ffi.rs:29-39)
void* orc_tiered_new() { return new TieredImpl(); }
void orc_tiered_free(void* p) { delete (TieredImpl*)p; }
void orc_tiered_insert(void* p, uint64_t idx, uint32_t v) { ((TieredImpl*)p)->insert((size_t)idx, v); }
void orc_tiered_remove(void* p, uint64_t idx) { ((TieredImpl*)p)->remove((size_t)idx); }
uint32_t orc_tiered_get(void* p, uint64_t idx) { return ((TieredImpl*)p)->get((size_t)idx); }
uint64_t orc_tiered_len(void* p) { return ((TieredImpl*)p)->len(); }

// Trie<3> (src/trie.rs:228-261)
void* orc_trie3_new() { return new Trie<3>(); }
void orc_trie3_free(void* p) { delete (Trie<3>*)p; }
int orc_trie3_insert(void* p, const uint8_t* b) { return ((Trie<3>*)p)->insert(b) ? 1 : 0; }
int orc_trie3_remove(void* p, const uint8_t* b) { return ((Trie<3>*)p)->remove(b) ? 1 : 0; }
int orc_trie3_contains(void* p, const uint8_t* b) { return ((Trie<3>*)p)->contains(b) ? 1 : 0; }
int orc_trie3_is_empty(void* p) { return ((Trie<3>*)p)->is_empty() ? 1 : 0; }
uint64_t orc_trie3_count(void* p) { return ((Trie<3>*)p)->count(); }
size_t orc_trie3_iter(void* p, uint8_t* out, size_t cap_items) {
    size_t n = 0;
    ((Trie<3>*)p)->for_each([&](const std::array<uint8_t, 3>& w) { if (n < cap_items) memcpy(out + 3 * n, w.data(), 3); n++; });
    return n;
}

// SlicedInt<3> (src/sliced_int.rs:143-169)
uint64_t orc_sliced3_roundtrip(uint64_t v) { return (uint64_t)SlicedInt<3>::from_int(v).get(); }
int orc_sliced3_cmp(uint64_t a, uint64_t b) { return SlicedInt<3>::from_int(a).cmp(SlicedInt<3>::from_int(b)); }

// WordSet<PREFIX_BITS, 8> i.e. BYTES = 1 (src/wordset/mod.rs:452-533, set_ops.rs:424-836)
void* orc_ws1_new(int prefix_bits, int suffix_bits) { try { return new WordSet<1>(prefix_bits, suffix_bits); } catch (...) { return nullptr; } }
void orc_ws1_free(void* p) { delete (WordSet<1>*)p; }
int orc_ws1_insert(void* p, uint64_t w) { return ((WordSet<1>*)p)->insert(w) ? 1 : 0; }
int orc_ws1_remove(void* p, uint64_t w) { return ((WordSet<1>*)p)->remove(w) ? 1 : 0; }
int orc_ws1_contains(void* p, uint64_t w) { return ((WordSet<1>*)p)->contains(w) ? 1 : 0; }
uint64_t orc_ws1_count(void* p) { return ((WordSet<1>*)p)->count(); }
int orc_ws1_is_empty(void* p) { return ((WordSet<1>*)p)->is_empty() ? 1 : 0; }
static std::vector<u128> widen(const uint64_t* w, size_t n) { std::vector<u128> v(n); for (size_t i = 0; i < n; i++) v[i] = w[i]; return v; }
void orc_ws1_insert_batch(void* p, const uint64_t* w, size_t n) { auto v = widen(w, n); ((WordSet<1>*)p)->insert_batch(v.data(), n); }
void orc_ws1_remove_batch(void* p, const uint64_t* w, size_t n) { auto v = widen(w, n); ((WordSet<1>*)p)->remove_batch(v.data(), n); }
void orc_ws1_contains_batch(void* p, const uint64_t* w, size_t n, uint8_t* out) {
    auto v = widen(w, n);
    std::vector<uint8_t> r;
    ((WordSet<1>*)p)->contains_batch(v.data(), n, r);
    memcpy(out, r.data(), r.size());
}
size_t orc_ws1_iter(void* p, uint64_t* out, size_t cap) {
    size_t n = 0;
    ((WordSet<1>*)p)->for_each_word([&](u128 w) { if (n < cap) out[n] = (uint64_t)w; n++; });
    return n;
}
void* orc_ws1_binary_op(int op, void* a, void* b) {
    return new WordSet<1>(WordSet<1>::binary_op((WordSet<1>::Op)op, *(WordSet<1>*)a, *(WordSet<1>*)b));
}
void orc_ws1_assign_op(int op, void* a, void* b) {
    auto* x = (WordSet<1>*)a; auto* y = (WordSet<1>*)b;
    switch (op) { case 0: x->or_assign(*y); break; case 1: x->and_assign(*y); break; case 2: x->sub_assign(*y); break; default: x->xor_assign(*y); }
}
void* orc_ws1_merge_many(void** hs, size_t n, int intersect) {
    std::vector<WordSet<1>*> v;
    for (size_t i = 0; i < n; i++) v.push_back((WordSet<1>*)hs[i]);
    return new WordSet<1>(intersect ? WordSet<1>::intersect(v) : WordSet<1>::merge(v));
}

}  // extern "C"
