// extern "C" surface of libcbl_gpu (include/cbl_gpu.h).  Nothing but status codes crosses the ABI.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/cbl_gpu.h"
#include "cbl_index.cuh"

using namespace cbl;

#define CBL_STR2(x) #x
#define CBL_STR(x) CBL_STR2(x)

struct cbl_handle {
    std::unique_ptr<IIndex> ix;
    std::string err;
};

namespace {
thread_local std::string g_err;

template <class F> int32_t guard(cbl_handle* h, F&& f) {
    try {
        f();
        return CBL_OK;
    } catch (const Error& e) {
        (h ? h->err : g_err) = e.what();
        g_err = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        (h ? h->err : g_err) = "host allocation failed";
        return CBL_ENOMEM;
    } catch (const std::exception& e) {
        (h ? h->err : g_err) = e.what();
        return CBL_ECUDA;
    } catch (...) {
        (h ? h->err : g_err) = "unknown error";
        return CBL_ECUDA;
    }
}
cbl_handle* mut(const cbl_handle* h) { return const_cast<cbl_handle*>(h); }
void need(const void* p, const char* what) {
    if (!p) throw Error(CBL_EINVAL, std::string("null pointer: ") + what);
}

// ---- bincode 1.3 varint (DefaultOptions::with_varint_encoding, src/cbl.rs:132-135) -------------
struct Writer {
    std::vector<uint8_t>* out;  // null => only count
    size_t n = 0;
    void u8(uint8_t v) { if (out) out->push_back(v); n++; }
    void varint(uint64_t v) {
        if (v < 251) u8((uint8_t)v);
        else if (v <= 0xFFFF) { u8(251); for (int i = 0; i < 2; i++) u8((uint8_t)(v >> (8 * i))); }
        else if (v <= 0xFFFFFFFFull) { u8(252); for (int i = 0; i < 4; i++) u8((uint8_t)(v >> (8 * i))); }
        else { u8(253); for (int i = 0; i < 8; i++) u8((uint8_t)(v >> (8 * i))); }
    }
};
struct Reader {
    const uint8_t* p;
    size_t n, i = 0;
    uint8_t u8() { if (i >= n) throw Error(CBL_EIO, "unexpected end of serialized index"); return p[i++]; }
    uint64_t varint() {
        uint8_t t = u8();
        if (t < 251) return t;
        int nb = t == 251 ? 2 : t == 252 ? 4 : t == 253 ? 8 : -1;
        if (nb < 0) throw Error(CBL_EIO, "bad varint tag in serialized index");
        uint64_t v = 0;
        for (int k = 0; k < nb; k++) v |= (uint64_t)u8() << (8 * k);
        return v;
    }
};

// layout: src/cbl.rs:40-54 (canonical, wordset) ; src/wordset/mod.rs:382-401 (map prefix -> TrieVec) ;
// src/trievec/mod.rs:8-15 (enum Vec | Trie) ; src/sliced_int.rs:110-114 (bytes) — SURVEY section 8 row f1
void serialize_index(IIndex* ix, Writer& w) {
    const KParams& P = ix->params();
    const int BYTES = (P.suffix_bits + 7) / 8;
    uint64_t nb = 0;
    ix->bucket_sizes(nullptr, nullptr, 0, &nb);
    std::vector<uint32_t> prefixes(nb ? nb : 1), sizes(nb ? nb : 1);
    if (nb) ix->bucket_sizes(prefixes.data(), sizes.data(), nb, &nb);
    w.u8(ix->config().canonical ? 1 : 0);
    w.varint(nb);
    const uint64_t CH = 1 << 22;
    std::vector<uint64_t> lo(CH), hi(CH);
    uint64_t have = 0, pos = 0, next = 0;  // buffered words [next-have+pos ..)
    const uint64_t total = ix->count();
    const unsigned __int128 smask = (((unsigned __int128)1) << P.suffix_bits) - 1;
    for (uint64_t r = 0; r < nb; r++) {
        w.varint(prefixes[r]);
        w.varint(0);  // TrieOrVec::Vec
        w.varint(sizes[r]);
        for (uint32_t e = 0; e < sizes[r]; e++) {
            if (pos == have) {
                uint64_t got = 0;
                ix->export_words(next, std::min<uint64_t>(CH, total - next), 0, lo.data(), hi.data(), &got);
                if (!got) throw Error(CBL_ECUDA, "internal: export ran dry");
                have = got; pos = 0; next += got;
            }
            unsigned __int128 word = ((unsigned __int128)hi[pos] << 64) | lo[pos];
            pos++;
            unsigned __int128 s = word & smask;
            w.varint(BYTES);
            for (int b = 0; b < BYTES; b++) w.u8((uint8_t)(s >> (8 * b)));
        }
    }
}

void read_trie(Reader& r, int depth, int BYTES, unsigned __int128 acc, unsigned __int128 prefix_part, int suffix_bits,
               std::vector<uint64_t>& lo, std::vector<uint64_t>& hi) {
    // node = { bv: seq of set indices (u8), children: seq of nodes } (src/trie.rs:53-57, bitvector/tiny/mod.rs:97-105)
    uint64_t k = r.varint();
    std::vector<uint8_t> idx(k);
    for (auto& b : idx) b = r.u8();
    uint64_t c = r.varint();
    if (depth == BYTES - 1 || c == 0) {
        for (uint8_t b : idx) {
            unsigned __int128 s = (acc << 8) | b;  // big-endian bytes, most significant first
            unsigned __int128 word = prefix_part | s;
            lo.push_back((uint64_t)word);
            hi.push_back((uint64_t)(word >> 64));
        }
        if (c != 0) throw Error(CBL_EIO, "trie leaf with children in serialized index");
        return;
    }
    if (c != k) throw Error(CBL_EIO, "trie node with mismatched children in serialized index");
    for (uint64_t i = 0; i < c; i++) read_trie(r, depth + 1, BYTES, (acc << 8) | idx[i], prefix_part, suffix_bits, lo, hi);
}

// prefix_lo / prefix_hi: keep only the buckets with prefix in [prefix_lo, prefix_hi) (one rank's range of a sharded set)
cbl_handle* deserialize_index(const cbl_handle* proto, const uint8_t* data, size_t len, uint64_t prefix_lo = 0, uint64_t prefix_hi = ~0ull) {
    Reader r{data, len};
    const int canonical = r.u8() ? 1 : 0;
    std::unique_ptr<cbl_handle> h(new cbl_handle());
    h->ix.reset(proto->ix->new_empty(canonical));   // same K / T / PREFIX_BITS / device(s) (and splitters) as the prototype
    const KParams& P = h->ix->params();
    const int BYTES = (P.suffix_bits + 7) / 8;
    uint64_t nb = r.varint();
    std::vector<uint64_t> lo, hi;
    for (uint64_t e = 0; e < nb; e++) {
        uint64_t prefix = r.varint();
        if (prefix >> P.prefix_bits) throw Error(CBL_EIO, "prefix out of range in serialized index");
        unsigned __int128 pp = (unsigned __int128)prefix << P.suffix_bits;
        uint64_t variant = r.varint();
        const size_t keep_from = lo.size();
        const bool keep = prefix >= prefix_lo && prefix < prefix_hi;
        struct Drop { std::vector<uint64_t>&a, &b; size_t n; bool keep; ~Drop() { if (!keep) { a.resize(n); b.resize(n); } } } drop{lo, hi, keep_from, keep};
        if (variant == 0) {
            uint64_t m = r.varint();
            for (uint64_t i = 0; i < m; i++) {
                if (r.varint() != (uint64_t)BYTES) throw Error(CBL_EIO, "suffix width mismatch in serialized index (different K / PREFIX_BITS?)");
                unsigned __int128 s = 0;
                for (int b = 0; b < BYTES; b++) s |= (unsigned __int128)r.u8() << (8 * b);
                unsigned __int128 word = pp | s;
                lo.push_back((uint64_t)word);
                hi.push_back((uint64_t)(word >> 64));
            }
        } else if (variant == 1) {
            read_trie(r, 0, BYTES, 0, pp, P.suffix_bits, lo, hi);
            (void)r.varint();  // element count stored next to the trie (src/trievec/mod.rs:11)
        } else throw Error(CBL_EIO, "bad bucket variant in serialized index");
    }
    if (r.i != r.n) throw Error(CBL_EIO, "trailing bytes after serialized index");  // reject_trailing_bytes
    h->ix->load_sorted_words(lo.data(), hi.data(), lo.size());
    return h.release();
}

// k-way union / intersection (src/cbl.rs:108-124, src/wordset/set_ops.rs:11-75).  Every binary step is one streaming merge
// of two CSR indexes, so the order matters for traffic: unions pair the inputs up in a balanced tree (every element is
// rewritten ~log2 k times instead of up to k times by a left fold), intersections start from the two smallest sets (the
// running result only shrinks).
cbl_handle* fold_many(cbl_handle** hs, size_t n, int op) {
    if (!hs || n == 0) throw Error(CBL_EINVAL, "empty list of indexes");
    for (size_t i = 0; i < n; i++) need(hs[i], "index handle");
    std::vector<std::unique_ptr<IIndex>> own;      // intermediate results
    std::vector<IIndex*> cur;
    for (size_t i = 0; i < n; i++) cur.push_back(hs[i]->ix.get());
    if (n == 1) { own.emplace_back(cur[0]->clone()); cur[0] = own.back().get(); }
    if (op == SETOP_AND) {
        std::stable_sort(cur.begin(), cur.end(), [](IIndex* a, IIndex* b) { return a->count() < b->count(); });
        own.emplace_back(cur[0]->setop(op, cur[1 < n ? 1 : 0]));
        IIndex* acc = own.back().get();
        for (size_t i = 2; i < n; i++) acc->setop_assign(op, cur[i]);
        cur.assign(1, acc);
    }
    while (cur.size() > 1) {
        std::vector<IIndex*> next;
        for (size_t i = 0; i + 1 < cur.size(); i += 2) {
            own.emplace_back(cur[i]->setop(op, cur[i + 1]));
            next.push_back(own.back().get());
        }
        if (cur.size() & 1) next.push_back(cur.back());
        cur.swap(next);
    }
    // hand the final result over (it is owned by `own` unless n == 1 cloned it there too)
    std::unique_ptr<cbl_handle> acc(new cbl_handle());
    for (auto& o : own)
        if (o.get() == cur[0]) { acc->ix = std::move(o); break; }
    if (!acc->ix) acc->ix.reset(cur[0]->clone());
    return acc.release();
}
}  // namespace

extern "C" {

int32_t cbl_create(uint32_t k, uint32_t word_bits, uint32_t prefix_bits, int32_t canonical, int32_t device, cbl_t** out) {
    return guard(nullptr, [&] {
        need(out, "out");
        Config cfg{(int)k, (int)word_bits, (int)prefix_bits, canonical ? 1 : 0, device};
        std::unique_ptr<cbl_handle> h(new cbl_handle());
        h->ix.reset(make_index(cfg));
        *out = h.release();
    });
}
int32_t cbl_create_sharded_ex(uint32_t k, uint32_t word_bits, uint32_t prefix_bits, int32_t canonical, int32_t n_gpus, const int32_t* devices,
                              const uint32_t* splitters, cbl_t** out) {
    return guard(nullptr, [&] {
        need(out, "out"); need(devices, "devices");
        Config cfg{(int)k, (int)word_bits, (int)prefix_bits, canonical ? 1 : 0, devices[0]};
        std::unique_ptr<cbl_handle> h(new cbl_handle());
        h->ix.reset(make_sharded_index(cfg, devices, n_gpus, splitters));
        *out = h.release();
    });
}
int32_t cbl_create_sharded(uint32_t k, uint32_t word_bits, uint32_t prefix_bits, int32_t canonical, int32_t n_gpus, const int32_t* devices, cbl_t** out) {
    return cbl_create_sharded_ex(k, word_bits, prefix_bits, canonical, n_gpus, devices, nullptr, out);
}
int32_t cbl_sharded_splitters(cbl_t* h, uint32_t* out, size_t cap, size_t* n_out) {
    return guard(h, [&] {
        need(h, "handle"); need(n_out, "n_out");
        std::vector<uint32_t> sp;
        if (!sharded_splitters(h->ix.get(), sp)) throw Error(CBL_EINVAL, "not a sharded handle");
        *n_out = sp.size();
        if (out) {
            if (cap < sp.size()) throw Error(CBL_EINVAL, "output buffer too small");
            for (size_t i = 0; i < sp.size(); i++) out[i] = sp[i];
        }
    });
}
int32_t cbl_destroy(cbl_t* h) {
    return guard(nullptr, [&] { delete h; });
}
int32_t cbl_clone(cbl_t* h, cbl_t** out) {
    return guard(h, [&] {
        need(h, "handle"); need(out, "out");
        std::unique_ptr<cbl_handle> c(new cbl_handle());
        c->ix.reset(h->ix->clone());
        *out = c.release();
    });
}
const char* cbl_last_error(const cbl_t* h) { return h ? h->err.c_str() : g_err.c_str(); }
const char* cbl_last_global_error(void) { return g_err.c_str(); }

int32_t cbl_count(const cbl_t* h, uint64_t* out) { return guard(mut(h), [&] { need(h, "handle"); need(out, "out"); *out = h->ix->count(); }); }
int32_t cbl_is_empty(const cbl_t* h, int32_t* out) {
    return guard(mut(h), [&] { need(h, "handle"); need(out, "out"); *out = h->ix->is_empty_reference_semantics() ? 1 : 0; });
}
int32_t cbl_is_canonical(const cbl_t* h, int32_t* out) {
    return guard(mut(h), [&] { need(h, "handle"); need(out, "out"); *out = h->ix->config().canonical; });
}
int32_t cbl_num_buckets(const cbl_t* h, uint64_t* out) {
    return guard(mut(h), [&] { need(h, "handle"); need(out, "out"); *out = h->ix->n_buckets(); });
}

int32_t cbl_insert_seq(cbl_t* h, const uint8_t* seq, size_t len) {
    return guard(h, [&] { need(h, "handle"); need(seq, "seq"); uint64_t off[2] = {0, len}; h->ix->insert_seqs(seq, off, 1, false); });
}
int32_t cbl_remove_seq(cbl_t* h, const uint8_t* seq, size_t len) {
    return guard(h, [&] { need(h, "handle"); need(seq, "seq"); uint64_t off[2] = {0, len}; h->ix->insert_seqs(seq, off, 1, true); });
}
int32_t cbl_contains_seq(cbl_t* h, const uint8_t* seq, size_t len, uint8_t* out, size_t* n_out) {
    return guard(h, [&] {
        need(h, "handle"); need(seq, "seq"); need(out, "out");
        uint64_t off[2] = {0, len};
        h->ix->contains_seqs(seq, off, 1, out);
        if (n_out) *n_out = (size_t)h->ix->last_produced;   // len - K + 1, fewer after non-ACGT bytes (F8)
    });
}
int32_t cbl_contains_all(cbl_t* h, const uint8_t* seq, size_t len, int32_t* out) {
    return guard(h, [&] {
        need(h, "handle"); need(seq, "seq"); need(out, "out");
        uint64_t off[2] = {0, len};
        if (len < (size_t)h->ix->config().k) h->ix->contains_seqs(seq, off, 1, nullptr);  // throws the short-sequence error
        std::vector<uint8_t> r(len - h->ix->config().k + 1);
        h->ix->contains_seqs(seq, off, 1, r.data());
        int all = 1;
        for (uint64_t i = 0; i < h->ix->last_produced; i++) if (!r[i]) { all = 0; break; }
        *out = all;
    });
}

int32_t cbl_last_kmer_count(const cbl_t* h, uint64_t* out) {
    return guard(mut(h), [&] { need(h, "handle"); need(out, "out"); *out = h->ix->last_produced; });
}

int32_t cbl_insert_seqs(cbl_t* h, const uint8_t* buf, const uint64_t* offsets, size_t n_seqs) {
    return guard(h, [&] { need(h, "handle"); need(buf, "buf"); need(offsets, "offsets"); h->ix->insert_seqs(buf, offsets, n_seqs, false); });
}
int32_t cbl_remove_seqs(cbl_t* h, const uint8_t* buf, const uint64_t* offsets, size_t n_seqs) {
    return guard(h, [&] { need(h, "handle"); need(buf, "buf"); need(offsets, "offsets"); h->ix->insert_seqs(buf, offsets, n_seqs, true); });
}
int32_t cbl_contains_seqs(cbl_t* h, const uint8_t* buf, const uint64_t* offsets, size_t n_seqs, uint8_t* out) {
    return guard(h, [&] { need(h, "handle"); need(buf, "buf"); need(offsets, "offsets"); need(out, "out"); h->ix->contains_seqs(buf, offsets, n_seqs, out); });
}
int32_t cbl_insert_seqs_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs) {
    return guard(h, [&] { need(h, "handle"); need(d_buf, "d_buf"); need(offsets, "offsets"); h->ix->insert_seqs_dev(d_buf, offsets[n_seqs], offsets, n_seqs); });
}
int32_t cbl_remove_seqs_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs) {
    return guard(h, [&] { need(h, "handle"); need(d_buf, "d_buf"); need(offsets, "offsets"); h->ix->remove_seqs_dev(d_buf, offsets[n_seqs], offsets, n_seqs); });
}
int32_t cbl_contains_seqs_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs, uint8_t* d_out) {
    return guard(h, [&] {
        need(h, "handle"); need(d_buf, "d_buf"); need(offsets, "offsets"); need(d_out, "d_out");
        h->ix->contains_seqs_dev(d_buf, offsets[n_seqs], offsets, n_seqs, d_out);
    });
}
int32_t cbl_count_kmers(const cbl_t* h, const uint64_t* offsets, size_t n_seqs, uint64_t* out) {
    return guard(mut(h), [&] { need(h, "handle"); need(offsets, "offsets"); need(out, "out"); *out = total_kmers_of(h->ix->config(), offsets, n_seqs); });
}

int32_t cbl_contains_kmers(cbl_t* h, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) {
    return guard(h, [&] { need(h, "handle"); need(lo, "lo"); need(out, "out"); h->ix->kmers_op(0, lo, hi, n, out); });
}
int32_t cbl_insert_kmers(cbl_t* h, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) {
    return guard(h, [&] { need(h, "handle"); need(lo, "lo"); h->ix->kmers_op(1, lo, hi, n, out); });
}
int32_t cbl_remove_kmers(cbl_t* h, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) {
    return guard(h, [&] { need(h, "handle"); need(lo, "lo"); h->ix->kmers_op(2, lo, hi, n, out); });
}

int32_t cbl_setop(int32_t op, cbl_t* a, cbl_t* b, cbl_t** out) {
    return guard(a, [&] {
        need(a, "a"); need(b, "b"); need(out, "out");
        if (op < 0 || op > 3) throw Error(CBL_EINVAL, "unknown set operation");
        std::unique_ptr<cbl_handle> r(new cbl_handle());
        r->ix.reset(a->ix->setop(op, b->ix.get()));
        *out = r.release();
    });
}
int32_t cbl_setop_assign(int32_t op, cbl_t* a, cbl_t* b) {
    return guard(a, [&] {
        need(a, "a"); need(b, "b");
        if (op < 0 || op > 3) throw Error(CBL_EINVAL, "unknown set operation");
        a->ix->setop_assign(op, b->ix.get());
    });
}
int32_t cbl_merge_many(cbl_t** hs, size_t n, cbl_t** out) {
    return guard(n && hs ? hs[0] : nullptr, [&] { need(out, "out"); *out = fold_many(hs, n, SETOP_OR); });
}
int32_t cbl_intersect_many(cbl_t** hs, size_t n, cbl_t** out) {
    return guard(n && hs ? hs[0] : nullptr, [&] { need(out, "out"); *out = fold_many(hs, n, SETOP_AND); });
}

int32_t cbl_export_words(cbl_t* h, uint64_t start, uint64_t* lo, uint64_t* hi, size_t cap, size_t* n_out) {
    return guard(h, [&] {
        need(h, "handle"); need(lo, "lo"); need(n_out, "n_out");
        uint64_t n = 0;
        h->ix->export_words(start, cap, 0, lo, hi, &n);
        *n_out = n;
    });
}
int32_t cbl_export_kmers(cbl_t* h, uint64_t start, uint64_t* lo, uint64_t* hi, size_t cap, size_t* n_out) {
    return guard(h, [&] {
        need(h, "handle"); need(lo, "lo"); need(n_out, "n_out");
        uint64_t n = 0;
        h->ix->export_words(start, cap, 1, lo, hi, &n);
        *n_out = n;
    });
}
int32_t cbl_bucket_sizes(cbl_t* h, uint32_t* prefixes, uint32_t* sizes, size_t cap, size_t* n_out) {
    return guard(h, [&] {
        need(h, "handle"); need(n_out, "n_out");
        uint64_t n = 0;
        h->ix->bucket_sizes(prefixes, sizes, cap, &n);
        *n_out = n;
    });
}

int32_t cbl_serialize_size(cbl_t* h, size_t* out) {
    return guard(h, [&] { need(h, "handle"); need(out, "out"); Writer w{nullptr}; serialize_index(h->ix.get(), w); *out = w.n; });
}
int32_t cbl_serialize(cbl_t* h, uint8_t* out, size_t cap, size_t* n_out) {
    return guard(h, [&] {
        need(h, "handle"); need(out, "out");
        std::vector<uint8_t> v;
        Writer w{&v};
        serialize_index(h->ix.get(), w);
        if (v.size() > cap) throw Error(CBL_EINVAL, "output buffer too small");
        memcpy(out, v.data(), v.size());
        if (n_out) *n_out = v.size();
    });
}
int32_t cbl_deserialize(const cbl_t* proto, const uint8_t* data, size_t len, cbl_t** out) {
    return guard(mut(proto), [&] { need(proto, "proto"); need(data, "data"); need(out, "out"); *out = deserialize_index(proto, data, len); });
}
int32_t cbl_deserialize_range(const cbl_t* proto, const uint8_t* data, size_t len, uint64_t prefix_lo, uint64_t prefix_hi, cbl_t** out) {
    return guard(mut(proto), [&] { need(proto, "proto"); need(data, "data"); need(out, "out"); *out = deserialize_index(proto, data, len, prefix_lo, prefix_hi); });
}
int32_t cbl_save_to_file(cbl_t* h, const char* path) {
    return guard(h, [&] {
        need(h, "handle"); need(path, "path");
        std::vector<uint8_t> v;
        Writer w{&v};
        serialize_index(h->ix.get(), w);
        std::ofstream f(path, std::ios::binary);
        if (!f) throw Error(CBL_EIO, std::string("Failed to create ") + path);  // src/cbl.rs:128-130
        f.write((const char*)v.data(), (std::streamsize)v.size());
        if (!f) throw Error(CBL_EIO, std::string("Failed to write index to ") + path);
    });
}
int32_t cbl_load_from_file(const cbl_t* proto, const char* path, cbl_t** out) {
    return guard(mut(proto), [&] {
        need(proto, "proto"); need(path, "path"); need(out, "out");
        std::ifstream f(path, std::ios::binary);
        if (!f) throw Error(CBL_EIO, std::string("Failed to open ") + path);  // src/cbl.rs:146-148
        std::vector<uint8_t> v((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        try {
            *out = deserialize_index(proto, v.data(), v.size());
        } catch (const Error& e) {
            throw Error(e.code, std::string("Failed to load index from ") + path + ": " + e.what());
        }
    });
}

int32_t cbl_seq_words_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs, void* d_words) {
    return guard(h, [&] {
        need(h, "handle"); need(d_buf, "d_buf"); need(offsets, "offsets"); need(d_words, "d_words");
        h->ix->seq_words_dev(d_buf, offsets[n_seqs], offsets, n_seqs, d_words, false);
    });
}
int32_t cbl_words_op_dev(cbl_t* h, int32_t op, const void* d_words, size_t n, uint8_t* d_out) {
    return guard(h, [&] {
        need(h, "handle");
        if (n) need(d_words, "d_words");
        if (op < 0 || op > 2) throw Error(CBL_EINVAL, "unknown word operation");
        if (op == 0 && n) need(d_out, "d_out");
        h->ix->words_op_dev(op, d_words, n, d_out);
        h->ix->sync();
    });
}
int32_t cbl_words_op_segments_dev(cbl_t* h, int32_t op, const void* const* seg, const uint64_t* seg_n, uint32_t n_seg) {
    return guard(h, [&] {
        need(h, "handle");
        if (n_seg) { need(seg, "seg"); need(seg_n, "seg_n"); }
        if (op < 1 || op > 2) throw Error(CBL_EINVAL, "words_op_segments: op must be 1 (insert) or 2 (remove)");
        h->ix->words_op_segments_dev(op, seg, seg_n, n_seg);
        h->ix->sync();
    });
}
int32_t cbl_words_contains_segments_dev(cbl_t* h, const void* const* seg, const uint64_t* seg_n, uint8_t* const* seg_out, uint32_t n_seg) {
    return guard(h, [&] {
        need(h, "handle");
        if (n_seg) { need(seg, "seg"); need(seg_n, "seg_n"); need(seg_out, "seg_out"); }
        h->ix->words_contains_segments_dev(seg, seg_n, seg_out, n_seg);
        h->ix->sync();
    });
}
int32_t cbl_export_words_dev(cbl_t* h, uint64_t start, uint64_t count, void* d_out) {
    return guard(h, [&] { need(h, "handle"); if (count) need(d_out, "d_out"); h->ix->export_words_dev(start, count, 0, d_out); h->ix->sync(); });
}
int32_t cbl_route_words_dev(cbl_t* h, const void* d_words, size_t n, const uint32_t* splitters, uint32_t n_splitters, void* d_send,
                            uint32_t* d_pos, uint64_t* counts) {
    return guard(h, [&] {
        need(h, "handle"); need(counts, "counts");
        if (n_splitters) need(splitters, "splitters");
        if (n) { need(d_words, "d_words"); need(d_send, "d_send"); }
        h->ix->route_words_dev(d_words, n, splitters, n_splitters, d_send, d_pos, counts);
    });
}
int32_t cbl_gather_u8_dev(cbl_t* h, const uint8_t* d_src, const uint32_t* d_pos, size_t n, uint8_t* d_out) {
    return guard(h, [&] {
        need(h, "handle");
        if (n) { need(d_src, "d_src"); need(d_pos, "d_pos"); need(d_out, "d_out"); }
        h->ix->gather_u8_dev(d_src, d_pos, n, d_out);
    });
}
int32_t cbl_route_counts_dev(cbl_t* h, const void* d_words, size_t n, const uint32_t* splitters, uint32_t n_splitters, uint64_t* counts) {
    return guard(h, [&] {
        need(h, "handle"); need(counts, "counts");
        if (n_splitters) need(splitters, "splitters");
        if (n) need(d_words, "d_words");
        h->ix->route_counts_dev(d_words, n, splitters, n_splitters, counts);
    });
}
int32_t cbl_route_scatter_dev(cbl_t* h, const void* d_words, size_t n, const uint32_t* splitters, uint32_t n_splitters,
                              void* const* peer_recv, const uint64_t* recv_offset, const uint64_t* counts, uint32_t* d_pos) {
    return guard(h, [&] {
        need(h, "handle"); need(peer_recv, "peer_recv"); need(recv_offset, "recv_offset"); need(counts, "counts");
        if (n_splitters) need(splitters, "splitters");
        if (n) need(d_words, "d_words");
        h->ix->route_scatter_dev(d_words, n, splitters, n_splitters, peer_recv, recv_offset, counts, d_pos);
    });
}
int32_t cbl_seq_route_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs, const uint32_t* splitters,
                          uint32_t n_splitters, void* const* peer_region, uint64_t cap, uint32_t* d_pos, uint64_t* counts) {
    return guard(h, [&] {
        need(h, "handle"); need(d_buf, "d_buf"); need(offsets, "offsets"); need(peer_region, "peer_region"); need(counts, "counts");
        if (n_splitters) need(splitters, "splitters");
        h->ix->seq_route_dev(d_buf, offsets[n_seqs], offsets, n_seqs, splitters, n_splitters, peer_region, cap, d_pos, counts);
    });
}
int32_t cbl_seq_contains_fused_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs, const uint32_t* splitters,
                                   uint32_t n_splitters, void* const* peer_region, void* const* peer_final, uint64_t cap, uint32_t* d_pos,
                                   void* const* recv_region, uint8_t* const* answer_region, const void* const* final_counts, uint32_t epoch,
                                   uint64_t* counts) {
    return guard(h, [&] {
        need(h, "handle"); need(d_buf, "d_buf"); need(offsets, "offsets"); need(peer_region, "peer_region"); need(peer_final, "peer_final");
        need(d_pos, "d_pos"); need(recv_region, "recv_region"); need(answer_region, "answer_region"); need(final_counts, "final_counts");
        need(counts, "counts");
        if (n_splitters) need(splitters, "splitters");
        FusedQuery q;
        q.splitters = splitters; q.n_split = n_splitters;
        q.peer_region = peer_region;
        q.peer_final = reinterpret_cast<unsigned long long* const*>(peer_final);
        q.cap = cap; q.d_pos = d_pos;
        q.recv_region = recv_region; q.answer_region = answer_region;
        q.final_ = reinterpret_cast<const unsigned long long* const*>(final_counts);
        q.epoch = epoch;
        h->ix->seq_contains_fused_dev(d_buf, offsets[n_seqs], offsets, n_seqs, q, counts);
    });
}
int32_t cbl_peer_fill(cbl_t* h, void* d_ptr, int32_t byte, size_t bytes) {
    return guard(h, [&] {
        need(h, "handle");
        CUDA_CHECK(cudaSetDevice(h->ix->config().device));
        if (d_ptr && bytes) { CUDA_CHECK(cudaMemsetAsync(d_ptr, byte, bytes, h->ix->stream())); CUDA_CHECK(cudaStreamSynchronize(h->ix->stream())); }
    });
}
int32_t cbl_peer_zero(cbl_t* h, void* d_ptr, size_t bytes) { return cbl_peer_fill(h, d_ptr, 0, bytes); }
// ---- peer memory: plain cudaMalloc blocks shared between the processes of one box with CUDA IPC ----
int32_t cbl_peer_alloc(cbl_t* h, size_t bytes, void** d_ptr, uint8_t* handle) {
    return guard(h, [&] {
        need(h, "handle"); need(d_ptr, "d_ptr"); need(handle, "handle bytes");
        static_assert(sizeof(cudaIpcMemHandle_t) == CBL_IPC_HANDLE_BYTES, "IPC handle size");
        CUDA_CHECK(cudaSetDevice(h->ix->config().device));
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1));
        cudaIpcMemHandle_t mh;
        cudaError_t e = cudaIpcGetMemHandle(&mh, p);
        if (e != cudaSuccess) { cudaFree(p); CUDA_CHECK(e); }
        memcpy(handle, &mh, sizeof mh);
        *d_ptr = p;
    });
}
int32_t cbl_peer_open(cbl_t* h, const uint8_t* handle, void** d_ptr) {
    return guard(h, [&] {
        need(h, "handle"); need(d_ptr, "d_ptr"); need(handle, "handle bytes");
        CUDA_CHECK(cudaSetDevice(h->ix->config().device));
        cudaIpcMemHandle_t mh;
        memcpy(&mh, handle, sizeof mh);
        void* p = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        *d_ptr = p;
    });
}
int32_t cbl_peer_close(cbl_t* h, void* d_ptr) {
    return guard(h, [&] { need(h, "handle"); CUDA_CHECK(cudaSetDevice(h->ix->config().device)); if (d_ptr) CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr)); });
}
int32_t cbl_peer_free(cbl_t* h, void* d_ptr) {
    return guard(h, [&] { need(h, "handle"); CUDA_CHECK(cudaSetDevice(h->ix->config().device)); if (d_ptr) CUDA_CHECK(cudaFree(d_ptr)); });
}
int32_t cbl_word_bytes(const cbl_t* h, int32_t* out) {
    return guard(mut(h), [&] {
        need(h, "handle"); need(out, "out");
        const KParams& P = h->ix->params();
        *out = (P.bits + P.pos_bits) <= 64 ? 8 : 16;
    });
}
int32_t cbl_suffix_bits(const cbl_t* h, int32_t* out) {
    return guard(mut(h), [&] { need(h, "handle"); need(out, "out"); *out = h->ix->params().suffix_bits; });
}

int32_t cbl_seq_words(cbl_t* h, const uint8_t* buf, const uint64_t* offsets, size_t n_seqs, uint64_t* lo, uint64_t* hi, int32_t brute) {
    return guard(h, [&] {
        need(h, "handle"); need(buf, "buf"); need(offsets, "offsets"); need(lo, "lo");
        h->ix->seq_words(buf, offsets, n_seqs, lo, hi, brute != 0);
    });
}
int32_t cbl_sync(cbl_t* h) { return guard(h, [&] { need(h, "handle"); h->ix->sync(); }); }
void* cbl_stream(const cbl_t* h) { return h ? (void*)h->ix->stream() : nullptr; }
uint64_t cbl_launch_count(void) { return g_launches.load(); }
uint64_t cbl_sort_fallback_count(void) { return g_sort_fallbacks.load(); }
void cbl_profile_enable(int32_t on) { g_prof_on.store(on ? 1 : 0); }
int32_t cbl_profile_report(char* out, size_t cap) {
    return guard(nullptr, [&] {
        need(out, "out");
        std::string r = prof_report();
        if (r.size() + 1 > cap) throw Error(CBL_EINVAL, "output buffer too small");
        memcpy(out, r.c_str(), r.size() + 1);
    });
}
int32_t cbl_set_sort_concentration(cbl_t* h, double factor) {
    return guard(h, [&] {
        need(h, "handle");
        h->ix->set_sort_concentration(factor);
    });
}
int32_t cbl_mem_trim(int32_t device) {
    return guard(nullptr, [&] { CUDA_CHECK(cudaSetDevice(device)); arena::trim(); });
}
uint64_t cbl_mem_cached_bytes(void) { return arena::cached_bytes(); }
const char* cbl_build_info(void) { return "libcbl_gpu sm_100a (cuda " CBL_STR(CUDART_VERSION) ")"; }

}  // extern "C"
