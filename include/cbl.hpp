// cbl.hpp — header-only C++ facade over the C ABI (cbl_gpu.h) mirroring the reference's public type
// CBL<K, T, PREFIX_BITS> (src/cbl.rs:40-569): same method names, argument meaning and failure
// behaviour (where the Rust code panics this throws cbl::Panic carrying the same message).
// The reference is a compiled (Rust) library whose toolchain is absent from the build image, so this
// is the compiled-language host side above the boundary; INTEGRATION.md shows the Rust binding.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "cbl_gpu.h"

namespace cbl {

struct Panic : std::runtime_error {
    int32_t code;
    Panic(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

using u128 = unsigned __int128;

// K, T (an unsigned integer type of 32/64/128 bits), PREFIX_BITS as in CBL::<K, T, PREFIX_BITS>
template <unsigned K, class T, unsigned PREFIX_BITS = 24>
class CBL {
    cbl_t* h_ = nullptr;
    explicit CBL(cbl_t* h) : h_(h) {}
    void chk(int32_t rc) const { if (rc) throw Panic(rc, cbl_last_error(h_)); }
    static void split(T kmer, uint64_t& lo, uint64_t& hi) {
        lo = (uint64_t)kmer;
        if constexpr (sizeof(T) > 8) hi = (uint64_t)(kmer >> 64); else hi = 0;
    }
    static T join(uint64_t lo, uint64_t hi) {
        if constexpr (sizeof(T) > 8) return ((T)hi << 64) | (T)lo; else { (void)hi; return (T)lo; }
    }

public:
    using Kmer = T;  // IntKmer<K, T>: first base most significant, A=0 C=1 T=2 G=3 (src/kmer.rs:11)

    explicit CBL(bool canonical = false, int device = 0) {  // new() / new_canonical()  (src/cbl.rs:71-79)
        int32_t rc = cbl_create(K, sizeof(T) * 8, PREFIX_BITS, canonical, device, &h_);
        if (rc) throw Panic(rc, cbl_last_global_error());
    }
    static CBL new_canonical(int device = 0) { return CBL(true, device); }
    // the same set prefix-sharded over several GPUs of this process (cbl_create_sharded): every method below works unchanged
    static CBL sharded(const std::vector<int>& devices, bool canonical = false) {
        cbl_t* h = nullptr;
        std::vector<int32_t> d(devices.begin(), devices.end());
        int32_t rc = cbl_create_sharded(K, sizeof(T) * 8, PREFIX_BITS, canonical, (int32_t)d.size(), d.data(), &h);
        if (rc) throw Panic(rc, cbl_last_global_error());
        return CBL(h);
    }
    ~CBL() { if (h_) cbl_destroy(h_); }
    CBL(CBL&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    CBL& operator=(CBL&& o) noexcept { if (this != &o) { if (h_) cbl_destroy(h_); h_ = o.h_; o.h_ = nullptr; } return *this; }
    CBL(const CBL& o) { int32_t rc = cbl_clone(o.h_, &h_); if (rc) throw Panic(rc, cbl_last_error(o.h_)); }  // Clone
    CBL& operator=(const CBL& o) { if (this != &o) { CBL t(o); std::swap(h_, t.h_); } return *this; }
    cbl_t* handle() const { return h_; }
    // Default-like: an empty set laid out like this one (same canonical flag, device(s) and splitters)
    CBL new_like() const {
        CBL c(*this);
        c -= c;   // x - x = {}  (src/cbl.rs:488-510)
        return c;
    }

    bool is_canonical() const { int32_t v; chk(cbl_is_canonical(h_, &v)); return v != 0; }
    size_t count() const { uint64_t v; chk(cbl_count(h_, &v)); return (size_t)v; }
    bool is_empty() const { int32_t v; chk(cbl_is_empty(h_, &v)); return v != 0; }

    // src/cbl.rs:219-235
    bool contains(Kmer kmer) const { uint64_t lo, hi; split(kmer, lo, hi); uint8_t f; chk(cbl_contains_kmers(h_, &lo, &hi, 1, &f)); return f; }
    bool insert(Kmer kmer) { uint64_t lo, hi; split(kmer, lo, hi); uint8_t f; chk(cbl_insert_kmers(h_, &lo, &hi, 1, &f)); return !f; }
    bool remove(Kmer kmer) { uint64_t lo, hi; split(kmer, lo, hi); uint8_t f; chk(cbl_remove_kmers(h_, &lo, &hi, 1, &f)); return f; }

    // src/cbl.rs:293-354
    bool contains_all(const uint8_t* seq, size_t len) { int32_t v; chk(cbl_contains_all(h_, seq, len, &v)); return v != 0; }
    std::vector<uint8_t> contains_seq(const uint8_t* seq, size_t len) {
        std::vector<uint8_t> out(len >= K ? len - K + 1 : 1);
        size_t n = 0;
        chk(cbl_contains_seq(h_, seq, len, out.data(), &n));
        out.resize(n);
        return out;
    }
    void insert_seq(const uint8_t* seq, size_t len) { chk(cbl_insert_seq(h_, seq, len)); }
    void remove_seq(const uint8_t* seq, size_t len) { chk(cbl_remove_seq(h_, seq, len)); }
    // whole record loops in one call (examples/cbl.rs:160-163)
    void insert_seqs(const uint8_t* buf, const uint64_t* offsets, size_t n) { chk(cbl_insert_seqs(h_, buf, offsets, n)); }
    void remove_seqs(const uint8_t* buf, const uint64_t* offsets, size_t n) { chk(cbl_remove_seqs(h_, buf, offsets, n)); }
    void contains_seqs(const uint8_t* buf, const uint64_t* offsets, size_t n, uint8_t* out) { chk(cbl_contains_seqs(h_, buf, offsets, n, out)); }

    // src/cbl.rs:358-360 — ascending word order
    std::vector<Kmer> iter() const {
        std::vector<Kmer> out;
        const size_t CH = 1 << 20;
        std::vector<uint64_t> lo(CH), hi(CH);
        for (uint64_t start = 0;;) {
            size_t n = 0;
            chk(cbl_export_kmers(h_, start, lo.data(), hi.data(), CH, &n));
            if (!n) break;
            for (size_t i = 0; i < n; i++) out.push_back(join(lo[i], hi[i]));
            start += n;
        }
        return out;
    }
    std::vector<std::pair<size_t, size_t>> buckets_sizes() const {  // src/cbl.rs:370-372
        size_t n = 0;
        chk(cbl_bucket_sizes(h_, nullptr, nullptr, 0, &n));
        std::vector<uint32_t> p(n ? n : 1), s(n ? n : 1);
        if (n) chk(cbl_bucket_sizes(h_, p.data(), s.data(), n, &n));
        std::vector<std::pair<size_t, size_t>> r(n);
        for (size_t i = 0; i < n; i++) r[i] = {p[i], s[i]};
        return r;
    }
    // src/cbl.rs:374-396: bucket size -> number of buckets / share of the stored k-mers
    std::map<size_t, size_t> buckets_size_count() const {
        std::map<size_t, size_t> m;
        for (auto& ps : buckets_sizes()) m[ps.second]++;
        return m;
    }
    std::map<size_t, double> buckets_load_repartition() const {
        const auto sc = buckets_size_count();
        double total = 0;
        for (auto& kv : sc) total += (double)(kv.first * kv.second);
        std::map<size_t, double> r;
        for (auto& kv : sc) r[kv.first] = total > 0 ? (double)(kv.first * kv.second) / total : 0.0;
        return r;
    }
    double prefix_load() const { uint64_t nb; chk(cbl_num_buckets(h_, &nb)); return (double)nb / (double)(1ull << PREFIX_BITS); }

    // src/cbl.rs:411-569
    friend CBL operator|(CBL& a, CBL& b) { return a.binary(CBL_OP_OR, b); }
    friend CBL operator&(CBL& a, CBL& b) { return a.binary(CBL_OP_AND, b); }
    friend CBL operator-(CBL& a, CBL& b) { return a.binary(CBL_OP_SUB, b); }
    friend CBL operator^(CBL& a, CBL& b) { return a.binary(CBL_OP_XOR, b); }
    CBL& operator|=(CBL& o) { chk(cbl_setop_assign(CBL_OP_OR, h_, o.h_)); return *this; }
    CBL& operator&=(CBL& o) { chk(cbl_setop_assign(CBL_OP_AND, h_, o.h_)); return *this; }
    CBL& operator-=(CBL& o) { chk(cbl_setop_assign(CBL_OP_SUB, h_, o.h_)); return *this; }
    CBL& operator^=(CBL& o) { chk(cbl_setop_assign(CBL_OP_XOR, h_, o.h_)); return *this; }
    // src/cbl.rs:108-124
    static CBL merge(const std::vector<CBL*>& cbls) { return many(cbls, false); }
    static CBL intersect(const std::vector<CBL*>& cbls) { return many(cbls, true); }

    // src/cbl.rs:127-160
    void save_to_file(const char* path) { chk(cbl_save_to_file(h_, path)); }
    static CBL load_from_file(const char* path, int device = 0) {
        CBL proto(false, device);
        return proto.load_like(path);
    }
    // load into a set laid out like this one (same device, or same devices + splitters for a sharded set)
    CBL load_like(const char* path) const {
        cbl_t* out = nullptr;
        int32_t rc = cbl_load_from_file(h_, path, &out);
        if (rc) throw Panic(rc, cbl_last_error(h_));
        return CBL(out);
    }

private:
    CBL binary(int op, CBL& o) {
        cbl_t* out = nullptr;
        chk(cbl_setop(op, h_, o.h_, &out));
        return CBL(out);
    }
    static CBL many(const std::vector<CBL*>& cbls, bool inter) {
        if (cbls.empty()) throw Panic(CBL_EINVAL, "empty list of indexes");
        std::vector<cbl_t*> hs;
        for (auto* c : cbls) hs.push_back(c->h_);
        cbl_t* out = nullptr;
        int32_t rc = inter ? cbl_intersect_many(hs.data(), hs.size(), &out) : cbl_merge_many(hs.data(), hs.size(), &out);
        if (rc) throw Panic(rc, cbl_last_error(hs[0]));
        return CBL(out);
    }
};

}  // namespace cbl
