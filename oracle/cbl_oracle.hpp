// CPU ORACLE — TEST INFRASTRUCTURE ONLY.
//
// A C++17 restatement of the reference's (imartayan/CBL @ e6ca8a4) batched sequence path.
// It exists to CHECK the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
// `--impl reference` leg).  Nothing under cbl_b200/ may include, link or call it.
//
// Every block cites the reference file:line it follows (paths relative to /root/reference).
// The Rust half of the reference cannot be compiled in this image (no rustc/cargo), so the Rust
// logic is restated here; the reference's own C++ half (cxx/rank_bv.h, cxx/tiered_vec.h with the
// vendored sux + tiered-vector) is LINKED, not restated, when built with -DORACLE_USE_REFERENCE_CXX
// (oracle/Makefile target `ref`, output oracle/_ref/).  Without that flag two small stand-ins with
// the same observable behaviour are used so the oracle also builds where /root/reference is absent.
//
// Parity pinning: the known-answer vectors the reference's own unit tests hold for this path
// (tests/golden/reference_kats.json) are replayed against this file by tests/test_oracle_kats.py.
// The bincode file layout (serde) has no reference test at all => that part is "parity unpinned".
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#ifdef ORACLE_USE_REFERENCE_CXX
#include "rank_bv.h"     // /root/reference/cxx/rank_bv.h   (sux WordDynRankSel<FenwickByteL>)
#include "tiered_vec.h"  // /root/reference/cxx/tiered_vec.h (Seq::Tiered, Layer32)
#endif

namespace orc {

using u128 = unsigned __int128;

// ---------------------------------------------------------------------------------------------
// Integer helpers (stand in for num-traits PrimInt on u32/u64/u128)
// ---------------------------------------------------------------------------------------------
template <class T> constexpr int type_bits() { return int(sizeof(T) * 8); }

template <class T> inline int popcount_t(T x) {
    if constexpr (sizeof(T) <= 8) return __builtin_popcountll((unsigned long long)x);
    else return __builtin_popcountll((uint64_t)x) + __builtin_popcountll((uint64_t)(x >> 64));
}

template <class T> inline T swap_bytes_t(T x) {
    if constexpr (sizeof(T) == 4) return (T)__builtin_bswap32((uint32_t)x);
    else if constexpr (sizeof(T) == 8) return (T)__builtin_bswap64((uint64_t)x);
    else {
        uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
        return ((u128)__builtin_bswap64(lo) << 64) | (u128)__builtin_bswap64(hi);
    }
}

template <class T> inline T rep_byte(uint8_t b) {
    T r = 0;
    for (size_t i = 0; i < sizeof(T); i++) r = (T)((r << 8) | b);
    return r;
}

// ---------------------------------------------------------------------------------------------
// src/kmer.rs:11-24,206-225 — nucleotide <-> 2-bit code.  A=0 C=1 T=2 G=3, lower case accepted,
// everything else is "None" (returned here as -1).
// ---------------------------------------------------------------------------------------------
inline int from_nuc(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'T': case 't': return 2;
        case 'G': case 'g': return 3;
        default: return -1;
    }
}
inline uint8_t to_nuc(int code) { static const char L[4] = {'A', 'C', 'T', 'G'}; return (uint8_t)L[code & 3]; }
inline int complement(int code) { return code ^ 0b10; }  // src/kmer.rs:218-220

// ---------------------------------------------------------------------------------------------
// src/kmer.rs:41-160,200-240,293-348 — IntKmer<K,T> with K as a run-time value.
// ---------------------------------------------------------------------------------------------
template <class T>
struct KmerOps {
    int K;
    T mask;  // src/kmer.rs:228  MASK = (1 << 2K) - 1
    explicit KmerOps(int k) : K(k) {
        mask = (2 * k >= type_bits<T>()) ? (T)~(T)0 : (T)(((T)1 << (2 * k)) - 1);
    }
    T extend(T s, T base) const { return (T)((s << 2) | base); }            // :61-63
    T append(T s, T base) const { return (T)(((s << 2) | base) & mask); }   // :70-72
    // :110-112,133-135  fold of the valid bases among the given bytes (at most K of them)
    T from_nucs(const uint8_t* nucs, size_t n) const {
        T s = 0;
        int taken = 0;
        for (size_t i = 0; i < n && taken < K; i++) {
            int c = from_nuc(nucs[i]);
            if (c < 0) continue;
            s = extend(s, (T)c);
            taken++;
        }
        return s;
    }
    void to_nucs(T s, uint8_t* out) const {  // :121-129,139-141
        for (int i = 0; i < K; i++) { out[K - 1 - i] = to_nuc((int)(s & 3)); s >>= 2; }
    }
    bool is_canonical(T s) const { return popcount_t(s) % 2 == 0; }         // :93-95
    // :327-348 (x86 variant): byte swap, nibble swap, 2-bit swap, xor 0xAA.., shift down
    T rev_comp(T s) const {
        T res = swap_bytes_t<T>(s);
        const T m4 = rep_byte<T>(0x0F), m2 = rep_byte<T>(0x33), aa = rep_byte<T>(0xAA);
        res = (T)(((res >> 4) & m4) | ((res & m4) << 4));
        res = (T)(((res >> 2) & m2) | ((res & m2) << 2));
        res ^= aa;
        return (T)(res >> (2 * (type_bits<T>() / 2 - K)));
    }
    T canonical(T s) const { return is_canonical(s) ? s : rev_comp(s); }     // :99-106
};

// ---------------------------------------------------------------------------------------------
// src/necklace/mod.rs:13-31 — normative brute-force necklace and its inverse.
// ---------------------------------------------------------------------------------------------
template <class T>
inline std::pair<T, size_t> necklace_pos(T word, int BITS) {
    T necklace = word, rot = word;
    size_t pos = 0;
    for (int i = BITS - 1; i >= 0; i--) {
        rot = (T)(((rot & (T)1) << (BITS - 1)) | (rot >> 1));
        if (rot <= necklace) { necklace = rot; pos = (size_t)i; }
    }
    return {necklace, pos};
}
template <class T>
inline T revert_necklace_pos(T necklace, size_t pos, int BITS) {
    T mask = (T)(((T)1 << BITS) - 1);
    // pos == 0 shifts by BITS (< T::BITS for every valid K), as in the reference
    return (T)(((necklace << (BITS - pos)) & mask) | (necklace >> pos));
}

// ---------------------------------------------------------------------------------------------
// A small ring deque standing in for std::collections::VecDeque (only the calls the reference
// makes: push_front/push_back/pop_front/truncate/index/len/clear).
// ---------------------------------------------------------------------------------------------
template <class V>
class RingDeque {
    std::vector<V> buf_;
    size_t head_ = 0, len_ = 0, capmask_;
public:
    explicit RingDeque(size_t cap_hint = 8) {
        size_t c = 8;
        while (c < cap_hint + 4) c <<= 1;
        buf_.resize(c);
        capmask_ = c - 1;
    }
    size_t len() const { return len_; }
    bool is_empty() const { return len_ == 0; }
    void clear() { head_ = 0; len_ = 0; }
    void grow() {
        std::vector<V> nb(buf_.size() * 2);
        for (size_t i = 0; i < len_; i++) nb[i] = buf_[(head_ + i) & capmask_];
        buf_.swap(nb);
        capmask_ = buf_.size() - 1;
        head_ = 0;
    }
    void push_back(const V& v) { if (len_ == buf_.size()) grow(); buf_[(head_ + len_) & capmask_] = v; len_++; }
    void push_front(const V& v) { if (len_ == buf_.size()) grow(); head_ = (head_ + capmask_) & capmask_; buf_[head_] = v; len_++; }
    void pop_front() { if (len_) { head_ = (head_ + 1) & capmask_; len_--; } }
    void truncate(size_t n) { if (n < len_) len_ = n; }
    const V& operator[](size_t i) const { return buf_[(head_ + i) & capmask_]; }
};

// ---------------------------------------------------------------------------------------------
// src/necklace/minimizer.rs:5-92 — LexMinQueue<WIDTH,T>: monotone deque keeping every position that
// holds the window minimum (ties kept).
// ---------------------------------------------------------------------------------------------
template <class T>
class LexMinQueue {
    size_t WIDTH;
    RingDeque<std::pair<T, size_t>> deq;
    RingDeque<size_t> min_pos;
    size_t pos = 0;
    void refill_min_pos() {  // minimizer.rs:36-39,55-58,85-88
        while (min_pos.len() < deq.len() && deq[min_pos.len()].first == deq[0].first)
            min_pos.push_back(deq[min_pos.len()].second);
    }
public:
    explicit LexMinQueue(size_t width) : WIDTH(width), deq(width), min_pos(width) {}
    size_t width() const { return WIDTH; }
    size_t n_min() const { return min_pos.len(); }
    // minimizer.rs:17-21
    size_t min_pos_at(size_t i) const { return (min_pos[i] + WIDTH - pos) % WIDTH; }
    std::vector<size_t> iter_min_pos() const {
        std::vector<size_t> r;
        for (size_t i = 0; i < min_pos.len(); i++) r.push_back(min_pos_at(i));
        return r;
    }
    // minimizer.rs:23-40 — vals given oldest-first, consumed from the back
    void insert_full(const std::vector<T>& vals) {
        deq.clear();
        min_pos.clear();
        size_t n = vals.size();
        T minimizer = vals[n - 1];
        size_t p = (pos + WIDTH - 1) % WIDTH;
        deq.push_front({minimizer, p});
        size_t taken = 0;
        for (size_t idx = n - 1; idx-- > 0 && taken < WIDTH - 1; taken++) {
            T u = vals[idx];
            p = (p + WIDTH - 1) % WIDTH;
            if (u <= minimizer) { minimizer = u; deq.push_front({minimizer, p}); }
        }
        refill_min_pos();
    }
    // minimizer.rs:42-60
    void insert(T u) {
        if (!deq.is_empty() && deq[0].second == pos) { deq.pop_front(); min_pos.pop_front(); }
        size_t i = deq.len();
        while (i > 0 && deq[i - 1].first > u) i--;
        deq.truncate(i);
        min_pos.truncate(i);
        deq.push_back({u, pos});
        refill_min_pos();
        pos = (pos + 1) % WIDTH;
    }
    // minimizer.rs:62-91
    void insert2(T u, T v) {
        size_t next_pos = (pos + 1) % WIDTH;
        if (!deq.is_empty() && deq[0].second == pos) { deq.pop_front(); min_pos.pop_front(); }
        if (!deq.is_empty() && deq[0].second == next_pos) { deq.pop_front(); min_pos.pop_front(); }
        T w = std::min(u, v);
        size_t i = deq.len();
        while (i > 0 && deq[i - 1].first > w) i--;
        deq.truncate(i);
        min_pos.truncate(i);
        if (u <= v) deq.push_back({u, pos});
        deq.push_back({v, next_pos});
        refill_min_pos();
        pos = (next_pos + 1) % WIDTH;
    }
};

// ---------------------------------------------------------------------------------------------
// src/necklace/queue.rs:14-118 — streaming NecklaceQueue<BITS,T,WIDTH,REVERSE>.
// ---------------------------------------------------------------------------------------------
template <class T, bool REVERSE>
class NecklaceQueue {
    int BITS;
    size_t WIDTH;
    int M;       // queue.rs:26  M = BITS - WIDTH + 1
    T MASK;      // queue.rs:27
    T MIN_MASK;  // queue.rs:28
    T word = 0;
    LexMinQueue<T> min_queue;
    T rotation(size_t p) const {  // queue.rs:48-50
        if (p == 0) return (T)(word & MASK);  // (word >> BITS) is 0 for BITS < T::BITS
        return (T)(((word << p) & MASK) | (word >> (BITS - p)));
    }
public:
    NecklaceQueue(int bits, size_t width) : BITS(bits), WIDTH(width), min_queue(width) {
        M = BITS - (int)WIDTH + 1;
        MASK = (T)(((T)1 << BITS) - 1);
        MIN_MASK = (T)(((T)1 << M) - 1);
    }
    // queue.rs:53-79
    std::pair<T, size_t> get_necklace_pos() const {
        std::pair<T, size_t> best{(T)0, 0};
        bool have = false;
        for (size_t i = 0; i < min_queue.n_min(); i++) {
            size_t q = min_queue.min_pos_at(i);
            size_t p = REVERSE ? (WIDTH - 1 - q) : q;
            std::pair<T, size_t> c{rotation(p), p};
            if (!have || c < best) { best = c; have = true; }
        }
        for (size_t p = WIDTH; p < (size_t)BITS; p++) {
            std::pair<T, size_t> c{rotation(p), p};
            if (!have || c < best) { best = c; have = true; }
        }
        return best;
    }
    // queue.rs:82-96
    void insert_full(T w) {
        word = (T)(w & MASK);
        std::vector<T> vals(WIDTH);
        for (size_t p = 0; p < WIDTH; p++)
            vals[p] = REVERSE ? (T)((w >> p) & MIN_MASK) : (T)((w >> (BITS - (int)p - M)) & MIN_MASK);
        min_queue.insert_full(vals);
    }
    // queue.rs:99-107
    void insert(T x) {
        if (REVERSE) {
            word = (T)((word >> 1) | ((x & 1) << (BITS - 1)));
            min_queue.insert((T)(word >> (WIDTH - 1)));
        } else {
            word = (T)(((word << 1) & MASK) | (x & 1));
            min_queue.insert((T)(word & MIN_MASK));
        }
    }
    // queue.rs:110-118
    void insert2(T x) {
        if (REVERSE) {
            word = (T)((word >> 2) | ((x & 3) << (BITS - 2)));
            min_queue.insert2((T)((word >> (WIDTH - 2)) & MIN_MASK), (T)(word >> (WIDTH - 1)));
        } else {
            word = (T)(((word << 2) & MASK) | (x & 3));
            min_queue.insert2((T)((word >> 1) & MIN_MASK), (T)(word & MIN_MASK));
        }
    }
    T current_word() const { return word; }
    const LexMinQueue<T>& queue() const { return min_queue; }
};

// ---------------------------------------------------------------------------------------------
// src/sliced_int.rs:12-101 — SlicedInt<BYTES>: little-endian packed suffix, numeric Ord
// (most significant byte first).
// ---------------------------------------------------------------------------------------------
template <int BYTES>
struct SlicedInt {
    std::array<uint8_t, BYTES> b{};
    static SlicedInt from_int(u128 v) {  // :21-25,65-75
        SlicedInt r;
        for (int i = 0; i < BYTES; i++) r.b[i] = (uint8_t)(v >> (8 * i));
        return r;
    }
    u128 get() const {  // :54-62
        u128 v = 0;
        for (int i = 0; i < BYTES; i++) v |= (u128)b[i] << (8 * i);
        return v;
    }
    std::array<uint8_t, BYTES> to_be_bytes() const {  // :47-52
        std::array<uint8_t, BYTES> r;
        for (int i = 0; i < BYTES; i++) r[i] = b[BYTES - 1 - i];
        return r;
    }
    static SlicedInt from_be_bytes(const uint8_t* be) {  // :36-42
        SlicedInt r;
        for (int i = 0; i < BYTES; i++) r.b[i] = be[BYTES - 1 - i];
        return r;
    }
    bool operator==(const SlicedInt& o) const { return b == o.b; }
    bool operator!=(const SlicedInt& o) const { return !(b == o.b); }
    int cmp(const SlicedInt& o) const {  // :90-101
        for (int i = BYTES - 1; i >= 0; i--) {
            if (b[i] != o.b[i]) return b[i] < o.b[i] ? -1 : 1;
        }
        return 0;
    }
    bool operator<(const SlicedInt& o) const { return cmp(o) < 0; }
    bool operator>(const SlicedInt& o) const { return cmp(o) > 0; }
};

// ---------------------------------------------------------------------------------------------
// src/bitvector/tiny/mod.rs:10-95 — 256-bit set with rank and ascending iteration.
// ---------------------------------------------------------------------------------------------
struct TinyBitvector {
    uint64_t w[4] = {0, 0, 0, 0};
    bool is_empty() const { return !(w[0] | w[1] | w[2] | w[3]); }
    size_t count() const {
        return (size_t)(__builtin_popcountll(w[0]) + __builtin_popcountll(w[1]) + __builtin_popcountll(w[2]) +
                        __builtin_popcountll(w[3]));
    }
    bool contains(uint8_t i) const { return (w[i / 64] >> (i % 64)) & 1; }
    bool insert(uint8_t i) { uint64_t o = w[i / 64]; w[i / 64] = o | (1ULL << (i % 64)); return w[i / 64] != o; }
    bool remove(uint8_t i) { uint64_t o = w[i / 64]; w[i / 64] = o & ~(1ULL << (i % 64)); return w[i / 64] != o; }
    size_t rank(uint8_t i) const {  // exclusive
        size_t r = (size_t)__builtin_popcountll(w[i / 64] & ((1ULL << (i % 64)) - 1));
        for (int k = 0; k < i / 64; k++) r += (size_t)__builtin_popcountll(w[k]);
        return r;
    }
    std::vector<uint8_t> indices() const {  // :71-95 ascending
        std::vector<uint8_t> r;
        for (int k = 0; k < 4; k++) {
            uint64_t blk = w[k];
            while (blk) { int t = __builtin_ctzll(blk); blk &= blk - 1; r.push_back((uint8_t)(k * 64 + t)); }
        }
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// src/trie.rs:8-221 — 256-ary byte trie over big-endian suffix bytes.
// ---------------------------------------------------------------------------------------------
template <int BYTES>
struct TrieNode {
    TinyBitvector bv;
    std::vector<std::unique_ptr<TrieNode>> children;

    std::unique_ptr<TrieNode> clone() const {
        auto n = std::make_unique<TrieNode>();
        n->bv = bv;
        n->children.reserve(children.size());
        for (auto& c : children) n->children.push_back(c->clone());
        return n;
    }
    bool is_empty() const { return bv.is_empty(); }
    size_t count() const {  // :74-89
        if (children.empty()) return bv.count();
        size_t c = 0;
        for (auto& ch : children) c += ch->count();
        return c;
    }
    size_t count_nodes() const {  // :91-102
        size_t c = 1;
        for (auto& ch : children) c += ch->count_nodes();
        return c;
    }
    bool contains(const uint8_t* bytes) const {  // :104-116
        const TrieNode* t = this;
        for (int i = 0; i < BYTES - 1; i++) {
            if (!t->bv.contains(bytes[i])) return false;
            t = t->children[t->bv.rank(bytes[i])].get();
        }
        return t->bv.contains(bytes[BYTES - 1]);
    }
    bool insert(const uint8_t* bytes) {  // :118-131
        TrieNode* t = this;
        for (int i = 0; i < BYTES - 1; i++) {
            bool absent = t->bv.insert(bytes[i]);
            size_t r = t->bv.rank(bytes[i]);
            if (absent) t->children.insert(t->children.begin() + (ptrdiff_t)r, std::make_unique<TrieNode>());
            t = t->children[r].get();
        }
        return t->bv.insert(bytes[BYTES - 1]);
    }
    bool remove(const uint8_t* bytes) {  // :133-162
        TrieNode* t = this;
        std::vector<TrieNode*> parents;
        for (int i = 0; i < BYTES - 1; i++) {
            if (!t->bv.contains(bytes[i])) return false;
            size_t r = t->bv.rank(bytes[i]);
            parents.push_back(t);
            t = t->children[r].get();
        }
        t->bv.remove(bytes[BYTES - 1]);  // NB: the reference returns true even if the last byte was absent
        if (!t->bv.is_empty()) return true;
        for (int i = BYTES - 2; i >= 0; i--) {
            t = parents.back();
            parents.pop_back();
            size_t r = t->bv.rank(bytes[i]);
            t->children.erase(t->children.begin() + (ptrdiff_t)r);
            t->bv.remove(bytes[i]);
            if (!t->bv.is_empty()) return true;
        }
        return true;
    }
    // :176-221 ascending DFS
    template <class F> void for_each(std::array<uint8_t, BYTES>& word, int depth, F&& f) const {
        if (depth == BYTES - 1 || children.empty()) {
            // leaf level (children empty): emit one word per set index
            for (uint8_t idx : bv.indices()) { word[depth] = idx; f(word); }
            return;
        }
        size_t r = 0;
        for (uint8_t idx : bv.indices()) {
            word[depth] = idx;
            children[r++]->for_each(word, depth + 1, f);
        }
    }
};

template <int BYTES>
struct Trie {
    std::unique_ptr<TrieNode<BYTES>> root = std::make_unique<TrieNode<BYTES>>();
    Trie() = default;
    Trie(const Trie& o) : root(o.root->clone()) {}
    Trie& operator=(const Trie& o) { root = o.root->clone(); return *this; }
    Trie(Trie&&) = default;
    Trie& operator=(Trie&&) = default;
    bool is_empty() const { return root->is_empty(); }
    size_t count() const { return root->count(); }
    size_t count_nodes() const { return root->count_nodes(); }
    bool contains(const uint8_t* b) const { return root->contains(b); }
    bool insert(const uint8_t* b) { return root->insert(b); }
    bool remove(const uint8_t* b) { return root->remove(b); }
    template <class F> void for_each(F&& f) const {
        std::array<uint8_t, BYTES> w{};
        if (BYTES == 1) { for (uint8_t idx : root->bv.indices()) { w[0] = idx; f(w); } return; }
        root->for_each(w, 0, f);
    }
};

// ---------------------------------------------------------------------------------------------
// src/trievec/mod.rs:8-220 — bucket container: unsorted Vec (linear contains) or Trie + len.
// ---------------------------------------------------------------------------------------------
template <int BYTES>
class TrieVec {
public:
    using S = SlicedInt<BYTES>;
    bool is_trie = false;
    std::vector<S> vec;
    Trie<BYTES> trie;
    size_t trie_len = 0;

    TrieVec() = default;
    static TrieVec from_vec(std::vector<S> v) { TrieVec t; t.vec = std::move(v); return t; }
    size_t len() const { return is_trie ? trie_len : vec.size(); }
    size_t count_nodes() const { return is_trie ? trie.count_nodes() : vec.size(); }
    bool is_empty() const { return len() == 0; }
    void clear() { vec.clear(); if (is_trie) { is_trie = false; trie = Trie<BYTES>(); trie_len = 0; } }  // :53-62
    bool contains(const S& x) const {  // :64-70
        if (is_trie) { auto be = x.to_be_bytes(); return trie.contains(be.data()); }
        return std::find(vec.begin(), vec.end(), x) != vec.end();
    }
    bool insert(const S& x) {  // :72-90
        if (is_trie) {
            auto be = x.to_be_bytes();
            bool absent = trie.insert(be.data());
            if (absent) trie_len++;
            return absent;
        }
        if (std::find(vec.begin(), vec.end(), x) == vec.end()) { vec.push_back(x); return true; }
        return false;
    }
    bool remove(const S& x) {  // :92-108
        if (is_trie) {
            auto be = x.to_be_bytes();
            bool present = trie.remove(be.data());
            if (present) trie_len--;
            return present;
        }
        auto it = std::find(vec.begin(), vec.end(), x);
        if (it != vec.end()) { *it = vec.back(); vec.pop_back(); return true; }  // swap_remove
        return false;
    }
    void insert_sorted_iter(const std::vector<S>& it) {  // :117-136
        if (is_trie) { for (auto& x : it) insert(x); return; }
        size_t stop = vec.size(), i = 0;
        for (auto& x : it) {
            while (i < stop && x > vec[i]) i++;
            if (i == stop || x < vec[i]) vec.push_back(x);
        }
    }
    void remove_sorted_iter(const std::vector<S>& it) {  // :145-168
        if (is_trie) { for (auto& x : it) remove(x); return; }
        size_t stop = vec.size(), i = 0;
        std::vector<size_t> deletions;
        for (auto& x : it) {
            while (i < stop && x > vec[i]) i++;
            if (i < stop && x == vec[i]) deletions.push_back(i);
        }
        for (size_t k = deletions.size(); k-- > 0;) { vec[deletions[k]] = vec.back(); vec.pop_back(); }
    }
    void as_trie() {  // :170-178
        if (is_trie) return;
        trie = Trie<BYTES>();
        for (auto& x : vec) { auto be = x.to_be_bytes(); trie.insert(be.data()); }
        trie_len = vec.size();
        vec.clear(); vec.shrink_to_fit();
        is_trie = true;
    }
    void as_vec() {  // :180-188
        if (!is_trie) return;
        std::vector<S> v;
        v.reserve(trie_len);
        trie.for_each([&](const std::array<uint8_t, BYTES>& be) { v.push_back(S::from_be_bytes(be.data())); });
        vec = std::move(v);
        trie = Trie<BYTES>();
        trie_len = 0;
        is_trie = false;
    }
    // :198-207 iteration in container order (Vec: insertion order; Trie: ascending)
    std::vector<S> items() const {
        if (!is_trie) return vec;
        std::vector<S> v;
        v.reserve(trie_len);
        trie.for_each([&](const std::array<uint8_t, BYTES>& be) { v.push_back(S::from_be_bytes(be.data())); });
        return v;
    }
    // :209-220 iter_sorted — sorts a Vec bucket IN PLACE as a side effect
    std::vector<S> items_sorted() {
        if (!is_trie) std::sort(vec.begin(), vec.end());
        return items();
    }

    // ---- src/trievec/set_ops.rs:5-257 : two-pointer merges over iter_sorted() ----
    static TrieVec op_or(TrieVec& a, TrieVec& b) {  // :5-41
        auto x = a.items_sorted(), y = b.items_sorted();
        std::vector<S> out;
        size_t i = 0, j = 0;
        while (i < x.size() && j < y.size()) {
            int c = x[i].cmp(y[j]);
            if (c < 0) out.push_back(x[i++]);
            else if (c > 0) out.push_back(y[j++]);
            else { out.push_back(x[i]); i++; j++; }
        }
        while (i < x.size()) out.push_back(x[i++]);
        while (j < y.size()) out.push_back(y[j++]);
        return from_vec(std::move(out));
    }
    void or_assign(TrieVec& o) {  // :43-71
        auto x = items_sorted(), y = o.items_sorted();
        std::vector<S> ins;
        size_t i = 0, j = 0;
        while (i < x.size() && j < y.size()) {
            int c = x[i].cmp(y[j]);
            if (c < 0) i++;
            else if (c > 0) ins.push_back(y[j++]);
            else { i++; j++; }
        }
        while (j < y.size()) ins.push_back(y[j++]);
        insert_sorted_iter(ins);
    }
    static TrieVec op_and(TrieVec& a, TrieVec& b) {  // :73-99
        auto x = a.items_sorted(), y = b.items_sorted();
        std::vector<S> out;
        size_t i = 0, j = 0;
        while (i < x.size() && j < y.size()) {
            int c = x[i].cmp(y[j]);
            if (c < 0) i++;
            else if (c > 0) j++;
            else { out.push_back(x[i]); i++; j++; }
        }
        return from_vec(std::move(out));
    }
    void and_assign(TrieVec& o) {  // :101-131
        auto x = items_sorted(), y = o.items_sorted();
        std::vector<S> del;
        size_t i = 0, j = 0;
        while (i < x.size() && j < y.size()) {
            int c = x[i].cmp(y[j]);
            if (c < 0) del.push_back(x[i++]);
            else if (c > 0) j++;
            else { i++; j++; }
        }
        while (i < x.size()) del.push_back(x[i++]);
        remove_sorted_iter(del);
    }
    static TrieVec op_sub(TrieVec& a, TrieVec& b) {  // :133-161
        auto x = a.items_sorted(), y = b.items_sorted();
        std::vector<S> out;
        size_t i = 0, j = 0;
        while (i < x.size() && j < y.size()) {
            int c = x[i].cmp(y[j]);
            if (c < 0) out.push_back(x[i++]);
            else if (c > 0) j++;
            else { i++; j++; }
        }
        while (i < x.size()) out.push_back(x[i++]);
        return from_vec(std::move(out));
    }
    void sub_assign(TrieVec& o) {  // :163-189
        auto x = items_sorted(), y = o.items_sorted();
        std::vector<S> del;
        size_t i = 0, j = 0;
        while (i < x.size() && j < y.size()) {
            int c = x[i].cmp(y[j]);
            if (c < 0) i++;
            else if (c > 0) j++;
            else { del.push_back(x[i]); i++; j++; }
        }
        remove_sorted_iter(del);
    }
    static TrieVec op_xor(TrieVec& a, TrieVec& b) {  // :191-224
        auto x = a.items_sorted(), y = b.items_sorted();
        std::vector<S> out;
        size_t i = 0, j = 0;
        while (i < x.size() && j < y.size()) {
            int c = x[i].cmp(y[j]);
            if (c < 0) out.push_back(x[i++]);
            else if (c > 0) out.push_back(y[j++]);
            else { i++; j++; }
        }
        while (i < x.size()) out.push_back(x[i++]);
        while (j < y.size()) out.push_back(y[j++]);
        return from_vec(std::move(out));
    }
    void xor_assign(TrieVec& o) {  // :226-257
        auto x = items_sorted(), y = o.items_sorted();
        std::vector<S> ins, del;
        size_t i = 0, j = 0;
        while (i < x.size() && j < y.size()) {
            int c = x[i].cmp(y[j]);
            if (c < 0) i++;
            else if (c > 0) ins.push_back(y[j++]);
            else { del.push_back(x[i]); i++; j++; }
        }
        while (j < y.size()) ins.push_back(y[j++]);
        insert_sorted_iter(ins);
        remove_sorted_iter(del);
    }
};

// ---------------------------------------------------------------------------------------------
// The C++ half of the reference: rank bitvector + tiered vector.
// With ORACLE_USE_REFERENCE_CXX these ARE the reference's classes (cxx/rank_bv.h:14-42,
// cxx/tiered_vec.h:31-89).  Otherwise: stand-ins with the same observable behaviour
// (exclusive rank; set() returns the previous bit; count_ones() == rank(size-1), i.e. it ignores
// the last bit — SURVEY F2).
// ---------------------------------------------------------------------------------------------
#ifdef ORACLE_USE_REFERENCE_CXX
using RankBVImpl = ::RankBV;
using TieredImpl = ::TieredVec32;
#else
class RankBVImpl {
    size_t nbits;
    std::vector<uint64_t> words;
    std::vector<int64_t> fen;  // Fenwick tree over per-word popcounts (1-based)
    void add(size_t widx, int64_t d) { for (size_t i = widx + 1; i <= words.size(); i += i & (~i + 1)) fen[i] += d; }
    uint64_t prefix(size_t nwords) const { int64_t s = 0; for (size_t i = nwords; i > 0; i -= i & (~i + 1)) s += fen[i]; return (uint64_t)s; }
public:
    explicit RankBVImpl(size_t size) : nbits(size), words((size + 63) / 64, 0), fen((size + 63) / 64 + 1, 0) {}
    size_t size() const { return nbits; }
    bool get(size_t i) const { return (words[i / 64] >> (i % 64)) & 1; }
    bool set(size_t i) { bool was = get(i); if (!was) { words[i / 64] |= 1ULL << (i % 64); add(i / 64, 1); } return was; }
    bool clear(size_t i) { bool was = get(i); if (was) { words[i / 64] &= ~(1ULL << (i % 64)); add(i / 64, -1); } return was; }
    uint64_t rank(size_t i) const { return prefix(i / 64) + (uint64_t)__builtin_popcountll(words[i / 64] & ((1ULL << (i % 64)) - 1)); }
    size_t count_ones() const { return (size_t)rank(nbits - 1); }
    size_t num_blocks() const { return words.size(); }
    uint64_t get_block(size_t b) const { return words[b]; }
    void update_block(size_t b, uint64_t v) {
        int64_t d = (int64_t)__builtin_popcountll(v) - (int64_t)__builtin_popcountll(words[b]);
        words[b] = v;
        if (d) add(b, d);
    }
};
// sequence rank -> bucket id with positional insert/remove (two-level blocked vector)
class TieredImpl {
    static constexpr size_t BLK = 2048;
    std::vector<std::vector<uint32_t>> blocks;
    size_t n = 0;
    std::pair<size_t, size_t> locate(size_t idx) const {
        size_t b = 0;
        while (b + 1 < blocks.size() && idx >= blocks[b].size()) { idx -= blocks[b].size(); b++; }
        return {b, idx};
    }
public:
    size_t len() const { return n; }
    uint32_t get(size_t idx) const { auto [b, o] = locate(idx); return blocks[b][o]; }
    void insert(size_t idx, uint32_t v) {
        if (blocks.empty()) blocks.emplace_back();
        auto [b, o] = locate(idx);
        auto& blk = blocks[b];
        blk.insert(blk.begin() + (ptrdiff_t)o, v);
        if (blk.size() > 2 * BLK) {
            std::vector<uint32_t> tail(blk.begin() + BLK, blk.end());
            blk.resize(BLK);
            blocks.insert(blocks.begin() + (ptrdiff_t)b + 1, std::move(tail));
        }
        n++;
    }
    void remove(size_t idx) {
        auto [b, o] = locate(idx);
        blocks[b].erase(blocks[b].begin() + (ptrdiff_t)o);
        if (blocks[b].empty() && blocks.size() > 1) blocks.erase(blocks.begin() + (ptrdiff_t)b);
        n--;
    }
};
#endif

// ---------------------------------------------------------------------------------------------
// src/bitvector/mod.rs:12-99 + set_ops.rs:4-106 — Bitvector wrapper (sized 1 << bitlength).
// ---------------------------------------------------------------------------------------------
class Bitvector {
    std::unique_ptr<RankBVImpl> bv;
    int bitlength_;
public:
    explicit Bitvector(int bitlength) : bv(new RankBVImpl((size_t)1 << bitlength)), bitlength_(bitlength) {}
    Bitvector(const Bitvector& o) : bv(new RankBVImpl((size_t)1 << o.bitlength_)), bitlength_(o.bitlength_) {  // :87-99
        for (size_t i = 0; i < o.bv->num_blocks(); i++) { uint64_t b = o.bv->get_block(i); if (b) bv->update_block(i, b); }
    }
    Bitvector& operator=(const Bitvector& o) { if (this != &o) { Bitvector t(o); std::swap(bv, t.bv); bitlength_ = o.bitlength_; } return *this; }
    Bitvector(Bitvector&&) = default;
    Bitvector& operator=(Bitvector&&) = default;
    int bitlength() const { return bitlength_; }
    bool contains(size_t i) const { return bv->get(i); }
    bool insert(size_t i) { return !bv->set(i); }   // :35-37  true if it was absent
    bool remove(size_t i) { return bv->clear(i); }  // :40-42
    size_t rank(size_t i) const { return (size_t)bv->rank(i); }
    size_t count() const { return bv->count_ones(); }
    size_t num_blocks() const { return bv->num_blocks(); }
    uint64_t get_block(size_t b) const { return bv->get_block(b); }
    std::vector<size_t> indices() const {  // :64-85 ascending set bits
        std::vector<size_t> r;
        for (size_t b = 0; b < bv->num_blocks(); b++) {
            uint64_t blk = bv->get_block(b);
            while (blk) { int t = __builtin_ctzll(blk); blk &= blk - 1; r.push_back(b * 64 + (size_t)t); }
        }
        return r;
    }
    template <class F> void for_each(F&& f) const {
        for (size_t b = 0; b < bv->num_blocks(); b++) {
            uint64_t blk = bv->get_block(b);
            while (blk) { int t = __builtin_ctzll(blk); blk &= blk - 1; f(b * 64 + (size_t)t); }
        }
    }
    enum Op { OR, AND, SUB, XOR };
    void assign_op(Op op, const Bitvector& o) {  // set_ops.rs:19-28,47-56,75-84,103-112
        for (size_t i = 0; i < bv->num_blocks(); i++) {
            uint64_t a = bv->get_block(i), b = o.bv->get_block(i), r;
            switch (op) { case OR: r = a | b; break; case AND: r = a & b; break; case SUB: r = a & ~b; break; default: r = a ^ b; }
            bv->update_block(i, r);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// src/wordset/mod.rs:18-437 + src/wordset/set_ops.rs:11-410 — WordSet<PREFIX_BITS,SUFFIX_BITS>.
// Words are handled as u128 (the reference is generic over T; nothing depends on T's width).
// ---------------------------------------------------------------------------------------------
template <int BYTES>
class WordSet {
public:
    using S = SlicedInt<BYTES>;
    using TV = TrieVec<BYTES>;
    static constexpr size_t THRESHOLD = 1024;  // mod.rs:34

    int PREFIX_BITS, SUFFIX_BITS;
    Bitvector prefixes;
    std::unique_ptr<TieredImpl> tiered;
    std::vector<TV> suffix_containers;
    std::vector<size_t> empty_containers;

    WordSet(int prefix_bits, int suffix_bits)
        : PREFIX_BITS(prefix_bits), SUFFIX_BITS(suffix_bits), prefixes(prefix_bits), tiered(new TieredImpl()) {
        if (prefix_bits > 32) throw std::invalid_argument("PREFIX_BITS should be <= 32");  // mod.rs:37-41
        if (suffix_bits <= 0) throw std::invalid_argument("SUFFIX_BITS should be != 0");
        if ((suffix_bits + 7) / 8 != BYTES) throw std::invalid_argument("BYTES mismatch");
    }
    WordSet(const WordSet& o)  // mod.rs:364-380
        : PREFIX_BITS(o.PREFIX_BITS), SUFFIX_BITS(o.SUFFIX_BITS), prefixes(o.prefixes), tiered(new TieredImpl()),
          suffix_containers(o.suffix_containers), empty_containers(o.empty_containers) {
        for (size_t i = 0; i < o.tiered->len(); i++) tiered->insert(i, o.tiered->get(i));
    }
    WordSet(WordSet&&) = default;
    WordSet& operator=(WordSet&&) = default;

    size_t count() const { size_t c = 0; for (auto& t : suffix_containers) c += t.len(); return c; }  // :50-55
    bool is_empty() const { return prefixes.count() == 0; }                                             // :57-60 (F2)
    size_t n_buckets() const { return tiered->len(); }

    std::pair<size_t, S> split(u128 word) const {  // :63-71
        u128 smask = (((u128)1) << SUFFIX_BITS) - 1;
        return {(size_t)(word >> SUFFIX_BITS), S::from_int(word & smask)};
    }
    u128 merge_ps(size_t prefix, const S& suffix) const { return ((u128)prefix << SUFFIX_BITS) | suffix.get(); }  // :74-84

    bool contains(u128 word) const {  // :87-95
        auto [p, s] = split(word);
        if (!prefixes.contains(p)) return false;
        size_t id = tiered->get(prefixes.rank(p));
        return suffix_containers[id].contains(s);
    }
    bool insert(u128 word) {  // :97-120
        auto [p, s] = split(word);
        bool absent = prefixes.insert(p);
        size_t rank = prefixes.rank(p);
        if (absent) {
            if (!empty_containers.empty()) {
                size_t id = empty_containers.back();
                empty_containers.pop_back();
                suffix_containers[id].insert(s);
                tiered->insert(rank, (uint32_t)id);
            } else {
                size_t id = suffix_containers.size();
                TV t; t.vec.push_back(s);
                suffix_containers.push_back(std::move(t));
                tiered->insert(rank, (uint32_t)id);
            }
        } else {
            size_t id = tiered->get(rank);
            absent = suffix_containers[id].insert(s);
            adapt_grow(id);
        }
        return absent;
    }
    bool remove(u128 word) {  // :122-137
        auto [p, s] = split(word);
        bool present = prefixes.contains(p);
        if (present) {
            size_t rank = prefixes.rank(p);
            size_t id = tiered->get(rank);
            present = suffix_containers[id].remove(s);
            adapt_shrink(id);
            if (suffix_containers[id].is_empty()) {
                empty_containers.push_back(id);
                tiered->remove(rank);
                prefixes.remove(p);
            }
        }
        return present;
    }
    // chunk_by(|(p1,_),(p2,_)| p1 == p2): runs of CONSECUTIVE equal prefixes (:147,172,192,223)
    template <class F> void for_each_group(const std::vector<std::pair<size_t, S>>& ps, F&& f) const {
        size_t i = 0;
        while (i < ps.size()) {
            size_t j = i + 1;
            while (j < ps.size() && ps[j].first == ps[i].first) j++;
            f(i, j);
            i = j;
        }
    }
    std::vector<std::pair<size_t, S>> split_all(const u128* words, size_t n) const {
        std::vector<std::pair<size_t, S>> ps;
        ps.reserve(n);
        for (size_t i = 0; i < n; i++) ps.push_back(split(words[i]));
        return ps;
    }
    bool contains_all(const u128* words, size_t n) const {  // :139-161
        auto ps = split_all(words, n);
        bool ok = true;
        for_each_group(ps, [&](size_t a, size_t b) {
            if (!ok) return;
            size_t p = ps[a].first;
            if (!prefixes.contains(p)) { ok = false; return; }
            size_t id = tiered->get(prefixes.rank(p));
            for (size_t k = a; k < b; k++) if (!suffix_containers[id].contains(ps[k].second)) { ok = false; return; }
        });
        return ok;
    }
    void contains_batch(const u128* words, size_t n, std::vector<uint8_t>& res) const {  // :163-185
        auto ps = split_all(words, n);
        for_each_group(ps, [&](size_t a, size_t b) {
            size_t p = ps[a].first;
            if (!prefixes.contains(p)) { res.insert(res.end(), b - a, 0); return; }
            size_t id = tiered->get(prefixes.rank(p));
            for (size_t k = a; k < b; k++) res.push_back(suffix_containers[id].contains(ps[k].second) ? 1 : 0);
        });
    }
    void insert_batch(const u128* words, size_t n) {  // :187-216
        auto ps = split_all(words, n);
        for_each_group(ps, [&](size_t a, size_t b) {
            size_t p = ps[a].first;
            bool absent = prefixes.insert(p);
            size_t rank = prefixes.rank(p);
            size_t id;
            if (absent) {
                if (!empty_containers.empty()) {
                    id = empty_containers.back();
                    empty_containers.pop_back();
                    tiered->insert(rank, (uint32_t)id);
                } else {
                    id = suffix_containers.size();
                    suffix_containers.emplace_back();
                    tiered->insert(rank, (uint32_t)id);
                }
            } else {
                id = tiered->get(rank);
            }
            for (size_t k = a; k < b; k++) suffix_containers[id].insert(ps[k].second);
            adapt_grow(id);
        });
    }
    void remove_batch(const u128* words, size_t n) {  // :218-237
        auto ps = split_all(words, n);
        for_each_group(ps, [&](size_t a, size_t b) {
            size_t p = ps[a].first;
            if (!prefixes.contains(p)) return;
            size_t rank = prefixes.rank(p);
            size_t id = tiered->get(rank);
            for (size_t k = a; k < b; k++) suffix_containers[id].remove(ps[k].second);
            if (suffix_containers[id].is_empty()) {
                empty_containers.push_back(id);
                tiered->remove(rank);
                prefixes.remove(p);
            }
            adapt_shrink(id);
        });
    }
    void adapt_grow(size_t id) { if (suffix_containers[id].len() > THRESHOLD) suffix_containers[id].as_trie(); }    // :240-244
    void adapt_shrink(size_t id) { if (suffix_containers[id].len() <= THRESHOLD) suffix_containers[id].as_vec(); }  // :247-251

    // :298-362 — prefixes ascending; inside a bucket: container order (history dependent, F5)
    template <class F> void for_each_word(F&& f) const {
        size_t rank = 0;
        prefixes.for_each([&](size_t p) {
            size_t id = tiered->get(rank++);
            for (auto& s : suffix_containers[id].items()) f(merge_ps(p, s));
        });
    }
    std::vector<std::pair<size_t, size_t>> buckets_sizes() const {  // :258-263
        std::vector<std::pair<size_t, size_t>> r;
        size_t rank = 0;
        prefixes.for_each([&](size_t p) { r.push_back({p, suffix_containers[tiered->get(rank++)].len()}); });
        return r;
    }

    // ---------------- set_ops.rs ----------------
    void push_container(size_t prefix, TV&& c) {
        size_t rank = suffix_containers.size();
        suffix_containers.push_back(std::move(c));
        tiered->insert(rank, (uint32_t)rank);
        prefixes.insert(prefix);
    }
    // merge_join_by over the two ascending (rank, prefix) streams
    template <class L, class R, class B>
    static void merge_join(WordSet& a, WordSet& b, L&& left, R&& right, B&& both) {
        auto pa = a.prefixes.indices(), pb = b.prefixes.indices();
        size_t i = 0, j = 0;
        while (i < pa.size() || j < pb.size()) {
            if (j >= pb.size() || (i < pa.size() && pa[i] < pb[j])) { left(i, pa[i]); i++; }
            else if (i >= pa.size() || pb[j] < pa[i]) { right(j, pb[j]); j++; }
            else { both(i, j, pa[i]); i++; j++; }
        }
    }
    enum Op { OR = 0, AND = 1, SUB = 2, XOR = 3 };
    // out-of-place | & - ^  (set_ops.rs:78-121,159-190,241-279,319-364)
    static WordSet binary_op(Op op, WordSet& a, WordSet& b) {
        WordSet res(a.PREFIX_BITS, a.SUFFIX_BITS);
        merge_join(
            a, b,
            [&](size_t ra, size_t p) {
                if (op == AND) return;
                TV c = a.suffix_containers[a.tiered->get(ra)];
                res.push_container(p, std::move(c));
            },
            [&](size_t rb, size_t p) {
                if (op == AND || op == SUB) return;
                TV c = b.suffix_containers[b.tiered->get(rb)];
                res.push_container(p, std::move(c));
            },
            [&](size_t ra, size_t rb, size_t p) {
                TV& x = a.suffix_containers[a.tiered->get(ra)];
                TV& y = b.suffix_containers[b.tiered->get(rb)];
                TV c = op == OR ? TV::op_or(x, y) : op == AND ? TV::op_and(x, y) : op == SUB ? TV::op_sub(x, y) : TV::op_xor(x, y);
                if (op == OR || !c.is_empty()) res.push_container(p, std::move(c));
            });
        return res;
    }
    void or_assign(WordSet& o) {  // set_ops.rs:128-156
        auto mine = prefixes.indices();
        auto theirs = o.prefixes.indices();
        size_t pi = 0, rank = 0;
        for (size_t orank = 0; orank < theirs.size(); orank++) {
            size_t op = theirs[orank];
            while (pi < mine.size() && mine[pi] < op) { pi++; rank++; }
            if (pi < mine.size() && mine[pi] == op) {
                size_t id = tiered->get(rank), oid = o.tiered->get(orank);
                suffix_containers[id].or_assign(o.suffix_containers[oid]);
                pi++; rank++;
            } else {
                size_t id = suffix_containers.size(), oid = o.tiered->get(orank);
                suffix_containers.push_back(o.suffix_containers[oid]);
                tiered->insert(rank, (uint32_t)id);
                rank++;
            }
        }
        prefixes.assign_op(Bitvector::OR, o.prefixes);
    }
    void and_assign(WordSet& o) {  // set_ops.rs:197-238
        auto mine = prefixes.indices();
        auto theirs = o.prefixes.indices();
        size_t pi = 0, rank = 0;
        std::vector<size_t> empty_prefixes;
        for (size_t orank = 0; orank < theirs.size(); orank++) {
            size_t op = theirs[orank];
            while (pi < mine.size() && mine[pi] < op) {
                size_t id = tiered->get(rank);
                suffix_containers[id].clear();
                empty_containers.push_back(id);
                tiered->remove(rank);
                pi++;
            }
            if (pi < mine.size() && mine[pi] == op) {
                size_t id = tiered->get(rank), oid = o.tiered->get(orank);
                suffix_containers[id].and_assign(o.suffix_containers[oid]);
                if (suffix_containers[id].is_empty()) {
                    empty_containers.push_back(id);
                    tiered->remove(rank);
                    empty_prefixes.push_back(mine[pi]);
                } else rank++;
                pi++;
            }
        }
        while (pi < mine.size()) {
            size_t id = tiered->get(rank);
            suffix_containers[id].clear();
            empty_containers.push_back(id);
            tiered->remove(rank);
            pi++;
        }
        prefixes.assign_op(Bitvector::AND, o.prefixes);
        for (size_t p : empty_prefixes) prefixes.remove(p);
    }
    void sub_assign(WordSet& o) {  // set_ops.rs:286-316
        auto mine = prefixes.indices();
        auto theirs = o.prefixes.indices();
        size_t pi = 0, rank = 0;
        std::vector<size_t> nonempty;
        for (size_t orank = 0; orank < theirs.size(); orank++) {
            size_t op = theirs[orank];
            while (pi < mine.size() && mine[pi] < op) { pi++; rank++; }
            if (pi < mine.size() && mine[pi] == op) {
                size_t id = tiered->get(rank), oid = o.tiered->get(orank);
                suffix_containers[id].sub_assign(o.suffix_containers[oid]);
                if (suffix_containers[id].is_empty()) { empty_containers.push_back(id); tiered->remove(rank); }
                else { nonempty.push_back(mine[pi]); rank++; }
                pi++;
            }
        }
        prefixes.assign_op(Bitvector::SUB, o.prefixes);
        for (size_t p : nonempty) prefixes.insert(p);
    }
    void xor_assign(WordSet& o) {  // set_ops.rs:371-409
        auto mine = prefixes.indices();
        auto theirs = o.prefixes.indices();
        size_t pi = 0, rank = 0;
        std::vector<size_t> nonempty;
        for (size_t orank = 0; orank < theirs.size(); orank++) {
            size_t op = theirs[orank];
            while (pi < mine.size() && mine[pi] < op) { pi++; rank++; }
            if (pi < mine.size() && mine[pi] == op) {
                size_t id = tiered->get(rank), oid = o.tiered->get(orank);
                suffix_containers[id].xor_assign(o.suffix_containers[oid]);
                if (suffix_containers[id].is_empty()) { empty_containers.push_back(id); tiered->remove(rank); }
                else { nonempty.push_back(mine[pi]); rank++; }
                pi++;
            } else {
                size_t id = suffix_containers.size(), oid = o.tiered->get(orank);
                suffix_containers.push_back(o.suffix_containers[oid]);
                tiered->insert(rank, (uint32_t)id);
                rank++;
            }
        }
        prefixes.assign_op(Bitvector::XOR, o.prefixes);
        for (size_t p : nonempty) prefixes.insert(p);
    }
    // k-way merge / intersect (set_ops.rs:11-75).  iter-set-ops 0.2 (not vendored) provides
    // merge_iters_detailed_by / intersect_iters_detailed_by: k-way sorted union/intersection of the
    // ascending prefix streams reporting which inputs hold each item; restated from its documented
    // behaviour and pinned by src/wordset/set_ops.rs:656-680.
    static WordSet merge(std::vector<WordSet*>& sets) {
        WordSet res(sets[0]->PREFIX_BITS, sets[0]->SUFFIX_BITS);
        std::vector<std::vector<size_t>> pre;
        for (auto* s : sets) pre.push_back(s->prefixes.indices());
        std::vector<size_t> cur(sets.size(), 0);
        for (;;) {
            size_t best = SIZE_MAX;
            for (size_t i = 0; i < sets.size(); i++) if (cur[i] < pre[i].size()) best = std::min(best, pre[i][cur[i]]);
            if (best == SIZE_MAX) break;
            std::vector<std::pair<size_t, size_t>> details;  // (set index, rank)
            for (size_t i = 0; i < sets.size(); i++) if (cur[i] < pre[i].size() && pre[i][cur[i]] == best) { details.push_back({i, cur[i]}); cur[i]++; }
            TV container;
            if (details.size() == 1) {
                auto [i, rank] = details[0];
                container = sets[i]->suffix_containers[sets[i]->tiered->get(rank)];
            } else {
                std::vector<S> all;
                for (auto [i, rank] : details) {
                    auto v = sets[i]->suffix_containers[sets[i]->tiered->get(rank)].items_sorted();
                    all.insert(all.end(), v.begin(), v.end());
                }
                std::sort(all.begin(), all.end());
                all.erase(std::unique(all.begin(), all.end()), all.end());
                container.insert_sorted_iter(all);
            }
            res.push_container(best, std::move(container));
        }
        return res;
    }
    static WordSet intersect(std::vector<WordSet*>& sets) {
        WordSet res(sets[0]->PREFIX_BITS, sets[0]->SUFFIX_BITS);
        std::vector<std::vector<size_t>> pre;
        for (auto* s : sets) pre.push_back(s->prefixes.indices());
        std::vector<size_t> cur(sets.size(), 0);
        for (;;) {
            bool done = false;
            size_t mx = 0;
            for (size_t i = 0; i < sets.size(); i++) { if (cur[i] >= pre[i].size()) { done = true; break; } mx = std::max(mx, pre[i][cur[i]]); }
            if (done) break;
            bool all_eq = true;
            for (size_t i = 0; i < sets.size(); i++) {
                while (cur[i] < pre[i].size() && pre[i][cur[i]] < mx) cur[i]++;
                if (cur[i] >= pre[i].size()) { done = true; break; }
                if (pre[i][cur[i]] != mx) all_eq = false;
            }
            if (done) break;
            if (!all_eq) continue;
            std::vector<S> acc;
            for (size_t i = 0; i < sets.size(); i++) {
                auto v = sets[i]->suffix_containers[sets[i]->tiered->get(cur[i])].items_sorted();
                if (i == 0) acc = v;
                else {
                    std::vector<S> t;
                    std::set_intersection(acc.begin(), acc.end(), v.begin(), v.end(), std::back_inserter(t));
                    acc.swap(t);
                }
                cur[i]++;
            }
            TV container;
            container.insert_sorted_iter(acc);
            if (!container.is_empty()) res.push_container(mx, std::move(container));
        }
        return res;
    }
};

// ---------------------------------------------------------------------------------------------
// bincode 1.3 "varint" integer encoding (crate not vendored; restated from its published format):
// u < 251 -> 1 byte; 251 + u16 LE; 252 + u32 LE; 253 + u64 LE.  PARITY UNPINNED (no reference test
// serialises anything).
// ---------------------------------------------------------------------------------------------
struct ByteWriter {
    std::vector<uint8_t> out;
    void u8(uint8_t v) { out.push_back(v); }
    void varint(uint64_t v) {
        if (v < 251) u8((uint8_t)v);
        else if (v <= 0xFFFF) { u8(251); for (int i = 0; i < 2; i++) u8((uint8_t)(v >> (8 * i))); }
        else if (v <= 0xFFFFFFFFull) { u8(252); for (int i = 0; i < 4; i++) u8((uint8_t)(v >> (8 * i))); }
        else { u8(253); for (int i = 0; i < 8; i++) u8((uint8_t)(v >> (8 * i))); }
    }
};
struct ByteReader {
    const uint8_t* p; size_t n, i = 0;
    ByteReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
    uint8_t u8() { if (i >= n) throw std::runtime_error("eof"); return p[i++]; }
    uint64_t varint() {
        uint8_t t = u8();
        if (t < 251) return t;
        int nb = t == 251 ? 2 : t == 252 ? 4 : t == 253 ? 8 : -1;
        if (nb < 0) throw std::runtime_error("bad varint tag");
        uint64_t v = 0;
        for (int k = 0; k < nb; k++) v |= (uint64_t)u8() << (8 * k);
        return v;
    }
};

template <int BYTES> void ser_trie_node(ByteWriter& w, const TrieNode<BYTES>& n) {
    // bitvector/tiny/mod.rs:97-105 (seq of set indices as u8) ; trie.rs:53-57 (children seq)
    auto idx = n.bv.indices();
    w.varint(idx.size());
    for (auto b : idx) w.u8(b);
    w.varint(n.children.size());
    for (auto& c : n.children) ser_trie_node<BYTES>(w, *c);
}
template <int BYTES> std::unique_ptr<TrieNode<BYTES>> de_trie_node(ByteReader& r) {
    auto n = std::make_unique<TrieNode<BYTES>>();
    uint64_t k = r.varint();
    for (uint64_t i = 0; i < k; i++) n->bv.insert(r.u8());
    uint64_t c = r.varint();
    for (uint64_t i = 0; i < c; i++) n->children.push_back(de_trie_node<BYTES>(r));
    return n;
}

// ---------------------------------------------------------------------------------------------
// src/cbl.rs:40-569 — CBL<K,T,PREFIX_BITS> with K / PREFIX_BITS as run-time values.
// ---------------------------------------------------------------------------------------------
inline int pos_bits_for(int kmer_bits) {  // cbl.rs:66  next_power_of_two().ilog2()
    int p = 0;
    while ((1 << p) < kmer_bits) p++;
    return p;
}

template <class T, int BYTES>
class CBL {
public:
    static constexpr size_t CHUNK_SIZE = 2048;  // cbl.rs:67
    static constexpr int M = 9;                 // cbl.rs:16
    int K, PREFIX_BITS, KMER_BITS, POS_BITS, SUFFIX_BITS;
    bool canonical;
    KmerOps<T> ops;
    WordSet<BYTES> wordset;
    NecklaceQueue<T, false> queue;
    NecklaceQueue<T, true> queue_rev;

    static int queue_width(int k) { int w = 2 * k - (M - 1); return w < 0 ? 0 : w; }  // cbl.rs:24-26
    static int suffix_bits(int k, int p) { int s = 2 * k + pos_bits_for(2 * k) - p; return s < 0 ? 0 : s; }  // cbl.rs:29-32

    CBL(int k, int prefix_bits, bool canon)
        : K(k), PREFIX_BITS(prefix_bits), KMER_BITS(2 * k), POS_BITS(pos_bits_for(2 * k)),
          SUFFIX_BITS(suffix_bits(k, prefix_bits)), canonical(canon), ops(k),
          wordset(prefix_bits, suffix_bits(k, prefix_bits)), queue(2 * k, (size_t)queue_width(k)),
          queue_rev(2 * k, (size_t)queue_width(k)) {
        if (KMER_BITS + POS_BITS > type_bits<T>())  // cbl.rs:87-91
            throw std::invalid_argument("Cannot fit a " + std::to_string(k) + "-mer and its length in a " +
                                        std::to_string(type_bits<T>()) + "-bit integer");
    }
    CBL(const CBL& o)
        : K(o.K), PREFIX_BITS(o.PREFIX_BITS), KMER_BITS(o.KMER_BITS), POS_BITS(o.POS_BITS), SUFFIX_BITS(o.SUFFIX_BITS),
          canonical(o.canonical), ops(o.ops), wordset(o.wordset), queue(o.queue), queue_rev(o.queue_rev) {}
    CBL(int k, int prefix_bits, bool canon, WordSet<BYTES>&& ws)
        : K(k), PREFIX_BITS(prefix_bits), KMER_BITS(2 * k), POS_BITS(pos_bits_for(2 * k)),
          SUFFIX_BITS(suffix_bits(k, prefix_bits)), canonical(canon), ops(k), wordset(std::move(ws)),
          queue(2 * k, (size_t)queue_width(k)), queue_rev(2 * k, (size_t)queue_width(k)) {}

    T merge_necklace_pos(T necklace, size_t pos) const { return (T)((necklace << POS_BITS) | (T)pos); }  // :181-184
    std::pair<T, size_t> split_necklace_pos(T word) const {                                               // :188-195
        return {(T)(word >> POS_BITS), (size_t)(word & (T)(((T)1 << POS_BITS) - 1))};
    }
    T get_word(T kmer) const {  // :199-206 (brute-force necklace)
        auto [n, p] = necklace_pos<T>(canonical ? ops.canonical(kmer) : kmer, KMER_BITS);
        return merge_necklace_pos(n, p);
    }
    T recover_kmer(T word) const {  // :210-215
        auto [n, p] = split_necklace_pos(word);
        return revert_necklace_pos<T>(n, p, KMER_BITS);
    }
    bool contains(T kmer) const { return wordset.contains((u128)get_word(kmer)); }  // :219-221
    bool insert(T kmer) { return wordset.insert((u128)get_word(kmer)); }            // :226-228
    bool remove(T kmer) { return wordset.remove((u128)get_word(kmer)); }            // :233-235
    size_t count() const { return wordset.count(); }
    bool is_empty() const { return wordset.is_empty(); }

    // :239-243 — windows [s, min(s + 2048 + K - 1, len)), s stepping by 2048 over raw bytes
    template <class F> void for_each_chunk(const uint8_t* seq, size_t len, F&& f) const {
        for (size_t start = 0; start < len - (size_t)K + 1; start += CHUNK_SIZE) {
            size_t end = std::min(start + CHUNK_SIZE + (size_t)K - 1, len);
            f(seq + start, end - start);
        }
    }
    // :247-289 — streaming necklaces; canonical mode returns forward-canonical words first, then the
    // reverse-complemented ones (F6).  Non-ACGT bytes are skipped by filter_map (F8).
    void get_seq_words(const uint8_t* seq, size_t len, std::vector<u128>& out) {
        if (canonical) {
            std::vector<u128> res_rc;
            T kmer = ops.from_nucs(seq, (size_t)K);
            queue.insert_full(kmer);
            queue_rev.insert_full(ops.rev_comp(kmer));
            auto emit = [&]() {
                if (ops.is_canonical(kmer)) { auto [n, p] = queue.get_necklace_pos(); out.push_back((u128)merge_necklace_pos(n, p)); }
                else { auto [n, p] = queue_rev.get_necklace_pos(); res_rc.push_back((u128)merge_necklace_pos(n, p)); }
            };
            emit();
            for (size_t i = (size_t)K; i < len; i++) {
                int c = from_nuc(seq[i]);
                if (c < 0) continue;
                kmer = ops.append(kmer, (T)c);
                queue.insert2((T)c);
                queue_rev.insert2((T)complement(c));
                emit();
            }
            out.insert(out.end(), res_rc.begin(), res_rc.end());
        } else {
            T kmer = ops.from_nucs(seq, (size_t)K);
            queue.insert_full(kmer);
            { auto [n, p] = queue.get_necklace_pos(); out.push_back((u128)merge_necklace_pos(n, p)); }
            for (size_t i = (size_t)K; i < len; i++) {
                int c = from_nuc(seq[i]);
                if (c < 0) continue;
                queue.insert2((T)c);
                auto [n, p] = queue.get_necklace_pos();
                out.push_back((u128)merge_necklace_pos(n, p));
            }
        }
    }
    void check_len(size_t len) const {  // :294-299 etc.
        if (len < (size_t)K)
            throw std::invalid_argument("Sequence size (" + std::to_string(len) + ") is smaller than K (" + std::to_string(K) + ")");
    }
    // all words of a sequence, chunk by chunk (what insert_seq/contains_seq feed to the wordset)
    void seq_words(const uint8_t* seq, size_t len, std::vector<u128>& out) {
        check_len(len);
        for_each_chunk(seq, len, [&](const uint8_t* c, size_t n) { get_seq_words(c, n, out); });
    }
    bool contains_all(const uint8_t* seq, size_t len) {  // :293-308
        check_len(len);
        bool ok = true;
        for_each_chunk(seq, len, [&](const uint8_t* c, size_t n) {
            if (!ok) return;
            std::vector<u128> w;
            get_seq_words(c, n, w);
            if (!wordset.contains_all(w.data(), w.size())) ok = false;
        });
        return ok;
    }
    void contains_seq(const uint8_t* seq, size_t len, std::vector<uint8_t>& res) {  // :311-324
        check_len(len);
        for_each_chunk(seq, len, [&](const uint8_t* c, size_t n) {
            std::vector<u128> w;
            get_seq_words(c, n, w);
            wordset.contains_batch(w.data(), w.size(), res);
        });
    }
    void insert_seq(const uint8_t* seq, size_t len) {  // :328-339
        check_len(len);
        for_each_chunk(seq, len, [&](const uint8_t* c, size_t n) {
            std::vector<u128> w;
            get_seq_words(c, n, w);
            wordset.insert_batch(w.data(), w.size());
        });
    }
    void remove_seq(const uint8_t* seq, size_t len) {  // :343-354
        check_len(len);
        for_each_chunk(seq, len, [&](const uint8_t* c, size_t n) {
            std::vector<u128> w;
            get_seq_words(c, n, w);
            wordset.remove_batch(w.data(), w.size());
        });
    }
    // :358-360 — iteration in the reference's (history dependent) order, as words
    void iter_words(std::vector<u128>& out) const { wordset.for_each_word([&](u128 w) { out.push_back(w); }); }

    // serde (cbl.rs:40-54; wordset/mod.rs:382-437; trievec/mod.rs:8-15; sliced_int.rs:110-134)
    void serialize(ByteWriter& w) const {
        w.u8(canonical ? 1 : 0);
        w.varint(wordset.n_buckets());
        size_t rank = 0;
        wordset.prefixes.for_each([&](size_t p) {
            w.varint((uint64_t)(uint32_t)p);
            const auto& tv = wordset.suffix_containers[wordset.tiered->get(rank++)];
            if (!tv.is_trie) {
                w.varint(0);
                w.varint(tv.vec.size());
                for (auto& s : tv.vec) { w.varint(BYTES); for (int i = 0; i < BYTES; i++) w.u8(s.b[i]); }
            } else {
                w.varint(1);
                ser_trie_node<BYTES>(w, *tv.trie.root);
                w.varint(tv.trie_len);
            }
        });
    }
    static CBL deserialize(int k, int prefix_bits, ByteReader& r) {
        bool canon = r.u8() != 0;
        CBL c(k, prefix_bits, canon);
        uint64_t n = r.varint();
        for (uint64_t e = 0; e < n; e++) {  // wordset/mod.rs:411-426
            size_t prefix = (size_t)r.varint();
            TrieVec<BYTES> tv;
            uint64_t variant = r.varint();
            if (variant == 0) {
                uint64_t m = r.varint();
                tv.vec.reserve(m);
                for (uint64_t i = 0; i < m; i++) {
                    uint64_t bl = r.varint();
                    if (bl != (uint64_t)BYTES) throw std::runtime_error("bad SlicedInt length");
                    SlicedInt<BYTES> s;
                    for (int b = 0; b < BYTES; b++) s.b[b] = r.u8();
                    tv.vec.push_back(s);
                }
            } else if (variant == 1) {
                tv.is_trie = true;
                tv.trie.root = de_trie_node<BYTES>(r);
                tv.trie_len = (size_t)r.varint();
            } else throw std::runtime_error("bad TrieOrVec variant");
            size_t rank = c.wordset.suffix_containers.size();
            c.wordset.prefixes.insert(prefix);
            c.wordset.tiered->insert(rank, (uint32_t)rank);
            c.wordset.suffix_containers.push_back(std::move(tv));
        }
        if (r.i != r.n) throw std::runtime_error("trailing bytes");  // reject_trailing_bytes
        return c;
    }
};

}  // namespace orc
