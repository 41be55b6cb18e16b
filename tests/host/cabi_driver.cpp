// C++ host driver of the drop-in boundary: compiles include/cbl.hpp (the C++ mirror of CBL<K, T, PREFIX_BITS>,
// src/cbl.rs:40-569), links libcbl_gpu through the C ABI only and checks insert_seq / contains_seq / | & - ^ / iter /
// serde — on one GPU and on a handle sharded over several GPUs of this process — against expected results written by
// the test harness from the CPU oracle (tests/test_cabi_cpp_driver.py).  No torch, no Python in this process.
//   cabi_driver <case-file> <tmp-dir> <dev0,dev1,...>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "cbl.hpp"

using Set = cbl::CBL<25, uint64_t, 24>;

static std::vector<uint8_t> read_blob(std::ifstream& f) {
    uint64_t n = 0;
    f.read((char*)&n, 8);
    std::vector<uint8_t> v(n);
    if (n) f.read((char*)v.data(), (std::streamsize)n);
    return v;
}
static std::vector<uint64_t> read_u64s(std::ifstream& f) {
    auto b = read_blob(f);
    std::vector<uint64_t> v(b.size() / 8);
    if (!v.empty()) memcpy(v.data(), b.data(), v.size() * 8);
    return v;
}
static std::vector<uint64_t> words_of(const Set& s) {
    std::vector<uint64_t> out, lo(1 << 16), hi(1 << 16);
    for (uint64_t start = 0;;) {
        size_t n = 0;
        if (cbl_export_words(s.handle(), start, lo.data(), hi.data(), lo.size(), &n)) throw std::runtime_error(cbl_last_error(s.handle()));
        if (!n) break;
        out.insert(out.end(), lo.begin(), lo.begin() + n);
        start += n;
    }
    return out;
}
static int failures = 0;
#define CHECK(cond, what) do { if (!(cond)) { std::printf("FAIL %s: %s\n", tag.c_str(), what); failures++; } } while (0)

static void run_case(const std::string& tag, Set a, Set b, const std::vector<uint8_t>& A, const std::vector<uint8_t>& B, const std::vector<uint64_t>& wA,
                     const std::vector<uint64_t> (&wop)[4], const std::vector<uint8_t>& ansBinA, const std::string& tmp) {
    CHECK(a.is_empty() && a.count() == 0, "fresh set is empty");
    a.insert_seq(A.data(), A.size());
    b.insert_seq(B.data(), B.size());
    CHECK(a.count() == wA.size(), "count after insert_seq");
    CHECK(words_of(a) == wA, "stored words after insert_seq (ascending)");
    CHECK(a.contains_all(A.data(), A.size()), "contains_all of the inserted sequence");
    CHECK(a.contains_seq(B.data(), B.size()) == ansBinA, "contains_seq answers");
    // iter yields the k-mers of the stored words: re-inserting them into a fresh set gives the same set (src/cbl.rs:764-773)
    {
        Set c = a.new_like();
        auto kmers = a.iter();
        CHECK(kmers.size() == wA.size(), "iter length");
        for (size_t i = 0; i < kmers.size(); i += 97) CHECK(a.contains(kmers[i]), "iter k-mer is contained");
        CHECK(c.insert(kmers[0]) && !c.insert(kmers[0]) && c.count() == 1 && c.remove(kmers[0]) && c.is_empty(), "single k-mer insert / remove return values");
    }
    Set u = a | b, i = a & b, d = a - b, x = a ^ b;
    CHECK(words_of(u) == wop[0], "a | b");
    CHECK(words_of(i) == wop[1], "a & b");
    CHECK(words_of(d) == wop[2], "a - b");
    CHECK(words_of(x) == wop[3], "a ^ b");
    {
        Set t = a;  // Clone
        t |= b; CHECK(words_of(t) == wop[0], "a |= b");
        t = a; t &= b; CHECK(words_of(t) == wop[1], "a &= b");
        t = a; t -= b; CHECK(words_of(t) == wop[2], "a -= b");
        t = a; t ^= b; CHECK(words_of(t) == wop[3], "a ^= b");
        CHECK(words_of(a) == wA, "operands untouched by the assign forms on a clone");
    }
    {
        std::vector<Set*> both{&a, &b};
        CHECK(words_of(Set::merge(both)) == wop[0], "CBL::merge");
        CHECK(words_of(Set::intersect(both)) == wop[1], "CBL::intersect");
    }
    const std::string path = tmp + "/" + tag + ".cbl";
    u.save_to_file(path.c_str());
    Set back = a.load_like(path.c_str());
    CHECK(words_of(back) == wop[0] && back.count() == wop[0].size(), "save_to_file / load_from_file round trip");
    a.remove_seq(A.data(), A.size());
    CHECK(a.is_empty() && a.count() == 0, "remove_seq empties the set");
    bool threw = false;
    try { a.insert_seq(A.data(), 10); } catch (const cbl::Panic& p) { threw = std::string(p.what()).find("smaller than K") != std::string::npos; }
    CHECK(threw, "short sequence panics with the reference's message");
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: cabi_driver <case-file> <tmp-dir> <devices>\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    const auto A = read_blob(f), B = read_blob(f);
    const auto wA = read_u64s(f);
    std::vector<uint64_t> wop[4];
    for (auto& w : wop) w = read_u64s(f);
    const auto ans = read_blob(f);
    std::vector<int> devs;
    { std::stringstream ss(argv[3]); std::string t; while (std::getline(ss, t, ',')) devs.push_back(std::atoi(t.c_str())); }
    try {
        run_case("single", Set(false, devs[0]), Set(false, devs[0]), A, B, wA, wop, ans, argv[2]);
        run_case("sharded", Set::sharded(devs), Set::sharded(devs), A, B, wA, wop, ans, argv[2]);
    } catch (const std::exception& e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 1;
    }
    std::printf(failures ? "cabi_driver: %d failure(s)\n" : "cabi_driver: all checks passed\n", failures);
    return failures ? 1 : 0;
}
