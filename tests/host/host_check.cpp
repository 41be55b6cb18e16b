// Host-side check of the __host__ __device__ k-mer / necklace code the kernels use
// (cbl_b200/csrc/kmer_necklace.cuh) against the CPU oracle's restatement of the reference.
// Built and run by tests/test_host_device_functions.py (CPU only).
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../cbl_b200/csrc/kmer_necklace.cuh"
#include "../../oracle/cbl_oracle.hpp"

using cbl::u128;

template <class W> static W rnd(std::mt19937_64& g, int bits) {
    u128 v = ((u128)g() << 64) | g();
    return (W)(v & (((u128)1 << bits) - 1));
}

template <class W> static long check_bits(int bits, long n, std::mt19937_64& g) {
    long bad = 0;
    for (long it = 0; it < n; it++) {
        W w;
        int kind = it % 8;
        if (kind == 0) {  // periodic word
            int per = 1 + (int)(g() % (bits / 2 + 1));
            W unit = rnd<W>(g, per);
            w = 0;
            for (int s = 0; s < bits; s += per) w = (W)((w << per) | unit);
            w &= cbl::low_mask<W>(bits);
        } else if (kind == 1) {  // sparse
            w = 0;
            int nb = (int)(g() % 4);
            for (int b = 0; b < nb; b++) w |= (W)1 << (g() % bits);
        } else if (kind == 2) {  // dense
            w = cbl::low_mask<W>(bits);
            int nb = (int)(g() % 4);
            for (int b = 0; b < nb; b++) w &= ~((W)1 << (g() % bits));
        } else w = rnd<W>(g, bits);
        W n1, n2;
        int p1, p2;
        cbl::necklace_brute<W>(w, bits, n1, p1);
        cbl::necklace_fast<W>(w, bits, n2, p2);
        { W n3; int p3; cbl::necklace_runs<W>(w, bits, n3, p3); if (n3 != n1 || p3 != p1) { n2 = ~n1; } }
        auto ref = orc::necklace_pos<u128>((u128)w, bits);
        if (n1 != n2 || p1 != p2 || (u128)n1 != ref.first || (size_t)p1 != ref.second) {
            if (bad < 5) fprintf(stderr, "necklace mismatch bits=%d w=%llx%016llx brute=(..,%d) fast=(..,%d) ref=(..,%zu)\n", bits,
                                 (unsigned long long)((u128)w >> 64), (unsigned long long)w, p1, p2, ref.second);
            bad++;
        }
    }
    return bad;
}

int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 200000;
    std::mt19937_64 g(2024);
    long bad = 0;
    for (int bits : {2, 6, 14, 30, 34, 42, 50, 58}) bad += check_bits<uint64_t>(bits, n, g);   // >= 33: the periodic-window step
    for (int bits : {14, 50, 62, 64, 66, 100, 118}) bad += check_bits<u128>(bits, n, g);
    // revcomp + canonical word vs oracle KmerOps (x86 byte-swap formulation, src/kmer.rs:327-348)
    for (int k : {1, 3, 7, 11, 15, 25, 29}) {
        orc::KmerOps<uint64_t> o(k);
        for (long it = 0; it < n / 4; it++) {
            uint64_t x = rnd<uint64_t>(g, 2 * k);
            if (cbl::revcomp(x, k) != o.rev_comp(x)) { bad++; if (bad < 5) fprintf(stderr, "revcomp64 k=%d\n", k); }
        }
    }
    for (int k : {7, 31, 32, 33, 45, 59}) {
        orc::KmerOps<u128> o(k);
        for (long it = 0; it < n / 4; it++) {
            u128 x = rnd<u128>(g, 2 * k);
            if (cbl::revcomp(x, k) != o.rev_comp(x)) { bad++; if (bad < 5) fprintf(stderr, "revcomp128 k=%d\n", k); }
        }
    }
    // word <-> kmer round trip and oracle get_word
    struct Cfg { int k, p, canon; };
    for (Cfg c : {Cfg{7, 14, 0}, Cfg{25, 24, 0}, Cfg{25, 24, 1}, Cfg{29, 24, 1}}) {
        cbl::KParams P{c.k, 2 * c.k, orc::pos_bits_for(2 * c.k), c.p, 2 * c.k + orc::pos_bits_for(2 * c.k) - c.p, c.canon};
        orc::KmerOps<uint64_t> o(c.k);
        for (long it = 0; it < n / 4; it++) {
            uint64_t x = rnd<uint64_t>(g, 2 * c.k);
            uint64_t w = cbl::kmer_to_word<uint64_t>(x, P);
            auto ref = orc::necklace_pos<uint64_t>(c.canon ? o.canonical(x) : x, 2 * c.k);
            uint64_t rw = (ref.first << P.pos_bits) | ref.second;
            uint64_t back = cbl::word_to_kmer<uint64_t>(w, P);
            if (w != rw || back != (c.canon ? o.canonical(x) : x)) { bad++; if (bad < 5) fprintf(stderr, "word64 k=%d\n", c.k); }
        }
    }
    for (Cfg c : {Cfg{31, 24, 0}, Cfg{31, 24, 1}, Cfg{59, 28, 0}, Cfg{59, 24, 1}}) {
        cbl::KParams P{c.k, 2 * c.k, orc::pos_bits_for(2 * c.k), c.p, 2 * c.k + orc::pos_bits_for(2 * c.k) - c.p, c.canon};
        orc::KmerOps<u128> o(c.k);
        for (long it = 0; it < n / 4; it++) {
            u128 x = rnd<u128>(g, 2 * c.k);
            u128 w = cbl::kmer_to_word<u128>(x, P);
            auto ref = orc::necklace_pos<u128>(c.canon ? o.canonical(x) : x, 2 * c.k);
            u128 rw = (ref.first << P.pos_bits) | ref.second;
            u128 back = cbl::word_to_kmer<u128>(w, P);
            if (w != rw || back != (c.canon ? o.canonical(x) : x)) { bad++; if (bad < 5) fprintf(stderr, "word128 k=%d\n", c.k); }
        }
    }
    // pack4 / is_acgt
    const char* nucs = "ACGTacgt";
    for (int a = 0; a < 8; a++) for (int b = 0; b < 8; b++) for (int c = 0; c < 8; c++) for (int d = 0; d < 8; d++) {
        uint32_t w = (uint32_t)(uint8_t)nucs[a] | ((uint32_t)(uint8_t)nucs[b] << 8) | ((uint32_t)(uint8_t)nucs[c] << 16) | ((uint32_t)(uint8_t)nucs[d] << 24);
        uint32_t expect = (uint32_t)((orc::from_nuc(nucs[a]) << 6) | (orc::from_nuc(nucs[b]) << 4) | (orc::from_nuc(nucs[c]) << 2) | orc::from_nuc(nucs[d]));
        if (cbl::pack4(w) != expect) bad++;
    }
    for (int c = 0; c < 256; c++) if (cbl::is_acgt((uint8_t)c) != (orc::from_nuc((uint8_t)c) >= 0)) bad++;
    printf("host_check: %ld mismatches\n", bad);
    return bad ? 1 : 0;
}
