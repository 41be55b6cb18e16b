"""Shared helpers for the tests: seeded synthetic DNA and a tiny pure-Python statement of the
normative semantics used to cross-check small cases."""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACTG", dtype=np.uint8)  # code -> nucleotide (src/kmer.rs:11)
CODE = {ord("A"): 0, ord("C"): 1, ord("T"): 2, ord("G"): 3, ord("a"): 0, ord("c"): 1, ord("t"): 2, ord("g"): 3}


def random_dna(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return BASES[rng.integers(0, 4, size=n, dtype=np.uint8)]


def pos_bits(k: int) -> int:
    p = 0
    while (1 << p) < 2 * k:
        p += 1
    return p


def kmer_int(nucs: bytes) -> int:
    x = 0
    for c in nucs:
        x = (x << 2) | CODE[c]
    return x


def revcomp_int(x: int, k: int) -> int:
    r = 0
    for _ in range(k):
        r = (r << 2) | ((x & 3) ^ 2)
        x >>= 2
    return r


def necklace_pos_py(w: int, bits: int):
    mask = (1 << bits) - 1
    best, bp = w, 0
    for p in range(1, bits):
        r = ((w << p) & mask) | (w >> (bits - p))
        if r < best:
            best, bp = r, p
    return best, bp


def word_py(x: int, k: int, canonical: bool) -> int:
    if canonical and bin(x).count("1") % 2 == 1:
        x = revcomp_int(x, k)
    neck, pos = necklace_pos_py(x, 2 * k)
    return (neck << pos_bits(k)) | pos


def seq_words_py(seq: bytes, k: int, canonical: bool):
    """Words of an ACGT-only sequence in the reference's order (Appendix A item 7)."""
    n = len(seq) - k + 1
    out = []
    for start in range(0, n, 2048):
        m = min(2048, n - start)
        fwd, rc = [], []
        for i in range(start, start + m):
            x = kmer_int(seq[i : i + k])
            w = word_py(x, k, canonical)
            if canonical and bin(x).count("1") % 2 == 1:
                rc.append(w)
            else:
                fwd.append(w)
        out.extend(fwd + rc)
    return out


def to_int_list(lo, hi=None):
    if hi is None:
        return [int(v) for v in lo]
    return [int(a) | (int(b) << 64) for a, b in zip(lo, hi)]
