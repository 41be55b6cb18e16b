"""Pins the CPU oracle against every known-answer vector the reference's own unit tests hold for
the batched sequence path (tests/golden/reference_kats.json), and replays the
reference's property tests with seeded inputs.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import OracleCBL
import cbl_testutil as util

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def test_revcomp_kats(oracle_lib, kats):
    for case in kats["revcomp"]["cases"]:
        for tb in case["t_bits"]:
            src = np.frombuffer(case["nucs"].encode(), dtype=np.uint8).copy()
            out = np.zeros(case["k"], dtype=np.uint8)
            assert oracle_lib.orc_revcomp_nucs(case["k"], tb, src.ctypes.data_as(u8p), out.ctypes.data_as(u8p)) == 0
            assert out.tobytes().decode() == case["rc"]


def test_revcomp_involution(oracle_lib, kats):
    lo, hi = C.c_uint64(), C.c_uint64()
    for case in kats["revcomp_involution"]["cases"]:
        k, tb = case["k"], case["t_bits"]
        for i in range(0, case["n"], 7 if case["n"] > 20000 else 1):
            oracle_lib.orc_kmer_revcomp(k, tb, i, 0, C.byref(lo), C.byref(hi))
            assert lo.value == util.revcomp_int(i, k)
            oracle_lib.orc_kmer_revcomp(k, tb, lo.value, hi.value, C.byref(lo), C.byref(hi))
            assert lo.value == i and hi.value == 0


def _queue_get(L, q):
    a, b, p = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.orc_queue_get(q, C.byref(a), C.byref(b), C.byref(p))
    return [a.value | (b.value << 64), p.value]


def test_necklace_queue_kats(oracle_lib, kats):
    nq = kats["necklace_queue"]
    for name, rev in (("forward", 0), ("reverse", 1)):
        c = nq[name]
        q = oracle_lib.orc_queue_new(nq["bits"], nq["width"], rev, nq["t_bits"])
        oracle_lib.orc_queue_insert_full(q, c["start_word"], 0)
        assert _queue_get(oracle_lib, q) == c["expect_start"]
        oracle_lib.orc_queue_insert(q, c["insert_bit"])
        assert _queue_get(oracle_lib, q) == c["expect_after"]
        oracle_lib.orc_queue_free(q)


def test_lex_min_queue_kats(oracle_lib, kats):
    k = kats["lex_min_queue"]
    out = np.zeros(16, dtype=np.uint64)

    def minpos(q):
        n = oracle_lib.orc_lexmin_min_pos(q, out.ctypes.data_as(u64p), 16)
        return [int(v) for v in out[:n]]

    q = oracle_lib.orc_lexmin_new(k["width"])
    vals = np.array(k["insert_full"]["vals"], dtype=np.uint64)
    oracle_lib.orc_lexmin_insert_full(q, vals.ctypes.data_as(u64p), len(vals))
    assert minpos(q) == k["insert_full"]["min_pos"]
    oracle_lib.orc_lexmin_free(q)
    q = oracle_lib.orc_lexmin_new(k["width"])
    for step in k["insert_trace"]:
        oracle_lib.orc_lexmin_insert(q, step["insert"])
        assert minpos(q) == step["min_pos"]
    oracle_lib.orc_lexmin_free(q)


def test_necklace_properties(oracle_lib, kats):
    """src/necklace/mod.rs:46-98 with a seeded generator: revert∘necklace = id, streaming queue ==
    brute force (both directions), periodic 60-bit words (ties -> smallest pos)."""
    P = kats["necklace_properties"]
    rng = np.random.default_rng(12345)
    bits = P["revert"]["bits"]
    words = (rng.integers(0, 1 << 32, size=P["revert"]["n"], dtype=np.uint64) >> np.uint64(1)).astype(np.uint64)
    a, b, p = C.c_uint64(), C.c_uint64(), C.c_uint64()
    for w in words[:20000]:
        oracle_lib.orc_necklace_pos(bits, int(w), 0, C.byref(a), C.byref(b), C.byref(p))
        assert [a.value, p.value] == list(util.necklace_pos_py(int(w), bits))
        oracle_lib.orc_revert_necklace_pos(bits, a.value, 0, p.value, C.byref(a), C.byref(b))
        assert a.value == int(w)
    q = P["queue_equals_brute"]
    for rev in (0, 1):
        assert oracle_lib.orc_necklace_queue_vs_brute(q["bits"], q["width"], rev, words.ctypes.data_as(u64p), None, len(words)) == 0
    per = P["periodic"]
    half = rng.integers(0, 1 << 30, size=per["n"], dtype=np.uint64)
    pw = ((half << np.uint64(30)) | half).astype(np.uint64)
    for rev in (0, 1):
        assert oracle_lib.orc_necklace_queue_vs_brute(per["bits"], per["width"], rev, pw.ctypes.data_as(u64p), None, len(pw)) == 0
    # smallest-pos tie rule on a fully periodic word
    oracle_lib.orc_necklace_pos(60, int(pw[0]), 0, C.byref(a), C.byref(b), C.byref(p))
    assert p.value < 30


def test_sliced_int(oracle_lib, kats):
    s = kats["sliced_int"]
    for v in s["roundtrip"]:
        assert oracle_lib.orc_sliced3_roundtrip(v) == v
    for x, y in s["less"]:
        assert oracle_lib.orc_sliced3_cmp(x, y) < 0 and oracle_lib.orc_sliced3_cmp(y, x) > 0 and oracle_lib.orc_sliced3_cmp(x, x) == 0


def test_bitvector(oracle_lib, kats):
    b = kats["bitvector"]
    n = b["even_test_n"]
    bv = oracle_lib.orc_bv_new(b["bitlength"])
    for i in range(0, 2 * n, 2):
        assert oracle_lib.orc_bv_insert(bv, i) == 1
    for i in range(0, 2 * n, 2):
        assert oracle_lib.orc_bv_contains(bv, i) == 1 and oracle_lib.orc_bv_contains(bv, i + 1) == 0
        assert oracle_lib.orc_bv_rank(bv, i) == i // 2
    for i in range(0, 2 * n, 2):
        assert oracle_lib.orc_bv_count(bv) == n - i // 2
        oracle_lib.orc_bv_remove(bv, i)
    assert oracle_lib.orc_bv_count(bv) == 0
    oracle_lib.orc_bv_free(bv)
    bv = oracle_lib.orc_bv_new(b["bitlength"])
    for i in b["iter"]:
        oracle_lib.orc_bv_insert(bv, i)
    out = np.zeros(16, dtype=np.uint64)
    m = oracle_lib.orc_bv_iter(bv, out.ctypes.data_as(u64p), 16)
    assert [int(v) for v in out[:m]] == b["iter"]
    # F2: count_ones() == rank(size-1) ignores the last bit
    assert oracle_lib.orc_bv_count(bv) == len(b["iter"]) - 1
    oracle_lib.orc_bv_free(bv)


def test_trie(oracle_lib, kats):
    t = kats["trie"]
    tr = oracle_lib.orc_trie3_new()

    def arr(x):
        return (C.c_uint8 * 3)(*x)

    for w in t["membership"]["insert"]:
        oracle_lib.orc_trie3_insert(tr, arr(w))
    assert oracle_lib.orc_trie3_is_empty(tr) == 0 and oracle_lib.orc_trie3_count(tr) == 3
    for w in t["membership"]["insert"]:
        assert oracle_lib.orc_trie3_contains(tr, arr(w)) == 1
    for w in t["membership"]["absent"]:
        assert oracle_lib.orc_trie3_contains(tr, arr(w)) == 0
    for w in t["membership"]["insert"]:
        oracle_lib.orc_trie3_remove(tr, arr(w))
    assert oracle_lib.orc_trie3_is_empty(tr) == 1
    for w in t["insert_order"]:
        oracle_lib.orc_trie3_insert(tr, arr(w))
    out = np.zeros(3 * 8, dtype=np.uint8)
    m = oracle_lib.orc_trie3_iter(tr, out.ctypes.data_as(u8p), 8)
    assert out[: 3 * m].reshape(m, 3).tolist() == t["iter"]
    oracle_lib.orc_trie3_free(tr)


def test_tiered(oracle_lib, kats):
    t = kats["tiered"]
    tv = oracle_lib.orc_tiered_new()
    for i in t["insert"]:
        oracle_lib.orc_tiered_insert(tv, i, i)
    for i in t["insert"]:
        assert oracle_lib.orc_tiered_get(tv, i) == i
    oracle_lib.orc_tiered_remove(tv, t["remove"])
    assert oracle_lib.orc_tiered_get(tv, t["get_after"][0]) == t["get_after"][1]
    assert oracle_lib.orc_tiered_len(tv) == len(t["insert"]) - 1
    oracle_lib.orc_tiered_free(tv)


def test_wordset(oracle_lib, kats):
    w = kats["wordset"]
    ws = oracle_lib.orc_ws1_new(w["prefix_bits"], w["suffix_bits"])
    for x in w["iter_insert"]:
        assert oracle_lib.orc_ws1_insert(ws, x) == 1
    out = np.zeros(16, dtype=np.uint64)
    m = oracle_lib.orc_ws1_iter(ws, out.ctypes.data_as(u64p), 16)
    assert [int(v) for v in out[:m]] == w["iter_expect"]
    oracle_lib.orc_ws1_free(ws)
    # src/wordset/mod.rs:456-516 (shuffled single ops, then batch ops)
    n = w["batch_n"]
    rng = np.random.default_rng(42)
    v0 = np.arange(0, 2 * n, 2, dtype=np.uint64)
    v1 = v0 + np.uint64(1)
    ws = oracle_lib.orc_ws1_new(w["prefix_bits"], w["suffix_bits"])
    for x in rng.permutation(v0)[:20000]:
        assert oracle_lib.orc_ws1_insert(ws, int(x)) == 1
    assert oracle_lib.orc_ws1_count(ws) == 20000
    oracle_lib.orc_ws1_insert_batch(ws, v0.ctypes.data_as(u64p), n)
    assert oracle_lib.orc_ws1_count(ws) == n
    res = np.zeros(n, dtype=np.uint8)
    oracle_lib.orc_ws1_contains_batch(ws, v0.ctypes.data_as(u64p), n, res.ctypes.data_as(u8p))
    assert res.all()
    oracle_lib.orc_ws1_contains_batch(ws, v1.ctypes.data_as(u64p), n, res.ctypes.data_as(u8p))
    assert not res.any()
    oracle_lib.orc_ws1_remove_batch(ws, v0.ctypes.data_as(u64p), n)
    assert oracle_lib.orc_ws1_is_empty(ws) == 1 and oracle_lib.orc_ws1_count(ws) == 0
    oracle_lib.orc_ws1_free(ws)


def _ws_from(L, pb, sb, vals):
    ws = L.orc_ws1_new(pb, sb)
    a = np.ascontiguousarray(vals, dtype=np.uint64)
    L.orc_ws1_insert_batch(ws, a.ctypes.data_as(u64p), len(a))
    return ws


def _ws_items(L, ws):
    n = L.orc_ws1_count(ws)
    out = np.zeros(max(n, 1), dtype=np.uint64)
    m = L.orc_ws1_iter(ws, out.ctypes.data_as(u64p), len(out))
    return out[:m]


def test_wordset_set_ops(oracle_lib, kats):
    """src/wordset/set_ops.rs:424-668: residues mod 3 for the four ops (both forms) and the exact
    k-way merge sequence."""
    k = kats["wordset_set_ops"]
    pb, sb, n = k["prefix_bits"], k["suffix_bits"], k["mod3_n"]
    v = [np.arange(r, 3 * n, 3, dtype=np.uint64) for r in range(3)]
    import functools

    def pyset(a):
        return set(int(x) for x in a)

    A, B = np.concatenate([v[0], v[1]]), np.concatenate([v[1], v[2]])
    expect = {0: pyset(A) | pyset(B), 1: pyset(A) & pyset(B), 2: pyset(A) - pyset(B), 3: pyset(A) ^ pyset(B)}
    for op in range(4):
        a, b = _ws_from(oracle_lib, pb, sb, A), _ws_from(oracle_lib, pb, sb, B)
        r = oracle_lib.orc_ws1_binary_op(op, a, b)
        assert pyset(_ws_items(oracle_lib, r)) == expect[op]
        assert oracle_lib.orc_ws1_count(r) == len(expect[op])
        oracle_lib.orc_ws1_assign_op(op, a, b)
        assert pyset(_ws_items(oracle_lib, a)) == expect[op]
        for h in (a, b, r):
            oracle_lib.orc_ws1_free(h)
    mm = k["multi_merge"]
    c, n = mm["c"], mm["n"]
    sets = [_ws_from(oracle_lib, pb, sb, np.arange(i, c * n, c, dtype=np.uint64)) for i in range(c)]
    arr = (C.c_void_p * c)(*sets)
    merged = oracle_lib.orc_ws1_merge_many(arr, c, 0)
    assert np.array_equal(_ws_items(oracle_lib, merged), np.arange(c * n, dtype=np.uint64))
    inter = oracle_lib.orc_ws1_merge_many(arr, c, 1)
    assert oracle_lib.orc_ws1_is_empty(inter) == 1
    for h in sets + [merged, inter]:
        oracle_lib.orc_ws1_free(h)


def test_worked_vectors(oracle_lib, kats):
    wv = kats["cbl_worked_vectors"]
    c = OracleCBL(7, 32, 14, lib=oracle_lib)
    for e in wv["k7_p14"]:
        x = util.kmer_int(e["nucs"].encode())
        assert x == e["kmer"]
        assert c.get_word(x) == e["word"] == util.word_py(x, 7, False)
        assert e["word"] >> 4 == e["prefix"] and e["word"] & 15 == e["suffix"]
        assert c.recover_kmer(e["word"]) == x
    e = wv["k25_p24"]
    x = util.kmer_int(e["nucs"].encode())
    assert x == e["kmer"]
    c = OracleCBL(25, 64, 24, lib=oracle_lib)
    assert c.get_word(x) == e["word"] and e["word"] >> 32 == e["prefix"] and e["word"] & 0xFFFFFFFF == e["suffix"]
    cc = OracleCBL(25, 64, 24, canonical=True, lib=oracle_lib)
    assert cc.get_word(x) == e["canonical"]["word"] == util.word_py(x, 25, True)
    lo, hi = c.seq_words(e["nucs"].encode())
    assert util.to_int_list(lo, hi) == [e["word"]]


@pytest.mark.parametrize("k,tb,pb", [(7, 32, 14), (25, 64, 24), (31, 128, 24), (59, 128, 24), (59, 128, 28)])
@pytest.mark.parametrize("canonical", [False, True])
def test_streaming_words_equal_normative(oracle_lib, k, tb, pb, canonical):
    """get_seq_words (streaming queues, chunks, F6 order) == the normative per-k-mer definition."""
    seq = util.random_dna(5000, seed=k * 10 + canonical).tobytes()
    c = OracleCBL(k, tb, pb, canonical=canonical, lib=oracle_lib)
    lo, hi = c.seq_words(seq)
    assert util.to_int_list(lo, hi) == util.seq_words_py(seq, k, canonical)


def test_type_too_small(oracle_lib):
    with pytest.raises(pyoracle.OracleError, match="Cannot fit a 31-mer"):
        OracleCBL(31, 64, 24, lib=oracle_lib)  # SURVEY F3 / src/cbl.rs:87-91
    c = OracleCBL(25, 64, 24, lib=oracle_lib)
    with pytest.raises(pyoracle.OracleError, match="smaller than K"):
        c.insert_seq(b"ACGT")


@pytest.mark.parametrize("k,tb,pb", [(59, 128, 24), (25, 64, 24), (7, 32, 14)])
def test_cbl_batch_and_single_ops(oracle_lib, k, tb, pb):
    """src/cbl.rs:592-773 at reduced N with seeded inputs."""
    n = 30000
    seq = util.random_dna(n, seed=k).tobytes()
    kmers = [util.kmer_int(seq[i : i + k]) for i in range(0, n - k + 1, 37)]
    for canonical in (False, True):
        s = OracleCBL(k, tb, pb, canonical=canonical, lib=oracle_lib)
        s.insert_seq(seq)
        assert s.contains_seq(seq).all() and s.contains_all(seq)
        words = set(util.seq_words_py(seq, k, canonical))
        assert s.count() == len(words)
        lo, hi = s.iter_words(sorted_=True)
        assert util.to_int_list(lo, hi) == sorted(words)
        for x in kmers[:200]:
            assert s.contains(x)
            if canonical:
                assert s.contains(util.revcomp_int(x, k))
        s2 = s.clone()
        s.remove_seq(seq)
        assert s.is_empty() and s.count() == 0 and not s.contains_seq(seq).any()
        assert s2.count() == len(words)
        t = OracleCBL(k, tb, pb, canonical=canonical, lib=oracle_lib)
        uniq = sorted(set(kmers))
        for x in uniq:
            t.insert(x)
        fresh = [x for x in uniq if (not canonical) or util.revcomp_int(x, k) not in uniq or True]
        for x in fresh:
            assert t.contains(x)
        removed = sum(1 for x in uniq if t.remove(x))
        assert t.is_empty() and removed == t.count() + removed


def test_cbl_iter_kat(oracle_lib):
    """src/cbl.rs:764-773: kmers 0,7,14,.. < 1000 ; sorted(iter) == inserted."""
    s = OracleCBL(59, 128, 24, lib=oracle_lib)
    kmers = list(range(0, 1000, 7))
    for x in kmers:
        assert s.insert(x)
    lo, hi = s.iter_words(sorted_=False)
    rec = sorted(s.recover_kmer(w) for w in util.to_int_list(lo, hi))
    assert rec == kmers


@pytest.mark.parametrize("k,tb,pb", [(59, 128, 24), (7, 32, 14)])
def test_cbl_set_ops(oracle_lib, k, tb, pb):
    """src/cbl.rs:776-914 at reduced N: | & - ^ (both forms), k-way merge / intersect."""
    n = 20000
    a_seq, b_seq = util.random_dna(n, 1).tobytes(), util.random_dna(n, 2).tobytes()
    b_seq = a_seq[: n // 2] + b_seq[n // 2 :]
    wa, wb = set(util.seq_words_py(a_seq, k, False)), set(util.seq_words_py(b_seq, k, False))
    expect = {0: wa | wb, 1: wa & wb, 2: wa - wb, 3: wa ^ wb}
    for op in range(4):
        a, b = OracleCBL(k, tb, pb, lib=oracle_lib), OracleCBL(k, tb, pb, lib=oracle_lib)
        a.insert_seq(a_seq)
        b.insert_seq(b_seq)
        r = a.binary_op(op, b)
        assert set(util.to_int_list(*r.iter_words())) == expect[op] and r.count() == len(expect[op])
        a.assign_op(op, b)
        assert set(util.to_int_list(*a.iter_words())) == expect[op]
        assert set(util.to_int_list(*b.iter_words())) == wb
    parts = [OracleCBL(k, tb, pb, lib=oracle_lib) for _ in range(4)]
    for i, p in enumerate(parts):
        p.insert_seq(a_seq[i * 3000 : i * 3000 + 8000])
    sets = [set(util.seq_words_py(a_seq[i * 3000 : i * 3000 + 8000], k, False)) for i in range(4)]
    m = OracleCBL.merge(parts)
    assert set(util.to_int_list(*m.iter_words())) == set().union(*sets)
    it = OracleCBL.intersect(parts)
    assert set(util.to_int_list(*it.iter_words())) == set.intersection(*sets)
    c = OracleCBL(k, tb, pb, canonical=True, lib=oracle_lib)
    with pytest.raises(pyoracle.OracleError, match="canonical"):
        parts[0].binary_op(0, c)


def test_non_acgt_dropped(oracle_lib):
    """non-ACGT bytes are skipped by filter_map while chunking is on raw offsets."""
    c = OracleCBL(7, 32, 14, lib=oracle_lib)
    seq = b"ACGTNACGTACGTTTGA"
    lo, hi = c.seq_words(seq)
    # first window = valid bases among the first 7 bytes (6 bases), then one word per later valid base
    assert len(lo) == 1 + sum(1 for ch in seq[7:] if ch in b"ACGTacgt")


def test_serde_roundtrip(oracle_lib):
    """Appendix A item 12 — parity unpinned (no reference test), so only self-consistency."""
    seq = util.random_dna(60000, 5).tobytes()
    for k, tb, pb in [(25, 64, 24), (59, 128, 28), (7, 32, 14)]:
        c = OracleCBL(k, tb, pb, canonical=True, lib=oracle_lib)
        c.insert_seq(seq)
        blob = c.serialize()
        d = c.deserialize(blob)
        assert d.is_canonical() and d.count() == c.count()
        assert util.to_int_list(*d.iter_words()) == util.to_int_list(*c.iter_words())
        assert d.serialize() == blob
    with pytest.raises(pyoracle.OracleError):
        c.deserialize(blob + b"\0")


def test_non_acgt_equals_padded_clean_chunks(oracle_lib):
    """The equivalence the GPU slow path (cbl_b200/csrc/sanitize.cuh) is built on: a chunk with non-ACGT bytes
    yields the words of the clean string 'A'*(K-a) ++ valid(first K bytes) ++ valid(rest)."""
    valid = set(b"ACGTacgt")
    for k, tb, pb, canonical in [(7, 32, 14, False), (25, 64, 24, True), (31, 128, 24, False), (59, 128, 28, True)]:
        c = OracleCBL(k, tb, pb, canonical=canonical, lib=oracle_lib)
        rng = np.random.default_rng(k)
        s = util.random_dna(9000, seed=k).copy()
        s[rng.integers(0, len(s), size=120)] = ord("N")
        s[2040:2040 + k + 3] = ord("n")
        seq = s.tobytes()
        expect = util.to_int_list(*c.seq_words(seq))
        got = []
        for start in range(0, len(seq) - k + 1, 2048):
            chunk = seq[start:min(start + 2048 + k - 1, len(seq))]
            head = bytes(ch for ch in chunk[:k] if ch in valid)
            rest = bytes(ch for ch in chunk[k:] if ch in valid)
            clean = b"A" * (k - len(head)) + head + rest
            got += util.to_int_list(*c.seq_words(clean))
        assert got == expect
