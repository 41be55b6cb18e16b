"""GPU parity tests: the CUDA path (through the C ABI, include/cbl_gpu.h) against the CPU oracle on
the same seeded inputs — bit-exact words, set contents (ascending word order), per-k-mer
contains_seq answers (reference order), set operations, iteration and serde.  Run on the B200 box:
    python -m pytest tests -m gpu
"""
import os

import numpy as np
import pytest

import cbl_testutil as util

pytestmark = pytest.mark.gpu

CONFIGS = [(7, 32, 14), (25, 64, 24), (29, 64, 24), (31, 128, 24), (59, 128, 24), (59, 128, 28)]


@pytest.fixture(scope="module")
def gpu():
    import cbl_b200

    return cbl_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle

    return pyoracle


def ints(lo, hi):
    return util.to_int_list(lo, hi)


def first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return f"first difference at {i}: got {a[i]:#x} expected {b[i]:#x} (lens {len(a)}/{len(b)})"
    return f"length mismatch {len(a)} vs {len(b)}"


def assert_same(a, b, what=""):
    a, b = list(a), list(b)
    assert a == b, what + " " + first_diff(a, b)


def low_complexity(n, seed):
    """Mostly-A sequence: necklaces crowd into the first prefixes (large buckets, heavy skew)."""
    rng = np.random.default_rng(seed)
    s = np.full(n, ord("A"), dtype=np.uint8)
    idx = rng.integers(0, n, size=n // 12)
    s[idx] = util.BASES[rng.integers(0, 4, size=len(idx))]
    return s


# ------------------------------------------------------------------------------------------------
# k1 + k2: encode + necklace
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,tb,pb", CONFIGS)
@pytest.mark.parametrize("canonical", [False, True])
def test_seq_words_match_oracle(gpu, orc, k, tb, pb, canonical):
    g = gpu.CBL(k, tb, pb, canonical)
    o = orc.OracleCBL(k, tb, pb, canonical)
    lens = [k, k + 1, k + 31, k + 32, 100, 1023 + k, 1024 + k, 2047 + k, 2048 + k - 1, 2048 + k, 4096 + k - 1, 5000, 70001]
    recs = [util.random_dna(n, seed=1000 + i).tobytes() for i, n in enumerate(lens)]
    recs.append(b"A" * 300)
    recs.append(b"ACGT" * 200 + b"g" * 77)
    recs.append(util.random_dna(3000, 7).tobytes().lower())
    recs.append(low_complexity(9000, 3).tobytes())
    expect = []
    for r in recs:
        expect += ints(*o.seq_words(r))
    for brute in (False, True):
        lo, hi = g.seq_words(recs, brute=brute)
        assert_same(ints(lo, hi), expect, f"seq_words brute={brute}")
    # one record at a time, unaligned starts inside a shared buffer
    blob = b"T" * 3 + b"".join(recs)
    arr = np.frombuffer(blob, dtype=np.uint8)
    off = 3
    for r in recs[:6]:
        lo, hi = g.seq_words([arr[off : off + len(r)]])
        assert_same(ints(lo, hi), ints(*o.seq_words(r)), "single record")
        off += len(r)


def test_worked_vectors(gpu, kats):
    wv = kats["cbl_worked_vectors"]
    g = gpu.CBL(7, 32, 14)
    for e in wv["k7_p14"]:
        lo, hi = g.seq_words([e["nucs"].encode()])
        assert ints(lo, hi) == [e["word"]]
    e = wv["k25_p24"]
    assert ints(*gpu.CBL(25, 64, 24).seq_words([e["nucs"].encode()])) == [e["word"]]
    assert ints(*gpu.CBL(25, 64, 24, True).seq_words([e["nucs"].encode()])) == [e["canonical"]["word"]]


def test_revcomp_kats_through_canonical_words(gpu, kats):
    """src/kmer.rs:355-378 via the GPU: in canonical mode a k-mer and its reverse complement map to
    the same word."""
    for case in kats["revcomp"]["cases"]:
        k = case["k"]
        if k % 2 == 0:
            continue
        g = gpu.CBL(k, 64, 10, True)
        a = ints(*g.seq_words([case["nucs"].encode()]))
        b = ints(*g.seq_words([case["rc"].encode()]))
        assert a == b


# ------------------------------------------------------------------------------------------------
# insert / contains / remove against the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,tb,pb", CONFIGS)
@pytest.mark.parametrize("canonical", [False, True])
def test_insert_contains_remove_match_oracle(gpu, orc, k, tb, pb, canonical):
    g = gpu.CBL(k, tb, pb, canonical)
    o = orc.OracleCBL(k, tb, pb, canonical)
    assert g.is_empty() and g.count() == 0
    a = util.random_dna(40000, 11).tobytes()
    b = a[10000:30000] + util.random_dna(20000, 12).tobytes()
    c = low_complexity(30000, 13).tobytes()
    for s in (a, b, c, a):
        g.insert_seq(s)
        o.insert_seq(s)
        assert g.count() == o.count()
        assert_same(g.words(), ints(*o.iter_words()), "set contents after insert")
    probe = util.random_dna(20000, 14).tobytes() + a[5000:15000] + c[:5000]
    assert_same(g.contains_seq(probe).astype(np.uint8), o.contains_seq(probe), "contains_seq answers")
    assert g.contains_all(a) and o.contains_all(a)
    assert g.contains_all(probe) == o.contains_all(probe)
    pb_, sz = g.buckets_sizes()
    op, osz = o.bucket_sizes()
    assert_same(pb_, op, "bucket prefixes")
    assert_same(sz, osz, "bucket sizes")
    for s in (b, c):
        g.remove_seq(s)
        o.remove_seq(s)
        assert g.count() == o.count()
        assert_same(g.words(), ints(*o.iter_words()), "set contents after remove")
        assert_same(g.contains_seq(probe).astype(np.uint8), o.contains_seq(probe), "contains_seq after remove")
    g.remove_seq(a)
    o.remove_seq(a)
    assert g.count() == 0 == o.count() and g.is_empty() and o.is_empty()
    assert not g.contains_seq(a).any()
    g.insert_seq(b)  # a set emptied by removals is reusable
    o.insert_seq(b)
    assert_same(g.words(), ints(*o.iter_words()), "reinsert after emptying")


def test_batch_records_and_small_internal_batches(gpu, orc, monkeypatch):
    """Many records in one call, with the internal sort batch and host group size forced tiny so the
    multi-batch / multi-group code paths run."""
    monkeypatch.setenv("CBL_BATCH_KMERS", "6000")
    monkeypatch.setenv("CBL_GROUP_BYTES", "20000")
    k, tb, pb = 25, 64, 24
    g = gpu.CBL(k, tb, pb, True)
    o = orc.OracleCBL(k, tb, pb, True)
    rng = np.random.default_rng(5)
    recs = [util.random_dna(int(n), 100 + i) for i, n in enumerate(rng.integers(k, 9000, size=40))]
    buf, offs = gpu.concat_records(recs)
    g.insert_seqs(buf, offs)
    for r in recs:
        o.insert_seq(r)
    assert_same(g.words(), ints(*o.iter_words()), "batched insert")
    q = [util.random_dna(int(n), 300 + i) for i, n in enumerate(rng.integers(k, 9000, size=20))] + recs[::3]
    qbuf, qoffs = gpu.concat_records(q)
    got = g.contains_seqs(qbuf, qoffs)
    exp = np.concatenate([o.contains_seq(r) for r in q])
    assert_same(got, exp, "batched contains")
    g.remove_seqs(buf, offs)
    assert g.is_empty()


def test_kmer_api_and_iter(gpu, orc):
    """src/cbl.rs:764-773 (iter KAT) and the single-k-mer API (src/cbl.rs:219-235)."""
    g = gpu.CBL(59, 128, 24)
    kmers = list(range(0, 1000, 7))
    for x in kmers[:5]:
        assert g.insert(x) is True
        assert g.insert(x) is False
    before = g.insert_kmers(kmers)
    assert before[:5].all() and not before[5:].any()
    assert sorted(g.iter()) == kmers
    assert g.contains_kmers(kmers).all() and not g.contains_kmers([1, 2, 3]).any()
    assert g.remove(kmers[0]) is True and g.remove(kmers[0]) is False
    assert g.count() == len(kmers) - 1
    # canonical: both strands hit (src/cbl.rs:686-761)
    k = 31
    c = gpu.CBL.new_canonical(k, 128, 24)
    oc = orc.OracleCBL(k, 128, 24, True)
    seq = util.random_dna(3000, 21).tobytes()
    xs = [util.kmer_int(seq[i : i + k]) for i in range(0, 2000, 13)]
    c.insert_kmers(xs)
    for x in xs:
        oc.insert(x)
    assert c.contains_kmers(xs).all() and c.contains_kmers([util.revcomp_int(x, k) for x in xs]).all()
    assert_same(c.words(), ints(*oc.iter_words()), "canonical k-mer inserts")
    assert sorted(c.iter()) == sorted(oc.recover_kmer(w) for w in ints(*oc.iter_words()))


# ------------------------------------------------------------------------------------------------
# set operations
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,tb,pb", [(7, 32, 14), (25, 64, 24), (31, 128, 24), (59, 128, 28)])
def test_set_ops_match_oracle(gpu, orc, k, tb, pb):
    n = 30000
    a_seq = util.random_dna(n, 31).tobytes()
    b_seq = a_seq[: n // 2] + util.random_dna(n // 2, 32).tobytes() + low_complexity(5000, 33).tobytes()
    for op in range(4):
        ga, gb = gpu.CBL(k, tb, pb), gpu.CBL(k, tb, pb)
        oa, ob = orc.OracleCBL(k, tb, pb), orc.OracleCBL(k, tb, pb)
        for s_, t_ in ((ga, a_seq), (gb, b_seq), (oa, a_seq), (ob, b_seq)):
            s_.insert_seq(t_)
        gr = [ga | gb, ga & gb, ga - gb, ga ^ gb][op]
        orr = oa.binary_op(op, ob)
        assert gr.count() == orr.count()
        assert_same(gr.words(), ints(*orr.iter_words()), f"out-of-place op {op}")
        assert_same(ga.words(), ints(*oa.iter_words(True)), "left operand untouched")
        if op == 0:
            ga |= gb
        elif op == 1:
            ga &= gb
        elif op == 2:
            ga -= gb
        else:
            ga ^= gb
        oa.assign_op(op, ob)
        assert_same(ga.words(), ints(*oa.iter_words()), f"assign op {op}")
        assert_same(gb.words(), ints(*ob.iter_words()), "right operand untouched")
        # the result is a live index
        assert_same(ga.contains_seq(b_seq).astype(np.uint8), oa.contains_seq(b_seq), "contains after set op")
    empty = gpu.CBL(k, tb, pb)
    full = gpu.CBL(k, tb, pb)
    full.insert_seq(a_seq)
    assert (full | empty).words() == full.words() == (empty | full).words()
    assert (full & empty).count() == 0 and (empty & full).count() == 0
    assert (full - empty).words() == full.words() and (empty - full).count() == 0
    assert (full ^ empty).words() == full.words() and (full ^ full).count() == 0 and (full - full).is_empty()
    assert (full | full).words() == full.words() == (full & full).words()


def test_multi_merge_intersect(gpu, orc):
    """src/cbl.rs:866-914."""
    k, tb, pb = 7, 32, 14
    seq = util.random_dna(12000, 41).tobytes()
    parts = [seq[i * 2000 : i * 2000 + 5000] for i in range(4)]
    gs, os_ = [], []
    for p in parts:
        g, o = gpu.CBL(k, tb, pb), orc.OracleCBL(k, tb, pb)
        g.insert_seq(p)
        o.insert_seq(p)
        gs.append(g)
        os_.append(o)
    assert_same(gpu.CBL.merge(gs).words(), ints(*orc.OracleCBL.merge(os_).iter_words()), "k-way merge")
    assert_same(gpu.CBL.intersect(gs).words(), ints(*orc.OracleCBL.intersect(os_).iter_words()), "k-way intersect")
    # odd number of inputs of very different sizes (balanced-tree union, smallest-first intersection), one input, K=25
    k, tb, pb = 25, 64, 24
    base = util.random_dna(60000, 43)
    cuts = [(0, 60000), (100, 30100), (25000, 26000), (5000, 50000), (20000, 45000)]
    gs, os_ = [], []
    for a, b in cuts:
        g, o = gpu.CBL(k, tb, pb), orc.OracleCBL(k, tb, pb)
        g.insert_seq(base[a:b])
        o.insert_seq(base[a:b])
        gs.append(g)
        os_.append(o)
    assert_same(gpu.CBL.merge(gs).words(), ints(*orc.OracleCBL.merge(os_).iter_words()), "5-way merge")
    assert_same(gpu.CBL.intersect(gs).words(), ints(*orc.OracleCBL.intersect(os_).iter_words()), "5-way intersect")
    assert_same(gpu.CBL.merge(gs[:1]).words(), ints(*os_[0].iter_words()), "1-way merge")
    assert_same(gpu.CBL.intersect(gs[2:3]).words(), ints(*os_[2].iter_words()), "1-way intersect")
    assert_same(gs[0].words(), ints(*os_[0].iter_words()), "inputs untouched")


def test_stats_api(gpu, orc):
    """prefix_load / buckets_sizes / buckets_size_count / buckets_load_repartition / buckets_nodes / buckets_node_count
    (src/cbl.rs:364-396, src/wordset/mod.rs:254-295) against the same arithmetic on the oracle's bucket list; node counts of
    buckets above the trie threshold against a plain-Python byte trie."""
    k, tb, pb = 11, 32, 6          # 64 prefixes, necklace skew => several buckets far above 1024 suffixes
    g, o = gpu.CBL(k, tb, pb), orc.OracleCBL(k, tb, pb)
    seq = util.random_dna(200000, 77)
    g.insert_seq(seq)
    o.insert_seq(seq)
    op, osz = o.bucket_sizes()
    order = np.argsort(op, kind="stable")
    op, osz = op[order], osz[order]
    assert g.prefix_load() == len(op) / float(1 << pb)
    sc = {}
    for s_ in osz:
        sc[int(s_)] = sc.get(int(s_), 0) + 1
    assert g.buckets_size_count() == sc
    total = float(sum(a * b for a, b in sc.items()))
    rep = g.buckets_load_repartition()
    assert set(rep) == set(sc) and all(abs(rep[a] - a * b / total) < 1e-12 for a, b in sc.items())
    gp, gn = g.buckets_nodes()
    assert np.array_equal(gp.astype(np.uint64), op)
    sb = 2 * k + util.pos_bits(k) - pb
    nbytes = (sb + 7) // 8
    words = ints(*o.iter_words())
    assert max(osz) > g.TRIE_THRESHOLD
    for prefix, size, nodes in zip(op, osz, gn):
        if size <= g.TRIE_THRESHOLD:
            assert nodes == size
        else:
            suf = [w & ((1 << sb) - 1) for w in words if (w >> sb) == int(prefix)]
            trie = {(): None}
            for x in suf:
                bs = x.to_bytes(nbytes, "big")
                for d in range(1, nbytes):
                    trie[bs[:d]] = None
            assert nodes == len(trie), (int(prefix), int(size))
    nc = g.buckets_node_count()
    assert sum(nc.values()) == len(op)


# ------------------------------------------------------------------------------------------------
# serde (reference bincode layout) — both directions through the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,tb,pb", [(25, 64, 24), (59, 128, 28), (7, 32, 14)])
def test_serde_roundtrip_and_interop(gpu, orc, k, tb, pb, tmp_path):
    g = gpu.CBL(k, tb, pb, True)
    o = orc.OracleCBL(k, tb, pb, True)
    seqs = [util.random_dna(30000, 51).tobytes(), low_complexity(60000, 52).tobytes()]
    for s in seqs:
        g.insert_seq(s)
        o.insert_seq(s)
    blob = g.serialize()
    back = g.deserialize(blob)
    assert back.is_canonical() and back.words() == g.words()
    # GPU-written file is readable by the (restated) reference reader and vice versa; the oracle's own
    # file carries Trie buckets when a bucket holds > 1024 suffixes
    od = o.deserialize(blob)
    assert_same(ints(*od.iter_words()), g.words(), "oracle reads GPU file")
    gd = g.deserialize(o.serialize())
    assert_same(gd.words(), ints(*o.iter_words()), "GPU reads oracle file")
    path = str(tmp_path / "index.cbl")
    g.save_to_file(path)
    assert open(path, "rb").read() == blob
    assert g.load_from_file(path).words() == g.words()
    with pytest.raises(gpu.CBLError):
        g.deserialize(blob + b"\0")
    with pytest.raises(gpu.CBLError, match="Failed to open"):
        g.load_from_file(str(tmp_path / "missing.cbl"))


# ------------------------------------------------------------------------------------------------
# error behaviour (the reference panics; the ABI returns CBL_EINVAL with the same message)
# ------------------------------------------------------------------------------------------------
def test_errors(gpu):
    with pytest.raises(gpu.CBLError, match="Cannot fit a 31-mer"):
        gpu.CBL(31, 64, 24)
    with pytest.raises(gpu.CBLError):
        gpu.CBL(60, 128, 24)
    g = gpu.CBL(25, 64, 24)
    with pytest.raises(gpu.CBLError, match=r"Sequence size \(4\) is smaller than K \(25\)"):
        g.insert_seq(b"ACGT")
    with pytest.raises(gpu.CBLError, match="smaller than K"):
        g.contains_seq(b"ACGT")
    assert g.count() == 0  # a rejected call leaves the set untouched
    c = gpu.CBL.new_canonical(25, 64, 24)
    with pytest.raises(gpu.CBLError, match="One of the index is canonical while the other isn't"):
        g | c
    with pytest.raises(gpu.CBLError):
        g |= gpu.CBL(25, 64, 20)


def test_is_empty_reference_quirk(gpu, orc):
    """a set holding only the all-ones prefix (poly-G) reports is_empty() like the reference."""
    g, o = gpu.CBL(25, 64, 24), orc.OracleCBL(25, 64, 24)
    g.insert_seq(b"G" * 25)
    o.insert_seq(b"G" * 25)
    assert g.count() == 1 == o.count()
    assert g.is_empty() == o.is_empty() == True  # noqa: E712


# ------------------------------------------------------------------------------------------------
# size-independent properties at larger sizes (src/cbl.rs:665-683, 776-863)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,tb,pb,n", [(25, 64, 24, 6_000_000), (59, 128, 28, 2_000_000), (31, 128, 24, 2_000_000)])
def test_large_roundtrip_properties(gpu, k, tb, pb, n):
    seq = util.random_dna(n, 61)
    other = util.random_dna(n // 2, 62)
    g = gpu.CBL(k, tb, pb)
    g.insert_seq(seq)
    cnt = g.count()
    assert 0 < cnt <= n - k + 1
    assert g.contains_seq(seq).all()
    hits = int(g.contains_seq(other).sum())
    assert hits < len(other) // 100  # random 25+-mers essentially never collide
    lo, hi = g.words_arrays()
    w = [int(a) | (int(b) << 64) for a, b in zip(lo[:200000], hi[:200000])]
    assert all(x < y for x, y in zip(w, w[1:])), "iteration must be strictly ascending"
    assert len(lo) == cnt
    h = gpu.CBL(k, tb, pb)
    h.insert_seq(other)
    u = g | h
    assert u.count() == cnt + h.count() - (g & h).count()
    assert (u - h).count() == (g - h).count() and (g ^ h).count() == u.count() - (g & h).count()
    g.insert_seq(seq)  # idempotent
    assert g.count() == cnt
    g.remove_seq(seq)
    assert g.is_empty() and g.count() == 0 and not g.contains_seq(seq[:100000]).any()


# ------------------------------------------------------------------------------------------------
# multi-GPU building blocks on one GPU: router (stable partition by owner rank) + answer gather
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,tb,pb", [(25, 64, 24), (59, 128, 28)])
def test_route_and_gather(gpu, k, tb, pb):
    import torch

    from cbl_b200.sharded import GpuEngine, equal_mass_splitters, route, word_prefixes

    eng = GpuEngine(k, tb, pb, False, 0)
    sb = 2 * k + util.pos_bits(k) - pb
    words = eng.sample_words(300_000, seed=5)
    pre = word_prefixes(words, sb, pb)
    for world in (1, 2, 8):
        sp = equal_mass_splitters(pre, world)
        send, pos, counts = eng.route(words, sp.cpu().numpy().astype(np.uint32))
        _, order, tcounts = route(pre, sp)
        assert counts.tolist() == tcounts.tolist()
        assert torch.equal(send, words.index_select(0, order)), "router must be a STABLE partition by owner"
        assert torch.equal(send[pos.long()], words), "pos maps every word to its slot"
        flags = (torch.arange(send.shape[0], device=send.device) % 251).to(torch.uint8)
        assert torch.equal(eng.gather(flags, pos), flags[pos.long()])
        if world > 1:
            assert counts.min() > 0.8 * counts.max()  # equal-mass splitters balance random DNA


def test_word_probe_tolerates_arbitrary_bit_patterns(gpu):
    """cbl_words_op_dev(op = contains) with words that cannot come from a k-mer (all ones = the exchange buffers' "nothing
    here" pattern, prefixes beyond 2^PREFIX_BITS, random bits): every answer is 0, nothing is read out of bounds."""
    import torch

    g = gpu.CBL(25, 64, 24)
    g.insert_seq(util.random_dna(50_000, seed=3).tobytes())
    rng = np.random.default_rng(9)
    junk = rng.integers(-(2 ** 63), 2 ** 63 - 1, size=200_000, dtype=np.int64) | np.int64(-(2 ** 62))   # top bits set: wider than a word
    junk[::7] = -1
    real_lo, _ = g.words_arrays()
    words = torch.from_numpy(np.concatenate([junk, real_lo[:1000].view(np.int64)])).cuda()
    out = torch.full((words.numel(),), 7, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    g.words_op_dev(0, words.data_ptr(), words.numel(), out.data_ptr())
    g.sync()
    res = out.cpu().numpy()
    assert not res[: len(junk)].any() and res[len(junk):].all()


# ------------------------------------------------------------------------------------------------
# non-ACGT bytes are dropped by filter_map while chunking stays on raw byte offsets
# (src/kmer.rs:133-135, src/cbl.rs:239-289) — reproduced on the GPU by the sanitising slow path
# ------------------------------------------------------------------------------------------------
def with_junk(n, seed, every, junk=b"NnRY-*\n"):
    """random DNA with a non-nucleotide byte roughly every `every` positions (plus the odd run of them)"""
    rng = np.random.default_rng(seed)
    s = util.random_dna(n, seed).copy()
    idx = rng.integers(0, n, size=max(1, n // every))
    s[idx] = np.frombuffer(junk, dtype=np.uint8)[rng.integers(0, len(junk), size=len(idx))]
    if n > 400:
        s[300:340] = ord("N")   # a run longer than K=25: a chunk whose first K bytes hold few valid bases
    return s.tobytes()


@pytest.mark.parametrize("k,tb,pb", [(7, 32, 14), (25, 64, 24), (31, 128, 24), (59, 128, 28)])
@pytest.mark.parametrize("canonical", [False, True])
def test_non_acgt_words_match_reference_behaviour(gpu, orc, k, tb, pb, canonical):
    g = gpu.CBL(k, tb, pb, canonical)
    o = orc.OracleCBL(k, tb, pb, canonical)
    recs = [with_junk(n, seed=40 + i, every=e) for i, (n, e) in enumerate([(k, 3), (k + 5, 4), (500, 50), (2048 + k - 1, 100), (2048 + k, 7),
                                                                           (4096 + k + 3, 300), (10000, 2000), (70001, 500)])]
    recs.append(b"N" * (k + 40))                       # nothing valid at all: one all-A word per chunk
    recs.append(b"N" * k + b"ACGT" * 20)               # empty first k-mer
    recs.append(util.random_dna(5000, 3).tobytes())    # a clean record inside a dirty batch
    for r in recs:                                      # one record per call
        exp = ints(*o.seq_words(r))
        lo, hi = g.seq_words([r])
        assert_same(ints(lo, hi), exp, f"F8 seq_words len={len(r)}")
        assert g.last_kmer_count() == len(exp)
    expect = []
    for r in recs:
        expect += ints(*o.seq_words(r))
    lo, hi = g.seq_words(recs)                          # the whole batch in one call
    assert_same(ints(lo, hi), expect, "F8 seq_words, batch")


@pytest.mark.parametrize("canonical", [False, True])
def test_non_acgt_insert_contains_remove(gpu, orc, canonical):
    k, tb, pb = 25, 64, 24
    g = gpu.CBL(k, tb, pb, canonical)
    o = orc.OracleCBL(k, tb, pb, canonical)
    a = with_junk(60000, seed=5, every=400)
    b = with_junk(30000, seed=6, every=150)
    clean = util.random_dna(20000, 9).tobytes()
    for r in (a, clean):
        g.insert_seq(r)
        o.insert_seq(r)
    assert g.count() == o.count()
    assert_same(g.words(), ints(*o.iter_words()), "set after inserting reads with N")
    for q in (a, b, a[100:9000] + b[:5000], clean):
        got = g.contains_seq(q)
        exp = o.contains_seq(q)
        assert len(got) == len(exp) and np.array_equal(got.astype(np.uint8), exp), "F8 contains_seq"
        assert g.contains_all(q) == bool(o.contains_all(q))
    assert g.contains_all(a)
    # batch entry points: answers compacted in reference order
    buf, offs = gpu.concat_records([a, clean, b])
    got = g.contains_seqs(buf, offs)
    exp = np.concatenate([o.contains_seq(r) for r in (a, clean, b)])
    assert np.array_equal(got, exp)
    g.remove_seq(a)
    o.remove_seq(a)
    assert g.count() == o.count()
    assert_same(g.words(), ints(*o.iter_words()), "set after removing reads with N")


# ------------------------------------------------------------------------------------------------
# k3 + k3b: hybrid batch sort (LSD passes on the top digits + segment sort) == plain LSD sort, on inputs that
# stress the segment sort: random reads, every k-mer repeated a few times (ties inside the bins), every k-mer
# repeated thousands of times (segments longer than a tile -> fallback to the LSD passes), mostly-A reads
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,tb,pb", [(25, 64, 24), (31, 128, 24), (59, 128, 28), (7, 32, 14)])
def test_hybrid_sort_equals_lsd_sort(gpu, orc, k, tb, pb, monkeypatch):
    rng = np.random.default_rng(77)
    block = util.random_dna(6000, 71)
    inputs = {
        "random": util.random_dna(1_500_000, 70),
        "repeat x6": np.tile(block, 6),
        "repeat x5000": np.tile(util.random_dna(400, 72), 5000),
        "mostly A": low_complexity(600_000, 73),
        "tiny": util.random_dna(k + 30, 74),
    }
    for name, seq in inputs.items():
        monkeypatch.setenv("CBL_SORT", "lsd")
        a = gpu.CBL(k, tb, pb)
        a.insert_seq(seq)
        monkeypatch.delenv("CBL_SORT")
        before = gpu.sort_fallback_count()
        b = gpu.CBL(k, tb, pb)
        b.insert_seq(seq)
        fell_back = gpu.sort_fallback_count() - before
        la, ha = a.words_arrays()
        lb, hb = b.words_arrays()
        assert a.count() == b.count() and np.array_equal(la, lb) and np.array_equal(ha, hb), name
        assert b.contains_seq(seq).all(), name
        if name == "random" and k >= 25:
            assert fell_back == 0, "random reads must not need the fallback"
        if name == "repeat x5000" and k >= 25:
            assert fell_back > 0, "a k-mer repeated 5000 times cannot fit a segment tile"
        if name in ("repeat x6", "tiny"):
            o = orc.OracleCBL(k, tb, pb)
            o.insert_seq(seq)
            assert_same(b.words(), ints(*o.iter_words()), name)
        b.remove_seq(seq)
        assert b.is_empty(), name
    del rng
