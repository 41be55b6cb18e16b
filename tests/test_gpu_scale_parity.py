"""GPU parity at BASELINE scale (VERDICT r01 "what's weak" 1): the CUDA path against the CPU oracle on inputs large
enough to take the production code paths — multi-pass hybrid sort with its production pass count, the run-time
segment-sort tile, multi-million-bucket directories, 32-bit look-back offsets — not only the toy sizes of
test_gpu_parity.py.

* C1 (BASELINE.json configs[0]): 10 x 1 Mbp, K=25 / u64 / PREFIX_BITS=24, plain and canonical: EVERY stored word
  (ascending), the count, the whole (prefix, size) bucket list and every contains_seq answer against the oracle.
* K=31 / u128 / 24 and K=59 / u128 / 28 builds of 50 M k-mers: every stored word, the count and the bucket list
  against the oracle's own dynamic set (src/cbl.rs:328-339 restated), plus hit / miss queries.

Mirrors the reference's own property tests at scale: src/cbl.rs:665-683 (insert => contains all => remove => empty)
and :764-773 (iter sorted == inserted).  Run on the B200 box:  python -m pytest tests -m gpu
"""
import numpy as np
import pytest

import cbl_testutil as util

pytestmark = pytest.mark.gpu

REC = 1_000_000


@pytest.fixture(scope="module")
def gpu():
    import cbl_b200

    return cbl_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle

    return pyoracle


def records(n, seed):
    return [util.random_dna(REC, seed=seed * 100_003 + i) for i in range(n)]


def check_set_equal(g, o):
    """every stored word (ascending), count and bucket list of the GPU index vs the oracle set"""
    assert g.count() == o.count()
    glo, ghi = g.words_arrays()
    olo, ohi = o.iter_words()
    assert len(glo) == len(olo)
    bad = np.flatnonzero((glo != olo) | (ghi != ohi))
    assert bad.size == 0, f"{bad.size} stored words differ, first at rank {bad[0]}: gpu {int(ghi[bad[0]]):#x}:{int(glo[bad[0]]):#x} oracle {int(ohi[bad[0]]):#x}:{int(olo[bad[0]]):#x}"
    gp, gs = g.buckets_sizes()
    op, os_ = o.bucket_sizes()
    order = np.argsort(op, kind="stable")
    assert np.array_equal(gp.astype(np.uint64), op[order]) and np.array_equal(gs.astype(np.uint64), os_[order]), "bucket (prefix, size) lists differ"


@pytest.mark.parametrize("canonical", [False, True])
def test_config1_exact(gpu, orc, canonical):
    """BASELINE configs[0]: cbl build on 10 Mbp, K=25, u64, PREFIX_BITS=24 — exact against the oracle."""
    k, tb, pb = 25, 64, 24
    recs = records(10, seed=1)
    g = gpu.CBL(k, tb, pb, canonical)
    o = orc.OracleCBL(k, tb, pb, canonical)
    buf, off = gpu.concat_records(recs)
    fb0 = gpu.sort_fallback_count()
    g.insert_seqs(buf, off)                      # one ABI call for the CLI's `for record { insert_seq }` loop
    for r in recs:
        o.insert_seq(r)
    check_set_equal(g, o)
    assert gpu.sort_fallback_count() == fb0      # the hybrid sort itself sorted the batch (no LSD fallback)
    # every contains_seq answer: the ten indexed records (all hits), two fresh records (misses), one half / half
    queries = recs + records(2, seed=77) + [np.concatenate([recs[3][:400_000], util.random_dna(600_000, seed=5)])]
    qbuf, qoff = gpu.concat_records(queries)
    ans = g.contains_seqs(qbuf, qoff)
    expect = np.concatenate([o.contains_seq(q) for q in queries])
    assert ans.shape == expect.shape
    bad = np.flatnonzero(ans != expect)
    assert bad.size == 0, f"{bad.size} contains_seq answers differ, first at k-mer {bad[0]}"
    assert int(expect[: 10 * (REC - k + 1)].sum()) == 10 * (REC - k + 1)
    # remove half, compare again, remove the rest => empty (src/cbl.rs:665-683)
    hbuf, hoff = gpu.concat_records(recs[:5])
    g.remove_seqs(hbuf, hoff)
    for r in recs[:5]:
        o.remove_seq(r)
    check_set_equal(g, o)
    rbuf, roff = gpu.concat_records(recs[5:])
    g.remove_seqs(rbuf, roff)
    assert g.count() == 0 and g.is_empty()


@pytest.mark.parametrize("k,tb,pb", [(31, 128, 24), (59, 128, 28)])
def test_u128_build_50m_matches_oracle(gpu, orc, k, tb, pb):
    """50 M k-mers in ONE sort batch (production pass count of the hybrid sort: 4 LSD passes + the segment sort for
    both tuples), 128-bit words: every stored word, count and bucket list against the oracle's dynamic set."""
    n_rec = 50
    recs = records(n_rec, seed=k)
    g = gpu.CBL(k, tb, pb, False)
    o = orc.OracleCBL(k, tb, pb, False)
    buf, off = gpu.concat_records(recs)
    fb0 = gpu.sort_fallback_count()
    g.insert_seqs(buf, off)
    for r in recs:
        o.insert_seq(r)
    assert gpu.sort_fallback_count() == fb0
    check_set_equal(g, o)
    queries = [recs[7], recs[n_rec - 1], util.random_dna(REC, seed=424242)]
    qbuf, qoff = gpu.concat_records(queries)
    ans = g.contains_seqs(qbuf, qoff)
    expect = np.concatenate([o.contains_seq(q) for q in queries])
    assert np.array_equal(ans, expect)
    # a second batch that overlaps the first (merge into a non-empty index), then removal of the first half
    more = recs[40:] + records(5, seed=k + 1)
    mbuf, moff = gpu.concat_records(more)
    g.insert_seqs(mbuf, moff)
    for r in more[10:]:
        o.insert_seq(r)
    hbuf, hoff = gpu.concat_records(recs[:25])
    g.remove_seqs(hbuf, hoff)
    for r in recs[:25]:
        o.remove_seq(r)
    check_set_equal(g, o)
