"""The sharded set behind the C ABI (cbl_create_sharded: several GPUs driven by ONE process, words exchanged through
peer memory) against the CPU oracle: every host-buffer entry point, both word widths, canonical mode, reads with
non-ACGT bytes, set operations between sharded handles, iteration order, serde interop with the oracle and with an
unsharded handle.  On a single-GPU box the shards share device 0 (same code path, peer access is trivially local)."""
import numpy as np
import pytest

import cbl_testutil as util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import cbl_b200

    return cbl_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle

    return pyoracle


def devices(n):
    import torch

    c = torch.cuda.device_count()
    return [i % c for i in range(n)]


def words(o):
    return util.to_int_list(*o.iter_words())


@pytest.mark.parametrize("k,tb,pb,canonical,n_sh", [(25, 64, 24, False, 2), (25, 64, 24, True, 3), (31, 128, 24, True, 2), (59, 128, 28, False, 4), (7, 32, 14, False, 2)])
def test_sharded_handle_matches_oracle(gpu, orc, k, tb, pb, canonical, n_sh):
    g = gpu.CBL.sharded(k, tb, pb, canonical, devices(n_sh))
    o = orc.OracleCBL(k, tb, pb, canonical)
    sp = g.shard_splitters()
    assert len(sp) == n_sh - 1 and all(sp[i] < sp[i + 1] for i in range(len(sp) - 1))
    assert g.is_empty() and g.count() == 0
    recs = [util.random_dna(n, seed=300 + i) for i, n in enumerate([90_000, k, 40_000, 2048 + k - 1, 150_000, 333])]
    buf, off = gpu.concat_records(recs)
    g.insert_seqs(buf, off)
    for r in recs:
        o.insert_seq(r)
    assert g.count() == o.count()
    assert g.words() == words(o), "shards in device order != ascending reference set"
    gp, gs = g.buckets_sizes()
    op, osz = o.bucket_sizes()
    order = np.argsort(op, kind="stable")
    assert np.array_equal(gp.astype(np.uint64), op[order]) and np.array_equal(gs.astype(np.uint64), osz[order])
    queries = [recs[0][10_000:70_000], util.random_dna(50_000, seed=9), np.concatenate([recs[4][:30_000], util.random_dna(20_000, seed=10)]), recs[1]]
    qbuf, qoff = gpu.concat_records(queries)
    ans = g.contains_seqs(qbuf, qoff)
    assert np.array_equal(ans, np.concatenate([o.contains_seq(q) for q in queries])), "per-k-mer answers (reference order)"
    assert g.contains_all(recs[2]) and g.contains_all(queries[1]) == o.contains_all(queries[1])
    # single k-mers: return values of insert / remove (src/cbl.rs:219-235)
    km = [util.kmer_int(queries[1][i : i + k].tobytes()) for i in (0, 17, 300)]
    assert [g.insert(x) for x in km] == [o.insert(x) for x in km]
    assert [g.insert(x) for x in km] == [o.insert(x) for x in km] == [False] * 3
    assert list(g.contains_kmers(km)) == [True] * 3
    assert [g.remove(x) for x in km] == [o.remove(x) for x in km]
    assert list(g) == [o.recover_kmer(w) for w in words(o)], "iter order"
    # remove
    rbuf, roff = gpu.concat_records([recs[0], recs[3]])
    g.remove_seqs(rbuf, roff)
    o.remove_seq(recs[0])
    o.remove_seq(recs[3])
    assert g.count() == o.count() and g.words() == words(o)
    assert np.array_equal(g.contains_seqs(qbuf, qoff), np.concatenate([o.contains_seq(q) for q in queries]))
    # device-pointer entry points do not apply to a sharded handle
    with pytest.raises(gpu.CBLError, match="not available on a sharded handle"):
        g.insert_seqs_dev(4096, off)   # (never dereferenced: the call is refused first)


def test_sharded_set_ops_serde_and_non_acgt(gpu, orc, tmp_path):
    k, tb, pb = 25, 64, 24
    devs = devices(2)
    a, b = gpu.CBL.sharded(k, tb, pb, False, devs), gpu.CBL.sharded(k, tb, pb, False, devs)
    oa, ob = orc.OracleCBL(k, tb, pb), orc.OracleCBL(k, tb, pb)
    A = util.random_dna(120_000, seed=1)
    B = np.concatenate([A[30_000:80_000], util.random_dna(70_000, seed=2)])
    a.insert_seq(A); oa.insert_seq(A)
    b.insert_seq(B); ob.insert_seq(B)
    for op in ("__or__", "__and__", "__sub__", "__xor__"):
        assert getattr(a, op)(b).words() == words(getattr(oa, op)(ob)), op
    for op in ("__ior__", "__iand__", "__isub__", "__ixor__"):
        c, oc = a.clone(), oa.clone()
        getattr(c, op)(b)
        getattr(oc, op)(ob)
        assert c.words() == words(oc), op
    assert gpu.CBL.merge([a, b, a]).words() == words(oa | ob)
    assert gpu.CBL.intersect([a, b]).words() == words(oa & ob)
    # a differently sharded / unsharded operand is refused
    with pytest.raises(gpu.CBLError):
        a | gpu.CBL(k, tb, pb)
    with pytest.raises(gpu.CBLError, match="sharded differently"):
        a | gpu.CBL.sharded(k, tb, pb, False, devs, splitters=[12345])
    # serde: sharded writer -> oracle reader, oracle writer -> sharded reader, sharded -> unsharded
    data = a.serialize()
    assert words(oa.deserialize(data)) == words(oa)
    assert a.deserialize(ob.serialize()).words() == words(ob)
    assert gpu.CBL(k, tb, pb).deserialize(data).words() == words(oa)
    path = str(tmp_path / "a.cbl")
    a.save_to_file(path)
    back = a.load_from_file(path)
    assert back.words() == words(oa) and list(back.shard_splitters()) == list(a.shard_splitters())
    # non-ACGT bytes: the reference's dropping behaviour through the sharded handle
    dirty = A[:30_000].copy()
    dirty[[5, 777, 2047, 2048, 9000, 29_999]] = ord("N")
    c, oc = gpu.CBL.sharded(k, tb, pb, False, devs), orc.OracleCBL(k, tb, pb)
    c.insert_seq(dirty); oc.insert_seq(dirty)
    assert c.words() == words(oc)
    assert np.array_equal(c.contains_seq(dirty).astype(np.uint8), oc.contains_seq(dirty))
    # emptied set is reusable; is_empty follows the reference
    a -= a
    assert a.is_empty() and a.count() == 0
    a.insert_seq(B)
    assert a.words() == words(ob)
