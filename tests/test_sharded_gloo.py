"""world_size-2 (and 3) gloo tests of the multi-GPU host logic in cbl_b200/sharded.py, on CPU tensors:
prefix extraction, equal-mass splitters, routing, the all-to-all-v exchange, the answer return path,
global count, the shard-local set algebra (| & - ^ and assign forms, clone), iteration order and the one-file serde
written rank after rank, with a stand-in engine (a Python set of words per rank fed by the CPU oracle's
seq_words).  The GPU engine itself is covered by the -m gpu tests."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

import cbl_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_word_prefixes_and_route():
    sys.path.insert(0, ROOT)
    from cbl_b200._lib import lib  # noqa: F401  (the package needs the native library to import)
    from cbl_b200.sharded import equal_mass_splitters, route, word_prefixes

    rng = np.random.default_rng(1)
    # 64-bit words (K=29: word uses all 64 bits incl. the sign bit)
    for k, pb in [(25, 24), (29, 24), (7, 14)]:
        sb = 2 * k + util.pos_bits(k) - pb
        words = [int(x) for x in rng.integers(0, 1 << 62, size=1000)] + [(1 << (2 * k + util.pos_bits(k))) - 1]
        t = torch.tensor(np.array(words, dtype=np.uint64).view(np.int64))
        got = word_prefixes(t, sb, pb).tolist()
        assert got == [(w >> sb) & ((1 << pb) - 1) for w in words]
    # 128-bit words as (lo, hi)
    for k, pb in [(31, 24), (59, 28), (59, 24), (33, 8)]:
        sb = 2 * k + util.pos_bits(k) - pb
        wb = 2 * k + util.pos_bits(k)
        words = [int.from_bytes(rng.bytes(16), "little") & ((1 << wb) - 1) for _ in range(1000)]
        arr = np.array([[w & 0xFFFFFFFFFFFFFFFF, w >> 64] for w in words], dtype=np.uint64).view(np.int64)
        got = word_prefixes(torch.tensor(arr), sb, pb).tolist()
        assert got == [w >> sb for w in words]
    pre = torch.tensor([5, 1, 9, 3, 3, 7, 0, 9])
    sp = torch.tensor([3, 8])
    dest, order, counts = route(pre, sp)
    assert dest.tolist() == [1, 0, 2, 1, 1, 1, 0, 2] and counts.tolist() == [2, 4, 2]
    assert pre[order].tolist() == [1, 0, 5, 3, 3, 7, 9, 9]  # stable grouping
    flat = ((0.0, 1.0), (1.0, 1.0))
    s = equal_mass_splitters(torch.arange(1000), 4, knots=flat)
    assert s.tolist() == [250, 500, 750]
    assert equal_mass_splitters(torch.zeros(100, dtype=torch.int64), 3).tolist() == [0, 1]
    # default: the measured probe-cost curve (PROBE_COST_KNOTS), ranges equalise its integral: the last range is the shortest
    from cbl_b200.sharded import PROBE_COST_KNOTS
    d = equal_mass_splitters(torch.arange(80000), 8)
    edges = [0] + d.tolist() + [80000]
    qs, ws = [k[0] for k in PROBE_COST_KNOTS], [k[1] for k in PROBE_COST_KNOTS]
    cost = [sum(float(np.interp(x / 80000.0, qs, ws)) for x in range(a, b)) for a, b in zip(edges, edges[1:])]
    assert max(cost) - min(cost) <= 0.002 * max(cost) and edges[1] > 10000 > edges[-1] - edges[-2]
    # cost-weighted splitters: cost rises from 1 to 1.5 over the first 62.5 % of the sorted sample, ranges equalise cost
    w = equal_mass_splitters(torch.arange(8000), 8, tail_cost=1.5)
    edges = [0] + w.tolist() + [8000]
    cost = [sum(1.0 + 0.5 * min(x / 5000.0, 1.0) for x in range(a, b)) for a, b in zip(edges, edges[1:])]
    assert max(cost) - min(cost) <= 3.0 and edges[1] > 1150 and edges[-1] - edges[-2] < 900


def test_cpp_splitters_match_python(tmp_path):
    """cbl_create_sharded (one process, C++: csrc/splitters.hpp) and ShardedCBL (one process per GPU, Python) must cut the same
    sample the same way: a set built by one host is then sharded like a set built by the other."""
    from cbl_b200.sharded import equal_mass_splitters

    exe = str(tmp_path / "splitters_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host", "splitters_check.cpp")], check=True)
    rng = np.random.default_rng(11)
    # skewed like necklace prefixes: most of the mass at small values, a long sparse tail
    sample = np.minimum((rng.exponential(60000.0, size=200_000)).astype(np.uint32) + rng.integers(0, 50, 200_000).astype(np.uint32), (1 << 24) - 1).astype(np.uint32)
    path = tmp_path / "sample.bin"
    sample.tofile(path)
    srt = np.sort(sample)
    for world in (2, 3, 4, 8, 16):
        out = subprocess.run([exe, str(path), str(world)], capture_output=True, text=True, check=True).stdout.split()
        cpp = np.array([int(x) for x in out], dtype=np.int64)
        py = equal_mass_splitters(torch.from_numpy(sample.astype(np.int64)), world).numpy()
        assert len(cpp) == world - 1 == len(py)
        # same cut up to one position of the sorted sample (the two hosts invert the same cost curve in floating point)
        for a, b in zip(cpp, py):
            ia, ib = np.searchsorted(srt, a, side="left"), np.searchsorted(srt, b, side="left")
            assert abs(int(ia) - int(ib)) <= 2, (world, cpp.tolist(), py.tolist())


WORKER = textwrap.dedent(
    """
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import cbl_testutil as util
    from oracle.pyoracle import OracleCBL
    from cbl_b200.sharded import ShardedCBL

    K, TB, PB, CANON = {k}, {tb}, {pb}, {canon}

    class SetEngine:
        '''stand-in shard: words from the CPU oracle, membership in a Python set'''
        device = torch.device("cpu")
        def __init__(self):
            self.o = OracleCBL(K, TB, PB, CANON)
            self.words = set()
            self.host = {{}}
        def _to_tensor(self, lo, hi):
            if 2 * K + util.pos_bits(K) <= 64:
                return torch.from_numpy(lo.view(np.int64).copy())
            return torch.from_numpy(np.stack([lo, hi], axis=1).view(np.int64).copy())
        def _ints(self, t):
            a = t.numpy().view(np.uint64)
            return [int(x) for x in a] if a.ndim == 1 else [int(l) | (int(h) << 64) for l, h in a]
        def seq_words(self, key, offsets):
            buf = self.host[key]
            los, his = [], []
            for i in range(len(offsets) - 1):
                lo, hi = self.o.seq_words(buf[int(offsets[i]):int(offsets[i + 1])])
                los.append(lo); his.append(hi)
            return self._to_tensor(np.concatenate(los), np.concatenate(his))
        def words_op(self, op, words, want_flags):
            ws = self._ints(words)
            flags = torch.tensor([w in self.words for w in ws], dtype=torch.uint8) if want_flags else None
            if op == 1: self.words.update(ws)
            if op == 2: self.words.difference_update(ws)
            return flags
        def count(self):
            return len(self.words)
        # shard-local set algebra / clone / export / serde (what ShardedCBL composes rank by rank)
        def _with(self, words):
            e = SetEngine(); e.words = set(words); return e
        def setop(self, op, other):
            a, b = self.words, other.words
            return self._with([a | b, a & b, a - b, a ^ b][op])
        def setop_assign(self, op, other):
            self.words = self.setop(op, other).words
        def clone(self):
            return self._with(self.words)
        def words_list(self):
            return sorted(self.words)
        def kmers_list(self):
            return [self.o.recover_kmer(w) for w in sorted(self.words)]
        def _oracle_of(self):
            o = OracleCBL(K, TB, PB, CANON)
            for w in self.words:
                o.insert(self.o.recover_kmer(w))
            return o
        def serialize(self):
            return self._oracle_of().serialize()
        def deserialize_range(self, data, lo, hi):
            sb = 2 * K + util.pos_bits(K) - PB
            return self._with([w for w in util.to_int_list(*self.o.deserialize(data).iter_words()) if lo <= (w >> sb) < hi])
        def sample_words(self, n, seed):
            self.host["sample"] = util.random_dna(n, seed)
            return self.seq_words("sample", np.array([0, n], dtype=np.uint64))

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = SetEngine()
    sh = ShardedCBL(K, TB, PB, CANON, engine=eng, sample_bases=60000)
    assert sh.splitters.numel() == world - 1
    # every rank has its own reads; the union is the global set
    reads = [util.random_dna(20000 + 1000 * r, seed=50 + r) for r in range(world)]
    mine = reads[rank]
    eng.host["mine"] = mine
    offs = np.array([0, 7000, len(mine)], dtype=np.uint64)
    sh.insert_seqs_dev("mine", offs)
    ref = OracleCBL(K, TB, PB, CANON)
    for r in range(world):
        ref.insert_seq(reads[r][:7000]); ref.insert_seq(reads[r][7000:])
    assert sh.count() == ref.count(), (sh.count(), ref.count())
    # shards are disjoint, ordered by rank, and together equal the reference set
    local = sorted(eng.words)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    flat = [w for part in gathered for w in part]
    assert flat == sorted(flat) and flat == util.to_int_list(*ref.iter_words()), "concatenation of shards != ascending reference set"
    sizes = [len(p) for p in gathered]
    assert min(sizes) > 0.5 * max(sizes), f"unbalanced shards {{sizes}}"   # equal-mass splitters (SURVEY F4)
    # contains: rank-local queries (mix of hits from OTHER ranks' reads and misses), answers in local order
    q = np.concatenate([reads[(rank + 1) % world][3000:9000], util.random_dna(5000, seed=900 + rank)])
    eng.host["q"] = q
    got = sh.contains_seqs_dev("q", np.array([0, len(q)], dtype=np.uint64)).numpy()
    exp = ref.contains_seq(q)
    assert np.array_equal(got, exp), "sharded contains_seq != reference answers"
    assert 0 < got.sum() < len(got)
    # ---- set algebra between two sets sharded alike: shard-local, result = concatenation by rank (src/cbl.rs:411-569)
    other_reads = [np.concatenate([reads[r][2000:9000], util.random_dna(6000, seed=700 + r)]) for r in range(world)]
    sh2 = sh._derive(SetEngine())
    sh2.engine.host["o"] = other_reads[rank]
    sh2.insert_seqs_dev("o", np.array([0, len(other_reads[rank])], dtype=np.uint64))
    ref2 = OracleCBL(K, TB, PB, CANON)
    for r in range(world):
        ref2.insert_seq(other_reads[r])
    W = lambda o: util.to_int_list(*o.iter_words())
    assert (sh | sh2).words() == W(ref | ref2) and (sh & sh2).words() == W(ref & ref2)
    assert (sh - sh2).words() == W(ref - ref2) and (sh ^ sh2).words() == W(ref ^ ref2)
    c = sh.clone(); c |= sh2; assert c.words() == W(ref | ref2)
    c = sh.clone(); c &= sh2; assert c.words() == W(ref & ref2) and c.count() == (ref & ref2).count()
    c = sh.clone(); c -= sh2; assert c.words() == W(ref - ref2)
    c = sh.clone(); c ^= sh2; assert c.words() == W(ref ^ ref2)
    assert sh.words() == W(ref), "operands untouched"
    assert list(sh.iter()) == [ref.recover_kmer(w) for w in W(ref)], "iter: ascending word order, rank after rank"
    try:
        sh | ShardedCBL(K, TB, PB, CANON, engine=SetEngine(), splitters=[s + 1 for s in sh.splitters_u32.tolist()])
        raise SystemExit("differently sharded operand accepted")
    except ValueError:
        pass
    # ---- serde: ONE file in the reference's layout, written rank after rank; the oracle reads it; a sharded set reloads it
    path = os.path.join({tmp!r}, "sharded.cbl")
    sh.save_to_file(path)
    assert W(ref.deserialize(open(path, "rb").read())) == W(ref), "oracle cannot read the sharded file"
    back = sh.load_from_file(path)
    assert back.words() == W(ref) and back.local_words() == sh.local_words()
    # remove what rank 0 inserted, everywhere
    eng.host["r0"] = reads[0]
    if rank == 0:
        sh.remove_seqs_dev("r0", np.array([0, 7000, len(reads[0])], dtype=np.uint64))
    else:
        eng.host["empty"] = reads[0][:K]
        sh.remove_seqs_dev("empty", np.array([0, K], dtype=np.uint64))
    ref.remove_seq(reads[0][:7000]); ref.remove_seq(reads[0][7000:]); 
    assert sh.count() == ref.count()
    dist.destroy_process_group()
    print("rank", rank, "ok")
    """
)


@pytest.mark.parametrize("world,k,tb,pb,canon", [(2, 25, 64, 24, False), (2, 59, 128, 28, True), (3, 31, 128, 24, False)])
def test_sharded_routing_gloo(tmp_path, world, k, tb, pb, canon):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, k=k, tb=tb, pb=pb, canon=canon, tmp=str(tmp_path)))
    port = 29600 + (os.getpid() + world * 7 + k) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world
