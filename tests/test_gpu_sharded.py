"""The sharded GPU path end to end on ONE device: two processes share cuda:0, the control plane (count
matrix, IPC handles, barriers) runs over gloo and the data path is the fused route + peer-memory exchange
(CUDA IPC mappings of the other process's buffers; the query runs as route kernel -> count exchange -> one probe launch,
or, "fused" modes, as the one-kernel fused query of csrc/shard_query.cuh), checked bit-exactly
against the oracle.  The same code
runs one process per GPU over NVLink on a multi-GPU box (bench.py --gpus N)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import cbl_testutil as util
    from oracle.pyoracle import OracleCBL
    from cbl_b200.sharded import ShardedCBL

    K, TB, PB, CANON, MODE = {k}, {tb}, {pb}, {canon}, {mode!r}
    os.environ["CBL_EXCHANGE"] = MODE.split("-")[0]
    if "tight" in MODE:   # regions too small at first: the overflow / retry path of the fused route
        os.environ["CBL_ROUTE_SLACK"] = "0.7"
    if "pipe" in MODE:    # pipelined query: 3 sub-batches, route of b + 1 overlapping the probe of b, two buffer sets
        os.environ["CBL_PIPE"] = "3"
    # "fused" (the default of ShardedCBL): one kernel per rank whose warps alternate between routing and probing
    # (cbl_seq_contains_fused_dev); otherwise: route kernel, count exchange, ONE probe launch
    os.environ["CBL_FUSED"] = "1" if "fused" in MODE else "0"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(0)
    sh = ShardedCBL(K, TB, PB, CANON, device=0, sample_bases=200000)
    assert (sh.peer is not None) == MODE.startswith("peer")
    dev = torch.device("cuda", 0)
    reads = [util.random_dna(300000 + 10000 * r, seed=70 + r) for r in range(world)]
    mine = torch.from_numpy(reads[rank]).to(dev)
    offs = np.array([0, 100000, len(reads[rank])], dtype=np.uint64)
    sh.insert_seqs_dev(mine.data_ptr(), offs)
    ref = OracleCBL(K, TB, PB, CANON)
    for r in range(world):
        ref.insert_seq(reads[r][:100000]); ref.insert_seq(reads[r][100000:])
    assert sh.count() == ref.count(), (sh.count(), ref.count())
    # shards are disjoint ascending ranges: concatenation by rank == ascending reference set
    local = sh.engine.cbl.words()
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    flat = [w for part in gathered for w in part]
    assert flat == util.to_int_list(*ref.iter_words()), "concatenation of shards != ascending reference set"
    sizes = [len(p) for p in gathered]
    assert min(sizes) > 0.5 * max(sizes), f"unbalanced shards {{sizes}}"
    for rep in range(2):   # second round reuses the mapped buffers
        q = np.concatenate([reads[(rank + 1) % world][30000:90000 + 1000 * rep], util.random_dna(50000, seed=900 + rank + 10 * rep)])
        qd = torch.from_numpy(q).to(dev)
        got = sh.contains_seqs_dev(qd.data_ptr(), np.array([0, len(q)], dtype=np.uint64)).cpu().numpy()
        exp = ref.contains_seq(q)
        assert np.array_equal(got, exp), "sharded contains_seq != reference answers"
        assert 0 < got.sum() < len(got)
    # a larger batch forces the peer buffers to grow (collective re-map)
    big = util.random_dna(900000, seed=5 + rank)
    bd = torch.from_numpy(big).to(dev)
    got = sh.contains_seqs_dev(bd.data_ptr(), np.array([0, 200000, 500000, 650000, len(big)], dtype=np.uint64)).cpu().numpy()
    exp = np.concatenate([ref.contains_seq(big[a:b]) for a, b in ((0, 200000), (200000, 500000), (500000, 650000), (650000, len(big)))])
    assert np.array_equal(got, exp)
    # set algebra / clone / serde of two sets sharded alike (shard-local kernels, no communication)
    others = [np.concatenate([reads[r][50000:200000], util.random_dna(80000, seed=400 + r)]) for r in range(world)]
    sh2 = sh._derive(type(sh.engine)(K, TB, PB, CANON, 0))
    od = torch.from_numpy(others[rank]).to(dev)
    sh2.insert_seqs_dev(od.data_ptr(), np.array([0, len(others[rank])], dtype=np.uint64))
    ref2 = OracleCBL(K, TB, PB, CANON)
    for r in range(world):
        ref2.insert_seq(others[r])
    W = lambda o: util.to_int_list(*o.iter_words())
    assert (sh | sh2).words() == W(ref | ref2) and (sh & sh2).words() == W(ref & ref2)
    assert (sh - sh2).words() == W(ref - ref2) and (sh ^ sh2).words() == W(ref ^ ref2)
    c = sh.clone(); c ^= sh2
    assert c.words() == W(ref ^ ref2) and c.count() == (ref ^ ref2).count() and sh.words() == W(ref)
    path = os.path.join({tmp!r}, "sharded_gpu.cbl")
    sh.save_to_file(path)
    assert W(ref.deserialize(open(path, "rb").read())) == W(ref)
    assert sh.load_from_file(path).words() == W(ref)
    sh2.close(); c.close()
    r0 = torch.from_numpy(reads[0]).to(dev)
    if rank == 0:
        sh.remove_seqs_dev(r0.data_ptr(), np.array([0, 100000, len(reads[0])], dtype=np.uint64))
    else:
        sh.remove_seqs_dev(r0.data_ptr(), np.array([0, K], dtype=np.uint64))
    ref.remove_seq(reads[0][:100000]); ref.remove_seq(reads[0][100000:])
    assert sh.count() == ref.count()
    # a rank without reads still takes part in the query (it answers the words the others send it); the query after a
    # mutation finds the exchange buffers used by another path and cleans them first
    q = np.concatenate([reads[1][150000:260000], util.random_dna(40000, seed=77)])
    qd = torch.from_numpy(q).to(dev)
    if "fused" not in MODE:
        pass
    elif rank == 0:
        got = sh.contains_seqs_dev(qd.data_ptr(), np.array([0, 70000, len(q)], dtype=np.uint64)).cpu().numpy()
        exp = np.concatenate([ref.contains_seq(q[:70000]), ref.contains_seq(q[70000:])])
        assert np.array_equal(got, exp) and 0 < got.sum() < len(got)
    else:
        assert sh.contains_seqs_dev(qd.data_ptr(), np.array([0], dtype=np.uint64)).numel() == 0
    if "fused" in MODE:
        # reads with a non-nucleotide byte on ONE rank: the fused kernel rejects them, and every rank must see the call fail
        # (nobody is left waiting in a collective); the set keeps working afterwards
        q2 = np.concatenate([reads[1][150000:200000], util.random_dna(30000, seed=78)])
        bad = q2.copy()
        bad[12345] = ord("N")
        t_bad = torch.from_numpy(bad if rank == world - 1 else q2).to(dev)
        try:
            sh.contains_seqs_dev(t_bad.data_ptr(), np.array([0, len(q2)], dtype=np.uint64))
            raised = False
        except Exception:
            raised = True
        assert raised, "a rank with non-ACGT reads must fail the call on every rank"
        t_ok = torch.from_numpy(q2).to(dev)
        got = sh.contains_seqs_dev(t_ok.data_ptr(), np.array([0, len(q2)], dtype=np.uint64)).cpu().numpy()
        assert np.array_equal(got, ref.contains_seq(q2))
    sh.close()
    dist.destroy_process_group()
    print("rank", rank, "ok")
    """
)


@pytest.mark.gpu
@pytest.mark.parametrize("k,tb,pb,canon,mode", [(25, 64, 24, False, "peer"), (31, 128, 24, True, "peer"), (59, 128, 28, False, "peer"),
                                                 (25, 64, 24, True, "peer-tight"), (25, 64, 24, False, "peer-pipe"),
                                                 (31, 128, 24, True, "peer-pipe-tight"), (25, 64, 24, True, "peer-fused"),
                                                 (59, 128, 28, True, "peer-fused-tight"), (31, 128, 24, False, "peer-fused")])
def test_sharded_two_ranks_one_gpu(tmp_path, k, tb, pb, canon, mode):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, k=k, tb=tb, pb=pb, canon=canon, mode=mode, tmp=str(tmp_path)))
    port = 29900 + (os.getpid() + k) % 90
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
