"""Experiment: how much faster is the word-level probe (MODE 3) when the words arrive grouped by prefix slab?
(what a router that bins by (owner, slab) instead of by owner alone would deliver).  N=1, K=25, 500 M-k-mer index."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, cbl_b200
dev = torch.device("cuda", 0)
K, T, P = 25, 64, 24
rec = 1_000_000
index, i_off, query, q_off = bench.make_workload(torch, dev, 500_000_000, 1_000_000_000, rec, 0)
c = cbl_b200.CBL(K, T, P, False, 0)
c.insert_seqs_dev(index.data_ptr(), i_off)
n = c.count_kmers(q_off)
words = torch.empty(n, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
c.seq_words_dev(query.data_ptr(), q_off, words.data_ptr())
flags = torch.empty(n, dtype=torch.uint8, device=dev)
st = torch.cuda.ExternalStream(c.stream_ptr(), device=dev)
def timed(w, label):
    torch.cuda.synchronize()
    c.words_op_dev(0, w.data_ptr(), n, flags.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3):
        c.words_op_dev(0, w.data_ptr(), n, flags.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    print(f"{label:40s} {e0.elapsed_time(e1)/3:8.3f} ms per 1 G words, hits {int(flags.sum())}", flush=True)
timed(words, "arrival order (read order)")
prefix = words >> 32
sample = prefix[torch.randint(0, n, (4_000_000,), device=dev)].sort().values
for S in (4, 8, 16, 32, 64, 256):
    sp = sample[(torch.arange(1, S, device=dev) * sample.numel() // S)]
    slab = torch.bucketize(prefix, sp, right=True)
    order = torch.argsort(slab, stable=True)
    ws = words[order].contiguous()
    del order, slab
    timed(ws, f"grouped into {S} equal-mass prefix slabs")
    del ws
ws = words.sort().values
timed(ws, "fully sorted")
