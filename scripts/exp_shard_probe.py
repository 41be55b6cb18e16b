"""1-GPU experiment: the shard one rank of an N-GPU run holds (all k-mers of N x 500 Mbp of reads whose prefix falls into
ONE of N equal-mass regions), probed by that region's share of fresh queries (word-level probe, MODE 3).
python scripts/exp_shard_probe.py [N] [regions, comma separated]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench, cbl_b200
from cbl_b200.sharded import equal_mass_splitters, word_prefixes

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
REGIONS = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,6").split(",")]
dev = torch.device("cuda", 0)
K, rec, NB = 25, 1_000_000, 500   # batches of 500 records = what one rank contributes
n_b = NB * (rec - K + 1)
suffix_bits = 2 * K + 6 - 24
router = cbl_b200.CBL(K, 64, 24, canonical=False, device=0)
offs = np.arange(NB + 1, dtype=np.uint64) * np.uint64(rec)
cap = (int(n_b / G * 1.3) + 4096 + 1023) // 1024 * 1024
recv = torch.full((G * cap,), -1, dtype=torch.int64, device=dev)
back = torch.empty(G * cap, dtype=torch.uint8, device=dev)
ctrl = torch.zeros(G + 8, dtype=torch.int64, device=dev)
pos = torch.empty(n_b, dtype=torch.int32, device=dev)
regions = [recv.data_ptr() + d * cap * 8 for d in range(G)]
answers = [back.data_ptr() + d * cap for d in range(G)]
finals = [ctrl.data_ptr() + d * 8 for d in range(G)]
sp = None
shards = {d: cbl_b200.CBL(K, 64, 24, canonical=False, device=0) for d in REGIONS}
keep = {}

def route(seed):
    global sp
    reads = bench.device_dna(torch, NB * rec, seed, dev)
    if sp is None:
        sample = torch.empty(4 * (rec - K + 1), dtype=torch.int64, device=dev)
        router.seq_words_dev(reads.data_ptr(), offs[:5], sample.data_ptr())
        sp = equal_mass_splitters(word_prefixes(sample, suffix_bits, 24), G).cpu().numpy().astype(np.uint32)
    recv.fill_(-1)
    torch.cuda.synchronize()
    os.environ["CBL_SQ_FLAGS"] = "1"   # produce only: the words stay in the regions
    counts = router.seq_contains_fused_dev(reads.data_ptr(), offs, sp, regions, finals, cap, pos.data_ptr(), regions, answers, finals, 1 + seed % 1000)
    del os.environ["CBL_SQ_FLAGS"]
    return counts

for b in range(G):   # the reads of "rank b"
    counts = route(100 + b)
    for d in REGIONS:
        shards[d].words_op_dev(1, regions[d], int(counts[d]))
        if b == 0:
            keep[d] = recv[d * cap : d * cap + int(counts[d])].clone()   # hits
    torch.cuda.synchronize()
for d in REGIONS:
    print(f"shard of region {d}: {shards[d].count()} k-mers, {shards[d].num_buckets()} buckets", flush=True)
counts = route(999)   # misses
flags = torch.empty(2 * cap, dtype=torch.uint8, device=dev)
cbl_b200.profile_enable(True)
for d in REGIONS:
    # arrival order is kept (consecutive k-mers of a read make related words: the probe lives on that locality); hits and
    # misses alternate in chunks of 1 M words
    miss = recv[d * cap : d * cap + int(counts[d])]
    parts = []
    for c0 in range(0, max(keep[d].numel(), miss.numel()), 1 << 20):
        parts += [keep[d][c0 : c0 + (1 << 20)], miss[c0 : c0 + (1 << 20)]]
    q = torch.cat(parts)
    torch.cuda.synchronize()
    print(f"region {d}: hits part {keep[d].numel()}, misses part {int(counts[d])}", flush=True)
    for it in range(2):
        shards[d].words_op_dev(0, q.data_ptr(), q.numel(), flags.data_ptr())
        if it == 0:
            cbl_b200.profile_report()
    torch.cuda.synchronize()
    rep = cbl_b200.profile_report()
    ms = sum(v["ms"] / v["n"] for k, v in rep.items() if "seq_words_kernel" in k)
    print(f"region {d}: {q.numel()} words, probe {ms:.3f} ms = {ms / q.numel() * 1e9:.2f} ms per 1 G words, hits {int(flags[:q.numel()].sum())}", flush=True)
