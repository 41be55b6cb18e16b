"""1-GPU experiment: probe cost per prefix region.  The query words are routed by equal-mass splitters into G regions
(produce-only run of the fused kernel), then every region is probed on its own (word-level probe, MODE 3): ms per 1 G
words as a function of where in the prefix space the words fall.   python scripts/exp_region_probe.py [G] [index k-mers]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench, cbl_b200
from cbl_b200.sharded import equal_mass_splitters, word_prefixes

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N_INDEX = int(float(sys.argv[2])) if len(sys.argv) > 2 else int(2e9)
N_QUERY = int(float(sys.argv[3])) if len(sys.argv) > 3 else int(1000e6)
dev = torch.device("cuda", 0)
K, rec = 25, 1_000_000
index, i_off, query, q_off = bench.make_workload(torch, dev, N_INDEX, N_QUERY, rec, seed_base=0)
n_q = (len(q_off) - 1) * (rec - K + 1)
cbl = cbl_b200.CBL(K, 64, 24, canonical=False, device=0)
step = 500
for r0 in range(0, len(i_off) - 1, step):
    sub = i_off[r0 : r0 + step + 1]
    cbl.insert_seqs_dev(index.data_ptr() + int(sub[0]), sub - sub[0])
print("index k-mers", cbl.count(), "buckets", cbl.num_buckets(), flush=True)
del index
suffix_bits = 2 * K + 6 - 24
sample = torch.empty(4 * (rec - K + 1), dtype=torch.int64, device=dev)
cbl.seq_words_dev(query.data_ptr(), q_off[:5], sample.data_ptr())
sp = equal_mass_splitters(word_prefixes(sample, suffix_bits, 24), G).cpu().numpy().astype(np.uint32)
cap = (int(n_q / G * 1.3) + 4096 + 1023) // 1024 * 1024
recv = torch.full((G * cap,), -1, dtype=torch.int64, device=dev)
back = torch.empty(G * cap, dtype=torch.uint8, device=dev)
ctrl = torch.zeros(G + 8, dtype=torch.int64, device=dev)
pos = torch.empty(n_q, dtype=torch.int32, device=dev)
regions = [recv.data_ptr() + d * cap * 8 for d in range(G)]
answers = [back.data_ptr() + d * cap for d in range(G)]
finals = [ctrl.data_ptr() + d * 8 for d in range(G)]
os.environ["CBL_SQ_FLAGS"] = "1"   # produce only: the words stay in the regions
counts = cbl.seq_contains_fused_dev(query.data_ptr(), q_off, sp, regions, finals, cap, pos.data_ptr(), regions, answers, finals, 1)
del os.environ["CBL_SQ_FLAGS"]
flags = torch.empty(cap, dtype=torch.uint8, device=dev)
cbl_b200.profile_enable(True)
res = []
for d in range(G):
    n = int(counts[d])
    for it in range(2):
        cbl.words_op_dev(0, regions[d], n, flags.data_ptr())
        if it == 0:
            cbl_b200.profile_report()
    torch.cuda.synchronize()
    rep = cbl_b200.profile_report()
    ms = sum(v["ms"] / v["n"] for k, v in rep.items() if "seq_words_kernel" in k)
    res.append((d, n, ms, ms / n * 1e9, int(flags[:n].sum())))
for d, n, ms, per_g, hits in res:
    print(f"region {d}: {n} words, probe {ms:.3f} ms = {per_g:.2f} ms per 1 G words, hits {hits}")
print("splitters", sp.tolist())
