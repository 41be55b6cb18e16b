"""1-GPU test bed of the fused sharded query kernel (csrc/shard_query.cuh) with the per-GPU load of a G-rank run:
the words of the reads are routed by real equal-mass splitters into G regions of a LOCAL receive buffer (region d plays
"my region at owner d" and "the region source d wrote at me" at once), so one kernel produces 1 G words and probes 1 G
words, exactly what every rank of a G-GPU weak-scaling step does, minus NVLink.  Answers are checked against the
single-GPU fused probe.   python scripts/exp_fused_g1.py [G] [index k-mers] [query k-mers]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench, cbl_b200
from cbl_b200.sharded import equal_mass_splitters, word_prefixes

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N_INDEX = int(float(sys.argv[2])) if len(sys.argv) > 2 else int(500e6)
N_QUERY = int(float(sys.argv[3])) if len(sys.argv) > 3 else int(1000e6)
REPS = int(os.environ.get("REPS", "3"))
dev = torch.device("cuda", 0)
K, rec = 25, 1_000_000
index, i_off, query, q_off = bench.make_workload(torch, dev, N_INDEX, N_QUERY, rec, seed_base=0)
n_q = (len(q_off) - 1) * (rec - K + 1)
cbl = cbl_b200.CBL(K, 64, 24, canonical=False, device=0)
cbl.insert_seqs_dev(index.data_ptr(), i_off)
suffix_bits = 2 * K + 6 - 24
# splitters from a sample of the query words
sample = torch.empty(4 * (rec - K + 1), dtype=torch.int64, device=dev)
cbl.seq_words_dev(query.data_ptr(), q_off[:5], sample.data_ptr())
sp = equal_mass_splitters(word_prefixes(sample, suffix_bits, 24), G).cpu().numpy().astype(np.uint32) if G > 1 else np.zeros(0, dtype=np.uint32)
BLOCK = 1024
cap = (int(n_q / G * 1.3) + 4096 + BLOCK - 1) // BLOCK * BLOCK
recv = torch.full((G * cap,), -1, dtype=torch.int64, device=dev)   # sentinel-filled
back = torch.empty(G * cap, dtype=torch.uint8, device=dev)
ctrl = torch.zeros(G + 8, dtype=torch.int64, device=dev)
pos = torch.empty(n_q, dtype=torch.int32, device=dev)
out = torch.empty(n_q, dtype=torch.uint8, device=dev)
ref = torch.empty(n_q, dtype=torch.uint8, device=dev)
regions = [recv.data_ptr() + d * cap * 8 for d in range(G)]
answers = [back.data_ptr() + d * cap for d in range(G)]
finals = [ctrl.data_ptr() + d * 8 for d in range(G)]
cbl_b200.profile_enable(True)
for it in range(REPS):
    torch.cuda.synchronize()
    counts = cbl.seq_contains_fused_dev(query.data_ptr(), q_off, sp, regions, finals, cap, pos.data_ptr(), regions, answers, finals, it + 1)
    assert int(counts.max()) <= cap, (counts, cap)
    cbl.gather_u8_dev(back.data_ptr(), pos.data_ptr(), n_q, out.data_ptr())
    cbl.contains_seqs_dev(query.data_ptr(), q_off, ref.data_ptr())
    if it == 0:
        cbl_b200.profile_report()
torch.cuda.synchronize()
rep = cbl_b200.profile_report()
print(json.dumps({k.split("<")[0] + (k[k.find(",Suf,") + 5] if ",Suf," in k else ""): round(v["ms"] / v["n"], 3) for k, v in rep.items()}))
print("G", G, "counts", counts.tolist(), "answers equal:", bool(torch.equal(out, ref)), "hits", int(out.sum()),
      "receive buffer left clean:", bool((recv == -1).all()))
