"""1-GPU experiment: time the word-level probe (MODE 3 of the fused kernel, what an owner rank runs in the sharded
path) against the fused sequence probe (MODE 1) and the words kernel (MODE 0) on the bench workload."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, cbl_b200
dev = torch.device("cuda", 0)
K, rec = 25, 1_000_000
index, i_off, query, q_off = bench.make_workload(torch, dev, int(500e6), int(1000e6), rec, seed_base=0)
n_q = (len(q_off) - 1) * (rec - K + 1)
cbl = cbl_b200.CBL(K, 64, 24, canonical=False, device=0)
cbl.insert_seqs_dev(index.data_ptr(), i_off)
words = torch.empty(n_q, dtype=torch.int64, device=dev)
flags = torch.empty(n_q, dtype=torch.uint8, device=dev)
flags2 = torch.empty(n_q, dtype=torch.uint8, device=dev)
cbl_b200.profile_enable(True)
for it in range(3):
    cbl.seq_words_dev(query.data_ptr(), q_off, words.data_ptr())
    cbl.words_op_dev(0, words.data_ptr(), n_q, flags.data_ptr())
    cbl.contains_seqs_dev(query.data_ptr(), q_off, flags2.data_ptr())
    if it == 0:
        cbl_b200.profile_report()
torch.cuda.synchronize()
rep = cbl_b200.profile_report()
print(json.dumps({k: round(v["ms"] / v["n"], 3) for k, v in rep.items()}))
print("answers equal:", bool(torch.equal(flags, flags2)), "hits", int(flags.sum()))
# sorted-word probe (locality upper bound)
sw, _ = torch.sort(words)
cbl_b200.profile_report()
cbl.words_op_dev(0, sw.data_ptr(), n_q, flags.data_ptr())
torch.cuda.synchronize()
print("sorted words:", json.dumps({k: round(v["ms"] / v["n"], 3) for k, v in cbl_b200.profile_report().items()}))
