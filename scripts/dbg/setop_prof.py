import sys, os, numpy as np, torch
sys.path.insert(0, '.')
import bench as B, cbl_b200
dev = torch.device("cuda", 0)
k, t, p = (31, 128, 24) if len(sys.argv) < 2 else (25, 64, 24)
rec, n_rec = 1_000_000, 100
ra = B.device_dna(torch, n_rec * rec, 5, dev); rb = B.device_dna(torch, n_rec * rec, 6, dev)
rb[: 50 * rec].copy_(ra[: 50 * rec])
off = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(rec)
a, b = cbl_b200.CBL(k, t, p), cbl_b200.CBL(k, t, p)
a.insert_seqs_dev(ra.data_ptr(), off); b.insert_seqs_dev(rb.data_ptr(), off)
print("buckets", a.num_buckets(), b.num_buckets(), "count", a.count())
for op in range(4):
    c = a.clone(); c._assign(op, b)
cbl_b200.profile_enable(True); cbl_b200.profile_report()
for op in range(4):
    c = a.clone(); c._assign(op, b)
torch.cuda.synchronize()
pr = cbl_b200.profile_report()
for kname, v in sorted(pr.items(), key=lambda kv: -kv[1]["ms"]): print(f"{kname[:70]:70s} x{v['n']} {v['ms']:.3f} ms")
