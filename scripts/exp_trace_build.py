import sys, os, time
sys.path.insert(0, "/root/repo")
import torch, bench, cbl_b200
dev = torch.device("cuda", 0)
index, i_off, _, _ = bench.make_workload(torch, dev, int(500e6), int(2e6), 1_000_000, seed_base=0)
w = cbl_b200.CBL(25, 64, 24, canonical=False, device=0); w.insert_seqs_dev(index.data_ptr(), i_off); del w
c = cbl_b200.CBL(25, 64, 24, canonical=False, device=0)
os.environ["CBL_TRACE"] = "1"
torch.cuda.synchronize(); t0 = time.perf_counter()
c.insert_seqs_dev(index.data_ptr(), i_off)
torch.cuda.synchronize(); print("wall ms", (time.perf_counter() - t0) * 1e3)
