"""1-GPU data point for BASELINE configs[2] (K=59, T=u128, PREFIX_BITS=28 build) and configs[3]-like K=31: insert_seq of
synthetic reads resident in HBM, per-kernel device times, hybrid sort vs plain LSD passes (CBL_SORT=lsd)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, cbl_b200
dev = torch.device("cuda", 0)
rec = 1_000_000
mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 250
for (K, T, P) in [(59, 128, 28), (31, 128, 24)]:
    index, i_off, _, _ = bench.make_workload(torch, dev, int(mbp * 1e6), int(2e6), rec, seed_base=7)
    n = (len(i_off) - 1) * (rec - K + 1)
    for mode in ("hybrid", "lsd"):
        if mode == "lsd":
            os.environ["CBL_SORT"] = "lsd"
        else:
            os.environ.pop("CBL_SORT", None)
        w = cbl_b200.CBL(K, T, P, canonical=False, device=0)
        w.insert_seqs_dev(index.data_ptr(), i_off)     # warm the arena
        del w
        c = cbl_b200.CBL(K, T, P, canonical=False, device=0)
        cbl_b200.profile_enable(True); cbl_b200.profile_report()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        c.insert_seqs_dev(index.data_ptr(), i_off)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        rep = cbl_b200.profile_report(); cbl_b200.profile_enable(False)
        kms = sum(v["ms"] for v in rep.values())
        print(json.dumps({"K": K, "T": T, "P": P, "sort": mode, "kmers": n, "stored": c.count(), "wall_ms": round(dt * 1e3, 1), "kernel_ms": round(kms, 1),
                          "insert_Gkmers_per_s_kernels": round(n / kms / 1e6, 2), "fallbacks": cbl_b200.sort_fallback_count(),
                          "kernels": {k.split("<")[0]: (v["n"], round(v["ms"], 1)) for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:6]}}))
        del c
    del index
    torch.cuda.empty_cache()
