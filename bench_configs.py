"""bench.py --config 1 | 3 | 4 | 5: the other BASELINE.json configurations (run by hand; outputs kept under profiles/).
Same JSON contract as the headline line (bench.py), same synthetic-data and parity helpers.

  C1  cbl build on 10 Mbp, K=25 / u64 / 24: full input on the GPU (device-resident and host buffers) and on the CPU
  C3  K=59 / u128 / PREFIX_BITS=28 build of a 3 Gbp-class FASTA, sharded over N GPUs (--index-mbp = Mbp per GPU,
      default 3000 / N: the total stays 3 Gbp, strong scaling)
  C4  |= &= -= ^= of two indexes sharing half of their reads, K=31 / u128 / 24, sharded over N GPUs
      (--index-mbp = Mbp per index per GPU, default 250 = 2 G k-mers per index on 8 GPUs)
  C5  mixed stream: >= 300 batches of 1 Mbp cycling insert_seq (new reads) / contains_seq (half known, half new) /
      remove_seq (reads inserted earlier) on a resident index (--index-mbp per GPU, default 500), K=31 / u128 / 24
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

import bench as B


def _setup(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU leg")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    return torch, dist, world, rank, local, device


class Ctx:
    def __init__(self, args):
        import cbl_b200

        self.lib = cbl_b200
        self.torch, self.dist, self.world, self.rank, self.local, self.device = _setup(args)
        self.k, self.t_bits, self.pb = B.CONFIG_PARAMS[args.config]
        self.wide, self.suffix_bits = B.word_geometry(self.k, self.pb)
        self.W = 16 if self.wide else 8
        self.S = 4 if self.suffix_bits <= 32 else (8 if self.suffix_bits <= 64 else 16)
        self.peak, self.peak_src = B.peaks()
        self.threads = args.parity_threads or max(1, (os.cpu_count() or 8) // max(1, self.world))
        self._splitters = None

    def new_index(self):
        if self.world > 1:
            from cbl_b200.sharded import ShardedCBL

            if self._splitters is None:
                s = ShardedCBL(self.k, self.t_bits, self.pb, canonical=False, device=self.local)
                self._splitters = [int(x) for x in s.splitters_u32]
                return s
            return ShardedCBL(self.k, self.t_bits, self.pb, canonical=False, device=self.local, splitters=self._splitters)
        return self.lib.CBL(self.k, self.t_bits, self.pb, canonical=False, device=self.local)

    def close(self, c):
        if self.world > 1:
            c.close()

    def local_of(self, c):
        return c.engine.cbl if self.world > 1 else c

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: int) -> int:
        if self.world == 1:
            return int(x)
        t = self.torch.tensor([int(x)], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t)
        return int(t.item())

    def stream(self, c):
        return self.torch.cuda.ExternalStream(c.stream_ptr(), device=self.device)

    def timed(self, c, fn):
        """fn() bracketed by CUDA events on c's library stream (calls are host-synchronous): device ms, max over ranks"""
        st = self.stream(c)
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(st)
        fn()
        e1.record(st)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def set_parity(self, c, reads_dev, offsets):
        """the shard's stored words == sorted distinct oracle words of the reads (all ranks), word by word"""
        t0 = time.perf_counter()
        expect = B.expected_shard_set(self.torch, self.dist, c, self.world, self.rank, self.device, self.k, self.t_bits, self.pb, False, reads_dev, offsets,
                                      self.threads)
        got = B.exported_shard_words(self.torch, self.local_of(c), self.device, self.wide)
        mism = B.count_set_mismatches(got, expect, self.wide)
        res = {"set_words_checked": self.sum_over_ranks(expect.shape[0]), "set_mismatches": self.sum_over_ranks(mism),
               "count_expected": self.sum_over_ranks(expect.shape[0]), "count_got": self.sum_over_ranks(got.shape[0]), "mismatches": self.sum_over_ranks(mism),
               "seconds": round(time.perf_counter() - t0, 2), "ranks": self.world,
               "how": "oracle words of every read of every rank, routed to the owner shard, sorted + deduplicated, compared word by word with the stored set"}
        return res, expect

    def line(self, args, metric_cfg, value, ms, cfg, roof, cpu, e2e, launches, clocks, parity, extra, scaling="weak"):
        return {"metric": B.metric_name(args.config, args.metric), "value": value, "unit": B.UNIT, "n_gpus": self.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": f"u{self.t_bits}", "data": "synthetic",
                "config": cfg, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity_check": parity, "extra": extra}


def _cpu_build(k, t_bits, pb, mbp, rec):
    from oracle import pyoracle

    L = pyoracle.load()
    n_i = max(1, int(mbp * 1e6) // rec)
    index = np.concatenate([B.host_dna(rec, 1000 + i) for i in range(n_i)])
    i_off = np.arange(n_i + 1, dtype=np.uint64) * np.uint64(rec)
    o = pyoracle.OracleCBL(k, t_bits, pb, lib=L)
    t = o.time_insert_seqs(index, i_off)
    n = n_i * (rec - k + 1)
    kind = "reference" if L.orc_uses_reference_cxx() else "port"
    return {"value": n / t, "unit": B.UNIT, "cores": 1, "kind": kind, "sample": f"build of {n_i} x {rec} bp ({n} k-mers) into an empty set, {t:.1f} s, 1 thread",
            "host_cores_available": os.cpu_count()}


def _build_steps(cx: Ctx, args, reads, off, n_kmers_total, host_e2e=True):
    """steps x (fresh set, insert_seqs_dev of all reads): device ms per step (max over ranks), launches, clocks, e2e"""
    torch = cx.torch
    ms, launches, sampler = [], 0, None
    keep = None
    for s in range(args.warmup + args.steps):
        c = cx.new_index()
        if cx.rank == 0:
            sampler = B.sampler_at(s, args.warmup, sampler, cx.world, cx.local, str(torch.cuda.get_device_properties(cx.local).uuid))
        l0 = cx.lib.launch_count()
        t = cx.timed(c, lambda: c.insert_seqs_dev(reads.data_ptr(), off))
        if s >= args.warmup:
            ms.append(t)
            launches += cx.lib.launch_count() - l0
        if s == args.warmup + args.steps - 1:
            keep = c
        else:
            cx.close(c)
            del c
    clocks = sampler.stop() if sampler else None
    e2e = None
    if host_e2e and not args.no_e2e:
        h = torch.empty(reads.numel(), dtype=torch.uint8, pin_memory=True)
        h.copy_(reads)
        hn = h.numpy()
        ts = []
        for s in range(1 + min(args.steps, 3)):
            c = cx.new_index()
            cx.barrier()
            t0 = time.perf_counter()
            c.insert_seqs(hn, off)
            n_c = c.count()
            cx.barrier()
            if s >= 1:
                ts.append(time.perf_counter() - t0)
            cx.close(c)
            del c
        t_e = cx.max_over_ranks(sum(ts))
        e2e = {"value": n_kmers_total * len(ts) / t_e, "unit": B.UNIT, "h2d_bytes_per_step": int(reads.numel()) * cx.world, "d2h_bytes_per_step": 8 * cx.world,
               "ms_per_step": 1e3 * t_e / len(ts), "count_read_back": n_c}
    return keep, ms, launches, clocks, e2e


def _build_profile(cx: Ctx, reads, off):
    if cx.world > 1:
        return None
    c = cx.new_index()
    cx.lib.profile_enable(True)
    cx.lib.profile_report()
    c.insert_seqs_dev(reads.data_ptr(), off)
    cx.torch.cuda.synchronize()
    prof = cx.lib.profile_report()
    cx.lib.profile_enable(False)
    del c
    return prof


def run_build(args, config):
    """C1 (10 Mbp, K=25) and C3 (3 Gbp-class, K=59, sharded): one step = one build of the whole input into an empty set"""
    cx = Ctx(args)
    torch = cx.torch
    rec = args.record_bp
    if config == 1:
        mbp = args.index_mbp or 10.0
        scaling = "weak"
    else:
        mbp = args.index_mbp or 3000.0 / cx.world
        scaling = "strong" if args.index_mbp is None else "weak"
    n_rec = max(1, int(mbp * 1e6) // rec)
    reads = B.device_dna(torch, n_rec * rec, 1 + 100 * cx.rank + (4 if config == 3 else 0), cx.device)
    off = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(rec)
    n_k = n_rec * (rec - cx.k + 1)
    total = n_k * cx.world
    c, ms, launches, clocks, e2e = _build_steps(cx, args, reads, off, total)
    elapsed = sum(ms) / 1e3
    value = total * len(ms) / elapsed
    stored = c.count()
    prof = _build_profile(cx, reads, off) if not args.no_build_profile else None
    roof_all = B.build_phase_roofline(prof, n_k, stored, cx.peak, W=cx.W, S=cx.S, prefix_bits=cx.pb) if prof else None
    roof = None
    if roof_all:
        dk = max((k_ for k_ in roof_all if not k_.startswith("_")), key=lambda k_: roof_all[k_]["ms"])
        r = roof_all[dk]
        roof = {"bound": "hbm", "achieved": r["achieved_GBps"], "peak": cx.peak, "unit": "GB/s", "frac": r["frac_of_hbm_peak"], "traffic": None, "kernel": dk,
                "ms_per_launch": r["ms"] / r["launches"], "peak_source": cx.peak_src, "algorithmic_bytes_per_launch": r["algorithmic_bytes_per_launch"],
                "aggregate": roof_all.get("_aggregate")}
    parity = None
    if not args.no_parity:
        try:
            parity, _ = cx.set_parity(c, reads, off)
        except Exception as e:
            parity = {"error": repr(e), "mismatches": None}
    cpu = None
    if cx.rank == 0 and cx.world == 1 and not args.no_cpu_baseline:
        cpu = _cpu_build(cx.k, cx.t_bits, cx.pb, mbp if config == 1 else min(args.cpu_index_mbp, 20.0), min(rec, 1_000_000))
    if cx.rank == 0:
        cfg = {"workload": f"configs[{config - 1}]: build (insert_seq into an empty set) of {n_rec} x {rec} bp per GPU, K={cx.k}, T=u{cx.t_bits}, PREFIX_BITS={cx.pb}",
               "records_per_gpu": n_rec, "record_bp": rec, "stored_kmers": stored, "buckets": c.num_buckets(),
               "l2_policy": f"inputs larger than L2: {reads.numel() / 1e6:.0f} MB of reads, {n_k * cx.W / 1e6:.0f} MB of words per GPU" if n_k * cx.W > 130e6 else "L2 flushed between steps by the fresh set's buffers (sort buffers + suffix array exceed L2)",
               "parallelism": "1 GPU" if cx.world == 1 else f"prefix-range sharded x{cx.world}, fused route over NVLink peer memory",
               "suffix_bytes_stored": cx.S, "word_bytes_device": cx.W}
        extra = {"kernel_ms": prof, "build_roofline": roof_all, "ms_steps": ms}
        print(json.dumps(cx.line(args, None, value, 1e3 * elapsed / len(ms), cfg, roof, cpu, e2e, launches, clocks, parity, extra, scaling=scaling)))
    cx.close(c)
    if cx.world > 1:
        cx.dist.destroy_process_group()


def _torch_setop(torch, a, b, op, wide):
    """set algebra on two ascending distinct word tensors (the checker's arithmetic): | & - ^"""
    allw = torch.cat([a, b])
    flag = torch.cat([torch.ones(a.shape[0], dtype=torch.int64, device=a.device), torch.full((b.shape[0],), 2, dtype=torch.int64, device=a.device)])
    if wide:
        lo_u = allw[:, 0] ^ (-(2 ** 63))
        o1 = torch.argsort(lo_u, stable=True)
        o2 = torch.argsort(allw[o1][:, 1], stable=True)
        o = o1[o2]
    else:
        o = torch.argsort(allw, stable=True)
    s, f = allw[o], flag[o]
    first = torch.ones(s.shape[0], dtype=torch.bool, device=a.device)
    first[1:] = (s[1:] != s[:-1]).any(dim=1) if wide else (s[1:] != s[:-1])
    run = torch.cumsum(first.to(torch.int64), 0) - 1
    mask = torch.zeros(int(run[-1].item()) + 1 if s.shape[0] else 0, dtype=torch.int64, device=a.device)
    mask.scatter_add_(0, run, f)          # 1 = only in a, 2 = only in b, 3 = in both
    uniq = s[first]
    keep = {0: mask > 0, 1: mask == 3, 2: mask == 1, 3: mask != 3}[op]
    return uniq[keep]


def run_setops(args):
    """C4: a op= b for op in | & - ^ on two indexes that share half of their reads; one step = the four operations"""
    cx = Ctx(args)
    torch = cx.torch
    rec = args.record_bp
    mbp = args.index_mbp or 250.0
    n_rec = max(2, int(mbp * 1e6) // rec)
    ra = B.device_dna(torch, n_rec * rec, 5 + 100 * cx.rank, cx.device)
    rb = B.device_dna(torch, n_rec * rec, 6 + 100 * cx.rank, cx.device)
    half = (n_rec // 2) * rec
    rb[:half].copy_(ra[:half])                       # the first half of the records is shared: 50 % overlap
    off = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(rec)
    a, b = cx.new_index(), cx.new_index()
    a.insert_seqs_dev(ra.data_ptr(), off)
    b.insert_seqs_dev(rb.data_ptr(), off)
    na, nb_ = a.count(), b.count()
    names = ["|=", "&=", "-=", "^="]
    per_op_ms = {n: [] for n in names}
    outs = {}
    launches, sampler = 0, None
    for s in range(args.warmup + args.steps):
        if cx.rank == 0:
            sampler = B.sampler_at(s, args.warmup, sampler, cx.world, cx.local, str(torch.cuda.get_device_properties(cx.local).uuid))
        for op, name in enumerate(names):
            c = a.clone()
            l0 = cx.lib.launch_count()
            t = cx.timed(c, lambda: c._assign(op, b))
            if s >= args.warmup:
                per_op_ms[name].append(t)
                launches += cx.lib.launch_count() - l0
            outs[name] = c.count()
            if s == args.warmup + args.steps - 1 and not args.no_parity:
                outs["set" + name] = c
            else:
                cx.close(c)
                del c
    clocks = sampler.stop() if sampler else None
    step_ms = [sum(per_op_ms[n][i] for n in names) for i in range(args.steps)]
    elapsed = sum(step_ms) / 1e3
    value = 4 * (na + nb_) * args.steps / elapsed          # operand k-mers streamed per second
    # roofline (SURVEY 8d set op): (N_a + N_b + N_out) * S + 3 * (4B + B/8) bytes per operation and shard
    la, lb = cx.local_of(a).count(), cx.local_of(b).count()
    roofs = {}
    for name in names:
        lo_out = cx.local_of(outs["set" + name]).count() if ("set" + name) in outs else outs[name] // cx.world
        byts = (la + lb + lo_out) * cx.S + 3 * ((1 << cx.pb) * 4 + (1 << cx.pb) // 8)
        msm = float(np.mean(per_op_ms[name]))
        roofs[name] = {"ms": msm, "algorithmic_bytes": byts, "achieved_GBps": byts / (msm * 1e-3) / 1e9, "frac_of_hbm_peak": byts / (msm * 1e-3) / 1e9 / cx.peak,
                       "out_kmers": outs[name]}
    worst = max(names, key=lambda n: roofs[n]["ms"])
    roof = {"bound": "hbm", "achieved": roofs[worst]["achieved_GBps"], "peak": cx.peak, "unit": "GB/s", "frac": roofs[worst]["frac_of_hbm_peak"], "traffic": None,
            "kernel": f"merge_apply_kernel (CSR x CSR), op {worst}", "ms_per_launch": roofs[worst]["ms"], "peak_source": cx.peak_src,
            "algorithmic_bytes_per_launch": roofs[worst]["algorithmic_bytes"], "model": "SURVEY 8d set op: (N_a + N_b + N_out) * S + 3 * (4B + B/8)", "per_op": roofs}
    parity = None
    if not args.no_parity:
        try:
            t0 = time.perf_counter()
            pa, ea = cx.set_parity(a, ra, off)
            pb_, eb = cx.set_parity(b, rb, off)
            mism = pa["set_mismatches"] + pb_["set_mismatches"]
            checked = pa["set_words_checked"] + pb_["set_words_checked"]
            for op, name in enumerate(names):
                exp = _torch_setop(torch, ea, eb, op, cx.wide)
                got = B.exported_shard_words(torch, cx.local_of(outs["set" + name]), cx.device, cx.wide)
                mism += cx.sum_over_ranks(B.count_set_mismatches(got, exp, cx.wide))
                checked += cx.sum_over_ranks(exp.shape[0])
                del exp, got
            parity = {"set_words_checked": checked, "set_mismatches": mism, "mismatches": mism, "seconds": round(time.perf_counter() - t0, 2), "ranks": cx.world,
                      "how": "operands: stored words == sorted distinct oracle words of their reads; results of |= &= -= ^=: stored words == the same algebra on the "
                             "oracle-derived operand sets (torch sort / scatter as the checker's arithmetic), word by word, every shard"}
        except Exception as e:
            parity = {"error": repr(e), "mismatches": None}
    cpu = None
    if cx.rank == 0 and cx.world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle

        L = pyoracle.load()
        n_c = 10
        oa, ob = pyoracle.OracleCBL(cx.k, cx.t_bits, cx.pb, lib=L), pyoracle.OracleCBL(cx.k, cx.t_bits, cx.pb, lib=L)
        shared = [B.host_dna(min(rec, 1_000_000), 50 + i) for i in range(n_c // 2)]
        for r in shared:
            oa.insert_seq(r)
            ob.insert_seq(r)
        for i in range(n_c // 2):
            oa.insert_seq(B.host_dna(min(rec, 1_000_000), 60 + i))
            ob.insert_seq(B.host_dna(min(rec, 1_000_000), 70 + i))
        t0 = time.perf_counter()
        for op in range(4):
            c = oa.clone()
            c.assign_op(op, ob)
        t = time.perf_counter() - t0
        cpu = {"value": 4 * (oa.count() + ob.count()) / t, "unit": B.UNIT, "cores": 1, "kind": "reference" if L.orc_uses_reference_cxx() else "port",
               "sample": f"two sets of {oa.count()} k-mers sharing half of their reads, |= &= -= ^= on clones (clone time included), {t:.1f} s, 1 thread",
               "host_cores_available": os.cpu_count()}
    if cx.rank == 0:
        cfg = {"workload": f"configs[3]: |= &= -= ^= of two indexes of {n_rec} x {rec} bp per GPU sharing half of their reads, K={cx.k}, T=u{cx.t_bits}, PREFIX_BITS={cx.pb}",
               "kmers_a": na, "kmers_b": nb_, "out_kmers": {n: outs[n] for n in names},
               "l2_policy": f"operands larger than L2: {la * cx.S / 1e6:.0f} + {lb * cx.S / 1e6:.0f} MB of suffixes per shard",
               "parallelism": "1 GPU" if cx.world == 1 else f"prefix-range sharded x{cx.world}: the operations are shard-local, no communication"}
        e2e = {"value": value, "unit": B.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 32 * cx.world,
               "note": "set operations take resident sets and leave a resident set (src/cbl.rs:411-569): the public call moves no bulk data; 8 bytes of totals come back per operation"}
        print(json.dumps(cx.line(args, None, value, 1e3 * elapsed / args.steps, cfg, roof, cpu, e2e, launches, clocks, parity, {"per_op_ms": per_op_ms})))
    for c in [a, b] + [outs[k_] for k_ in outs if k_.startswith("set")]:
        cx.close(c)
    if cx.world > 1:
        cx.dist.destroy_process_group()


def run_stream(args):
    """C5: a stream of 1 Mbp batches cycling insert_seq (new reads) / contains_seq (half known, half new) / remove_seq (reads
    inserted two cycles earlier) on a resident index; one step = the whole stream (default 300 batches per rank)"""
    cx = Ctx(args)
    torch = cx.torch
    rec = args.record_bp
    mbp = args.index_mbp or 500.0
    n_rec = max(1, int(mbp * 1e6) // rec)
    n_batches = int(os.environ.get("CBL_STREAM_BATCHES", 300))
    base = B.device_dna(torch, n_rec * rec, 7 + 100 * cx.rank, cx.device)
    off = np.arange(n_rec + 1, dtype=np.uint64) * np.uint64(rec)
    fresh = B.device_dna(torch, (n_batches // 3 + 2) * rec, 8 + 100 * cx.rank, cx.device).view(-1, rec)
    one = np.array([0, rec], dtype=np.uint64)
    c = cx.new_index()
    c.insert_seqs_dev(base.data_ptr(), off)
    n0 = c.count()
    qbuf = torch.empty(rec, dtype=torch.uint8, device=cx.device)
    ans = torch.empty(rec - cx.k + 1, dtype=torch.uint8, device=cx.device)
    nk = rec - cx.k + 1

    def stream_once():
        hits = 0
        for i in range(n_batches):
            j = i // 3
            if i % 3 == 0:
                c.insert_seqs_dev(fresh[j].data_ptr(), one)
            elif i % 3 == 1:
                qbuf[: rec // 2].copy_(base[j * rec : j * rec + rec // 2])
                qbuf[rec // 2 :].copy_(fresh[j + 1][: rec - rec // 2])        # not inserted yet: misses
                torch.cuda.current_stream().synchronize()
                if cx.world > 1:
                    a = c.contains_seqs_dev(qbuf.data_ptr(), one)
                else:
                    c.contains_seqs_dev(qbuf.data_ptr(), one, ans.data_ptr())
                    a = ans
                hits += int(a.sum(dtype=torch.int64).item())
            elif j >= 2:
                c.remove_seqs_dev(fresh[j - 2].data_ptr(), one)
        return hits

    ms, launches, sampler, hits = [], 0, None, 0
    for s in range(min(args.warmup, 1) + args.steps):
        if cx.rank == 0:
            sampler = B.sampler_at(s, min(args.warmup, 1), sampler, cx.world, cx.local, str(torch.cuda.get_device_properties(cx.local).uuid))
        l0 = cx.lib.launch_count()
        box = {}
        t = cx.timed(c, lambda: box.update(h=stream_once()))
        if s >= min(args.warmup, 1):
            ms.append(t)
            launches += cx.lib.launch_count() - l0
            hits = box["h"]
    clocks = sampler.stop() if sampler else None
    elapsed = sum(ms) / 1e3
    total = n_batches * nk * cx.world
    value = total * len(ms) / elapsed
    # after a full stream: base + the fresh records whose removal has not come round yet
    parity = None
    if not args.no_parity:
        try:
            n_ins = (n_batches + 2) // 3
            n_rem = max(0, sum(1 for i in range(n_batches) if i % 3 == 2 and i // 3 >= 2))
            removed = set(i // 3 - 2 for i in range(n_batches) if i % 3 == 2 and i // 3 >= 2)
            live = [j for j in range(n_ins) if j not in removed]
            reads = torch.cat([base] + [fresh[j] for j in live])
            roff = np.arange(n_rec + len(live) + 1, dtype=np.uint64) * np.uint64(rec)
            parity, _ = cx.set_parity(c, reads, roff)
            parity["stream"] = {"batches": n_batches, "inserted_records": n_ins, "removed_records": n_rem, "contains_hits_last_pass": hits}
        except Exception as e:
            parity = {"error": repr(e), "mismatches": None}
    if cx.rank == 0:
        local_n = cx.local_of(c).count()
        bytes_per_mutation = 2 * local_n * cx.S + nk * cx.W * 8 + 3 * (1 << cx.pb) * 4
        cfg = {"workload": f"configs[4]: {n_batches} batches of {rec} bp per GPU cycling insert_seq / contains_seq / remove_seq on a resident index of {n_rec} x {rec} bp per GPU, "
                           f"K={cx.k}, T=u{cx.t_bits}, PREFIX_BITS={cx.pb}",
               "base_kmers": n0, "final_kmers": c.count(), "batches": n_batches,
               "l2_policy": f"every mutation streams the whole shard ({local_n * cx.S / 1e6:.0f} MB of suffixes, larger than L2)",
               "parallelism": "1 GPU" if cx.world == 1 else f"prefix-range sharded x{cx.world}, fused route over NVLink peer memory"}
        ms_batch = 1e3 * elapsed / len(ms) / n_batches
        roof = {"bound": "hbm", "achieved": bytes_per_mutation / (ms_batch * 1e-3) / 1e9, "peak": cx.peak, "unit": "GB/s",
                "frac": bytes_per_mutation / (ms_batch * 1e-3) / 1e9 / cx.peak, "traffic": None, "kernel": "merge_apply_kernel (full rewrite of the shard per mutation batch)",
                "ms_per_launch": ms_batch, "peak_source": cx.peak_src, "algorithmic_bytes_per_launch": bytes_per_mutation,
                "model": "per mutation batch: read + write of the shard's suffixes (2 N S) + the batch's sort traffic + per-prefix counters; averaged over all batches of the "
                         "stream incl. the contains batches (an upper bound of the mutation kernels' achieved rate)",
                "note": "small batches on a large shard are dominated by the full rewrite (SURVEY section 7 'dynamic updates'): amortised deltas are not built"}
        print(json.dumps(cx.line(args, None, value, 1e3 * elapsed / len(ms), cfg, roof, None, None, launches, clocks, parity, {"ms_per_batch": ms_batch, "ms_steps": ms})))
    cx.close(c)
    if cx.world > 1:
        cx.dist.destroy_process_group()


def run(args):
    if args.config in (1, 3):
        run_build(args, args.config)
    elif args.config == 4:
        run_setops(args)
    else:
        run_stream(args)
