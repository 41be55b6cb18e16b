#!/usr/bin/env python
"""bench.py — benchmarks of the batched sequence path on the BASELINE.json configurations.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--config C] [--metric contains_seq|insert_seq]

Default (what the driver runs): --config 2 = BASELINE configs[1] — contains_seq of a 1 Gbp synthetic FASTA (1000 records
x 1 Mbp, every other record a copy of an index record => ~50 % hits) against a 500M-k-mer index (500 records x 1 Mbp),
K=25, T=u64, PREFIX_BITS=24; per GPU when N > 1 (weak scaling, prefix-range sharded index).  One "step" = one
contains_seq pass over the whole query batch.  The index build (insert_seq) is measured as well and reported in
`extra.insert_seq`; `--metric insert_seq` makes it the headline line (one step = one build of the 500M-k-mer index).

Other configurations (run by hand, outputs kept under profiles/):
  --config 1   cbl build on 10 Mbp, K=25/u64/24, full input on the GPU and on the CPU
  --config 3   K=59 / u128 / PREFIX_BITS=28 build, sharded over N GPUs (size per GPU: --index-mbp)
  --config 4   | & - ^ of two indexes sharing half their reads, K=31 / u128 / 24, sharded over N GPUs
  --config 5   mixed stream of insert_seq / contains_seq / remove_seq batches on a resident index, K=31 / u128 / 24

Prints ONE JSON line (see the task contract): value = k-mers/s with the inputs resident in HBM (CUDA events on the
library's stream), e2e = the same through the host-buffer C-ABI call (pinned host memory in, answers out, copies inside
the timed region), roofline of the dominant kernel, cpu_baseline = the CPU oracle timed on a bounded sample,
parity_check = the step's own answers and the built set checked against the oracle.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "k-mers/s"
# (K, T bits, PREFIX_BITS) per BASELINE config (K=31 needs T=u128, finding F3)
CONFIG_PARAMS = {1: (25, 64, 24), 2: (25, 64, 24), 3: (59, 128, 28), 4: (31, 128, 24), 5: (31, 128, 24)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration, 1-based (2 = configs[1], the headline)")
    ap.add_argument("--metric", default="contains_seq", choices=["contains_seq", "insert_seq"], help="headline of --config 2")
    ap.add_argument("--index-mbp", type=float, default=None, help="index size in Mbp (per GPU when N>1)")
    ap.add_argument("--query-mbp", type=float, default=None, help="query size in Mbp per step (per GPU when N>1)")
    ap.add_argument("--record-bp", type=int, default=1_000_000)
    ap.add_argument("--cpu-index-mbp", type=float, default=100.0, help="CPU baseline sample: index size (BASELINE.md section 3)")
    ap.add_argument("--cpu-query-mbp", type=float, default=10.0, help="CPU baseline sample: query size per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-build-profile", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-threads", type=int, default=0, help="host threads of the oracle legs of parity_check (0 = cores / ranks)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def metric_name(config: int, metric: str) -> str:
    k, t, p = CONFIG_PARAMS[config]
    what = {1: "insert_seq k-mers/s, cbl build of 10 Mbp", 2: f"{metric} k-mers/s; 1 Gbp query vs 500M-k-mer index",
            3: "insert_seq k-mers/s, build of a 3 Gbp-class synthetic FASTA (sharded)", 4: "set-op k-mers/s (| & - ^ of two sharded indexes)",
            5: "mixed insert/remove/contains stream k-mers/s on a resident index"}[config]
    if config == 2:
        return f"{metric} k-mers/s (K={k}, u{t}, PREFIX_BITS={p}; 1 Gbp query vs 500M-k-mer index)"
    return f"{what} (K={k}, u{t}, PREFIX_BITS={p})"


# ------------------------------------------------------------------------------------------------
# synthetic data (uniform ACGT, seeded).  numpy on the host: used by the CPU legs and tiny runs.
# ------------------------------------------------------------------------------------------------
BASES = np.frombuffer(b"ACTG", dtype=np.uint8)


def host_dna(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return BASES[rng.integers(0, 4, size=n, dtype=np.uint8)]


def device_dna(torch, n: int, seed: int, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACTG"), dtype=torch.uint8, device=device)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    step = 1 << 26
    for s in range(0, n, step):
        m = min(step, n - s)
        codes = torch.randint(0, 4, (m,), generator=g, device=device, dtype=torch.int32)
        out[s : s + m] = lut[codes]
    return out


def make_workload(torch, device, index_bp: int, query_bp: int, rec: int, seed_base: int):
    """index = n_i records; query = n_q records where even records are copies of index records
    (hits) and odd records are fresh (misses)."""
    n_i, n_q = max(1, index_bp // rec), max(1, query_bp // rec)
    index = device_dna(torch, n_i * rec, seed_base + 2, device)
    query = device_dna(torch, n_q * rec, seed_base + 3, device)
    iv, qv = index.view(n_i, rec), query.view(n_q, rec)
    for q in range(0, n_q, 2):
        qv[q].copy_(iv[(q // 2) % n_i])
    i_off = np.arange(n_i + 1, dtype=np.uint64) * np.uint64(rec)
    q_off = np.arange(n_q + 1, dtype=np.uint64) * np.uint64(rec)
    return index, i_off, query, q_off


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an in-process NVML thread (5 ms period; the
    timed region of this bench is tens of milliseconds, too short for an `nvidia-smi -lms` child to start up),
    falling back to `nvidia-smi -lms` when NVML is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.rows, self.proc, self.nv, self.samples = [], None, None, []
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nv, self.h = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
                except Exception:
                    pw = None
                self.samples.append((mhz, rs, pw))
            except Exception:
                pass
            self._stop.wait(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nv is not None:
            self._stop.set()
            self.t.join(timeout=1)
            nv = self.nv
            flags = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
                     "hw_power_brake_slowdown": nv.nvmlClocksEventReasonHwPowerBrakeSlowdown}
            reasons = sorted(k for k, f in flags.items() if any(rs & f for _, rs, _ in self.samples))
            sm = [m for m, _, _ in self.samples]
            pw = [p for _, _, p in self.samples if p is not None]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                    "power_w_max": max(pw) if pw else None, "source": "nvml, 5 ms period, timed region only"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}



_POLL_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
u, idx, period = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
try:
    h = nv.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
except Exception:
    h = nv.nvmlDeviceGetHandleByIndex(idx)
print("max", float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)), flush=True)
while True:
    try:
        pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
    except Exception:
        pw = -1.0
    print(time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)), pw, flush=True)
    time.sleep(period)
"""


def sampler_at(step_i: int, first_timed: int, sampler, world: int, local: int, uuid: str):
    """Clock sampler of a timed loop, called by rank 0 at the top of every iteration: N = 1 -> the in-process NVML thread,
    started with the first timed step; N > 1 -> the child-process sampler, started with the first iteration (it needs a moment
    to come up) and armed with the first timed step."""
    if world > 1:
        if step_i == 0 and sampler is None:
            sampler = ExternalClockSampler(local, uuid)
        if step_i == first_timed and sampler is not None:
            sampler.begin()
    elif step_i == first_timed:
        sampler = ClockSampler(local, uuid)
    return sampler


class ExternalClockSampler:
    """The same samples taken by a CHILD process (NVML, 5 ms period), used when N > 1: in the sharded step every phase is
    host-synchronous, and an in-process NVML thread (driver locks, GIL hand-overs) was seen to hold rank 0 back by tens of
    milliseconds in some steps, which every other rank then waits for.  Start it early (the child needs a moment to come
    up), call begin() right before the timed region; stop() keeps the samples taken between begin() and stop()."""

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.rows, self.t0 = [], None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _POLL_SRC, uuid or "", str(gpu_index), "0.005"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.split())

    def begin(self):
        self.t0 = time.time()

    def stop(self):
        t1 = time.time()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml child unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mx = next((float(r[1]) for r in self.rows if r and r[0] == "max"), None)
        rows = [r for r in self.rows if len(r) == 4 and r[0] != "max" and (self.t0 or 0) <= float(r[0]) <= t1]
        try:
            import pynvml as nv

            flags = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
                     "hw_power_brake_slowdown": nv.nvmlClocksEventReasonHwPowerBrakeSlowdown}
        except Exception:
            flags = {}
        reasons = sorted(k for k, f in flags.items() if any(int(r[2]) & f for r in rows))
        sm = [float(r[1]) for r in rows]
        pw = [float(r[3]) for r in rows if float(r[3]) >= 0]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm),
                "power_w_max": max(pw) if pw else None, "source": "nvml in a child process, 5 ms period, timed region only"}

# ------------------------------------------------------------------------------------------------
# CPU legs (the oracle is the checker / baseline, never the product)
# ------------------------------------------------------------------------------------------------
def cpu_leg(k, t_bits, prefix_bits, index_mbp: float, query_mbp: float, rec: int, steps: int = 1, warmup: int = 0):
    """The reference's CPU path (restated oracle, 1 thread: the reference is single-threaded) on a bounded
    sample of the configs[1] workload: build an index of index_mbp, then `steps` contains_seq passes over query_mbp."""
    from oracle import pyoracle

    L = pyoracle.load()
    kind_note = "restated reference (C++ port of the Rust path) linked against the reference's own sux/tiered-vector C++" if L.orc_uses_reference_cxx() else "restated reference (C++ port), stand-in bitvector/tiered vector"
    rec = min(rec, 1_000_000)
    n_i, n_q = max(1, int(index_mbp * 1e6) // rec), max(1, int(query_mbp * 1e6) // rec)
    index = np.concatenate([host_dna(rec, 1000 + i) for i in range(n_i)])
    qrecs = [index[(q // 2 % n_i) * rec : (q // 2 % n_i + 1) * rec] if q % 2 == 0 else host_dna(rec, 5000 + q) for q in range(n_q)]
    query = np.concatenate(qrecs)
    i_off = np.arange(n_i + 1, dtype=np.uint64) * np.uint64(rec)
    q_off = np.arange(n_q + 1, dtype=np.uint64) * np.uint64(rec)
    o = pyoracle.OracleCBL(k, t_bits, prefix_bits, lib=L)
    t_ins = o.time_insert_seqs(index, i_off)
    n_ins = n_i * (rec - k + 1)
    n_q_kmers = n_q * (rec - k + 1)
    times = []
    pos = 0
    for s in range(warmup + steps):
        t, pos = o.time_contains_seqs(query, q_off)
        if s >= warmup:
            times.append(t)
    t_q = float(np.mean(times))
    return {
        "contains_kmers_per_s": n_q_kmers / t_q,
        "insert_kmers_per_s": n_ins / t_ins,
        "ms_per_step": 1e3 * t_q,
        "insert_s": t_ins,
        "sample": f"index {n_i} x {rec} bp ({n_ins} k-mers, build {t_ins:.1f} s), query {n_q} x {rec} bp per step (50% hit records), 1 thread",
        "kind_note": kind_note,
        "kind": "reference" if L.orc_uses_reference_cxx() else "port",
        "positives": pos,
        "n_q_kmers": n_q_kmers,
    }


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (restated oracle linked against the reference's
    own C++ half when oracle/_ref is present; the Rust crate cannot be built in this image), single-threaded because the
    reference is.  One step = contains_seq (or, --metric insert_seq, one build) of the bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k, t_bits, pb = CONFIG_PARAMS[args.config]
    insert_line = args.config in (1, 3) or (args.config == 2 and args.metric == "insert_seq")
    if insert_line:
        # one step = one build of the sample (a fresh oracle set per step)
        from oracle import pyoracle

        L = pyoracle.load()
        rec = min(args.record_bp, 1_000_000)
        mbp = 10.0 if args.config == 1 else min(args.cpu_index_mbp, 20.0)
        n_i = max(1, int(mbp * 1e6) // rec)
        index = np.concatenate([host_dna(rec, 1000 + i) for i in range(n_i)])
        i_off = np.arange(n_i + 1, dtype=np.uint64) * np.uint64(rec)
        n_ins = n_i * (rec - k + 1)
        times = []
        for s in range(min(args.warmup, 1) + max(1, min(args.steps, 5))):
            o = pyoracle.OracleCBL(k, t_bits, pb, lib=L)
            t = o.time_insert_seqs(index, i_off)
            del o
            if s >= min(args.warmup, 1):
                times.append(t)
        t_b = float(np.mean(times))
        value, ms = n_ins / t_b, 1e3 * t_b
        sample = f"build of {n_i} x {rec} bp ({n_ins} k-mers) into an empty set per step, 1 thread"
        kind = "reference" if L.orc_uses_reference_cxx() else "port"
        note = "restated reference linked against the reference's own C++ half" if kind == "reference" else "restated reference (C++ port)"
        extra = {}
    else:
        r = cpu_leg(k, t_bits, pb, args.cpu_index_mbp, args.cpu_query_mbp, args.record_bp, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        value, ms, sample, kind, note = r["contains_kmers_per_s"], r["ms_per_step"], r["sample"], r["kind"], r["kind_note"]
        extra = {"insert_seq_kmers_per_s": r["insert_kmers_per_s"]}
    line = {
        "impl": "reference", "metric": metric_name(args.config, args.metric), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": f"u{t_bits}", "data": "synthetic",
        "config": {"workload": f"configs[{args.config - 1}] K={k} u{t_bits} PREFIX_BITS={pb} (bounded CPU sample of the same workload)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "note": note, "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "extra": extra,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# rooflines (SURVEY section 8d algorithmic bytes)
# ------------------------------------------------------------------------------------------------
def build_phase_roofline(build_prof, n_kmers, stored, peak, W=8, S=4, prefix_bits=24):
    """Per-kernel HBM roofline of the build (insert_seq) phases: algorithmic bytes of SURVEY 8d (device word W bytes,
    stored suffix S bytes) divided by the kernel's CUDA-event time.  The words kernel is integer-bound
    (its roofline is the ALU pipe, see profiles/), listed for completeness."""
    if not build_prof:
        return None
    per_launch = {
        "seq_words_kernel": n_kmers * (1 + W),            # ASCII in, word out (digit histograms ride along)
        "radix_hist_kernel": n_kmers * W,                  # one read of the words (only when not folded into the words kernel)
        "radix_pass_kernel": n_kmers * 2 * W,              # read + scatter per pass
        "seg_sort_kernel": n_kmers * 2 * W,                # read + write, whatever the number of remaining bits
        "merge_apply_kernel": n_kmers * W + stored * S + 3 * (1 << prefix_bits) * 4,   # batch words in, new suffixes out, per-prefix counters
    }
    out = {}
    total_bytes, total_ms = 0.0, 0.0
    for name, rec in build_prof.items():
        for key, nbytes in per_launch.items():
            if name.startswith(key) and rec.get("ms"):
                gbs = nbytes * rec["n"] / (rec["ms"] * 1e-3) / 1e9
                out[key] = {"launches": rec["n"], "ms": rec["ms"], "algorithmic_bytes_per_launch": nbytes, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
                total_bytes += nbytes * rec["n"]
        total_ms += rec.get("ms", 0.0)
    if total_ms:
        out["_aggregate"] = {"algorithmic_bytes": total_bytes, "kernel_ms": total_ms, "achieved_GBps": total_bytes / (total_ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": total_bytes / (total_ms * 1e-3) / 1e9 / peak, "bytes_per_kmer": total_bytes / max(1, n_kmers)}
    return out


# ------------------------------------------------------------------------------------------------
# parity_check: the run's own results against the oracle (test infrastructure: oracle words on the host cores,
# torch sort / searchsorted and NCCL on the GPUs as the set arithmetic of the CHECKER — none of it is on the timed path)
# ------------------------------------------------------------------------------------------------
def oracle_words(host_buf: np.ndarray, offsets: np.ndarray, recs, k, t_bits, pb, canonical, threads: int):
    """oracle (restated src/cbl.rs:247-289) words of the given records, one (lo, hi) pair of arrays per record"""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import pyoracle

    L = pyoracle.load()
    tls = threading.local()

    def one(r):
        o = getattr(tls, "o", None)
        if o is None:
            o = tls.o = pyoracle.OracleCBL(k, t_bits, pb, canonical, lib=L)   # get_seq_words uses per-handle scratch queues
        return o.seq_words(host_buf[int(offsets[r]) : int(offsets[r + 1])])

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        return list(ex.map(one, recs))


def words_to_torch(torch, device, parts, wide: bool):
    """list of (lo, hi) numpy pairs -> int64 tensor (n,) or, for 128-bit words, (n, 2) = [lo, hi] on the device"""
    n = sum(len(p[0]) for p in parts)
    out = torch.empty((n, 2) if wide else (n,), dtype=torch.int64, device=device)
    at = 0
    for lo, hi in parts:
        m = len(lo)
        if wide:
            out[at : at + m, 0] = torch.from_numpy(lo.view(np.int64)).to(device)
            out[at : at + m, 1] = torch.from_numpy(hi.view(np.int64)).to(device)
        else:
            out[at : at + m] = torch.from_numpy(lo.view(np.int64)).to(device)
        at += m
    return out


def sort_unique(torch, w):
    """ascending distinct words (unsigned order; words stay below 2^63 per half except the low half of 128-bit words)"""
    if w.dim() == 1:
        s, _ = torch.sort(w)          # all words < 2^63 (word bits <= 62 for u64 configs)
        keep = torch.ones_like(s, dtype=torch.bool)
        keep[1:] = s[1:] != s[:-1]
        return s[keep]
    lo, hi = w[:, 0], w[:, 1]
    lo_u = lo ^ (-(2 ** 63))          # unsigned order of the low half under signed compares
    o1 = torch.argsort(lo_u, stable=True)
    o2 = torch.argsort(hi[o1], stable=True)
    o = o1[o2]
    s = w[o]
    keep = torch.ones(s.shape[0], dtype=torch.bool, device=w.device)
    keep[1:] = (s[1:] != s[:-1]).any(dim=1)
    return s[keep]


def word_prefix(torch, w, suffix_bits: int):
    if w.dim() == 1:
        return w >> suffix_bits
    lo, hi = w[:, 0], w[:, 1]
    if suffix_bits >= 64:
        return hi >> (suffix_bits - 64)
    return (hi << (64 - suffix_bits)) | ((lo >> suffix_bits) & ((1 << (64 - suffix_bits)) - 1))


def membership(torch, sorted_set, q):
    """q in sorted_set (both int64 (n,) or (n, 2) [lo, hi])"""
    if sorted_set.shape[0] == 0:
        return torch.zeros(q.shape[0], dtype=torch.bool, device=q.device)
    if q.dim() == 1:
        i = torch.searchsorted(sorted_set, q).clamp_(max=sorted_set.shape[0] - 1)
        return sorted_set[i] == q
    # 128-bit: search on hi, then scan the (short) run of equal hi for lo — words of one prefix share hi only rarely;
    # do it exactly with a combined key of the rank of (hi, lo) pairs instead
    allw = torch.cat([sorted_set, q])
    flag = torch.cat([torch.zeros(sorted_set.shape[0], dtype=torch.int64, device=q.device), torch.ones(q.shape[0], dtype=torch.int64, device=q.device)])
    lo_u = allw[:, 0] ^ (-(2 ** 63))
    o1 = torch.argsort(flag, stable=True)                      # set elements before queries among equals
    o2 = torch.argsort(lo_u[o1], stable=True)
    o3 = torch.argsort(allw[o1][o2][:, 1], stable=True)
    o = o1[o2][o3]
    s, f = allw[o], flag[o]
    same_as_prev = torch.zeros(s.shape[0], dtype=torch.bool, device=q.device)
    same_as_prev[1:] = (s[1:] == s[:-1]).all(dim=1)
    # a query is present iff the run of equal words it sits in starts with a set element
    run_start = ~same_as_prev
    run_id = torch.cumsum(run_start.to(torch.int64), 0) - 1
    run_has_set = torch.zeros(int(run_id[-1].item()) + 1, dtype=torch.bool, device=q.device)
    run_has_set[run_id[f == 0]] = True
    res = torch.zeros(allw.shape[0], dtype=torch.bool, device=q.device)
    res[o] = run_has_set[run_id]
    return res[sorted_set.shape[0] :]


def word_geometry(k: int, pb: int):
    word_bits = 2 * k + (2 * k - 1).bit_length()
    return word_bits > 64, word_bits - pb          # (128-bit device words?, SUFFIX_BITS)


def expected_shard_set(torch, dist, cbl, world, rank, device, k, t_bits, pb, canonical, reads_dev, offsets, threads):
    """oracle words of EVERY record of `reads_dev` (all ranks), routed to the owner rank by the shard's splitters, sorted +
    deduplicated: the words this rank's shard must hold after insert_seq of those reads into an empty set"""
    wide, suffix_bits = word_geometry(k, pb)
    h = reads_dev.cpu().numpy()
    parts = oracle_words(h, offsets, range(len(offsets) - 1), k, t_bits, pb, canonical, threads)
    mine = words_to_torch(torch, device, parts, wide)
    del parts, h
    if world > 1:
        sp = cbl.splitters.to(device)
        dest = torch.bucketize(word_prefix(torch, mine, suffix_bits), sp, right=True)
        order = torch.argsort(dest, stable=True)
        send = mine[order].contiguous()
        sc = torch.bincount(dest, minlength=world)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
        recv = send.new_empty((int(rc.sum().item()),) + tuple(send.shape[1:]))
        dist.all_to_all_single(recv, send, output_split_sizes=rc.tolist(), input_split_sizes=sc.tolist())
        del send, mine, order, dest
    else:
        recv = mine
    return sort_unique(torch, recv)


def exported_shard_words(torch, local, device, wide: bool):
    """the shard's stored words in ascending order, left on the device (int64 (n,) or (n, 2) = [lo, hi])"""
    n_loc = local.count()
    got = torch.empty((n_loc, 2) if wide else (n_loc,), dtype=torch.int64, device=device)
    torch.cuda.synchronize()
    at, CH = 0, 1 << 27
    while at < n_loc:
        m = min(CH, n_loc - at)
        local.export_words_dev(at, m, got.data_ptr() + at * (16 if wide else 8))
        at += m
    return got


def count_set_mismatches(got, expect, wide: bool) -> int:
    if got.shape[0] == expect.shape[0]:
        return int((got != expect).any(dim=1).sum().item()) if wide else int((got != expect).sum().item())
    m = min(got.shape[0], expect.shape[0])
    d = got[:m] != expect[:m]
    return abs(got.shape[0] - expect.shape[0]) + int((d.any(dim=1) if wide else d).sum().item())


def parity_check(torch, dist, cbl, world, rank, device, k, t_bits, pb, canonical, index_dev, i_off, query_dev, q_off, answers_dev, threads,
                 sample_records=None):
    """(1) the WHOLE built set: oracle words of every index record (all ranks), routed to the owner rank by the shard's
    splitters, sorted + deduplicated == the shard's stored words in ascending order (and the count);
    (2) the step's own answers for >= 1 % of the query records (hit and miss records alike) == membership of their
    oracle words in that oracle-derived set.  Returns a dict for the JSON line."""
    t0 = time.perf_counter()
    wide, suffix_bits = word_geometry(k, pb)
    n_i = len(i_off) - 1
    expect = expected_shard_set(torch, dist, cbl, world, rank, device, k, t_bits, pb, canonical, index_dev, i_off, threads)
    t_words = time.perf_counter() - t0
    local = cbl.engine.cbl if world > 1 else cbl
    got = exported_shard_words(torch, local, device, wide)
    n_loc = got.shape[0]
    set_mismatch = count_set_mismatches(got, expect, wide)
    del got
    # ---- answers of sampled query records
    n_q = len(q_off) - 1
    if sample_records is None:
        n_s = max(2, int(math.ceil(0.01 * n_q)))
        n_s += n_s % 2
        half = n_s // 2
        ev = [2 * (i * max(1, (n_q // 2) // half)) for i in range(half)]                      # hit records (even)
        od = [min(n_q - 1, e + 1) for e in ev]                                                 # miss records (odd)
        sample_records = sorted(set(r for r in ev + od if r < n_q))
    h_query = np.concatenate([query_dev[int(q_off[r]) : int(q_off[r + 1])].cpu().numpy() for r in sample_records])
    s_off = np.zeros(len(sample_records) + 1, dtype=np.uint64)
    s_off[1:] = np.cumsum([int(q_off[r + 1] - q_off[r]) for r in sample_records])
    qparts = oracle_words(h_query, s_off, range(len(sample_records)), k, t_bits, pb, canonical, threads)
    qw = words_to_torch(torch, device, qparts, wide)
    if world > 1:
        sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([qw.shape[0]], dtype=torch.int64, device=device))
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes)
        pad = qw.new_zeros((mx,) + tuple(qw.shape[1:]))
        pad[: qw.shape[0]] = qw
        allq = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(allq, pad)
        memb = torch.stack([membership(torch, expect, a).to(torch.uint8) for a in allq])      # (world, mx): is it in MY range's set
        dist.all_reduce(memb, op=dist.ReduceOp.MAX)
        exp_ans = memb[rank, : qw.shape[0]]
    else:
        exp_ans = membership(torch, expect, qw).to(torch.uint8)
    got_ans = torch.cat([answers_dev[int(q_off[r]) - r * (k - 1) : int(q_off[r + 1]) - (r + 1) * (k - 1)] for r in sample_records])
    ans_mismatch = int((got_ans != exp_ans).sum().item())
    res = {"records": len(sample_records), "kmers": int(qw.shape[0]), "mismatches": ans_mismatch, "hits_expected": int(exp_ans.sum().item()),
           "set_words_checked": int(expect.shape[0]), "set_mismatches": set_mismatch, "count_expected": int(expect.shape[0]), "count_got": int(n_loc),
           "index_records": n_i, "oracle_words_s": round(t_words, 2), "seconds": round(time.perf_counter() - t0, 2)}
    if world > 1:   # every rank checked its own shard and its own sampled records: sum the verdicts
        v = torch.tensor([res["records"], res["kmers"], res["mismatches"], res["set_words_checked"], res["set_mismatches"], res["count_expected"],
                          res["count_got"], res["hits_expected"]], dtype=torch.int64, device=device)
        dist.all_reduce(v)
        res.update(records=int(v[0]), kmers=int(v[1]), mismatches=int(v[2]), set_words_checked=int(v[3]), set_mismatches=int(v[4]),
                   count_expected=int(v[5]), count_got=int(v[6]), hits_expected=int(v[7]), ranks=world)
    res["how"] = ("oracle words (restated src/cbl.rs:247-289) of EVERY index record of every rank, routed to the owner shard, sorted + deduplicated, "
                  "compared word by word with the shard's stored set; the timed step's answers for the sampled records compared with membership "
                  "of their oracle words in that set")
    return res


# ------------------------------------------------------------------------------------------------
# ours: config 2 (headline) — contains_seq / insert_seq, K=25 u64 PREFIX_BITS=24
# ------------------------------------------------------------------------------------------------
def run_config2(args):
    import torch
    import torch.distributed as dist

    import cbl_b200

    K, T_BITS, PREFIX_BITS = CONFIG_PARAMS[2]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU leg")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    rec = args.record_bp
    index_bp, query_bp = int((args.index_mbp or 500.0) * 1e6), int((args.query_mbp or 1000.0) * 1e6)
    peak, peak_src = peaks()

    def new_index():
        if world > 1:
            from cbl_b200.sharded import ShardedCBL

            return ShardedCBL(K, T_BITS, PREFIX_BITS, canonical=False, device=local)
        return cbl_b200.CBL(K, T_BITS, PREFIX_BITS, canonical=False, device=local)

    cbl = new_index()
    index, i_off, query, q_off = make_workload(torch, device, index_bp, query_bp, rec, seed_base=100 * rank)
    n_q_kmers = (len(q_off) - 1) * (rec - K + 1)
    n_i_kmers = (len(i_off) - 1) * (rec - K + 1)
    answers = torch.empty(n_q_kmers, dtype=torch.uint8, device=device)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def lib_stream(c):
        return torch.cuda.ExternalStream(c.stream_ptr(), device=device)

    # ---- build (insert_seq): the index the query steps run against; timed (cold: first use of the memory arena)
    barrier()
    t0 = time.perf_counter()
    cbl.insert_seqs_dev(index.data_ptr(), i_off)
    barrier()
    t_build_first = time.perf_counter() - t0
    stored = cbl.count()
    nb = cbl.num_buckets()

    # ---- insert_seq steps: one step = one build of the whole index into a FRESH empty set (device-resident reads);
    #      CUDA events on the handle's stream around each build, max over ranks of the sum
    insert_headline = args.metric == "insert_seq"
    n_ins_steps = args.steps if insert_headline else min(args.steps, 3)
    n_ins_warm = args.warmup if insert_headline else 1
    ins_ms, ins_wall = [], []
    launches_ins = 0
    sampler = None
    for s in range(n_ins_warm + n_ins_steps):
        scratch = new_index()
        barrier()
        if insert_headline and rank == 0:
            sampler = sampler_at(s, n_ins_warm, sampler, world, local, str(torch.cuda.get_device_properties(local).uuid))
        st = lib_stream(scratch)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = cbl_b200.launch_count()
        e0.record(st)
        t0 = time.perf_counter()
        scratch.insert_seqs_dev(index.data_ptr(), i_off)
        e1.record(st)
        barrier()
        w = time.perf_counter() - t0
        if s >= n_ins_warm:
            ins_ms.append(max_over_ranks(e0.elapsed_time(e1)))
            ins_wall.append(w)
            launches_ins += cbl_b200.launch_count() - l0
        if world > 1:
            scratch.close()
        del scratch
    clocks_ins = sampler.stop() if sampler else None
    ins_elapsed = sum(ins_ms) / 1e3
    insert_value = n_i_kmers * world * n_ins_steps / ins_elapsed

    # per-kernel attribution of the build: a separate, untimed build with CUDA events around every launch
    build_prof = None
    if world == 1 and not args.no_build_profile:
        scratch = new_index()
        cbl_b200.profile_enable(True)
        cbl_b200.profile_report()
        scratch.insert_seqs_dev(index.data_ptr(), i_off)
        torch.cuda.synchronize()
        build_prof = cbl_b200.profile_report()
        cbl_b200.profile_enable(False)
        del scratch
    build_roof = None
    try:
        build_roof = build_phase_roofline(build_prof, n_i_kmers, stored, peak)
    except Exception as e:  # reporting only: never lose the bench line over it
        build_roof = {"error": repr(e)}

    # insert_seq end to end: pinned host buffers through cbl_insert_seqs (H2D inside the timed region, count read back)
    ins_e2e = None
    if not args.no_e2e:
        h_index = torch.empty(index.numel(), dtype=torch.uint8, pin_memory=True)
        h_index.copy_(index)
        hi_np = h_index.numpy()
        ts = []
        for s in range(1 + min(n_ins_steps, 3)):
            scratch = new_index()
            barrier()
            t0 = time.perf_counter()
            scratch.insert_seqs(hi_np, i_off)
            c = scratch.count()
            barrier()
            if s >= 1:
                ts.append(time.perf_counter() - t0)
            assert c == stored, f"host-buffer build stored {c} k-mers, device-resident build {stored}"
            if world > 1:
                scratch.close()
            del scratch
        t_e = max_over_ranks(sum(ts))
        ins_e2e = {"value": n_i_kmers * world * len(ts) / t_e, "unit": UNIT, "h2d_bytes_per_step": int(index.numel()) * world, "d2h_bytes_per_step": 8 * world,
                   "ms_per_step": 1e3 * t_e / len(ts)}
        del h_index, hi_np

    # ---- timed contains_seq steps (query resident in HBM), profiling OFF
    box = {"answers": answers}

    def step():
        if world > 1:
            box["answers"] = cbl.contains_seqs_dev(query.data_ptr(), q_off)  # sharded: returns the answers tensor
        else:
            cbl.contains_seqs_dev(query.data_ptr(), q_off, answers.data_ptr())

    sampler = None
    if rank == 0 and not insert_headline and world > 1:
        sampler = ExternalClockSampler(local, str(torch.cuda.get_device_properties(local).uuid))   # comes up during the warm-up
    for _ in range(args.warmup):
        step()
    barrier()
    if rank == 0 and not insert_headline and world == 1:
        sampler = ClockSampler(local, str(torch.cuda.get_device_properties(local).uuid))
    if isinstance(sampler, ExternalClockSampler):
        sampler.begin()
    launches0 = cbl_b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the stream the library launches on (sharded: every library call in the step is host-synchronous, so two events on
    # the shard's stream bracket the whole region on the device timeline, host gaps included)
    stream = lib_stream(cbl)
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = cbl_b200.launch_count() - launches0
    clocks_q = sampler.stop() if sampler else None
    elapsed = max_over_ranks(dev_ms / 1e3)
    hits = int(box["answers"].sum(dtype=torch.int64).item())
    total_q = n_q_kmers * world
    value = total_q * args.steps / elapsed

    # ---- per-kernel attribution of the query step: a separate, untimed pass with CUDA events around every launch
    cbl_b200.profile_enable(True)
    cbl_b200.profile_report()
    step()
    torch.cuda.synchronize()
    prof = cbl_b200.profile_report()
    cbl_b200.profile_enable(False)
    barrier()

    # ---- roofline of the dominant kernel of the step (fused encode + necklace + probe)
    # Algorithmic bytes per k-mer of the realised branch of SURVEY 8d's probe figure (independent random lookups, queries not
    # sorted): 1 B ASCII in + 1 B answer out + 8 B directory word + 8 B bucket range + 2 B interpolation corrections + ONE
    # 32-byte suffix window (the minimum a lookup must touch).  Extra windows after a mispredicted slot and table re-fetches
    # after L2 misses are waste: they show up in `traffic` (measured DRAM bytes per launch from the ncu capture named in
    # profiles/roofline_traffic.json).  The kernel is instruction-issue bound, not byte bound: the
    # fraction says how far the realised algorithm is from the HBM roof, and frac_vs_sec8d_streaming_bound how far from the
    # streaming branch of SURVEY 8d's min().
    dom = None
    for name, rec_ in prof.items():
        if ("seq_words_kernel" in name or "shard_query_kernel" in name) and (dom is None or rec_["ms"] > prof[dom]["ms"]):
            dom = name
    roof = None
    if dom:
        ms_per_launch = prof[dom]["ms"] / max(1, prof[dom]["n"])
        bytes_per_kmer = 1 + 1 + 8 + 8 + 2 + 32
        fused_sharded = "shard_query_kernel" in dom
        if fused_sharded:
            # the fused sharded query also moves every word once out (8 B store into the owner's region: remote HBM for 7 of 8
            # words, but every GPU receives as much as it sends), once in (8 B) and resets its slot (8 B), and writes the 4-byte
            # answer slot of every k-mer for the gather that follows
            bytes_per_kmer += 8 + 8 + 8 + 4
        kmers_per_launch = n_q_kmers / max(1, prof[dom]["n"])
        achieved = bytes_per_kmer * kmers_per_launch / (ms_per_launch * 1e-3) / 1e9
        stream_bpk = 8 + 1 + 8 + min(stored * 4 / max(1, n_q_kmers), math.ceil(math.log2(stored / max(1, nb) + 1)) * 32)
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if fused_sharded:
                    tj = tj["fused_sharded"]   # captured on the 1-GPU test bed of the kernel (ncu cannot replay a multi-rank kernel)
                traffic = tj["dram_bytes_per_kmer"] * kmers_per_launch
                traffic_src = tj.get("source")
            except Exception:
                pass
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": dom, "ms_per_launch": ms_per_launch, "peak_source": peak_src,
                "algorithmic_bytes_per_kmer": bytes_per_kmer, "algorithmic_bytes_per_launch": bytes_per_kmer * kmers_per_launch,
                "traffic_source": traffic_src,
                "model": "1 B ASCII + 1 B answer + 8 B directory word + 8 B bucket range + 2 B corrections + one 32 B suffix window per k-mer "
                         "(realised branch of SURVEY 8d's probe figure: unsorted queries, random lookups)"
                         + (" + 8 B word out + 8 B word in + 8 B slot reset + 4 B answer slot (fused sharded query: shard_query.cuh)" if fused_sharded else ""),
                "frac_vs_sec8d_streaming_bound": stream_bpk * kmers_per_launch / (ms_per_launch * 1e-3) / 1e9 / peak,
                "sec8d_streaming_bytes_per_kmer": stream_bpk,
                "note": "instruction-issue bound (518 thread-instructions per k-mer: necklace ~250, probe ~270), not HBM-byte bound; the 60 % "
                        "HBM target is missed — see DESIGN.md section 6",
                "mean_bucket": stored / max(1, nb),
                "kernel_share_of_step": prof[dom]["ms"] / (1e3 * elapsed / args.steps)}
    if insert_headline and build_roof:
        # headline = the build: roofline of its dominant kernel (radix pass), aggregate in extra.build_roofline
        dk = max((k_ for k_ in build_roof if not k_.startswith("_")), key=lambda k_: build_roof[k_]["ms"], default=None)
        if dk:
            r = build_roof[dk]
            roof = {"bound": "hbm", "achieved": r["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": r["frac_of_hbm_peak"], "traffic": None,
                    "kernel": dk, "ms_per_launch": r["ms"] / r["launches"], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": r["algorithmic_bytes_per_launch"],
                    "aggregate": build_roof.get("_aggregate"),
                    "model": "SURVEY 8d per-phase bytes: words n(1+W); radix pass 2nW; segment sort 2nW; merge nW + N'S + 3*2^P*4"}

    # ---- e2e: host buffers through the C ABI (pinned memory; H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        h_query = torch.empty(query.numel(), dtype=torch.uint8, pin_memory=True)
        h_query.copy_(query)
        h_ans = torch.empty(n_q_kmers, dtype=torch.uint8, pin_memory=True)
        hq, ha = h_query.numpy(), h_ans.numpy()
        for _ in range(min(args.warmup, 2)):
            cbl.contains_seqs(hq, q_off, out=ha)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cbl.contains_seqs(hq, q_off, out=ha)
        barrier()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        assert int(ha.sum(dtype=np.int64)) == hits, "e2e answers differ from the device-resident run"
        e2e = {"value": total_q * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(query.numel()) * world,
               "d2h_bytes_per_step": int(n_q_kmers) * world, "ms_per_step": 1e3 * t_e2e / args.steps}
        # what the host links allow: the same bytes copied in and out by every rank at the same time, nothing else running
        # (two streams, pinned memory).  e2e cannot beat this; on a box whose GPUs share host bandwidth it is what limits e2e at N > 1
        d_in, d_out = torch.empty_like(query), torch.empty(n_q_kmers, dtype=torch.uint8, device=device)
        s_a, s_b = torch.cuda.Stream(device), torch.cuda.Stream(device)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            with torch.cuda.stream(s_a):
                d_in.copy_(h_query, non_blocking=True)
            with torch.cuda.stream(s_b):
                h_ans.copy_(d_out, non_blocking=True)
        s_a.synchronize()
        s_b.synchronize()
        barrier()
        t_link = max_over_ranks(time.perf_counter() - t0)
        e2e["host_link_floor"] = {"ms_per_step": 1e3 * t_link / args.steps,
                                  "GBps_all_ranks_both_directions": (int(query.numel()) + int(n_q_kmers)) * world * args.steps / t_link / 1e9,
                                  "what": "pure H2D + D2H of the step's bytes on every rank at once (pinned memory, two streams): the floor of e2e on this box"}
        del h_query, h_ans, d_in, d_out

    # ---- parity_check: the built set and the step's own answers against the oracle (every rank)
    parity = None
    if not args.no_parity:
        threads = args.parity_threads or max(1, (os.cpu_count() or 8) // max(1, world))
        try:
            parity = parity_check(torch, dist, cbl, world, rank, device, K, T_BITS, PREFIX_BITS, False, index, i_off, query, q_off, box["answers"], threads)
        except Exception as e:   # a failed CHECK must be visible, never silently dropped
            parity = {"error": repr(e), "mismatches": None}

    # ---- CPU baseline (rank 0, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_leg(K, T_BITS, PREFIX_BITS, args.cpu_index_mbp, args.cpu_query_mbp, rec)
        cpu = {"value": r["insert_kmers_per_s"] if insert_headline else r["contains_kmers_per_s"], "unit": UNIT, "cores": 1, "kind": r["kind"],
               "sample": r["sample"], "note": r["kind_note"], "contains_seq_kmers_per_s": r["contains_kmers_per_s"],
               "insert_seq_kmers_per_s": r["insert_kmers_per_s"], "host_cores_available": os.cpu_count()}

    # NVLink traffic of the fused sharded query: 8-byte words out to the other owners, 1-byte answers back, over the kernel's time
    nvlink = None
    C = getattr(cbl, "last_route_counts", None) if world > 1 else None
    if C is not None and dom and "shard_query_kernel" in dom:
        sent_away = int(C[rank].sum()) - int(C[rank][rank])
        recv_from_others = int(C[:, rank].sum()) - int(C[rank][rank])
        k_ms = prof[dom]["ms"] / max(1, prof[dom]["n"])
        nvlink = {"words_out": sent_away, "answers_out": recv_from_others,
                  "egress_GBps": (sent_away * 8 + recv_from_others) / (k_ms * 1e-3) / 1e9,
                  "ingress_GBps": (recv_from_others * 8 + sent_away) / (k_ms * 1e-3) / 1e9,
                  "kernel_ms": k_ms, "rank": rank,
                  "what": "rank 0 over its fused query kernel: 8-byte word stores to the other owners + 1-byte answers back (NVLink 5 / NVSwitch: 900 GB/s per direction)"}
    if rank == 0:
        insert_block = {"value": insert_value, "unit": UNIT, "steps": n_ins_steps, "ms_per_step": 1e3 * ins_elapsed / n_ins_steps,
                        "wall_ms_per_step": 1e3 * float(np.mean(ins_wall)), "first_build_s_cold_arena": t_build_first, "e2e": ins_e2e,
                        "kernel_ms": build_prof, "kernel_ms_sum": sum(v["ms"] for v in build_prof.values()) if build_prof else None,
                        "roofline": build_roof, "gpu_launches": launches_ins,
                        "what": "one step = insert_seq of the whole index (500 x 1 Mbp per GPU) into a fresh empty set, reads resident in HBM"}
        contains_block = {"value": value, "ms_per_step": 1e3 * elapsed / args.steps, "e2e": e2e, "nvlink": nvlink}
        cfg = {"workload": "configs[1]: contains_seq of 1 Gbp synthetic FASTA vs 500M-k-mer index, K=25, T=u64, PREFIX_BITS=24"
                           + (" — headline = the index build (insert_seq)" if insert_headline else ""),
               "index_records": len(i_off) - 1, "query_records": len(q_off) - 1, "record_bp": rec, "per_gpu": world > 1,
               "stored_kmers": stored, "buckets": nb, "hit_fraction": hits / max(1, n_q_kmers),
               "l2_policy": f"inputs larger than L2: {query.numel() / 1e6:.0f} MB query + {stored * 4 / 1e6:.0f} MB index per step",
               "parallelism": "1 GPU" if world == 1 else f"prefix-range sharded x{world}, one fused route + probe kernel per GPU over NVLink peer memory"}
        line = {
            "metric": metric_name(2, args.metric), "value": insert_value if insert_headline else value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": (1e3 * ins_elapsed / n_ins_steps) if insert_headline else (1e3 * elapsed / args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": cfg, "roofline": roof, "cpu_baseline": cpu, "e2e": ins_e2e if insert_headline else e2e,
            "gpu_launches": launches_ins if insert_headline else launches, "clocks": clocks_ins if insert_headline else clocks_q,
            "parity_check": parity,
            "extra": {"wall_s_timed_region": wall, "contains_seq": contains_block, "insert_seq": insert_block, "kernel_ms": prof,
                      "timing": "CUDA events on the library's stream, profiling off; per-kernel times from a separate untimed pass"},
        }
        print(json.dumps(line))
    if world > 1:
        cbl.close()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 2:
        run_config2(args)
    else:
        import bench_configs   # configs 1, 3, 4, 5 (scripts kept next to this file)

        bench_configs.run(args)


if __name__ == "__main__":
    main()
