#!/bin/bash
# small-scale smoke of bench.py --config 1/3/4/5 (and the reference arm) on one GPU
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" > gpurun_out/cfg.json 2> gpurun_out/cfg.err; echo "rc $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/cfg.json").read().strip().splitlines()[-1])
    print("value %.4e %s  ms/step %.3f" % (d["value"], d["unit"], d["ms_per_step"]))
    print(" e2e", json.dumps(d.get("e2e"))[:300]); print(" parity", json.dumps(d.get("parity_check"))[:400]); print(" roof", json.dumps(d.get("roofline"))[:500]); print(" cpu", json.dumps(d.get("cpu_baseline"))[:300])
except Exception as e:
    print("no json line:", e)
PY
tail -4 gpurun_out/cfg.err; }
run python bench.py --config 1 --steps 3 --warmup 2
run python bench.py --config 3 --index-mbp 100 --steps 2 --warmup 1 --no-cpu-baseline
run python bench.py --config 4 --index-mbp 50 --steps 2 --warmup 1
CBL_STREAM_BATCHES=30 run python bench.py --config 5 --index-mbp 100 --steps 1 --warmup 1
run python bench.py --impl reference --steps 2 --warmup 1 --cpu-index-mbp 20
run python bench.py --impl reference --metric insert_seq --steps 2 --warmup 1
