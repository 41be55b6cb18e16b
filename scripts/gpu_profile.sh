#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).  Usage: bash scripts/gpu_profile.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r01}; shift || true
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/launches_${TAG}.out 2>&1
echo "launch list exit $?"
# (2) full capture of the named kernels (one launch each)
for K in "$@"; do
  NAME=$(echo "$K" | tr -c 'a-zA-Z0-9_' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_${NAME} $B > gpurun_out/prof_${TAG}_${NAME}.out 2>&1
  echo "full capture $K exit $?"
done
ls -la gpurun_out | tail -20
