#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).
# Usage: bash scripts/gpu_profile.sh <tag> [--no-list] [demangled-kernel-regex ...]
set -u
TAG=${1:-r01}; shift || true
LIST=1
if [ "${1:-}" = "--no-list" ]; then LIST=0; shift; fi
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity ${PROFILE_BENCH_ARGS:-}"
if [ $LIST = 1 ]; then
  # (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/launches_${TAG}.out 2>&1
  echo "launch list exit $?"
fi
# (2) full capture of the named kernels (third matching launch of each)
i=0
for K in "$@"; do
  i=$((i+1))
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_$i $B > gpurun_out/prof_${TAG}_$i.out 2>&1
  echo "full capture [$i] $K exit $?"
done
ls -la gpurun_out | tail -20
