#!/bin/bash
# full ncu capture (with source) of one kernel of the bench; run under gpurun:  bash scripts/gpu_prof_kernel.sh <tag> '<regex>' [skip] [ENV=..]
set -u
TAG=$1; K=$2; SKIP=${3:-0}; shift 3 || true
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
env "$@" timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG} $B > gpurun_out/prof_${TAG}.out 2>&1
echo "capture exit $?"; tail -3 gpurun_out/prof_${TAG}.out
