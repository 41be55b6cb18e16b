#!/bin/bash
# compute-sanitizer over the GPU parity tests (small inputs).  Usage under gpurun:
#   bash scripts/gpu_sanitize.sh <tag> <tool: memcheck|racecheck|initcheck|synccheck> [pytest -k expression]
# Writes gpurun_out/sanitizer_<tag>_<tool>.log (full) and .summary (the lines the judge needs).
set -u
TAG=${1:-r02}; TOOL=${2:-memcheck}; KEXPR=${3:-"insert_contains_remove or set_ops or non_acgt_insert or hybrid_sort"}
mkdir -p gpurun_out
LOG=gpurun_out/sanitizer_${TAG}_${TOOL}.log
timeout 1500 compute-sanitizer --tool $TOOL --target-processes all --print-limit 20 --error-exitcode 66 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$KEXPR" > $LOG 2>&1
RC=$?
{ echo "tool=$TOOL rc=$RC kexpr=$KEXPR"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|Race reported|Uninitialized|at .*\.cuh:|at .*\.cu:" $LOG | head -60; } > gpurun_out/sanitizer_${TAG}_${TOOL}.summary
cat gpurun_out/sanitizer_${TAG}_${TOOL}.summary
