#!/bin/bash
# One GPU-box session: parity tests, smoke, a reduced bench, optional profiles.  Everything lands in
# gpurun_out/ (scratch).  Usage under gpurun:  bash scripts/gpu_check.sh [tests|bench|full]
set -u
mkdir -p gpurun_out
MODE=${1:-tests}
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
if [ "$MODE" = "tests" ] || [ "$MODE" = "full" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
  tail -n 60 gpurun_out/pytest_gpu.log
fi
if [ "$MODE" = "bench" ] || [ "$MODE" = "full" ]; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  tail -c 3000 gpurun_out/bench.json; tail -n 20 gpurun_out/bench.err
fi
