#!/bin/bash
# DRAM traffic / L2 hit rate of one kernel under different env knobs (run under gpurun, 1 GPU):
#   bash scripts/gpu_ncu_metrics.sh '<kernel regex>' "ENV1=.. ENV2=.." "ENV=.." ...
set -u
K=$1; shift
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-build-profile"
i=0
for E in "$@"; do
  i=$((i+1))
  echo "=== [$i] $E"
  env $E timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum \
    --clock-control none --kernel-name-base demangled -k "regex:$K" -s 1 -c 1 --csv $B 2>/dev/null | grep -E '"(dram__|lts__|gpu__|smsp__)' | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
