#!/bin/bash
# 1-GPU A/B of the fused sharded query kernel (scripts/exp_fused_g1.py):  bash scripts/gpu_fused_g1.sh "VARIANT ENV=.. ENV=.." ...
# VARIANT = name of a scripts/build_variant.sh library, or "-" for the in-tree libcbl_gpu.so
mkdir -p gpurun_out
for E in "$@"; do
  V=${E%% *}; R=${E#* }; [ "$R" = "$E" ] && R=""
  L=""; [ "$V" != "-" ] && L="CBL_GPU_LIB=$PWD/cbl_b200/csrc/libcbl_gpu_var_$V.so"
  echo "=== $E"
  env $L $R timeout 300 python scripts/exp_fused_g1.py ${G1_ARGS:-8} 2>&1 | grep -v "^W\|^$" | tail -3
done
