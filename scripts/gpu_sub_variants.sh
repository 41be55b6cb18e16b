#!/bin/bash
# A/B of correction-byte scaling variants: shard probe at 8-GPU density + the single-GPU bench kernels
for V in "$@"; do
  echo "== $V"
  L=$PWD/cbl_b200/csrc/libcbl_gpu_var_$V.so
  CBL_GPU_LIB=$L timeout 400 python scripts/exp_shard_probe.py 8 0,4,6,7 2>&1 | grep "^region.*probe" | awk '{printf "%s %s %s ms/G | ", $1, $2, $9} END{print ""}'
  true
done
