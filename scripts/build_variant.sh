#!/bin/bash
# developer A/B builds: one instantiation (u64 words, <= 32-bit suffixes) with extra -D flags
#   bash scripts/build_variant.sh <name> "<-Dflags>"   ->  cbl_b200/csrc/libcbl_gpu_var_<name>.so   (select with CBL_GPU_LIB=...)
set -e
cd "$(dirname "$0")/../cbl_b200/csrc"
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr \
  -DCBL_FAST_BUILD $2 -shared -o libcbl_gpu_var_$1.so cbl_index.cu c_api.cu sharded_index.cu inst_u64_u32.cu -cudart static
