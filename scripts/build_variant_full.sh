#!/bin/bash
# developer A/B build of the FULL library (all four instantiations) with extra -D flags, out of tree:
#   bash scripts/build_variant_full.sh <name> "<-Dflags>"  ->  cbl_b200/csrc/libcbl_gpu_var_<name>.so  (select with CBL_GPU_LIB=...)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
T=/tmp/cbl_var_$1
rm -rf $T && mkdir -p $T/cbl_b200 $T/include && cp -r $ROOT/cbl_b200/csrc $T/cbl_b200/ && cp $ROOT/include/*.h $T/include/
cd $T/cbl_b200/csrc && rm -f *.o *.so
make NVFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr $2" > /dev/null 2>&1 || { tail -20 build.log; exit 1; }
cp libcbl_gpu.so $ROOT/cbl_b200/csrc/libcbl_gpu_var_$1.so
echo built $1
