#!/bin/bash
# A/B of developer variant libraries (scripts/build_variant.sh) on the bench workload:  bash scripts/gpu_variants.sh name1 name2 ...
mkdir -p gpurun_out
for V in "$@"; do
  echo "=== variant $V"
  CBL_GPU_LIB=$PWD/cbl_b200/csrc/libcbl_gpu_var_$V.so timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity ${BENCH_ARGS:-} 2>gpurun_out/var_$V.err | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); x=d['extra']; print('contains %.3f ms/step | insert %.3f ms | query kernels %s' % (x['contains_seq']['ms_per_step'], x['insert_seq']['ms_per_step'], {k.split('<')[0]: round(v['ms']/v['n'],3) for k,v in x['kernel_ms'].items()}))
    else: print(l)
"
  tail -3 gpurun_out/var_$V.err
done
