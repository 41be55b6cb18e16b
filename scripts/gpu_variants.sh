#!/bin/bash
# A/B of developer variant libraries (scripts/build_variant.sh) on the bench workload:  bash scripts/gpu_variants.sh name1 name2 ...
mkdir -p gpurun_out
for V in "$@"; do
  echo "=== variant $V"
  CBL_GPU_LIB=$PWD/cbl_b200/csrc/libcbl_gpu_var_$V.so timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline ${BENCH_ARGS:-} 2>gpurun_out/var_$V.err | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e k-mers/s  ms/step %.3f  hit_fraction %.7f stored %d build_s %.4f warm %s' % (d['value'], d['ms_per_step'], d['config']['hit_fraction'], d['config']['stored_kmers'], d['extra']['build_s'], d['extra'].get('build_s_warm_pool')))
        bk=d['extra'].get('build_kernel_ms') or {}; print('build kernels: '+', '.join('%s x%d %.2f' % (k.split('<')[0], v['n'], v['ms']) for k, v in sorted(bk.items(), key=lambda kv: -kv[1]['ms'])) + ' | sum %.1f ms' % sum(v['ms'] for v in bk.values()))
    else: print(l)
"
  tail -3 gpurun_out/var_$V.err
done
