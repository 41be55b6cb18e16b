#!/bin/bash
# quick A/B runs of the bench under different env knobs:  bash scripts/gpu_exp.sh "ENV1=.. ENV2=.." "ENV=.." ...
mkdir -p gpurun_out
i=0
for E in "$@"; do
  i=$((i+1))
  echo "=== [$i] $E"
  env $E timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e k-mers/s  ms/step %.2f  insert %.3e  build_s %.3f' % (d['value'], d['ms_per_step'], d['extra']['insert_seq_kmers_per_s'], d['extra']['build_s'])); print('warm build_s', d['extra'].get('build_s_warm_pool')); print(json.dumps(d['extra']['kernel_ms'])); bk=d['extra'].get('build_kernel_ms') or {}; print('build kernels: '+', '.join('%s x%d %.2f' % (k.split('<')[0], v['n'], v['ms']) for k, v in sorted(bk.items(), key=lambda kv: -kv[1]['ms'])) + ' | sum %.1f ms' % sum(v['ms'] for v in bk.values()))
    else: print(l)
"
done
