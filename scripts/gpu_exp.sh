#!/bin/bash
# quick A/B runs of the bench under different env knobs:  bash scripts/gpu_exp.sh "ENV1=.. ENV2=.." "ENV=.." ...
mkdir -p gpurun_out
i=0
for E in "$@"; do
  i=$((i+1))
  echo "=== [$i] $E"
  env $E timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-parity ${BENCH_ARGS:-} 2>gpurun_out/exp_$i.err | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); x=d['extra']; print('contains %.3e k-mers/s %.3f ms | insert %.3e k-mers/s %.3f ms (wall %.3f) first %.4f s' % (x['contains_seq']['value'], x['contains_seq']['ms_per_step'], x['insert_seq']['value'], x['insert_seq']['ms_per_step'], x['insert_seq']['wall_ms_per_step'], x['insert_seq']['first_build_s_cold_arena']))
        bk=x['insert_seq'].get('kernel_ms') or {}; print('build kernels: '+', '.join('%s x%d %.2f' % (k.split('<')[0], v['n'], v['ms']) for k, v in sorted(bk.items(), key=lambda kv: -kv[1]['ms'])) + ' | sum %.1f ms' % sum(v['ms'] for v in bk.values()))
        print('query kernels: '+', '.join('%s x%d %.2f' % (k.split('<')[0], v['n'], v['ms']) for k, v in sorted(x['kernel_ms'].items(), key=lambda kv: -kv[1]['ms'])))
    else: print(l)
"
  tail -3 gpurun_out/exp_$i.err
done
