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
        d=json.loads(l); print('value %.3e k-mers/s  ms/step %.2f  insert %.3e  build_s %.3f' % (d['value'], d['ms_per_step'], d['extra']['insert_seq_kmers_per_s'], d['extra']['build_s'])); print(json.dumps(d['extra']['kernel_ms']))
    else: print(l)
"
done
