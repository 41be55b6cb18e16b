#!/bin/bash
# A/B of the sharded bench under env knobs (run under gpurun --gpus N):  bash scripts/gpu_multi_exp.sh N "ENV=.. ENV=.." ...
N=$1; shift
mkdir -p gpurun_out
for E in "$@"; do
  echo "=== $E"
  env $E timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e  ms/step %.2f' % (d['value'], d['ms_per_step']), {k.split('<')[0]+k[-24:-20]:(v['n'],round(v['ms']/d['steps'],1)) for k,v in d['extra']['kernel_ms'].items()})
"
done
