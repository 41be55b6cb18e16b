#!/bin/bash
# A/B runs of the sharded bench on N GPUs:  bash scripts/gpu_multi_exp.sh N "ENV=.. ENV=.." ...   (BENCH_ARGS for extra flags)
N=$1; shift
mkdir -p gpurun_out
i=0
for E in "$@"; do
  i=$((i+1))
  echo "=== [$i] N=$N $E"
  env $E timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + i)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>gpurun_out/multi_$i.err | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); x=d['extra']; print('contains %.3e k-mers/s %.3f ms/step | e2e %s | insert %.3e %.3f ms | parity %s' % (x['contains_seq']['value'], x['contains_seq']['ms_per_step'], json.dumps(x['contains_seq']['e2e']), x['insert_seq']['value'], x['insert_seq']['ms_per_step'], json.dumps(d.get('parity_check'))[:300]))
        print('query kernels: '+', '.join('%s x%d %.2f' % (k.split('<')[0]+k[k.find(',Suf,')+5:k.find(',Suf,')+6] if ',Suf,' in k else k, v['n'], v['ms']) for k, v in sorted(x['kernel_ms'].items(), key=lambda kv: -kv[1]['ms'])))
    elif 'Error' in l or 'error' in l: print(l)
"
  grep -v "^W\|^$" gpurun_out/multi_$i.err | tail -4
done
