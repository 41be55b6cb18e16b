#!/bin/bash
# per-rank trace of the fused sharded query under different env knobs:  bash scripts/gpu_multi_trace.sh N "ENV=.." "ENV=.." ...
N=$1; shift
mkdir -p gpurun_out
i=0
for E in "$@"; do
  i=$((i+1))
  echo "=== [$i] N=$N $E"
  env CBL_SHARD_TRACE=1 $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + i)) bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-build-profile >gpurun_out/trace_$i.out 2>gpurun_out/trace_$i.err
  python - gpurun_out/trace_$i.out gpurun_out/trace_$i.err <<'PY'
import sys, json, re
out, err = open(sys.argv[1]).read().splitlines(), open(sys.argv[2]).read().splitlines()
for l in out:
    if l.startswith('{'):
        d = json.loads(l); x = d['extra']
        print('contains %.3e k-mers/s %.3f ms/step | insert %.3f ms | kernels %s' % (x['contains_seq']['value'], x['contains_seq']['ms_per_step'], x['insert_seq']['ms_per_step'], {k.split('<')[0]: round(v['ms'] / v['n'], 2) for k, v in x['kernel_ms'].items()}))
fq = [l for l in err + out if l.startswith('[fused query]')]
st = [l for l in err + out if l.startswith('[shard trace] rank')]
n = int(sys.argv[1].split('_')[-1].split('.')[0]) if False else None
last = {}
for l in fq:
    m = re.search(r'device (\d+): production of (\d+) words complete ([\d.]+) ms .* done ([\d.]+) ms', l)
    if m: last[int(m.group(1))] = (int(m.group(2)), float(m.group(3)), float(m.group(4)))
recv = {}
for l in st:
    m = re.search(r'rank (\d+): sent (\d+) words, received (\d+)', l)
    if m: recv[int(m.group(1))] = int(m.group(3))
for d in sorted(last):
    print('  device %d: produced %d, production done at %.2f ms, kernel %.2f ms, received %s' % (d, last[d][0], last[d][1], last[d][2], recv.get(d)))
PY
  grep -v "^W\|^$\|OMP\|\*\*\*\|^\[" gpurun_out/trace_$i.err | tail -3
done
