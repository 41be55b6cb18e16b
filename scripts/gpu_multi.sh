#!/bin/bash
# multi-GPU check of the sharded bench (run under `gpurun --gpus N`):  bash scripts/gpu_multi.sh N [steps]
set -u
N=${1:-2}; STEPS=${2:-3}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"
tail -c 2500 gpurun_out/bench_n$N.json; tail -n 15 gpurun_out/bench_n$N.err
