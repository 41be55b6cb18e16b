#!/bin/bash
# round-end evidence on ONE GPU: parity tests, smoke, the default bench line, ncu launch list + full captures.
# Usage under gpurun:  bash scripts/gpu_round_final.sh <tag>
set -u
TAG=${1:-r02c}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_n1_$TAG.json
timeout 600 python bench.py --metric insert_seq > gpurun_out/bench_n1_insert_$TAG.json 2> gpurun_out/bench_n1_insert_$TAG.err; echo "bench insert exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench reference exit $?"; tail -c 600 gpurun_out/bench_ref_$TAG.json
bash scripts/gpu_profile.sh $TAG 'seq_words_kernel<unsigned long, unsigned int, \(int\)1' > gpurun_out/profile_$TAG.log 2>&1
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:shard_query -c 1 -f -o gpurun_out/prof_${TAG}_shard_query python scripts/exp_fused_g1.py 8 > gpurun_out/prof_${TAG}_shard_query.out 2>&1; echo "shard_query capture exit $?"; tail -2 gpurun_out/prof_${TAG}_shard_query.out
