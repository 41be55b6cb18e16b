#!/bin/bash
# full ncu capture of the fused probe kernel (MODE 1) and the words kernel (MODE 0); run under gpurun
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-build-profile"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:seq_words_kernel<unsigned long, unsigned int, \(int\)1' -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_probe $B > gpurun_out/prof_${TAG}_probe.out 2>&1
echo "probe capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:seq_words_kernel<unsigned long, unsigned int, \(int\)0' -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_words $B > gpurun_out/prof_${TAG}_words.out 2>&1
echo "words capture exit $?"
ls -la gpurun_out | tail
