for V in "" mg8 mg6; do
  if [ -n "$V" ]; then export CBL_GPU_LIB=$PWD/cbl_b200/csrc/libcbl_gpu_var_$V.so; else unset CBL_GPU_LIB; fi
  echo "=== variant [${V:-default mg12}]"
  python bench.py --config 4 --index-mbp 100 --steps 3 --warmup 1 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4 value %.3e ms/step %.3f' % (d['value'], d['ms_per_step'])); print({k:round(v['ms'],3) for k,v in d['roofline']['per_op'].items()}, {k:round(v['frac_of_hbm_peak'],3) for k,v in d['roofline']['per_op'].items()})"
  python bench.py --config 3 --index-mbp 100 --steps 3 --warmup 1 --no-cpu-baseline --no-parity --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 value %.3e ms/step %.3f' % (d['value'], d['ms_per_step'])); print({k.split('<')[0]:round(v['ms'],3) for k,v in d['extra']['kernel_ms'].items() if v['ms']>0.1})"
done
unset CBL_GPU_LIB
bash scripts/gpu_profile.sh r02b --no-list "seq_words_kernel<unsigned long, unsigned int, \(int\)1" "seq_words_kernel<unsigned long, unsigned int, \(int\)0, \(bool\)0" > gpurun_out/profile_r02b.log 2>&1; cat gpurun_out/profile_r02b.log | head -5
