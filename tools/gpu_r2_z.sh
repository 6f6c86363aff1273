#!/bin/bash
# round 2, call Z: k_ewald with the factorised exponential -- parity suites, A/B timing
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_random.py tests/test_gpu_announce.py tests/test_gpu_fullsize.py tests/test_gpu_multirank.py -m gpu -q ) > gpurun_out/pytest_gpu_z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_z.log
grep -E "passed|failed|rc=|real|FAILED" gpurun_out/pytest_gpu_z.log | tail -6
for v in ex0 product; do
  if [ $v = product ]; then unset GASOLINE_B200_LIB; else export GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so; fi
  echo $v
  timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1 | cut -c1-140
  timeout 300 python tools/quick_perf.py --workload periodic --n 256 --theta 0.5 --reps 2 2>&1 | tail -1 | cut -c1-140
done
unset GASOLINE_B200_LIB
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 2 --parity-buckets 96 > gpurun_out/bench_z_c4.json 2> gpurun_out/bench_z_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_z_c4.json'))
b=d['roofline']['step_breakdown_ms']; p=d.get('parity',{})
print('C4 step %.1f walk %.1f scat %.1f eval %.1f ewald %.1f e2e %.1f | acc rms %.2e max %.2e pot rms %.2e max %.2e ok %s' % (d['ms_per_step'], b['k_walk'], b['scan+k_scatter'], b['k_eval'], b['k_ewald'], d['e2e']['ms_per_step'], p.get('acc_rel_rms',0), p.get('acc_rel_max',0), p.get('pot_rel_rms',0), p.get('pot_rel_max',0), p.get('ok')))
PY
