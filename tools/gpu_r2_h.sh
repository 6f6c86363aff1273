#!/bin/bash
# round 2, call H: concurrent Ewald (own low-priority stream) -- parity tests, A/B timing
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real|FAILED" gpurun_out/pytest_gpu.log | tail -6
for f in 1 0; do
  echo "GG_EWALD_ASYNC=$f"
  GG_EWALD_ASYNC=$f timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1
  GG_EWALD_ASYNC=$f timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 2 --parity-buckets 96 > gpurun_out/bench_async_$f.json 2> gpurun_out/bench_async_$f.err
  python - $f <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/bench_async_{sys.argv[1]}.json'))
b=d['roofline']['step_breakdown_ms']; p=d.get('parity',{})
print('C4 async=%s step %.1f walk %.1f scat %.1f eval %.1f ewald %.1f frac %.3f e2e %.1f | acc rms %.2e max %.2e pot rms %.2e max %.2e ok %s' % (sys.argv[1], d['ms_per_step'], b['k_walk'], b['scan+k_scatter'], b['k_eval'], b['k_ewald'], d['roofline']['frac'], d['e2e']['ms_per_step'], p.get('acc_rel_rms',0), p.get('acc_rel_max',0), p.get('pot_rel_rms',0), p.get('pot_rel_max',0), p.get('ok')))
PY
done
