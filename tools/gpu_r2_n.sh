#!/bin/bash
# round 2, call N: gather variants (pieces of a moment record per lane) -- parity + Plummer 1M, periodic 128^3, configs[3]
mkdir -p gpurun_out
for v in g1 g2 g4; do
  GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/variant_check.py 2>&1 | grep "^\[" | sed "s#$PWD/gpurun_variants/##" | tail -3
  GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1 | cut -c1-120
  GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 2 --parity-buckets 96 > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python - $v <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/bench_var_{sys.argv[1]}.json'))
b=d['roofline']['step_breakdown_ms']; p=d.get('parity',{})
print(sys.argv[1], 'step %.1f walk %.1f scat %.1f eval %.1f ewald %.1f frac %.3f | acc rms %.2e max %.2e pot rms %.2e max %.2e ok %s' % (d['ms_per_step'], b['k_walk'], b['scan+k_scatter'], b['k_eval'], b['k_ewald'], d['roofline']['frac'], p.get('acc_rel_rms',0), p.get('acc_rel_max',0), p.get('pot_rel_rms',0), p.get('pot_rel_max',0), p.get('ok')))
PY
done 2>&1 | tee gpurun_out/variants_g.log
