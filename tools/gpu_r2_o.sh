#!/bin/bash
# round 2, call O: gg_announce (early Ewald beside the upload) -- its tests, the suites that share code with it, bench on configs[3] and [2]
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_announce.py tests/test_gpu_delivery.py tests/test_gpu_parity.py tests/test_gpu_sun.py tests/test_gpu_multirank.py tests/test_gpu_multirank_host.py tests/test_gpu_dropin.py -m gpu -q -s ) > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
grep -E "passed|failed|rc=|real|FAILED|128\^3:|sun with" gpurun_out/pytest_gpu_o.log | tail -12
timeout 900 python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/bench_o_c4.json 2> gpurun_out/bench_o_c4.err
timeout 900 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 --workload periodic:128:0.7 > gpurun_out/bench_o_c3.json 2> gpurun_out/bench_o_c3.err
python - <<'PY'
import json
for f in ('bench_o_c4','bench_o_c3'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    b=d['roofline']['step_breakdown_ms']; p=d.get('parity',{})
    print(f, 'step %.1f walk %.1f scat %.1f eval %.1f ewald %.1f frac %.3f e2e %.1f | acc rms %.2e pot rms %.2e ok %s' % (d['ms_per_step'], b['k_walk'], b['scan+k_scatter'], b['k_eval'], b['k_ewald'], d['roofline']['frac'], d['e2e']['ms_per_step'], p.get('acc_rel_rms',0), p.get('pot_rel_rms',0), p.get('ok')))
PY
