#!/bin/bash
# round 2, call AC (final state of round 2): opening criteria + nReplicas 4/5 -- whole GPU suite, k_walk A/B timing, smoke, default bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_ac.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ac.log
grep -E "passed|failed|rc=|real|FAILED|Error" gpurun_out/pytest_gpu_ac.log | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ac.log 2>&1; echo "smoke rc=$?"
( time timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ac.json 2> gpurun_out/bench_ac.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ac.json'))
b=d['roofline']['step_breakdown_ms']; p=d.get('parity',{})
print('C4 step %.1f walk %.1f scat %.1f eval %.1f ewald %.1f e2e %.1f frac %.3f | acc rms %.2e ok %s | c1 %.2f ms | shim %.1f ms' % (d['ms_per_step'], b['k_walk'], b['scan+k_scatter'], b['k_eval'], b['k_ewald'], d['e2e']['ms_per_step'], d['roofline']['frac'], p.get('acc_rel_rms',0), p.get('ok'), d['configs1']['ms_per_step'], d['e2e_pkdGravAll']['ms_per_step']))
PY
