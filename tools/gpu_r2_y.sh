#!/bin/bash
# round 2, call Y: whole GPU suite, smoke and the default bench line at HEAD
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real|FAILED" gpurun_out/pytest_gpu.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
( time timeout 1200 python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err ) 2> gpurun_out/bench_r2_final_n1.time; tail -3 gpurun_out/bench_r2_final_n1.time | head -1
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final_n1.json'))
b=d['roofline']['step_breakdown_ms']; p=d.get('parity',{})
print('C4 value %.3e step %.1f walk %.1f scat %.1f eval %.1f ewald %.1f frac %.3f traffic %s e2e %.1f | acc rms %.2e pot rms %.2e ok %s' % (d['value'], d['ms_per_step'], b['k_walk'], b['scan+k_scatter'], b['k_eval'], b['k_ewald'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['ms_per_step'], p.get('acc_rel_rms',0), p.get('pot_rel_rms',0), p.get('ok')))
c=d['configs1']; print('C2 step %.2f frac %.3f e2e %.2f' % (c['ms_per_step'], c['roofline']['frac'], c['e2e']['ms_per_step']))
print('from particles', d['e2e_from_particles']['ms_per_step'], 'kdk', d['e2e_from_particles']['resident_kdk_step']['ms_per_step'])
print('pkdGravAll', {k:v for k,v in d['e2e_pkdGravAll'].items() if k!='what'})
PY
