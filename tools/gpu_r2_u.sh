#!/bin/bash
# round 2, call U: the driver's sequence on one GPU at HEAD -- reference arm, default bench (all legs), launch list + ncu full of k_eval on configs[3]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref_n1.json 2> gpurun_out/bench_r2_ref_n1.err ) 2> gpurun_out/bench_r2_ref_n1.time; tail -3 gpurun_out/bench_r2_ref_n1.time | head -1
( time timeout 1200 python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err ) 2> gpurun_out/bench_r2_final_n1.time; tail -3 gpurun_out/bench_r2_final_n1.time | head -1
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final_n1.json'))
b=d['roofline']['step_breakdown_ms']; p=d.get('parity',{})
print('C4 value %.3e step %.1f walk %.1f scat %.1f eval %.1f ewald %.1f frac %.3f e2e %.1f | acc rms %.2e pot rms %.2e ok %s' % (d['value'], d['ms_per_step'], b['k_walk'], b['scan+k_scatter'], b['k_eval'], b['k_ewald'], d['roofline']['frac'], d['e2e']['ms_per_step'], p.get('acc_rel_rms',0), p.get('pot_rel_rms',0), p.get('ok')))
c=d['configs1']; print('C2 step %.2f frac %.3f e2e %.2f' % (c['ms_per_step'], c['roofline']['frac'], c['e2e']['ms_per_step']))
print('from particles', d['e2e_from_particles']['ms_per_step'], 'kdk', d['e2e_from_particles']['resident_kdk_step']['ms_per_step'])
print('pkdGravAll', {k:v for k,v in d['e2e_pkdGravAll'].items() if k!='what'})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
r=json.load(open('gpurun_out/bench_r2_ref_n1.json')); print('ref', r['value'], r.get('ms_per_step'), r['cpu_baseline']['sample'][:80])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-parity > gpurun_out/bench_ncu_r02.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -f -o gpurun_out/prof_r02_k_eval_c4 python tools/quick_perf.py --workload periodic --n 256 --theta 0.5 --reps 2 > gpurun_out/ncu_r02_eval_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ewald -s 1 -c 1 -f -o gpurun_out/prof_r02_k_ewald_c3 python tools/quick_perf.py --workload periodic --n 128 --reps 2 > gpurun_out/ncu_r02_ewald_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
