#!/bin/bash
# round 2, call W (8 GPUs): the driver's sequence at HEAD -- reference arm and the default bench line (all legs) on 8 x B200
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_r2_ref_n8.json 2> gpurun_out/bench_r2_ref_n8.err ) 2> gpurun_out/bench_r2_ref_n8.time; tail -3 gpurun_out/bench_r2_ref_n8.time | head -1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2_final_n8.json 2> gpurun_out/bench_r2_final_n8.err ) 2> gpurun_out/bench_r2_final_n8.time; tail -3 gpurun_out/bench_r2_final_n8.time | head -1
tail -3 gpurun_out/bench_r2_final_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final_n8.json'))
b=d['roofline']['step_breakdown_ms']; p=d.get('parity') or {}
print('N8 value %.3e step %.2f | %s | e2e %.2f %s | frac %.3f | parity %s acc %.2e' % (d['value'], d['ms_per_step'], {k[:22]: round(v,2) for k,v in b.items() if isinstance(v,float)}, d['e2e']['ms_per_step'], {k: round(v,2) for k,v in d['e2e'].get('phases_ms_rank0',{}).items()}, d['roofline']['frac'], p.get('ok'), p.get('acc_rel_rms',0)))
print('from particles', {k:v for k,v in d.get('e2e_from_particles',{}).items() if k in ('ms_per_step','tree_build_device_ms_rank0')})
print('pkdGravAll', {k:v for k,v in d.get('e2e_pkdGravAll',{}).items() if k!='what'})
print('dd', d['domain_decomposition'])
r=json.load(open('gpurun_out/bench_r2_ref_n8.json')); print('ref', r.get('value'), r.get('ms_per_step'), r.get('cpu_baseline',{}).get('cores'))
PY
