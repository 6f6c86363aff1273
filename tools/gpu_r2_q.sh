#!/bin/bash
# round 2, call Q (8 GPUs): configs[3] (256^3 strong scaling), configs[2] (128^3), configs[4] (512^3) on 8 x B200
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1; nproc >> gpurun_out/topo_n8.txt; free -g >> gpurun_out/topo_n8.txt
run() { # name port args...
  name=$1; port=$2; shift 2
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err ) 2> gpurun_out/$name.time
  tail -3 gpurun_out/$name.time | head -1; tail -2 gpurun_out/$name.err | cut -c1-300
  python - $name <<'PY'
import json, sys
try:
    d=json.load(open(f'gpurun_out/{sys.argv[1]}.json'))
    b=d['roofline']['step_breakdown_ms']; p=d.get('parity') or {}
    print(sys.argv[1], 'value %.3e step %.2f ms | %s | e2e %.2f ms %s | frac %.3f | parity %s acc %.2e | clocks %s' % (d['value'], d['ms_per_step'], {k[:22]: round(v,2) for k,v in b.items() if isinstance(v,float)}, d['e2e']['ms_per_step'], {k: round(v,2) for k,v in d['e2e'].get('phases_ms_rank0',{}).items()}, d['roofline']['frac'], p.get('ok'), p.get('acc_rel_rms',0), d['clocks']))
except Exception as e:
    print(sys.argv[1], 'no line:', e)
PY
}
run bench_q_c4_n8 29541 --steps 10 --warmup 3 --no-extra --no-cpu-baseline
run bench_q_c3_n8 29542 --steps 10 --warmup 3 --no-extra --no-cpu-baseline --workload periodic:128:0.7
run bench_q_c5_n8 29543 --steps 3 --warmup 2 --no-extra --no-cpu-baseline --workload periodic:512:0.7
