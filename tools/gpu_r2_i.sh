#!/bin/bash
# round 2, call I (8 GPUs): bench without the extra legs (sampler fix check)
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/bench_r2_n8b.json 2> gpurun_out/bench_r2_n8b.err ) 2> gpurun_out/bench_r2_n8b.time
tail -3 gpurun_out/bench_r2_n8b.time; tail -3 gpurun_out/bench_r2_n8b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n8b.json'))
print(d['value'], d['ms_per_step'], d['rank_balance'], d['e2e']['ms_per_step'], d['roofline']['step_breakdown_ms'], d['clocks'])
PY
