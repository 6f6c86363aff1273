#!/bin/bash
# round 2, call D: full GPU suite (no -x), quick timings, short C4 bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real|FAILED" gpurun_out/pytest_gpu.log | tail -12
timeout 300 python tools/quick_perf.py --n 1000000 --reps 3 2>&1 | tail -1
timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_quick_c4.json 2> gpurun_out/bench_quick_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick_c4.json'))
print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_breakdown_ms'], d.get('parity',{}).get('ok'), d.get('parity',{}).get('pot_rel_rms'), d.get('parity',{}).get('acc_rel_rms'))
PY
