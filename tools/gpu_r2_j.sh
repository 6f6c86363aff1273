#!/bin/bash
# round 2, call J: full GPU suite at HEAD, smoke, round-2 ncu captures (k_eval on C4 and C2, k_walk on C4), launch list of the bench command
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real|FAILED" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -f -o gpurun_out/prof_r02_k_eval_c4 python tools/quick_perf.py --workload periodic --n 256 --theta 0.5 --reps 2 > gpurun_out/ncu_r02_eval_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 1 -c 1 -f -o gpurun_out/prof_r02_k_walk_c4 python tools/quick_perf.py --workload periodic --n 256 --theta 0.5 --reps 2 > gpurun_out/ncu_r02_walk_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -f -o gpurun_out/prof_r02_k_eval_c2 python tools/quick_perf.py --n 1000000 --reps 2 > gpurun_out/ncu_r02_eval_c2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-parity > gpurun_out/bench_ncu_r02.log 2>&1
ls -la gpurun_out | tail -12
