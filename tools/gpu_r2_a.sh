#!/bin/bash
# round 2, call A: full GPU test suite + quick timings (baseline for the round)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real" gpurun_out/pytest_gpu.log | tail -5
timeout 300 python tools/quick_perf.py --n 1000000 --reps 4 2>&1 | tail -2
timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -2
