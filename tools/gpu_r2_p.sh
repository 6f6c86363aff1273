#!/bin/bash
# round 2, call P: no host round trip between walk and evaluation (k_guard) -- whole GPU suite, A/B timing with GG_SYNC_WALK=1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real|FAILED" gpurun_out/pytest_gpu.log | tail -8
( GG_SYNC_WALK=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q 2>&1 | tail -2 )
for s in 1 0; do
  echo "GG_SYNC_WALK=$s"
  GG_SYNC_WALK=$s timeout 300 python tools/quick_perf.py --n 1000000 --reps 4 2>&1 | tail -2 | cut -c1-110
  GG_SYNC_WALK=$s timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 4 2>&1 | tail -2 | cut -c1-110
done
