#!/bin/bash
# round 2, call S: sliced local upload (gg_local_*), shim pipelining, Ewald rsqrt -- tests + the C-host leg A/B
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_announce.py tests/test_gpu_dropin.py tests/test_gpu_multirank_host.py tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -s ) > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s.log
grep -E "passed|failed|rc=|real|FAILED|128\^3:|pstGravity" gpurun_out/pytest_gpu_s.log | tail -12
for sl in 0 1; do
  echo "GG_SHIM_SLICED=$sl"
  GG_SHIM_SLICED=$sl timeout 600 python tools/c_host_leg.py --workload plummer:1000000:0.7 --steps 5 --warmup 2 2>/dev/null | tail -1 | cut -c1-700
  GG_SHIM_SLICED=$sl timeout 900 python tools/c_host_leg.py --workload periodic:256:0.5 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-700
done
timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1 | cut -c1-120
