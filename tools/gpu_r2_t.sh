#!/bin/bash
# round 2, call T: chunked evaluation (gg_gravity_chunked), sliced upload; shim A/B on the C-host leg
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|real|FAILED" gpurun_out/pytest_gpu.log | tail -8
for ch in 1 8; do
  echo "GG_SHIM_CHUNKS=$ch"
  GG_SHIM_CHUNKS=$ch timeout 600 python tools/c_host_leg.py --workload plummer:1000000:0.7 --steps 5 --warmup 2 2>/dev/null | tail -1 | cut -c1-330
  GG_SHIM_CHUNKS=$ch timeout 900 python tools/c_host_leg.py --workload periodic:256:0.5 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-330
done
