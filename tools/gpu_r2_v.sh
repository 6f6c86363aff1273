#!/bin/bash
# round 2, call V: bisection of the domain decomposition on the device (gg_orb_bisect)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_orb.py -m gpu -q -s ) > gpurun_out/pytest_gpu_orb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_orb.log
grep -E "passed|failed|rc=|real|FAILED|ORB 1 M|Error" gpurun_out/pytest_gpu_orb.log | tail -12
