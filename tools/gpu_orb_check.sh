#!/bin/bash
# GPU check of the ORB services: parity tests, then memcheck over the small cases
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_orb.py -x -q -s > gpurun_out/pytest_gpu_orb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_orb.log
tail -25 gpurun_out/pytest_gpu_orb.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_orb.py -x -q -k "services or reference_domains" > gpurun_out/sanitize_orb.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_orb.log
tail -6 gpurun_out/sanitize_orb.log
