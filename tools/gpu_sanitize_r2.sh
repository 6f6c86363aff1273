#!/bin/bash
# compute-sanitizer over the code paths added in round 2 (under gpurun): memcheck on the announced / sliced upload with the
# Ewald correction on the side stream, the chunked evaluation, the walk -> evaluation guard, the library exchange, the
# device-resident bisection; racecheck on the bisection's shared-memory kernels and on k_eval / k_walk (31-cell FP64 blocks).
# Exit code 9 = a finding.
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_announce.py tests/test_gpu_delivery.py \
    tests/test_gpu_comm.py tests/test_gpu_multirank_host.py -x -q -k "not slices and not full" > gpurun_out/sanitize_r2_memcheck.log 2>&1 || { tail -30 gpurun_out/sanitize_r2_memcheck.log; exit 9; }
tail -3 gpurun_out/sanitize_r2_memcheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_orb.py -x -q -k "bisection_on_the_device_equals" > gpurun_out/sanitize_r2_orb.log 2>&1 || { tail -30 gpurun_out/sanitize_r2_orb.log; exit 9; }
tail -3 gpurun_out/sanitize_r2_orb.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_orb.py tests/test_gpu_parity.py -x -q -k "bisection_on_the_device_equals or periodic10_order4_ewald3 or seams" > gpurun_out/sanitize_r2_racecheck.log 2>&1 || { tail -30 gpurun_out/sanitize_r2_racecheck.log; exit 9; }
tail -3 gpurun_out/sanitize_r2_racecheck.log
