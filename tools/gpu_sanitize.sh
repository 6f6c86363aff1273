#!/bin/bash
# compute-sanitizer passes over the small GPU tests (under gpurun): memcheck on everything small, racecheck on the
# shared-memory tree-build kernels.  Exit code 9 = a finding.
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tree_build.py tests/test_gpu_sun.py \
    tests/test_gpu_state.py tests/test_gpu_multirank.py tests/test_gpu_dropin.py -x -q \
    -k "tiny or duplicates or periodic16 or inside or rungs or bit_exact or sequence or (device_built and periodic) or plummer3000 or jitter_r3" || exit $?
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tree_build.py -x -q \
    -k "tiny or duplicates or periodic16 or plummer20k" || exit $?
# the ORB services (shared-memory slot tables, shared atomics, per-warp partial weights)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_orb.py -x -q -k "services or reference_domains" || exit $?
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_orb.py -x -q -k "services or reference_domains or weights" || exit $?
