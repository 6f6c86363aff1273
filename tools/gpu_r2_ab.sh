#!/bin/bash
# round 2, calls AB (--gpus 2, then --gpus 4): the collective bisection (gg_orb_bisect_all) -- thread ranks on one GPU (test),
# NCCL on N GPUs (tools/dd_collective_check.py: same domains as the host path, decomposition time of both)
N=${1:-2}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_orb.py -m gpu -x -q -k collective ) > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ab.log
grep -E "passed|failed|rc=|real|FAILED|Error" gpurun_out/pytest_gpu_ab.log | tail -8
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dd_collective_check.py 2000000 > gpurun_out/dd_collective_n$N.json 2> gpurun_out/dd_collective_n$N.err; echo "dd rc=$?"
tail -c 900 gpurun_out/dd_collective_n$N.json; tail -5 gpurun_out/dd_collective_n$N.err | cut -c1-300
