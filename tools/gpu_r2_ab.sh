#!/bin/bash
# round 2, call AB (--gpus 2): the collective bisection -- thread ranks on one GPU (test), NCCL on two GPUs (script)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_orb.py tests/test_gpu_comm.py -m gpu -x -q ) > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ab.log
grep -E "passed|failed|rc=|real|FAILED|Error" gpurun_out/pytest_gpu_ab.log | tail -8
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dd_collective_check.py 2000000 > gpurun_out/dd_collective_n2.json 2> gpurun_out/dd_collective_n2.err; echo "dd rc=$?"
tail -c 900 gpurun_out/dd_collective_n2.json; tail -5 gpurun_out/dd_collective_n2.err | cut -c1-300
