#!/bin/bash
# round 2, call AB (--gpus 2): the collective bisection -- thread ranks on one GPU (test), NCCL on two GPUs (script)
mkdir -p gpurun_out
true
true
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/dd_collective_check.py 2000000 > gpurun_out/dd_collective_n4.json 2> gpurun_out/dd_collective_n4.err; echo "dd rc=$?"
tail -c 900 gpurun_out/dd_collective_n4.json; tail -5 gpurun_out/dd_collective_n4.err | cut -c1-300
