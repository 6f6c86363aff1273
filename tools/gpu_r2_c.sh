#!/bin/bash
# round 2, call C (2 GPUs): NCCL transport test + bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
( time timeout 900 python -m pytest tests/test_dist_nccl.py -m gpu -x -q -s ) > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl.log
grep -E "passed|failed|rc=|real|over NCCL" gpurun_out/pytest_nccl.log | tail -12
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err ) 2> gpurun_out/bench_r2_n2.time
tail -3 gpurun_out/bench_r2_n2.time; tail -8 gpurun_out/bench_r2_n2.err; head -c 400 gpurun_out/bench_r2_n2.json
