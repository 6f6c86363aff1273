#!/bin/bash
# round 2, call G (8 GPUs): the default bench (configs[3], strong scaling) as the driver launches it
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
free -g | head -2 >> gpurun_out/topo_n8.txt; nproc >> gpurun_out/topo_n8.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err ) 2> gpurun_out/bench_r2_n8.time
tail -3 gpurun_out/bench_r2_n8.time; tail -5 gpurun_out/bench_r2_n8.err; head -c 300 gpurun_out/bench_r2_n8.json
