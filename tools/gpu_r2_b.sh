#!/bin/bash
# round 2, call B: bench (both arms) on the default workload, 1 GPU
mkdir -p gpurun_out
( time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err ) 2> gpurun_out/bench_r2_n1.time
tail -3 gpurun_out/bench_r2_n1.time; tail -5 gpurun_out/bench_r2_n1.err; head -c 600 gpurun_out/bench_r2_n1.json
