#!/bin/bash
# ncu full captures of the two tree kernels on the bench workload (Plummer 1M) + Ewald on periodic 64^3.  Under gpurun.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o gpurun_out/prof_eval python tools/quick_perf.py --n 1000000 --reps 2 > gpurun_out/ncu_eval.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 1 -c 1 -o gpurun_out/prof_walk python tools/quick_perf.py --n 1000000 --reps 2 > gpurun_out/ncu_walk.log 2>&1
timeout 300 python tools/quick_perf.py --n 1000000 --reps 4 > gpurun_out/perf_plummer1m.log 2>&1
ls -la gpurun_out
