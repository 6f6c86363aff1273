#!/bin/bash
# launch list + ncu full captures of the tree kernels on the bench workload (Plummer 1M).  Under gpurun.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 14 -c 40 --csv --log-file gpurun_out/launches_quick.csv python tools/quick_perf.py --n 1000000 --reps 3 > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o gpurun_out/prof_eval python tools/quick_perf.py --n 1000000 --reps 2 > gpurun_out/ncu_eval.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 1 -c 1 -o gpurun_out/prof_walk python tools/quick_perf.py --n 1000000 --reps 2 > gpurun_out/ncu_walk.log 2>&1
ls -la gpurun_out
