#!/bin/bash
# ncu full capture of one tree kernel on the bench workload (Plummer 1M).  Under gpurun.  usage: gpu_prof.sh k_eval|k_walk|...
mkdir -p gpurun_out
for k in "$@"; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_${k} -f python tools/quick_perf.py --n 1000000 --reps 2 > gpurun_out/ncu_${k}.log 2>&1
done
ls -la gpurun_out
