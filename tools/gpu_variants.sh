#!/bin/bash
# usage: gpu_variants.sh NAME...   times every gpurun_variants/NAME.so on the periodic 128^3 box and the 1 M Plummer sphere
for v in "$@"; do echo "== $v"; GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1; GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/quick_perf.py --n 1000000 --reps 3 2>&1 | tail -1; done
