#!/bin/bash
# usage: gpu_variants.sh [-p] NAME...   (-p: periodic 128^3 only)
for v in "$@"; do echo "== $v"; GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1; done
