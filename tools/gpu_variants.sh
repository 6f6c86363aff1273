#!/bin/bash
# usage: gpu_variants.sh SCRIPT-ARGS -- NAME...   runs tools/$TOOL with each variant library
for v in "$@"; do echo "== $v"; GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/build_perf.py --reps 3 --gravity 0 2>&1 | tail -1; GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/build_perf.py --workload periodic --n 128 --reps 3 --gravity 0 2>&1 | tail -1; done
