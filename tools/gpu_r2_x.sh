#!/bin/bash
# round 2, call X: k_walk code size -- append / distribute as real functions (instruction-cache misses are its 2nd stall)
mkdir -p gpurun_out
for v in w0 w1 w2; do
  GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/variant_check.py 2>&1 | grep "^\[" | sed "s#$PWD/gpurun_variants/##" | tail -3
  GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1 | cut -c1-120
  GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/quick_perf.py --workload periodic --n 256 --theta 0.5 --reps 2 2>&1 | tail -1 | cut -c1-120
done 2>&1 | tee gpurun_out/variants_w.log
