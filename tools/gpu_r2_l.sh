#!/bin/bash
# round 2, call L: source-per-lane cell evaluation (GG_EVAL_T) -- variants: parity vs the oracle + Plummer 1M timing, periodic 128^3 timing
mkdir -p gpurun_out
for v in base b4 t5 t4 t4rb4; do
  GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/variant_check.py 2>&1 | grep "^\[" | sed "s#$PWD/gpurun_variants/##" 
  GASOLINE_B200_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1
done 2>&1 | tee gpurun_out/variants_t.log
