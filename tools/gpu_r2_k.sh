#!/bin/bash
# round 2, call K: LDS/FFMA mix microbenchmark; the bDoSun multi-domain test after its tolerance fix
mkdir -p gpurun_out
./tools/micro/lds_mix > gpurun_out/lds_mix.txt 2>&1; cat gpurun_out/lds_mix.txt
timeout 600 python -m pytest tests/test_gpu_sun.py -m gpu -q -s 2>&1 | tail -12
