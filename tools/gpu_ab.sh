mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_multirank.py -x -q 2>&1 | tail -2
timeout 300 python tools/quick_perf.py --n 1000000 --reps 4 2>&1 | tail -1
timeout 300 python tools/quick_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -1
