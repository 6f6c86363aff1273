mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_device_moments.py tests/test_gpu_delivery.py tests/test_gpu_multirank.py -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_fullsize.py -k "c2 and sampled" -x -q -s 2>&1 | grep -E "acc rms|passed|failed"
echo NEW; timeout 300 python tools/quick_perf.py --n 1000000 --reps 3 2>&1 | tail -2
echo OLD; GG_EVAL_OLD=1 timeout 300 python tools/quick_perf.py --n 1000000 --reps 3 2>&1 | tail -2
