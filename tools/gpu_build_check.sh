mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tree_build.py -x -q 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_fullsize.py -k device_tree -x -q -s 2>&1 | tail -5
timeout 300 python tools/build_perf.py --reps 3 2>&1 | tail -3
timeout 300 python tools/build_perf.py --workload periodic --n 128 --reps 3 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_build.csv python tools/build_perf.py --reps 1 --gravity 0 > gpurun_out/build_ncu.log 2>&1
