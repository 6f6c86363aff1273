#!/bin/bash
# One GPU session: tests, smoke, bench (both arms), ncu launch list of the bench command + full captures of the kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python tools/quick_perf.py --workload periodic --n 64 --theta 0.7 --reps 3 > gpurun_out/perf_periodic64.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
bash tools/gpu_prof.sh k_eval k_walk k_scatter
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ewald -s 0 -c 1 -f -o gpurun_out/prof_k_ewald python tools/quick_perf.py --workload periodic --n 64 --reps 1 > gpurun_out/ncu_ewald.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -f -o gpurun_out/prof_k_eval_mono64 python tools/quick_perf.py --workload periodic --n 128 --reps 2 > gpurun_out/ncu_mono.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_build.csv python tools/build_perf.py --reps 1 --gravity 0 > gpurun_out/build_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_split -s 40 -c 1 -f -o gpurun_out/prof_k_split python tools/build_perf.py --reps 2 --gravity 0 > gpurun_out/ncu_split.log 2>&1
timeout 300 python tools/build_perf.py --reps 3 > gpurun_out/build_perf.log 2>&1
timeout 300 python tools/build_perf.py --workload periodic --n 128 --reps 3 >> gpurun_out/build_perf.log 2>&1
ls -la gpurun_out
