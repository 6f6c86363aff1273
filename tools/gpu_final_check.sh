set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -3 gpurun_out/pytest_gpu_final.log; tail -2 gpurun_out/smoke_final.log; head -c 600 gpurun_out/bench_final.json
