#!/bin/bash
# quick GPU check: parity tests + device timing of the two standard workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "acc rms|passed|failed|Error|error|assert" | head -40
timeout 300 python tools/quick_perf.py --n 1000000 --reps 4 2>&1 | tail -3
timeout 300 python tools/quick_perf.py --workload periodic --n 64 --reps 3 2>&1 | tail -2
