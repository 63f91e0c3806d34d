#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_2.log 2>&1; echo "pytest multi exit $?"; tail -3 gpurun_out/pytest_multi_2.log
timeout 600 python bench.py --no-cpu --no-e2e --no-extra > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | grep -o '"ms_per_step": [0-9.]*, \|"phases_ms": {[^}]*}'
N=2 bash scripts/gpu_multi_timing.sh
