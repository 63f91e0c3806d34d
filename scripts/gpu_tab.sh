#!/bin/bash
# fast-path tests, then A/B of environment variants on the C3 bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fast_path.py -m gpu -x -q > gpurun_out/pytest_fast.log 2>&1; echo "pytest fast exit $?"; tail -5 gpurun_out/pytest_fast.log
bash scripts/gpu_ab.sh "$@"
