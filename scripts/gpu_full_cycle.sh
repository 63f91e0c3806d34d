#!/bin/bash
# fast-path tests, A/B of env variants, then an ncu --set full capture of the pair kernel (default variant)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fast_path.py -m gpu -x -q > gpurun_out/pytest_fast.log 2>&1; echo "pytest fast exit $?"; tail -5 gpurun_out/pytest_fast.log
bash scripts/gpu_ab.sh "$@"
bash scripts/gpu_prof_pair.sh
