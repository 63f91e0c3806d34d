#!/bin/bash
# whole GPU suite, then the C3 bench (quick) under the given env variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -5 gpurun_out/pytest_gpu.log
bash scripts/gpu_ab.sh "$@"
