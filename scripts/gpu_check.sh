#!/bin/bash
# Runs on the GPU box via gpurun: parity tests, smoke, a short bench. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --frames ${FRAMES:-512} --steps 5 --warmup 3 > gpurun_out/bench_small.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/bench_small.log
tail -5 gpurun_out/bench_small.log
