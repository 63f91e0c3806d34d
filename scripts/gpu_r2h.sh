#!/bin/bash
# k_points_pair: correctness first (fast-path + parity suites), then A/B against k_points_col
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fast_path.py -m gpu -x -q > gpurun_out/pytest_fast.log 2>&1; echo "pytest fast exit $?"; tail -15 gpurun_out/pytest_fast.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu --no-extra"
i=0
for v in "GV_FAST_KIND=0" "GV_FAST_KIND=3"; do
  i=$((i+1))
  env $v timeout 600 python bench.py $B > gpurun_out/bench_p$i.log 2>&1; echo "[$v] exit $?"
  tail -1 gpurun_out/bench_p$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d.get('grid_crc'), d.get('parity_sample'))"
done
