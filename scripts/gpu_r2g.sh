#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
B="--steps 10 --warmup 3 --no-e2e --no-cpu --no-extra"
i=0
for v in "GV_X=0" "GV_OVERLAP=0"; do
  i=$((i+1))
  env $v timeout 600 python bench.py $B > gpurun_out/bench_v$i.log 2>&1; echo "[$v] exit $?"
  tail -1 gpurun_out/bench_v$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d.get('grid_crc'), d.get('parity_sample'))"
done
PB="--frames 1024 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
export GV_OVERLAP=0
python bench.py $PB > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base function --kernel-name regex:'^k_points_(tma|fast|col)' -s 1 -c 1 \
    -o gpurun_out/prof_points -f python bench.py $PB > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
