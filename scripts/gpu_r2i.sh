#!/bin/bash
# k_points_pair v2: tests, bench, ncu full capture of the pair kernel with source counters
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fast_path.py -m gpu -x -q > gpurun_out/pytest_fast.log 2>&1; echo "pytest fast exit $?"; tail -5 gpurun_out/pytest_fast.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu --no-extra"
i=0
for v in "GV_FAST_KIND=0"; do
  i=$((i+1))
  env $v timeout 600 python bench.py $B > gpurun_out/bench_p$i.log 2>&1; echo "[$v] exit $?"
  tail -1 gpurun_out/bench_p$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d.get('grid_crc'), d.get('parity_sample'))"
done
PB="--frames 1024 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
python bench.py $PB > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base function --kernel-name regex:'^k_points_(pair|col)' -s 1 -c 1 \
    -o gpurun_out/prof_pair -f python bench.py $PB > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof_pair.ncu-rep --page raw --csv > gpurun_out/prof_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_pair.ncu-rep --page source --csv --print-source sass > gpurun_out/src_pair.csv 2>/dev/null
ls -la gpurun_out | head -5
