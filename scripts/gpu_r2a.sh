#!/bin/bash
# round 2, call A: parity of the fast path, atomic ceilings, A/B of the kernel variants, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/microbench.py > gpurun_out/microbench.json 2> gpurun_out/microbench.err; echo "microbench exit $?"; cat gpurun_out/microbench.json
B="--steps 5 --warmup 3 --no-e2e --no-cpu --no-extra"
for v in "GV_NO_FAST=1" "GV_FAST_U=1" "GV_FAST_U=2" "GV_FAST_U=4"; do
  env $v timeout 600 python bench.py $B > gpurun_out/bench_$v.log 2>&1; echo "$v exit $?"
  tail -1 gpurun_out/bench_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])"
done
PB="--frames 1024 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
python bench.py $PB > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -c 200 \
    --csv --log-file gpurun_out/launches.csv python bench.py $PB > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
python bench.py $PB > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base function --kernel-name regex:'^k_points' -s 1 -c 2 \
    -o gpurun_out/prof_points -f python bench.py $PB > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/ | head -50
