#!/bin/bash
mkdir -p gpurun_out
PB="--frames 4096 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
python bench.py $PB > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base function --kernel-name regex:'^k_sweep_walk' -s 1 -c 1 \
    -o gpurun_out/prof_walk -f python bench.py $PB > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof_walk.ncu-rep --page raw --csv > gpurun_out/prof_walk_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_walk.ncu-rep --page source --csv --print-source sass > gpurun_out/src_walk.csv 2>/dev/null
