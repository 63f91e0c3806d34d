#!/bin/bash
mkdir -p gpurun_out
PB="--frames 1024 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
python bench.py $PB > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base function --kernel-name regex:'^k_points_(pair|col)' -s 1 -c 1 \
    -o gpurun_out/prof_pair -f python bench.py $PB > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof_pair.ncu-rep --page raw --csv > gpurun_out/prof_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_pair.ncu-rep --page source --csv --print-source sass > gpurun_out/src_pair.csv 2>/dev/null
