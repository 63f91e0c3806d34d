#!/bin/bash
# Light profile refresh: plain run, ncu launch list, ncu --set full of the hot kernels (C3, 4096 frames).
mkdir -p gpurun_out
A="--frames 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra"
python bench.py $A > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function --kernel-name regex:'^k_' -c 100 \
    --csv --log-file gpurun_out/launches.csv python bench.py $A > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base function --kernel-name regex:'^k_(points|sweep_compact|sweep_walk|finalize)' -s 4 -c 4 \
    -o gpurun_out/prof -f python bench.py $A > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
