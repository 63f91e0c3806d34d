#!/bin/bash
mkdir -p gpurun_out
python scripts/c5_probe.py > gpurun_out/c5_plain.log 2>&1; tail -2 gpurun_out/c5_plain.log
ncu --metrics gpu__time_duration.sum,lts__t_sectors_op_red.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --kernel-name-base function --kernel-name regex:'^k_' -c 200 \
    --csv --log-file gpurun_out/c5_launches.csv python scripts/c5_probe.py > gpurun_out/c5_ncu.log 2>&1
echo "ncu exit $?"
