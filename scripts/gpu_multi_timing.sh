#!/bin/bash
N=${N:-8}
mkdir -p gpurun_out
for mode in p2p nccl; do
GV_TIMING=1 GV_MERGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/timing_n${N}_$mode.log 2>&1; echo "exit $?"
grep -E "finalize_multi stages" gpurun_out/timing_n${N}_$mode.log
tail -1 gpurun_out/timing_n${N}_$mode.log | cut -c1-150
done
