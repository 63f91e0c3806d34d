#!/bin/bash
# 8-GPU box: parity at 2/4/8 ranks in both merge modes, bench at 8 / 4 / 2 ranks. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_8.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest multi exit $?"
tail -5 gpurun_out/pytest_multi_8.log
run() { # n mode extra
  GV_MERGE=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$1$4 bench.py --gpus $1 --steps 10 --warmup 3 --no-cpu $3 > gpurun_out/bench_n$1_$2.log 2>&1; echo "bench n=$1 $2 exit $?"
  tail -1 gpurun_out/bench_n$1_$2.log | cut -c1-160
}
run 8 p2p "" 1
run 8 nccl "--no-e2e" 2
run 4 p2p "" 3
run 2 p2p "--no-e2e" 4
