#!/bin/bash
# 8-GPU box: parity at 2/4/8 ranks, bench at 8 and 4 ranks. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_8.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest multi exit $?"
tail -5 gpurun_out/pytest_multi_8.log
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n$n.log 2>&1; echo "bench n=$n exit $?"
  tail -1 gpurun_out/bench_n$n.log | cut -c1-200
done
