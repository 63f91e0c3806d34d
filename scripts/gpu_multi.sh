#!/bin/bash
# Multi-GPU check: parity test + bench at N ranks. Logs -> gpurun_out/.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest multi exit $?"
tail -15 gpurun_out/pytest_multi_$N.log
for n in 1 $N; do
  if [ $n = 1 ]; then CMD="python"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533"; fi
  timeout 900 $CMD bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n$n.log 2>&1; echo "bench n=$n exit $?"
  tail -2 gpurun_out/bench_n$n.log | cut -c1-1500
done
GV_MERGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n${N}_nccl.log 2>&1; echo "bench nccl n=$N exit $?"
tail -1 gpurun_out/bench_n${N}_nccl.log | cut -c1-300
