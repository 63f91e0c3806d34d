#!/bin/bash
# 8-GPU box: parity at 2/4/8 ranks in both merge modes, bench at 8 ranks (merge mode x overlap), 4 and 2 ranks
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_8.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest multi exit $?"
tail -3 gpurun_out/pytest_multi_8.log
run() { # n mode overlap tag extra
  GV_TIMING=1 GV_MERGE=$2 GV_OVERLAP=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29$5 bench.py --gpus $1 --steps 20 --warmup 3 --no-cpu --no-extra $6 > gpurun_out/bench_n$1_$4.log 2>&1; echo "bench n=$1 $4 exit $?"
  grep -a "finalize_multi stages" gpurun_out/bench_n$1_$4.log | head -1
  grep -a '^{' gpurun_out/bench_n$1_$4.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['phases_ms'].get('fuse_bin'), d['phases_ms'].get('raycast_merge_finalize'), d.get('grid_crc'), (d.get('e2e') or {}).get('value'))"
}
run 8 p2p 1 p2p 511 "--no-e2e"
run 8 p2p 0 p2p_noov 512 "--no-e2e"
run 8 nccl 1 nccl 513 "--no-e2e"
run 8 nccl 0 nccl_noov 514 "--no-e2e"
run 4 p2p 1 p2p 515 "--no-e2e"
run 2 p2p 1 p2p 516 "--no-e2e"
