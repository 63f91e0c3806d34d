#!/bin/bash
# 4-GPU box: parity at 2/4 ranks, bench at 4 ranks (merge mode x overlap)
mkdir -p gpurun_out
N=${N:-4}
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest multi exit $?"
tail -3 gpurun_out/pytest_multi_$N.log
run() { # n mode overlap tag port timing
  GV_TIMING=$6 GV_MERGE=$2 GV_OVERLAP=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29$5 bench.py --gpus $1 --steps 20 --warmup 3 --no-cpu --no-extra --no-e2e > gpurun_out/bench_n$1_$4.log 2>&1; echo "bench n=$1 $4 exit $?"
  grep -a "finalize_multi stages" gpurun_out/bench_n$1_$4.log | head -1
  grep -a '^{' gpurun_out/bench_n$1_$4.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['phases_ms'].get('fuse_bin'), d['phases_ms'].get('raycast_merge_finalize'), d.get('grid_crc'))"
}
run $N p2p 1 p2p 511 0
run $N p2p 1 p2p_timed 512 1
run $N p2p 0 p2p_noov 513 0
run $N nccl 1 nccl 514 0
run $N nccl 0 nccl_noov 515 0
