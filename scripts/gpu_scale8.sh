#!/bin/bash
# 8-GPU box: the driver's scaling sequence (default merge) + the peer-memory merge at 8 ranks
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_8.txt
run() { # n mode tag port extra
  GV_MERGE=$2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29$4 bench.py --gpus $1 --steps 20 --warmup 3 --no-cpu --no-extra $5 > gpurun_out/scale_n$1_$3.log 2>&1; echo "bench n=$1 $3 exit $?"
  grep -a '^{' gpurun_out/scale_n$1_$3.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['phases_ms'].get('fuse_bin'), d['phases_ms'].get('raycast_merge_finalize'), d.get('grid_crc'), d.get('grid_crc_ranks_agree'), (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('ms_per_step'))"
}
run 8 nccl nccl 511 ""
run 8 p2p p2p 512 "--no-e2e"
run 4 nccl nccl 513 "--no-e2e"
run 2 nccl nccl 514 "--no-e2e"
python bench.py --steps 20 --warmup 3 --no-cpu --no-extra --no-e2e > gpurun_out/scale_n1.log 2>&1
grep -a '^{' gpurun_out/scale_n1.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['phases_ms'].get('fuse_bin'), d['phases_ms'].get('raycast_merge_finalize'), d.get('grid_crc'))"
