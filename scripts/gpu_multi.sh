#!/bin/bash
# Multi-GPU check: parity test + bench at N ranks (+ stage timings). Logs -> gpurun_out/.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest multi exit $?"
tail -15 gpurun_out/pytest_multi_$N.log
run() {  # name, env..., then bench args
  local name=$1; shift
  env "$@" GV_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n${N}_$name.log 2>&1; echo "bench $name n=$N exit $?"
  grep -h "finalize_multi stages" gpurun_out/bench_n${N}_$name.log | tail -1
  tail -1 gpurun_out/bench_n${N}_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms']['fuse_bin'], d['phases_ms']['raycast_merge_finalize'], d['merge'][:20], d.get('grid_crc'), d.get('grid_crc_ranks_agree'))"
}
run p2p GV_X=0
run p2p_nooverlap GV_OVERLAP=0
run nccl GV_MERGE=nccl
run nccl_nooverlap GV_MERGE=nccl GV_OVERLAP=0
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_n1.log 2>&1; echo "bench n=1 exit $?"
tail -1 gpurun_out/bench_n1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms']['fuse_bin'], d['phases_ms']['raycast_merge_finalize'], d.get('grid_crc'))"
