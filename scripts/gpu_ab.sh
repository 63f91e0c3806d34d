#!/bin/bash
# A/B of environment variants on the C3 bench: usage gpu_ab.sh "VAR=1" "VAR=2 OTHER=3" ...
mkdir -p gpurun_out
B="--steps ${STEPS:-5} --warmup 3 --no-e2e --no-cpu --no-extra"
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py $B > gpurun_out/bench_ab$i.log 2>&1; echo "[$v] exit $?"
  tail -1 gpurun_out/bench_ab$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms']['fuse_bin'], d['phases_ms']['raycast_merge_finalize'], d['roofline']['frac'], d.get('grid_crc'), d.get('parity_sample',{}).get('grids_equal'))"
done
