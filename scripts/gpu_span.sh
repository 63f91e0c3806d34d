#!/bin/bash
mkdir -p gpurun_out
GV_SPAN_CHUNKS=128 timeout 1500 python -m pytest tests/test_gpu_fast_path.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_span128.log 2>&1; echo "pytest span128 exit $?"; tail -3 gpurun_out/pytest_span128.log
for c in 32 64 128; do
  GV_SPAN_CHUNKS=$c python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_span$c.log 2>&1; echo "[$c] exit $?"
  tail -1 gpurun_out/bench_span$c.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['phases_ms']['fuse_bin'], d['phases_ms']['raycast_merge_finalize'], d['grid_crc'])
for k in ('C1','C2','C5','C3_adversarial'):
    v=d['other_configs'][k]; print(' ', k, round(v['ms'],4), round(v.get('ms_raycast_finalize',0),4), '%.3g' % v['points_per_s'])
"
done
