#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log
for sg in 64; do
  GV_WALK_SEG=$sg python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import grid_vision_b200 as gv
from grid_vision_b200 import synth
import bench
dev = torch.device("cuda", 0)
ctx = gv.Context(0); ctx.use_torch_stream()
out = []
for wl in (synth.C1, synth.C2):
    ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], synth.camera_extrinsics(1))
    r = bench.timed_batch(torch, gv, synth, ctx, dev, wl, 1, 100)
    g = bench.timed_scan_graph(torch, gv, synth, dev, wl, 300)
    out.append((wl.name[:2], round(r["ms"] * 1e3, 1), round(r["ms_raycast_finalize"] * 1e3, 1), round(g["ms"] * 1e3, 1)))
print("GV_WALK_SEG", os.environ.get("GV_WALK_SEG"), out)
PY
done
