"""One C5 batch (8192^2 @0.05 m, 128 scans, 120 m rays) through process_batch + finalize, a few times:
run under `ncu --metrics gpu__time_duration.sum` to see where the raycast of the large map goes."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_vision_b200 as gv
from grid_vision_b200 import synth
import bench

dev = torch.device("cuda", 0)
ctx = gv.Context(0)
ctx.use_torch_stream()
wl = synth.C5
ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], synth.camera_extrinsics(1))
r = bench.timed_batch(torch, gv, synth, ctx, dev, wl, wl.frames, 3)
print({k: r[k] for k in ("ms", "ms_points_kernel", "ms_raycast_finalize", "distinct_end_cells")})
print(ctx.stats())
