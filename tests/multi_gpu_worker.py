"""torchrun worker for tests/test_gpu_multi.py: every rank processes its shard of frames on its
own GPU, gv_grid_finalize_multi merges over NCCL, every rank checks its full grid against the
single-rank CPU oracle (bit-exact counts-derived log-odds)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import grid_vision_b200 as gv  # noqa: E402
from grid_vision_b200 import sharding, synth  # noqa: E402
from oracle import gv_oracle as orc  # noqa: E402


def main():
    rank, world, local = sharding.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = synth.C3.scaled(rings=16, azimuth=1024, grid_nx=1000, grid_ny=1000, resolution=0.2)
    F, P = 9, wl.points_per_frame
    xyz = synth.make_scans(wl, frames=F).numpy()
    per_frame = [synth.make_boxes(wl, frame=f) for f in range(F)]
    Tc, Tb = synth.camera_extrinsics(1), synth.T_base_lidar()
    f0, f1 = sharding.shard_frames(F, rank, world)
    ctx = gv.Context(local)
    ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], Tc)
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
    ctx.set_base_transform(Tb)
    sharding.init_context_comm(ctx, torch.device("cuda", local))
    mode = os.environ.get("GV_MULTI_MODE", "nccl")
    if mode == "p2p":
        assert sharding.enable_p2p(ctx, torch.device("cuda", local)), "peer mapping failed"
    prm = dict(occ_mode=gv.OCC_LABELLED, r_max=wl.r_max)
    corners = orc.pose_corners(synth.make_footprints(wl, n=6))
    g = orc.Grid.from_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
    for rounds in range(2):  # two merged batches: state carries over between them
        if f1 > f0:
            boxes = np.concatenate(per_frame[f0:f1])
            bo = np.cumsum([0] + [len(b) for b in per_frame[f0:f1]]).astype(np.int32)
            fo = (np.arange(f1 - f0 + 1) * P).astype(np.uint64)
            ctx.process_batch(*np.ascontiguousarray(xyz[:, f0 * P:f1 * P]), fo, boxes, bo, gv.accum_params(**prm))
        ctx.grid_finalize(F, corners, multi=True)
        lo, oc = ctx.grid_download()
        for f in range(F):
            fx = xyz[:, f * P:(f + 1) * P]
            cam = orc.transform_points(Tc[0], *fx)
            lab, _, _, _ = orc.project_label(wl.K(), wl.image_w, wl.image_h, *cam, per_frame[f])
            g.accumulate(Tb, *fx, lab, **prm)
        g.finalize(F, corners)
        assert np.array_equal(lo.view(np.uint32), g.log_odds.view(np.uint32)), f"rank {rank} round {rounds}: log_odds"
        assert np.allclose(oc, g.occupancy, rtol=1e-5, atol=0), f"rank {rank}: occupancy"
        hit, miss = ctx.grid_counts()
        assert not hit.any() and not miss.any()
    # three more batches back to back, nothing read in between: batch k+1 is binned into the second
    # end-cell plane while batch k's merge (peer barriers, sweep, finalise) is still in flight
    for rounds in range(3):
        if f1 > f0:
            ctx.process_batch(*np.ascontiguousarray(xyz[:, f0 * P:f1 * P]), fo, boxes, bo, gv.accum_params(**prm))
        ctx.grid_finalize(F, corners, multi=True)
        for f in range(F):
            fx = xyz[:, f * P:(f + 1) * P]
            cam = orc.transform_points(Tc[0], *fx)
            lab, _, _, _ = orc.project_label(wl.K(), wl.image_w, wl.image_h, *cam, per_frame[f])
            g.accumulate(Tb, *fx, lab, **prm)
        g.finalize(F, corners)
    lo, oc = ctx.grid_download()
    assert np.array_equal(lo.view(np.uint32), g.log_odds.view(np.uint32)), f"rank {rank}: log_odds after pipelined batches"
    ctx.close()
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} mode={mode}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
