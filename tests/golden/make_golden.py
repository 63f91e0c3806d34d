#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/ (small .npz files, committed).

Inputs are seeded; expected outputs come from oracle/gv_oracle.c, which is itself checked
bit-for-bit against the reference's own compiled code (tests/test_oracle_vs_ref.py) wherever
/root/reference is available.  Re-run after any deliberate change of the X1-X3 specification:
    python tests/golden/make_golden.py
tests/test_golden.py replays them against the oracle (CPU) and against the CUDA path (GPU).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from grid_vision_b200 import synth  # noqa: E402
from oracle import gv_oracle as orc  # noqa: E402


def fuse_case(name, wl, frame, nboxes):
    xyz = synth.make_scans(wl, frame0=frame, frames=1).numpy()
    boxes = synth.make_boxes(wl, frame=frame, n=nboxes)
    T = synth.camera_extrinsics(1)[0]
    cam = orc.transform_points(T, *xyz)
    lab, pix, u, v = orc.project_label(wl.K(), wl.image_w, wl.image_h, *cam, boxes)
    uvz = orc.project_kdtree(wl.K(), *cam)
    np.savez_compressed(os.path.join(HERE, name), xyz=xyz, boxes=boxes, T_cam=T, K=wl.K(),
                        wh=np.array([wl.image_w, wl.image_h]), label=lab, pix=pix, u=u, v=v, uvz=uvz)


def grid_case(name, wl, frame, **prm):
    xyz = synth.make_scans(wl, frame0=frame, frames=1, adversarial=prm.pop("adversarial", False)).numpy()
    T = synth.T_base_lidar()
    g = orc.Grid.from_cells(wl.grid_nx, wl.grid_ny, wl.resolution, wl.pos_x, wl.pos_y)
    rng = np.random.default_rng(frame)
    labels = rng.integers(-1, 3, xyz.shape[1]).astype(np.int16)
    upd, cells, flags = g.accumulate(T, *xyz, labels, **prm)
    hit, miss = g.hit.copy(), g.miss.copy()
    corners = orc.pose_corners(synth.make_footprints(wl, frame=frame, n=5))
    g.finalize(2, corners)
    np.savez_compressed(os.path.join(HERE, name), xyz=xyz, labels=labels, T_base=T,
                        grid=np.array([wl.grid_nx, wl.grid_ny]), res=wl.resolution,
                        pos=np.array([wl.pos_x, wl.pos_y]), occ_mode=prm.get("occ_mode", 0),
                        z_gate=np.array(prm.get("z_gate") or [np.nan, np.nan]), r_max=prm.get("r_max", 0.0),
                        cells=cells, flags=flags, hit=hit, miss=miss, updates=upd, corners=corners,
                        log_odds=g.log_odds.copy(), occupancy=g.occupancy.copy())


def update_case(name):
    rng = np.random.default_rng(5)
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    lo0 = rng.uniform(-2.5, 4.0, g.nx * g.ny).astype(np.float32)
    g.log_odds[:] = lo0
    poses = np.stack([rng.uniform(-12, 44, 9), rng.uniform(-12, 12, 9), rng.uniform(0.3, 6, 9),
                      rng.uniform(0.3, 3, 9)], 1)
    xy = np.stack([rng.uniform(-12, 44, 9), rng.uniform(-12, 12, 9)], 1)
    labels = rng.integers(0, 11, 9).astype(np.int32)
    g.update_map()
    s1 = g.log_odds.copy()
    g.update_map_poses(poses)
    s2 = g.log_odds.copy()
    g.update_map_points(xy, labels)
    np.savez_compressed(os.path.join(HERE, name), lo0=lo0, poses=poses, xy=xy, labels=labels, after_r7=s1,
                        after_r8=s2, after_r9=g.log_odds.copy(), occupancy=g.occupancy.copy())


if __name__ == "__main__":
    c1 = synth.C1.scaled(rings=16, azimuth=512)
    fuse_case("fuse_c1_small.npz", c1, 0, 20)
    fuse_case("fuse_c2_small.npz", synth.C2.scaled(rings=16, azimuth=512), 3, 50)
    grid_case("grid_c1_small.npz", c1.scaled(grid_nx=200, grid_ny=200), 1)
    grid_case("grid_c3_labelled.npz", synth.C3.scaled(rings=8, azimuth=512, grid_nx=256, grid_ny=192, resolution=0.4),
              2, occ_mode=orc.OCC_LABELLED, z_gate=(0.2, 2.5), r_max=30.0, adversarial=True)
    update_case("updates_reference_grid.npz")
    print(sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
