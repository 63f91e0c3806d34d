"""N > 1 host logic on CPU: two gloo ranks shard a batch of frames, bin + raycast their share
with the oracle standing in for the device, merge the integer planes with an all-reduce and
finalise; the result must be bit-identical to the single-rank run (the property that makes
gv_grid_finalize_multi exact).  Also covers shard_frames and the unique-id broadcast."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grid_vision_b200 import sharding, synth
from oracle import gv_oracle as orc


def test_shard_frames_partition():
    for n in (0, 1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            parts = [sharding.shard_frames(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_frames(4, 2, 2)


WL = synth.C3.scaled(rings=8, azimuth=256, grid_nx=256, grid_ny=256, resolution=0.8)
FRAMES = 5


def _planes(f0, f1):
    g = orc.Grid.from_cells(WL.grid_nx, WL.grid_ny, WL.resolution)
    xyz = synth.make_scans(WL, frames=FRAMES).numpy()
    P = WL.points_per_frame
    Tc, Tb = synth.camera_extrinsics(1)[0], synth.T_base_lidar()
    for f in range(f0, f1):
        fx = xyz[:, f * P:(f + 1) * P]
        cam = orc.transform_points(Tc, *fx)
        lab, _, _, _ = orc.project_label(WL.K(), WL.image_w, WL.image_h, *cam, synth.make_boxes(WL, frame=f))
        g.accumulate(Tb, *fx, lab, occ_mode=orc.OCC_LABELLED, r_max=WL.r_max)
    return g


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = sharding.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
    assert uid == bytes(range(128))
    f0, f1 = sharding.shard_frames(FRAMES, rank, world)
    g = _planes(f0, f1)
    hit, miss = torch.from_numpy(g.hit.copy()), torch.from_numpy(g.miss.copy())
    dist.all_reduce(hit)
    dist.all_reduce(miss)
    g.hit[:] = hit.numpy()
    g.miss[:] = miss.numpy()
    g.finalize(FRAMES)
    if rank == 0:
        np.savez(out, lo=g.log_odds, oc=g.occupancy, hit=hit.numpy(), miss=miss.numpy())
    dist.destroy_process_group()


def test_two_rank_merge_is_bit_identical(tmp_path):
    out = str(tmp_path / "r0.npz")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    g = _planes(0, FRAMES)
    assert np.array_equal(got["hit"], g.hit) and np.array_equal(got["miss"], g.miss)
    assert g.hit.sum() > 0 and g.miss.sum() > g.hit.sum()
    g.finalize(FRAMES)
    assert np.array_equal(got["lo"].view(np.uint32), g.log_odds.view(np.uint32))
    assert np.array_equal(got["oc"].view(np.uint32), g.occupancy.view(np.uint32))
