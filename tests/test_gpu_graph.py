"""CUDA-graph replay of the per-scan call sequence (gv_graph_*): a captured gv_process_batch_dev +
gv_grid_finalize, replayed on refilled device buffers, gives bit for bit what the direct calls give
(labels, hit/miss-derived log-odds, occupancy) -- and both match the CPU oracle."""
import numpy as np
import pytest

import grid_vision_b200 as gv
from grid_vision_b200 import synth
from tests.helpers import oracle_fuse, oracle_grid, small

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wl", [small(synth.C1, rings=32, azimuth=1024), synth.C2], ids=["C1_small", "C2_full"])
def test_graph_replay_equals_direct_calls(wl):
    import torch
    dev = torch.device("cuda", 0)
    P = wl.points_per_frame
    nscan = 3
    scans = synth.make_scans(wl, frames=nscan).numpy()
    boxes = [synth.make_boxes(wl, frame=f) for f in range(nscan)]
    fo = np.array([0, P], np.uint64)
    bo = np.array([0, wl.boxes_per_camera], np.int32)
    Tc, Tb = synth.camera_extrinsics(1), synth.T_base_lidar()
    prm = gv.accum_params()

    def setup(ctx):
        ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], Tc)
        ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
        ctx.set_base_transform(Tb)

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        xyz = torch.empty((3, P), dtype=torch.float32, device=dev)
        d_boxes = torch.empty(wl.boxes_per_camera * synth.BOX_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        lab = torch.empty(P, dtype=torch.int16, device=dev)

        def fill(i):
            xyz.copy_(torch.from_numpy(scans[:, i * P:(i + 1) * P].copy()))
            d_boxes.copy_(torch.from_numpy(boxes[i].view(np.uint8).copy()))

        def one(ctx):
            ctx.process_batch(xyz[0], xyz[1], xyz[2], fo, d_boxes, bo, prm, lab)
            ctx.grid_finalize(1)

        out = {}
        for mode in ("direct", "graph"):
            with gv.Context(0) as ctx:
                setup(ctx)
                gid = None
                if mode == "graph":
                    fill(0)
                    one(ctx)              # un-captured first: tables, scratch buffers
                    stream.synchronize()
                    ctx.grid_reset()
                    ctx.graph_begin()
                    one(ctx)              # captured, not executed
                    gid = ctx.graph_end()
                res = []
                for i in range(nscan):
                    fill(i)
                    if gid is None:
                        one(ctx)
                    else:
                        ctx.graph_launch(gid)
                    stream.synchronize()
                    lo, oc = ctx.grid_download()
                    res.append((lab.cpu().numpy().copy(), lo.copy(), oc.copy()))
                if gid is not None:
                    assert ctx.stats()["kernel_launches"] > 0
                    ctx.graph_destroy(gid)
                out[mode] = res
    for i in range(nscan):
        for a, b in zip(out["direct"][i], out["graph"][i]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"scan {i}"
    # and both are the oracle's: per-scan accumulate + finalize(1) on one persistent grid
    g = oracle_grid(wl)
    for i in range(nscan):
        x = scans[:, i * P:(i + 1) * P]
        elab, _, _, _ = oracle_fuse(wl, x, boxes[i], Tc[0])
        g.accumulate(Tb, *x, elab, want_cells=False)
        g.finalize(1)
        assert np.array_equal(out["graph"][i][0], elab), f"labels, scan {i}"
        assert np.array_equal(out["graph"][i][1].view(np.uint32), g.log_odds.view(np.uint32)), f"log-odds, scan {i}"


def test_graph_capture_errors_are_clean():
    import torch
    with gv.Context(0) as ctx:
        wl = small(synth.C1, rings=16, azimuth=512)
        ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], synth.camera_extrinsics(1))
        ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
        ctx.set_base_transform(synth.T_base_lidar())
        with pytest.raises(gv.GridVisionError):
            ctx.graph_end()           # no capture open
        with pytest.raises(gv.GridVisionError):
            ctx.graph_launch(7)       # no such graph
        ctx.graph_begin()
        with pytest.raises(gv.GridVisionError):
            ctx.graph_begin()         # already capturing
        gid = ctx.graph_end()         # an empty graph is legal
        ctx.graph_launch(gid)
        ctx.synchronize()
