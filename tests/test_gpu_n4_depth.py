"""N4 on the GPU: gv_box_depths (cloud_detections::computeDepthForBoundingBoxes, ref
src/cloud_detections.cpp:43-87) and gv_pixels_to_3d (pixelTo3D, :89-103) against the oracle, and the
drop-in shim against the reference's own compiled function."""
import ctypes as C
import os

import numpy as np
import pytest

from grid_vision_b200 import synth
from oracle import gv_oracle as orc
from tests.test_oracle_vs_ref import K416, REF_SO, depth_case, p, random_boxes

pytestmark = pytest.mark.gpu
f32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_SO = os.path.join(ROOT, "oracle", "_ref", "libgv_shim.so")


@pytest.mark.parametrize("k", [1, 4, 10, 64])
def test_box_depths_matches_oracle(ctx, k):
    rng = np.random.default_rng(70 + k)
    for m, nb in ((200000, 50), (4097, 7), (3, 5), (0, 3)):
        uvz, boxes = depth_case(rng, m, nb) if m else (np.zeros((0, 3), f32), random_boxes(rng, nb, 416, 416))
        got = ctx.box_depths(uvz, boxes, k)
        exp = orc.box_depths(uvz, boxes, k)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (m, nb, k)


def test_box_depths_on_projected_scan(ctx):
    """The node's own sequence (src/grid_vision_node.cpp:170-180): buildKDTree projection of a C2 scan,
    then the depth of every static detection with k = k_near = 4."""
    wl = synth.C2
    xyz = synth.make_scans(wl, frames=1).numpy()
    cam = orc.transform_points(synth.camera_extrinsics(1)[0], *xyz)
    ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], None)
    uvz = ctx.project_kdtree(0, *cam)
    assert np.array_equal(uvz.view(np.uint32), orc.project_kdtree(wl.K(), *cam).view(np.uint32))
    boxes = synth.make_boxes(wl)
    got = ctx.box_depths(uvz, boxes, 4)
    exp = orc.box_depths(uvz, boxes, 4)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    assert (exp > 0).all()
    Ki = np.linalg.inv(wl.K())
    pts = ctx.pixels_to_3d(boxes, got, Ki)
    for i, b in enumerate(boxes):
        cx = f32(b["x_min"] + (b["x_max"] - b["x_min"]) / 2.0)
        cy = f32(b["y_min"] + (b["y_max"] - b["y_min"]) / 2.0)
        assert np.array_equal(pts[i].view(np.uint64), orc.pixel_to_3d(Ki, cx, cy, got[i]).view(np.uint64))


@pytest.mark.skipif(not (os.path.exists(SHIM_SO) and os.path.exists(REF_SO)), reason="oracle/_ref not built")
def test_compute_depth_for_bounding_boxes_drop_in():
    import torch  # noqa: F401
    ref, shim = C.CDLL(REF_SO), C.CDLL(SHIM_SO)
    rng = np.random.default_rng(12)
    for m, nb, k in ((150000, 40, 4), (9, 6, 10), (0, 3, 4)):
        uvz, boxes = depth_case(rng, m, nb) if m else (np.zeros((0, 3), f32), random_boxes(rng, nb, 416, 416))
        outs = []
        for lib in (ref, shim):
            d = np.empty(nb, f32)
            lib.ref_box_depths(p(uvz), C.c_size_t(len(uvz)), p(boxes), C.c_int(nb), C.c_int(k), p(d))
            outs.append(d)
        assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32)), (m, nb, k)
