"""Drop-in boundary test: grid_vision_b200/shim/*.cpp (the reference's C++ signatures over the
C ABI) against the reference's own compiled code, through the SAME harness entry points
(oracle/ref_build/ref_harness.cpp).  libgv_ref.so = reference sources, libgv_shim.so = shim;
both are built by `make ref` where /root/reference exists and travel prebuilt to the GPU box.
These tests read like tests of the reference itself: build a cloud and boxes, call
extractCloudPerBBox / buildKDTree / OccupancyGridMap::updateMap, compare."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.test_oracle_vs_ref import (K416, KYAML, REF_SO, RefGrid, p, random_boxes, random_cloud,
                                      ref_labels)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_SO = os.path.join(ROOT, "oracle", "_ref", "libgv_shim.so")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(SHIM_SO) and os.path.exists(REF_SO)),
                                 reason="oracle/_ref not built")]
f32 = np.float32


def _load(path):
    import torch  # noqa: F401  libnccl / CUDA runtime first
    lib = C.CDLL(path)
    lib.ref_build_kdtree.restype = C.c_size_t
    lib.ref_grid_new.restype = C.c_void_p
    lib.ref_grid_free.argtypes = [C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def libs():
    return _load(REF_SO), _load(SHIM_SO)


@pytest.mark.parametrize("K,W,H", [(K416, 416, 416), (KYAML, 640, 480)], ids=["yolo416", "yaml640x480"])
def test_extract_cloud_per_bbox_drop_in(libs, K, W, H):
    ref, shim = libs
    rng = np.random.default_rng(W)
    xyz = random_cloud(rng, 300000)
    for boxes in (random_boxes(rng, 37, W, H), random_boxes(rng, 90, W, H, integer=False),
                  np.zeros(0, random_boxes(rng, 1, W, H).dtype)):
        exp = ref_labels(ref, xyz, K, boxes, W, H)     # asserts order/width/height/is_dense inside
        got = ref_labels(shim, xyz, K, boxes, W, H)
        assert np.array_equal(got, exp)


def test_build_kdtree_drop_in(libs):
    ref, shim = libs
    rng = np.random.default_rng(3)
    xyz = random_cloud(rng, 150000)
    x, y, z = (np.ascontiguousarray(a) for a in xyz)
    K = np.ascontiguousarray(K416)
    outs = []
    for lib in (ref, shim):
        uvz = np.empty((x.size, 3), f32)
        m = lib.ref_build_kdtree(p(x), p(y), p(z), C.c_size_t(x.size), p(K), p(uvz))
        outs.append(uvz[:m].copy())
    a, b = outs
    assert a.shape == b.shape and len(a) > 70000
    ab, bb = a.view(np.uint32).copy(), b.view(np.uint32).copy()
    nan = np.isnan(a) & np.isnan(b)
    ab[nan] = bb[nan] = 0
    assert np.array_equal(ab, bb)


def test_compute_bbox_pose_empty_convention(libs):
    ref, shim = libs
    assert ref.ref_compute_bbox_pose_empty() == shim.ref_compute_bbox_pose_empty() == 0


@pytest.mark.parametrize("like_node", [False, True], ids=["heap", "optional-move-like-the-node"])
def test_occupancy_grid_map_drop_in(libs, like_node):
    """like_node: the map is built the way src/grid_vision_node.cpp:35 builds it (a temporary moved
    into a std::optional), so the object that is updated is not the one the constructor ran on."""
    ref, shim = libs
    a, b = RefGrid(ref, 50, 20, 0.1, like_node), RefGrid(shim, 50, 20, 0.1, like_node)
    assert (a.nx, a.ny, a.len, a.pos) == (b.nx, b.ny, b.len, b.pos)
    rng = np.random.default_rng(4)
    lo0 = rng.uniform(-2.5, 4.0, a.nx * a.ny).astype(f32)
    a.write(lo0)
    b.write(lo0)   # host-side edit of the public member between calls must be honoured
    for k in range(12):
        kind = k % 3
        if kind == 0:
            for lib, g in ((ref, a), (shim, b)):
                lib.ref_update_map(g.h)
        elif kind == 1:
            n = int(rng.integers(0, 30))
            poses = np.ascontiguousarray(np.stack(
                [rng.uniform(-12, 44, n), rng.uniform(-12, 12, n), rng.uniform(0.3, 6, n),
                 rng.uniform(0.3, 3, n)], 1))
            for lib, g in ((ref, a), (shim, b)):
                lib.ref_update_map_poses(g.h, p(poses), C.c_int(n))
        else:
            n = int(rng.integers(1, 30))
            xy = np.ascontiguousarray(np.stack([rng.uniform(-12, 44, n), rng.uniform(-12, 12, n)], 1))
            lab = rng.integers(0, 11, n).astype(np.int32)
            for lib, g in ((ref, a), (shim, b)):
                lib.ref_update_map_points(g.h, p(xy), p(lab), C.c_int(n))
        if k == 7:  # another host-side edit mid-sequence (the shim must notice and re-upload)
            lo1 = rng.uniform(-2.5, 4.0, a.nx * a.ny).astype(f32)
            a.write(lo1)
            b.write(lo1)
            continue
        la, oa = a.read()
        lb, ob = b.read()
        assert np.array_equal(la.view(np.uint32), lb.view(np.uint32)), f"log_odds, step {k}"
        assert np.all(np.abs(oa - ob) <= 1e-5 * np.maximum(np.abs(oa), np.abs(ob))), f"occupancy, step {k}"


@pytest.mark.parametrize("is_dense", [False, True])
def test_transform_lidar_to_camera_drop_in(libs, is_dense):
    """gv_shim::transformPointCloud (the replacement for the pcl_ros call inside
    GridVision::transformLidarToCamera, src/grid_vision_node.cpp:304) vs the stand-in pcl_ros path."""
    from tests.test_oracle_vs_ref import rigid, transform_cloud
    ref, shim = libs
    rng = np.random.default_rng(6)
    xyz = random_cloud(rng, 200000)
    R, t = rigid(3)
    exp = transform_cloud(ref, xyz, R, t, is_dense)
    got = transform_cloud(shim, xyz, R, t, is_dense)
    for a, b in zip(got, exp):
        ab, bb = a.view(np.uint32).copy(), b.view(np.uint32).copy()
        nan = np.isnan(a) & np.isnan(b)
        ab[nan] = bb[nan] = 0
        assert np.array_equal(ab, bb)
