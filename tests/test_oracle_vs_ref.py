"""Pins oracle/gv_oracle.c against the REFERENCE'S OWN CODE.

oracle/_ref/libgv_ref.so is /root/reference/src/{occupancy_grid,cloud_detections}.cpp compiled
unmodified against the stand-in headers of oracle/ref_build/stubs (recipe:
oracle/ref_build/Makefile, run by `make ref` / __graft_entry__.build() wherever
/root/reference exists).  Every comparison is bit-exact.  The library travels to the GPU box
prebuilt; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import gv_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgv_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO),
                                reason="oracle/_ref not built (needs /root/reference at build time)")

K416 = np.array([[208.0, 0, 208.0], [0, 208.0, 208.0], [0, 0, 1.0]])
KYAML = np.array([[320.0, 0, 320.0], [0, 320.0, 240.0], [0, 0, 1.0]])  # config/grid_vision_cfg.yaml:16-19
f32 = np.float32


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(REF_SO)
    lib.ref_build_kdtree.restype = C.c_size_t
    lib.ref_grid_new.restype = C.c_void_p
    lib.ref_grid_free.argtypes = [C.c_void_p]
    return lib


def p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def ref_labels(ref, xyz, K, boxes, W, H):
    x, y, z = (np.ascontiguousarray(a, f32) for a in xyz)
    K = np.ascontiguousarray(K, np.float64)
    boxes = np.ascontiguousarray(boxes, orc.BOX_DTYPE)
    lab = np.empty(x.size, np.int16)
    ok = ref.ref_extract_cloud_per_bbox(p(x), p(y), p(z), C.c_size_t(x.size), p(K), p(boxes),
                                        C.c_int(len(boxes)), C.c_int(W), C.c_int(H), p(lab))
    assert ok == 1, "reference output clouds: order / width / height / is_dense"
    return lab


def random_cloud(rng, n):
    xyz = np.stack([rng.normal(size=n) * 8, rng.normal(size=n) * 5, rng.uniform(-5, 60, n)]).astype(f32)
    xyz[:, rng.integers(0, n, n // 100)] = np.nan
    xyz[rng.integers(0, 3, n // 200), rng.integers(0, n, n // 200)] = np.inf
    xyz[2, rng.integers(0, n, 20)] = f32(0.001)
    xyz[2, rng.integers(0, n, 20)] = np.nextafter(f32(0.001), f32(1))
    xyz[2, rng.integers(0, n, 20)] = 0.0
    return xyz


def random_boxes(rng, n, W, H, integer=True):
    w = rng.integers(8, W // 2, n)
    h = rng.integers(8, H // 2, n)
    x0 = rng.integers(0, W - w)
    y0 = rng.integers(0, H - h)
    b = np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.float64)
    if not integer:
        b += rng.uniform(-0.5, 0.5, b.shape)
    return orc.make_boxes(b, confidence=np.sort(rng.uniform(0.6, 1, n))[::-1], label=rng.integers(0, 11, n))


@pytest.mark.parametrize("K,W,H", [(K416, 416, 416), (KYAML, 640, 480)], ids=["yolo416", "yaml640x480"])
@pytest.mark.parametrize("integer_boxes", [True, False], ids=["int-boxes", "frac-boxes"])
def test_extract_cloud_per_bbox_labels(ref, K, W, H, integer_boxes):
    rng = np.random.default_rng(100 + W + integer_boxes)
    xyz = random_cloud(rng, 200000)
    boxes = random_boxes(rng, 37, W, H, integer_boxes)
    got = ref_labels(ref, xyz, K, boxes, W, H)
    exp, _, _, _ = orc.project_label(K, W, H, *xyz, boxes)
    assert np.array_equal(got, exp)
    assert (exp >= 0).sum() > 5000 and len(np.unique(exp)) > 20


def test_extract_cloud_per_bbox_edge_points(ref):
    """The KAT points of tests/test_oracle_kat.py, decided by the reference's own loop."""
    X = np.array([0, 0, 0, 1.0, np.nextafter(f32(1), f32(0)), 1 - 2.0 ** -22, -1.0,
                  np.nextafter(f32(-1), f32(-2)), 0.25, 0.25 + 2.0 ** -20, np.nan, 0, 0, -0.5,
                  np.nextafter(f32(-0.5), f32(-1))], f32)
    Y = np.zeros_like(X)
    Z = np.array([1, 0.001, np.nextafter(f32(0.001), f32(1)), 1, 1, 1, 1, 1, 1, 1, 1, np.inf, -1, 1, 1], f32)
    for boxes in (orc.make_boxes([[100, 50, 260, 300], [0, 0, 416, 416]]),
                  orc.make_boxes([[104, 0, 259.99999999, 416], [260.00000001, 0, 416, 416]]),
                  orc.make_boxes([[np.nan, 0, 416, 416], [300, 0, 100, 416], [-np.inf, -1e300, 1e300, np.inf]]),
                  np.zeros(0, orc.BOX_DTYPE)):
        got = ref_labels(ref, (X, Y, Z), K416, boxes, 416, 416)
        exp, _, _, _ = orc.project_label(K416, 416, 416, X, Y, Z, boxes)
        assert np.array_equal(got, exp), boxes


def test_extract_skewed_and_projective_K(ref):
    rng = np.random.default_rng(7)
    xyz = random_cloud(rng, 50000)
    boxes = random_boxes(rng, 20, 416, 416)
    for K in (np.array([[207.3, 0.7, 211.9], [0.0, 209.1, 203.3], [0, 0, 1.0]]),
              np.array([[200.0, 0, 208.0], [0, 200.0, 208.0], [1e-3, 0, 1.0]])):
        got = ref_labels(ref, xyz, K, boxes, 416, 416)
        exp, _, _, _ = orc.project_label(K, 416, 416, *xyz, boxes)
        assert np.array_equal(got, exp)


def test_build_kdtree_projection(ref):
    rng = np.random.default_rng(8)
    xyz = random_cloud(rng, 100000)
    x, y, z = (np.ascontiguousarray(a) for a in xyz)
    uvz = np.empty((x.size, 3), f32)
    K = np.ascontiguousarray(K416)
    m = ref.ref_build_kdtree(p(x), p(y), p(z), C.c_size_t(x.size), p(K), p(uvz))
    exp = orc.project_kdtree(K416, x, y, z)
    assert m == len(exp) and m > 50000
    a, b = uvz[:m].view(np.uint32).copy(), exp.view(np.uint32).copy()
    nan = np.isnan(uvz[:m]) & np.isnan(exp)
    a[nan] = b[nan] = 0
    assert np.array_equal(a, b)
    assert np.isnan(exp).any()  # NaN depths pass the z <= 0 predicate (src/cloud_detections.cpp:16)


def test_compute_bbox_pose_empty_cloud_convention(ref):
    assert ref.ref_compute_bbox_pose_empty() == 0  # src/cloud_detections.cpp:308-309


class RefGrid:
    def __init__(self, ref, gx, gy, res, like_node=False):
        self.ref = ref
        new = ref.ref_grid_new_like_node if like_node else ref.ref_grid_new
        new.restype = C.c_void_p
        self.h = C.c_void_p(new(C.c_uint8(gx), C.c_uint8(gy), C.c_double(res)))
        nx, ny, r = C.c_int(), C.c_int(), C.c_double()
        ln, ps = (C.c_double * 2)(), (C.c_double * 2)()
        ref.ref_grid_desc(self.h, C.byref(nx), C.byref(ny), ln, ps, C.byref(r))
        self.nx, self.ny, self.len, self.pos, self.res = nx.value, ny.value, tuple(ln), tuple(ps), r.value

    def read(self):
        lo = np.empty(self.nx * self.ny, f32)
        oc = np.empty(self.nx * self.ny, f32)
        self.ref.ref_grid_read(self.h, p(lo), p(oc))
        return lo, oc

    def write(self, lo):
        lo = np.ascontiguousarray(lo, f32)
        self.ref.ref_grid_write_log_odds(self.h, p(lo))

    def get_index(self, x, y):
        ix, iy = C.c_int(), C.c_int()
        ok = self.ref.ref_grid_get_index(self.h, C.c_double(x), C.c_double(y), C.byref(ix), C.byref(iy))
        return (ix.value, iy.value) if ok else None

    def __del__(self):
        self.ref.ref_grid_free(self.h)


def same_grid(rg, og):
    lo, oc = rg.read()
    assert np.array_equal(lo.view(np.uint32), og.log_odds.view(np.uint32)), "log_odds"
    assert np.array_equal(oc.view(np.uint32), og.occupancy.view(np.uint32)), "occupancy"


@pytest.mark.parametrize("gx,gy,res", [(50, 20, 0.1), (20, 20, 0.1), (255, 100, 0.25), (7, 3, 0.3)])
def test_constructor_geometry(ref, gx, gy, res):
    rg = RefGrid(ref, gx, gy, res)
    og = orc.Grid(reference_ctor=(gx, gy, res))
    assert (rg.nx, rg.ny) == (og.nx, og.ny)
    assert rg.len == (og.len_x, og.len_y) and rg.pos == (og.pos_x, og.pos_y) == (float(gx // 3), 0.0)
    same_grid(rg, og)
    rng = np.random.default_rng(gx)
    for px, py in zip(rng.uniform(og.pos_x - gx * 0.6, og.pos_x + gx * 0.6, 3000),
                      rng.uniform(-gy * 0.6, gy * 0.6, 3000)):
        assert rg.get_index(px, py) == og.get_index(px, py)
    for px, py in [(og.pos_x + og.len_x / 2, og.len_y / 2), (og.pos_x - og.len_x / 2, 0.0),
                   (np.nextafter(og.pos_x - og.len_x / 2, 1e9), 0.0), (float("nan"), 0.0)]:
        assert rg.get_index(px, py) == og.get_index(px, py)


def test_update_map_sequences(ref):
    rg = RefGrid(ref, 50, 20, 0.1)
    og = orc.Grid(reference_ctor=(50, 20, 0.1))
    rng = np.random.default_rng(9)
    lo0 = rng.uniform(-2.5, 4.0, og.nx * og.ny).astype(f32)
    rg.write(lo0)
    og.log_odds[:] = lo0
    for k in range(14):
        kind = k % 3
        if kind == 0:
            ref.ref_update_map(rg.h)
            og.update_map()
        elif kind == 1:
            n = int(rng.integers(0, 30))
            poses = np.ascontiguousarray(np.stack(
                [rng.uniform(-12, 44, n), rng.uniform(-12, 12, n), rng.uniform(0.3, 6, n),
                 rng.uniform(0.3, 3, n)], 1))
            ref.ref_update_map_poses(rg.h, p(poses), C.c_int(n))
            og.update_map_poses(poses)
        else:
            n = int(rng.integers(1, 30))
            xy = np.ascontiguousarray(np.stack([rng.uniform(-12, 44, n), rng.uniform(-12, 12, n)], 1))
            lab = rng.integers(0, 11, n).astype(np.int32)
            ref.ref_update_map_points(rg.h, p(xy), p(lab), C.c_int(n))
            og.update_map_points(xy, lab)
        same_grid(rg, og)
    lo, _ = rg.read()
    assert lo.min() == f32(-2.0) and lo.max() > 0


def test_finalize_restates_reference_updates(ref):
    """X3 with zero counts == the reference's updateMap(grid, poses), on the reference itself."""
    rg = RefGrid(ref, 50, 20, 0.1)
    og = orc.Grid(reference_ctor=(50, 20, 0.1))
    rng = np.random.default_rng(10)
    for _ in range(5):
        n = 15
        poses = np.ascontiguousarray(np.stack(
            [rng.uniform(-5, 40, n), rng.uniform(-9, 9, n), rng.uniform(0.3, 6, n), rng.uniform(0.3, 3, n)], 1))
        ref.ref_update_map_poses(rg.h, p(poses), C.c_int(n))
        og.finalize(1, orc.pose_corners(poses))
        same_grid(rg, og)


# ------------------------------------------------------------------------- N4: k-NN median depth, pixelTo3D
def depth_case(rng, m, nb, W=416, H=416):
    uvz = np.stack([rng.uniform(-50, W + 50, m), rng.uniform(-50, H + 50, m), rng.uniform(0.2, 80.0, m)], 1).astype(f32)
    uvz[rng.integers(0, m, max(1, m // 50)), 2] = np.nan      # NaN depths pass buildKDTree's z <= 0 test
    uvz[rng.integers(0, m, max(1, m // 80)), 0] = np.inf
    boxes = random_boxes(rng, nb, W, H, integer=False)
    return np.ascontiguousarray(uvz), boxes


@pytest.mark.parametrize("k", [1, 4, 10])
def test_box_depths_oracle_vs_reference(ref, k):
    """The reference's own computeDepthForBoundingBoxes (over the exact-search FLANN stand-in) and the
    oracle restatement: box centre in float, 3-D float metric with depth as a coordinate, median."""
    rng = np.random.default_rng(40 + k)
    for m, nb in ((5000, 37), (3, 5), (0, 4), (1, 2)):
        uvz, boxes = depth_case(rng, m, nb) if m else (np.zeros((0, 3), f32), random_boxes(rng, nb, 416, 416))
        got = np.empty(nb, f32)
        ref.ref_box_depths(p(uvz), C.c_size_t(len(uvz)), p(boxes), C.c_int(nb), C.c_int(k), p(got))
        exp = orc.box_depths(uvz, boxes, k)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (m, nb, k)
        if m == 0:
            assert np.all(exp == -1.0)


def test_box_depths_depth_is_a_coordinate(ref):
    """A far point right at the box centre loses to a near point a few pixels away: the tree is built
    over (u, v, depth), src/cloud_detections.cpp:28-30, and the query's third coordinate is 0 (:60)."""
    boxes = orc.make_boxes([[100, 100, 200, 200]])
    uvz = np.array([[150, 150, 30.0], [153, 150, 2.0], [150, 156, 2.5]], f32)
    exp = orc.box_depths(uvz, boxes, 1)
    got = np.empty(1, f32)
    ref.ref_box_depths(p(uvz), C.c_size_t(3), p(boxes), C.c_int(1), C.c_int(1), p(got))
    assert exp[0] == got[0] == f32(2.0)
    assert orc.box_depths(uvz, boxes, 3)[0] == f32(2.5)  # median of {2, 2.5, 30}: element 3 // 2 ascending


def test_pixel_to_3d_oracle_vs_reference(ref):
    rng = np.random.default_rng(9)
    Ki = np.ascontiguousarray(np.linalg.inv(K416))
    for _ in range(200):
        px, py, d = f32(rng.uniform(0, 416)), f32(rng.uniform(0, 416)), f32(rng.uniform(0.1, 90))
        got = np.empty(3, np.float64)
        ref.ref_pixel_to_3d(p(Ki), C.c_float(px), C.c_float(py), C.c_float(d), p(got))
        assert np.array_equal(got.view(np.uint64), orc.pixel_to_3d(Ki, px, py, d).view(np.uint64))


# ------------------------------------------------------------------------- R1 through the node's call
def transform_cloud(lib, xyz, R, t, is_dense):
    x, y, z = (np.ascontiguousarray(a, f32) for a in xyz)
    n = x.size
    ox, oy, oz, it = (np.empty(n, f32) for _ in range(4))
    R = np.ascontiguousarray(R, np.float64)
    t = np.ascontiguousarray(t, np.float64)
    lib.ref_transform_cloud(p(x), p(y), p(z), C.c_size_t(n), C.c_int(int(is_dense)), p(R), p(t), p(ox), p(oy), p(oz),
                            p(it))
    assert np.array_equal(it, np.arange(n, dtype=f32))  # the other fields of every point are copied
    return ox, oy, oz


def rigid(seed):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R, rng.uniform(-3, 3, 3)


@pytest.mark.parametrize("is_dense", [False, True])
def test_transform_lidar_to_camera_oracle_vs_pcl_ros_standin(ref, is_dense):
    """GridVision::transformLidarToCamera's compute call (src/grid_vision_node.cpp:296-304) through the
    pcl_ros stand-in vs the oracle's R1 (gvo_transform_points) on the narrowed float matrix."""
    rng = np.random.default_rng(5)
    xyz = random_cloud(rng, 50000)
    for seed in (1, 2):
        R, t = rigid(seed)
        got = transform_cloud(ref, xyz, R, t, is_dense)
        T = np.eye(4, dtype=f32)
        T[:3, :3] = R.astype(f32)
        T[:3, 3] = t.astype(f32)
        exp = orc.transform_points(T, *xyz, is_dense=is_dense)
        for a, b in zip(got, exp):
            ab, bb = a.view(np.uint32).copy(), b.view(np.uint32).copy()
            nan = np.isnan(a) & np.isnan(b)
            ab[nan] = bb[nan] = 0
            assert np.array_equal(ab, bb)
