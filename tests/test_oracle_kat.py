"""Known-answer tests that pin the CPU oracle (SURVEY.md §8.c list a-j).

The reference has no tests or golden vectors of its own, so each expectation here is
hand-derived from the reference source (citations relative to /root/reference) or
computed by an independent numpy float32/float64 mirror of the same arithmetic.
"""
import math

import numpy as np
import pytest

from oracle import gv_oracle as orc

K416 = np.array([[208.0, 0, 208.0], [0, 208.0, 208.0], [0, 0, 1.0]])
f32 = np.float32


def nextf(x, toward):
    return np.nextafter(f32(x), f32(toward))


def label1(x, y, z, boxes, K=K416, W=416, H=416):
    lab, pix, u, v = orc.project_label(K, W, H, [x], [y], [z], boxes)
    return int(lab[0]), int(pix[0]), u[0], v[0]


BOX_ALL = orc.make_boxes([[0, 0, 416, 416]])


# ---------------------------------------------------------------- (a) (b) (c)
def test_a_principal_point():
    lab, pix, u, v = label1(0.0, 0.0, 1.0, BOX_ALL)
    assert (u, v) == (f32(208.0), f32(208.0))
    assert pix == 208 * 416 + 208 and lab == 0


def test_b_depth_threshold_and_nonfinite():
    # src/cloud_detections.cpp:264: z <= 0.001f rejected
    assert label1(0, 0, f32(0.001), BOX_ALL)[0] == -1
    assert label1(0, 0, nextf(0.001, 1.0), BOX_ALL)[0] == 0
    for bad in (np.nan, np.inf, -np.inf):
        assert label1(bad, 0, 1, BOX_ALL)[0] == -1
        assert label1(0, bad, 1, BOX_ALL)[0] == -1
        assert label1(0, 0, bad, BOX_ALL)[0] == -1
    # rejected before projection -> u,v stay NaN, pix -1
    lab, pix, u, v = label1(0, 0, -1.0, BOX_ALL)
    assert pix == -1 and np.isnan(u) and np.isnan(v)


def test_c_image_bounds_half_open():
    # u = 208*X + 208 with Z = 1: X = 1 -> u = 416 == W rejected (:276), one ulp below accepted
    lab, pix, u, _ = label1(1.0, 0.0, 1.0, BOX_ALL)
    assert u == f32(416.0) and pix == -1 and lab == -1
    # one float ulp below X = 1 still rounds to u = 416.0f (ulp(416) = 2^-15 > 208 * 2^-24):
    # the double quotient is < 416 but the narrowing to float (:272) lands on W -> rejected
    lab, pix, u, _ = label1(nextf(1.0, 0.0), 0.0, 1.0, BOX_ALL)
    assert u == f32(416.0) and pix == -1
    lab, pix, u, _ = label1(f32(1.0 - 2.0 ** -22), 0.0, 1.0, BOX_ALL)
    assert u == nextf(416.0, 0.0) - f32(2.0 ** -15) and pix == 208 * 416 + 415 and lab == 0
    # u = 0 exactly (X = -1) accepted: the test is u < 0
    lab, pix, u, _ = label1(-1.0, 0.0, 1.0, BOX_ALL)
    assert u == f32(0.0) and pix == 208 * 416 and lab == 0
    assert label1(nextf(-1.0, -2.0), 0.0, 1.0, BOX_ALL)[1] == -1


# ---------------------------------------------------------------- (d) (e)
def test_d_box_edges_inclusive():
    # box x in [100, 260]: u = 260 exactly at X = 0.25 (208*0.25+208), inclusive (:282-283)
    boxes = orc.make_boxes([[100, 50, 260, 300]])
    assert label1(0.25, 0.0, 1.0, boxes)[0] == 0
    # one ulp of X above 0.25 still narrows to u = 260.0f -> still inside (float u decides)
    assert label1(nextf(0.25, 1.0), 0.0, 1.0, boxes)[0] == 0
    assert label1(f32(0.25 + 2.0 ** -20), 0.0, 1.0, boxes)[0] == -1
    # lower edge x_min = 100: X = (100-208)/208 is not exact; use u = 104 -> X = -0.5
    boxes = orc.make_boxes([[104, 50, 260, 300]])
    assert label1(-0.5, 0.0, 1.0, boxes)[0] == 0
    assert label1(nextf(-0.5, -1.0), 0.0, 1.0, boxes)[0] == -1
    # non-float-representable double bound: u (float) vs 260.00000001 (double)
    boxes = orc.make_boxes([[0, 0, 259.99999999, 416]])
    assert label1(0.25, 0.0, 1.0, boxes)[0] == -1
    boxes = orc.make_boxes([[260.00000001, 0, 416, 416]])
    assert label1(0.25, 0.0, 1.0, boxes)[0] == -1
    assert label1(f32(0.25 + 2.0 ** -20), 0.0, 1.0, boxes)[0] == 0


def test_e_first_box_wins():
    boxes = orc.make_boxes([[0, 0, 10, 10], [150, 150, 300, 300], [100, 100, 400, 400]])
    assert label1(0.0, 0.0, 1.0, boxes)[0] == 1
    boxes = boxes[[0, 2, 1]]
    assert label1(0.0, 0.0, 1.0, boxes)[0] == 1  # now the big box is index 1 and wins
    assert label1(0.0, 0.0, 1.0, boxes[:1])[0] == -1


# ---------------------------------------------------------------- (f) R1
def test_f_transform_identity_and_permutation():
    rng = np.random.default_rng(1)
    p = rng.normal(size=(3, 1000)).astype(f32) * 30
    ox, oy, oz = orc.transform_points(np.eye(4), *p)
    assert np.array_equal(ox.view(np.uint32), p[0].view(np.uint32))
    assert np.array_equal(oz.view(np.uint32), p[2].view(np.uint32))
    T = np.eye(4)
    T[:3, :3] = [[0, -1, 0], [0, 0, -1], [1, 0, 0]]
    ox, oy, oz = orc.transform_points(T, *p)
    assert np.array_equal(ox, -p[1]) and np.array_equal(oy, -p[2]) and np.array_equal(oz, p[0])


def test_f_transform_matches_numpy_float32_order():
    rng = np.random.default_rng(2)
    p = (rng.normal(size=(3, 4096)) * 40).astype(f32)
    T = np.eye(4, dtype=f32)
    a = 0.3
    T[:3, :3] = np.array([[math.cos(a), -math.sin(a), 0.01], [math.sin(a), math.cos(a), -0.02],
                          [0.005, 0.03, 1.0]], dtype=f32)
    T[:3, 3] = np.array([0.1, -0.4, 2.4], dtype=f32)
    got = orc.transform_points(T, *p)
    for r in range(3):
        # c0*x + (c1*y + (c2*z + c3)), every op rounded to float32
        exp = T[r, 0] * p[0] + (T[r, 1] * p[1] + (T[r, 2] * p[2] + T[r, 3]))
        assert exp.dtype == f32
        assert np.array_equal(got[r].view(np.uint32), exp.view(np.uint32))


def test_f_nonfinite_passthrough_when_not_dense():
    T = np.eye(4, dtype=f32)
    T[0, 3] = 5
    x = np.array([1.0, np.nan, 2.0], f32)
    y = np.array([1.0, 1.0, np.inf], f32)
    z = np.ones(3, f32)
    ox, oy, oz = orc.transform_points(T, x, y, z, is_dense=False)
    assert ox[0] == 6 and np.isnan(ox[1]) and oy[1] == 1 and ox[2] == 2 and np.isinf(oy[2])
    # dense clouds transform every point: 0*inf = NaN poisons all three outputs
    ox, oy, oz = orc.transform_points(T, x, y, z, is_dense=True)
    assert ox[0] == 6 and np.isnan(ox[2]) and np.isinf(oy[2]) and np.isnan(oz[2])


def test_projection_matches_numpy_mirror():
    rng = np.random.default_rng(3)
    n = 20000
    x = (rng.normal(size=n) * 10).astype(f32)
    y = (rng.normal(size=n) * 5).astype(f32)
    z = (rng.uniform(-2, 40, size=n)).astype(f32)
    boxes = orc.make_boxes(np.array([[10, 20, 200, 220], [150, 100, 400, 300.5], [0, 0, 416, 50]]))
    lab, pix, u, v = orc.project_label(K416, 416, 416, x, y, z, boxes)
    X, Y, Z = (a.astype(np.float64) for a in (x, y, z))
    ok = np.isfinite(x) & np.isfinite(y) & np.isfinite(z) & (z > f32(0.001))
    with np.errstate(all="ignore"):
        ue = (((208.0 * X + 0.0 * Y) + 208.0 * Z) / ((0 * X + 0 * Y) + 1.0 * Z)).astype(f32)
        ve = (((0.0 * X + 208.0 * Y) + 208.0 * Z) / ((0 * X + 0 * Y) + 1.0 * Z)).astype(f32)
    assert np.array_equal(u[ok].view(np.uint32), ue[ok].view(np.uint32))
    assert np.array_equal(v[ok].view(np.uint32), ve[ok].view(np.uint32))
    inimg = ok & (ue >= 0) & (ue < 416) & (ve >= 0) & (ve < 416)
    pe = np.where(inimg, ve.astype(np.int32) * 416 + ue.astype(np.int32), -1)
    assert np.array_equal(pix, pe)
    le = np.full(n, -1, np.int16)
    for i in range(len(boxes) - 1, -1, -1):
        b = boxes[i]
        m = inimg & (ue >= b["x_min"]) & (ue <= b["x_max"]) & (ve >= b["y_min"]) & (ve <= b["y_max"])
        le[m] = i
    assert np.array_equal(lab, le)
    assert (lab >= 0).sum() > 100


def test_kdtree_projection_predicate():
    # src/cloud_detections.cpp:16: only z <= 0 is skipped; no bounds test, order kept
    x = np.array([0, 5, 0, -100, 0], f32)
    y = np.array([0, 0, 0, 0, 0], f32)
    z = np.array([1, 1, 0, 2, -3], f32)
    uvz = orc.project_kdtree(K416, x, y, z)
    assert uvz.shape == (3, 3)
    assert np.allclose(uvz[:, 2], [1, 1, 2])
    assert uvz[1, 0] == f32(208 * 5 + 208) and uvz[2, 0] == f32(-100 * 208 / 2 + 208)


def test_aos_reference_shape_matches_soa_labels():
    rng = np.random.default_rng(4)
    n = 5000
    xyz = np.stack([(rng.normal(size=n) * 5), (rng.normal(size=n) * 5), rng.uniform(0.5, 30, n)]).astype(f32)
    boxes = orc.make_boxes([[100, 100, 300, 300], [0, 0, 416, 416]])
    lab, _, _, _ = orc.project_label(K416, 416, 416, *xyz, boxes)
    aos = np.zeros(n, dtype=orc.POINT_DTYPE)
    aos["x"], aos["y"], aos["z"], aos["intensity"] = xyz[0], xyz[1], xyz[2], np.arange(n)
    clouds = orc.extract_cloud_per_bbox_aos(aos, K416, boxes, 416, 416)
    for i, c in enumerate(clouds):
        assert np.array_equal(c["intensity"], np.flatnonzero(lab == i).astype(f32))  # stable order


# ---------------------------------------------------------------- (g) grid geometry
def test_g_reference_geometry_and_get_index():
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    # src/occupancy_grid.cpp:10-11: Length(50,20) @0.1 -> 500x200 cells, centre (50/3, 0) = (16, 0)
    assert (g.nx, g.ny) == (500, 200)
    assert (g.pos_x, g.pos_y) == (16.0, 0.0)
    assert np.all(g.log_odds == 0.0) and np.all(g.occupancy == 0.5)
    assert g.get_index(16.0, 0.0) == (250, 100)          # centre
    assert g.get_index(41.0, 10.0) == (0, 0)             # max-x / max-y corner is index 0
    assert g.get_index(np.nextafter(41.0, 42.0), 0.0) is None
    assert g.get_index(-9.0, 0.0) is None                 # q == length is outside (strict <)
    # one ulp inside the min edge passes the within-map test but (p - 25) - 16 rounds to -50,
    # index 500 -> out of range -> false: grid_map's two tests use different op orders
    assert g.get_index(np.nextafter(-9.0, 0.0), 0.0) is None
    assert g.get_index(-9.0 + 1e-9, -10.0 + 1e-9) == (499, 199)
    assert g.get_index(float("nan"), 0.0) is None
    assert g.get_index(1e300, 0.0) is None


def test_g_size_rounding():
    g = orc.Grid(10.04, 5.06, 0.1)   # round(100.4)=100, round(50.6)=51
    assert (g.nx, g.ny) == (100, 51)
    assert g.len_x == 100 * 0.1 and g.len_y == 51 * 0.1


def test_g_get_index_numpy_mirror():
    g = orc.Grid(51.2, 25.6, 0.05, 3.3, -1.7)
    rng = np.random.default_rng(5)
    P = np.stack([rng.uniform(-25, 31, 20000), rng.uniform(-16, 13, 20000)], 1)
    for px, py in P[:4000]:
        qx = -((px - g.pos_x) - 0.5 * g.len_x)
        qy = -((py - g.pos_y) - 0.5 * g.len_y)
        inside = qx >= 0 and qy >= 0 and qx < g.len_x and qy < g.len_y
        exp = None
        if inside:
            ix = int(-(((px - 0.5 * g.len_x) - g.pos_x) / g.res))
            iy = int(-(((py - 0.5 * g.len_y) - g.pos_y) / g.res))
            if 0 <= ix < g.nx and 0 <= iy < g.ny:
                exp = (ix, iy)
        assert g.get_index(px, py) == exp


# ---------------------------------------------------------------- (h) (i) updates
def test_h_offmap_corner_skips_footprint():
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    # object centred at x=40.5 with length 2 -> front corners at 41.5 > 41 (off-map): skipped
    g.update_map_poses([[40.5, 0.0, 2.0, 1.0]])
    assert np.all(g.log_odds == f32(-0.2))
    # fully inside: cells x in [39.5,40.5] -> ix 5..15, y in [-0.5,0.5] -> iy 95..105
    g2 = orc.Grid(reference_ctor=(50, 20, 0.1))
    g2.update_map_poses([[40.0, 0.0, 1.0, 1.0]])
    lo = g2.log_odds.reshape(g2.ny, g2.nx)  # [iy, ix] view of the column-major plane
    exp = f32(f32(0.0) + f32(-0.2)) + f32(0.85)
    inside = lo == exp
    iy, ix = np.nonzero(inside)
    assert ix.min() == g2.get_index(40.5, 0.0)[0] and ix.max() == g2.get_index(39.5, 0.0)[0]
    assert iy.min() == g2.get_index(40.0, 0.5)[1] and iy.max() == g2.get_index(40.0, -0.5)[1]
    assert np.all(lo[~inside] == f32(-0.2))


def test_i_decay_sequence_clamp_and_saturation():
    g = orc.Grid(reference_ctor=(5, 5, 0.5))
    l = f32(0.0)
    for k in range(12):
        g.update_map()
        l = f32(l + f32(-0.2))
        l = max(l, f32(-2.0))
        assert np.all(g.log_odds.view(np.uint32) == np.array(l, f32).view(np.uint32)), k
    assert np.all(g.log_odds == f32(-2.0))
    assert np.allclose(g.occupancy, 1.0 / (1.0 + math.exp(2.0)), rtol=1e-6)
    # +0.85 - 0.2 per frame saturates at exactly 3.6f
    g = orc.Grid(reference_ctor=(5, 5, 0.5))
    cx, cy = g.pos_x, g.pos_y
    l = f32(0.0)
    for k in range(8):
        g.update_map_poses([[cx, cy, 1.0, 1.0]])
        l = min(f32(f32(l + f32(-0.2)) + f32(0.85)), f32(3.6))
        c = g.get_index(cx, cy)
        assert g.log_odds[c[0] + c[1] * g.nx] == l
    assert l == f32(3.6)


def test_i_sigmoid_of_zero_and_r9_depths():
    g = orc.Grid(reference_ctor=(5, 5, 0.5))
    assert np.all(g.occupancy == 0.5)
    g.finalize(k_decay=0)
    assert np.all(g.log_odds == 0) and np.all(g.occupancy == f32(0.5))
    # src/occupancy_grid.cpp:185-196
    assert [orc.estimated_depth(l) for l in (9, 2, 0, 1, 3, 10)] == [3.5, pytest.approx(0.6), 2.5, 2.5, -1.0, -1.0]


def test_r9_points_update_matches_corner_form():
    g1 = orc.Grid(reference_ctor=(50, 20, 0.1))
    g2 = orc.Grid(reference_ctor=(50, 20, 0.1))
    xy = np.array([[20.0, 1.0], [30.0, -3.0], [39.0, 9.5], [10.0, 0.0]])
    labels = np.array([9, 2, 0, 5], np.int32)   # label 5 -> depth -1 (degenerate but valid rect)
    g1.update_map_points(xy, labels)
    g2.finalize(1, orc.point_corners(xy, labels))
    assert np.array_equal(g1.log_odds.view(np.uint32), g2.log_odds.view(np.uint32))
    assert np.array_equal(g1.occupancy.view(np.uint32), g2.occupancy.view(np.uint32))
    assert (g1.log_odds > 0).sum() > 0


def test_finalize_equals_r8_when_counts_are_zero():
    rng = np.random.default_rng(7)
    poses = np.stack([rng.uniform(-9, 41, 12), rng.uniform(-10, 10, 12), rng.uniform(1, 5, 12),
                      rng.uniform(0.5, 2.5, 12)], 1)
    g1 = orc.Grid(reference_ctor=(50, 20, 0.1))
    g2 = orc.Grid(reference_ctor=(50, 20, 0.1))
    for _ in range(4):
        g1.update_map_poses(poses)
        g2.finalize(1, orc.pose_corners(poses))
    assert np.array_equal(g1.log_odds.view(np.uint32), g2.log_odds.view(np.uint32))
    assert np.array_equal(g1.occupancy.view(np.uint32), g2.occupancy.view(np.uint32))


# ---------------------------------------------------------------- (j) Bresenham
def test_j_bresenham_cases():
    c = orc.bresenham_cells(2, 3, 7, 3)
    assert c.tolist() == [[x, 3] for x in range(2, 8)]
    c = orc.bresenham_cells(2, 3, 2, 0)
    assert c.tolist() == [[2, y] for y in (3, 2, 1, 0)]
    c = orc.bresenham_cells(0, 0, 4, 4)
    assert c.tolist() == [[i, i] for i in range(5)]
    c = orc.bresenham_cells(0, 0, -3, 3)
    assert c.tolist() == [[-i, i] for i in range(4)]
    # slope 1/2: den=4, num starts at 2, add 2 -> y steps at k=1 and k=3
    c = orc.bresenham_cells(0, 0, 4, 2)
    assert c.tolist() == [[0, 0], [1, 1], [2, 1], [3, 2], [4, 2]]
    c = orc.bresenham_cells(0, 0, -4, -2)
    assert c.tolist() == [[0, 0], [-1, -1], [-2, -1], [-3, -2], [-4, -2]]
    c = orc.bresenham_cells(0, 0, 2, 4)
    assert c.tolist() == [[0, 0], [1, 1], [1, 2], [2, 3], [2, 4]]
    assert orc.bresenham_cells(5, 5, 5, 5).tolist() == [[5, 5]]  # zero-length: n = 1


def test_j_bresenham_closed_form():
    # minor offset at step k is floor((floor(D/2) + k*A)/D) — the closed form the GPU may use
    rng = np.random.default_rng(8)
    for _ in range(200):
        sx, sy, ex, ey = rng.integers(-50, 50, 4)
        c = orc.bresenham_cells(int(sx), int(sy), int(ex), int(ey))
        dx, dy = abs(ex - sx), abs(ey - sy)
        D, A = max(dx, dy), min(dx, dy)
        k = np.arange(D + 1)
        minor = (D // 2 + k * A) // max(D, 1)
        sgx, sgy = (1 if ex >= sx else -1), (1 if ey >= sy else -1)
        if dx >= dy:
            exp = np.stack([sx + sgx * k, sy + sgy * minor], 1)
        else:
            exp = np.stack([sx + sgx * minor, sy + sgy * k], 1)
        assert np.array_equal(c, exp)
        assert c[-1].tolist() == [ex, ey]


# ---------------------------------------------------------------- X1-X3
def test_accumulate_single_beams_known_cells():
    g = orc.Grid.from_cells(20, 20, 1.0)          # cells are 1 m, centre (0,0); index 0 at +10
    T = np.eye(4, dtype=f32)
    # origin (0,0) -> continuous index 10 -> cell (10,10); point (3.5, 0.2) -> cell (6, 9)
    upd, cells, flags = g.accumulate(T, [3.5], [0.2], [0.0])
    assert g.get_index(0.0, 0.0) == (10, 10) and g.get_index(3.5, 0.2) == (6, 9)
    assert cells[0] == 6 + 9 * 20 and flags[0] == orc.F_VALID | orc.F_HIT
    line = orc.bresenham_cells(10, 10, 6, 9)
    assert upd == len(line) == 5
    assert g.hit.sum() == 1 and g.hit[6 + 9 * 20] == 1
    assert g.miss.sum() == 4
    for cx, cy in line[:-1]:
        assert g.miss[cx + cy * 20] == 1
    # off-map endpoint: clipped, no hit, every traversed cell (incl. the end) is a miss
    g = orc.Grid.from_cells(20, 20, 1.0)
    upd, cells, flags = g.accumulate(T, [100.0], [0.0], [0.0])
    assert flags[0] == orc.F_VALID | orc.F_CLIPPED and g.hit.sum() == 0
    assert cells[0] == 0 + 10 * 20 and g.miss.sum() == 11 and upd == 11
    # NaN beam dropped; zero-length beam = hit on the origin cell only
    g = orc.Grid.from_cells(20, 20, 1.0)
    upd, cells, flags = g.accumulate(T, [np.nan, -0.1], [0.0, -0.1], [0.0, 0.0])
    assert cells.tolist() == [-1, 10 + 10 * 20] and flags.tolist() == [0, 3]
    assert g.hit[10 + 10 * 20] == 1 and g.miss.sum() == 0 and upd == 1


def test_accumulate_range_cap_z_gate_and_labels():
    g = orc.Grid.from_cells(40, 40, 0.5)
    T = np.eye(4, dtype=f32)
    T[2, 3] = 2.0
    x = np.array([3.0, 8.0, 3.0, 3.0], f32)
    y = np.array([0.0, 0.0, 1.0, -1.0], f32)
    z = np.array([0.0, 0.0, -2.5, 0.0], f32)
    labels = np.array([0, 0, 0, -1], np.int16)
    upd, cells, flags = g.accumulate(T, x, y, z, labels, occ_mode=orc.OCC_LABELLED,
                                     z_gate=(0.0, 3.0), r_max=5.0)
    V, Hh, Cc, R = orc.F_VALID, orc.F_HIT, orc.F_CLIPPED, orc.F_RANGECAP
    assert flags.tolist() == [V | Hh, V | R, V, V]   # capped / z=-0.5 gated out / unlabelled
    # range-capped beam ends 5 m out along +x: cell of (5.0, 0)
    assert cells[1] == g.get_index(5.0, 0.0)[0] + g.get_index(5.0, 0.0)[1] * 40
    assert g.hit.sum() == 1


def test_finalize_count_order():
    g = orc.Grid.from_cells(4, 4, 1.0)
    g.hit[:] = np.arange(16)
    g.miss[:] = np.arange(16)[::-1] * 3
    l0 = np.linspace(-1, 1, 16).astype(f32)
    g.log_odds[:] = l0
    hit, miss = g.hit.copy(), g.miss.copy()
    g.finalize(k_decay=2)
    l = l0 + f32(2) * f32(-0.2)
    l = l + miss.astype(f32) * f32(-0.4)
    l = l + hit.astype(f32) * f32(1.2)
    l = np.minimum(np.maximum(l, f32(-2.0)), f32(3.6))
    assert l.dtype == f32
    assert np.array_equal(g.log_odds.view(np.uint32), l.view(np.uint32))
    assert np.all(g.hit == 0) and np.all(g.miss == 0)
    assert np.allclose(g.occupancy, 1 / (1 + np.exp(-l.astype(np.float64))), rtol=1e-6)


def test_to_occupancy_grid_order_and_scale():
    g = orc.Grid.from_cells(3, 2, 1.0)
    g.occupancy[:] = np.array([0.0, 0.5, 1.0, 0.119, 0.999, np.nan], f32)
    d = g.to_occupancy_grid()
    assert d.tolist() == [-1, 99, 11, 100, 50, 0]
