"""GPU parity tests: every call goes through the C ABI (libgridvision_b200.so) and is
compared with the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): bit-exact pixel indices, labels, cell indices, traversed-cell
counts and log-odds; occupancy (float sigmoid, GPU expf vs glibc expf) within 1e-5 relative.
"""
import numpy as np
import pytest

import grid_vision_b200 as gv
from grid_vision_b200 import synth
from oracle import gv_oracle as orc
from tests.helpers import assert_bits_equal, oracle_fuse, oracle_grid, rel_close, scan, small

pytestmark = pytest.mark.gpu

OCC_RTOL = 1e-5  # north_star: "within 1e-5 relative for accumulated float log-odds"


def set_camera(ctx, wl, T=True, ncam=1, K=None):
    K = wl.K() if K is None else K
    Ts = synth.camera_extrinsics(ncam) if T else None
    ctx.set_cameras(np.tile(np.asarray(K).reshape(1, 9), (ncam, 1)),
                    [[wl.image_w, wl.image_h]] * ncam, Ts)
    return Ts


# ------------------------------------------------------------------------- R1 + R3
@pytest.mark.parametrize("wl", [synth.C1, synth.C2], ids=["C1", "C2"])
def test_fuse_matches_oracle_full_config(ctx, wl):
    xyz = scan(wl)
    boxes = synth.make_boxes(wl)
    Ts = set_camera(ctx, wl)
    lab, pix, uv = ctx.fuse(*xyz, boxes)
    elab, epix, eu, ev = oracle_fuse(wl, xyz, boxes, Ts[0])
    assert np.array_equal(lab[0], elab)
    assert np.array_equal(pix[0], epix)
    assert_bits_equal(uv[0, 0], eu, "u")
    assert_bits_equal(uv[0, 1], ev, "v")
    assert (elab >= 0).sum() > 1000 and (epix >= 0).sum() > (elab >= 0).sum()


def test_fuse_camera_frame_is_extract_cloud_per_bbox(ctx):
    """No extrinsic set: gv_fuse is exactly the loop of extractCloudPerBBox."""
    wl = synth.C1
    rng = np.random.default_rng(11)
    n = 100003  # not a multiple of 4: exercises the scalar tail
    xyz = np.stack([rng.normal(size=n) * 8, rng.normal(size=n) * 4, rng.uniform(-3, 60, n)]).astype(np.float32)
    xyz[:, rng.integers(0, n, 500)] = np.nan
    xyz[2, rng.integers(0, n, 50)] = np.inf
    boxes = synth.make_boxes(wl, frame=3)
    set_camera(ctx, wl, T=False)
    lab, pix, uv = ctx.fuse(*xyz, boxes)
    elab, epix, eu, ev = oracle_fuse(wl, xyz, boxes, None)
    assert np.array_equal(lab[0], elab) and np.array_equal(pix[0], epix)
    assert_bits_equal(uv[0, 0], eu, "u")
    assert_bits_equal(uv[0, 1], ev, "v")
    # the per-box clouds the reference would build (stable order) from the labels
    idx, off = ctx.partition_by_label(lab[0], len(boxes))
    for b in range(len(boxes)):
        assert np.array_equal(idx[int(off[b]):int(off[b + 1])], np.flatnonzero(elab == b))


def test_fuse_kat_points_on_gpu(ctx):
    """The hand-derived edge cases of tests/test_oracle_kat.py, evaluated by the kernel."""
    wl = synth.C1
    set_camera(ctx, wl, T=False)
    f32 = np.float32
    X = np.array([0, 0, 0, 1.0, np.nextafter(f32(1), f32(0)), 1 - 2.0 ** -22, -1.0,
                  np.nextafter(f32(-1), f32(-2)), 0.25, 0.25 + 2.0 ** -20, np.nan, 0, 0], f32)
    Y = np.zeros_like(X)
    Z = np.array([1, 0.001, np.nextafter(f32(0.001), f32(1)), 1, 1, 1, 1, 1, 1, 1, 1, np.inf, -1], f32)
    boxes = orc.make_boxes([[100, 50, 260, 300], [0, 0, 416, 416]])
    lab, pix, uv = ctx.fuse(X, Y, Z, boxes)
    elab, epix, eu, ev = orc.project_label(wl.K(), 416, 416, X, Y, Z, boxes)
    assert lab[0].tolist() == elab.tolist() == [0, -1, 0, -1, -1, 1, 1, -1, 0, 1, -1, -1, -1]
    assert np.array_equal(pix[0], epix)
    assert_bits_equal(uv[0, 0], eu)


@pytest.mark.parametrize("K", [
    np.array([[320.0, 0, 320.0], [0, 320.0, 240.0], [0, 0, 1.0]]),       # the YAML default
    np.array([[207.3, 0.7, 211.9], [0.0, 209.1, 203.3], [0, 0, 1.0]]),   # skew, non-float entries
    np.array([[208.0, 0, 0.0], [0, 208.0, -0.0], [0, 0, 1.0]]),          # zero principal point
    np.array([[200.0, 0, 208.0], [0, 200.0, 208.0], [1e-3, 0, 1.0]]),    # non-affine bottom row
], ids=["yaml", "skew", "zero-pp", "projective"])
def test_fuse_generic_intrinsics(ctx, K):
    wl = small(synth.C1, rings=32, azimuth=1024).scaled(image_w=640, image_h=480)
    xyz = scan(wl)
    xyz[0, ::97] = 0.0   # X == 0 / Y == 0 rows exercise the signed-zero path
    xyz[1, ::89] = -0.0
    boxes = synth.make_boxes(wl.scaled(image_w=416, image_h=416))
    Ts = synth.camera_extrinsics(1)
    ctx.set_cameras(K.reshape(1, 9), [[640, 480]], Ts)
    lab, pix, uv = ctx.fuse(*xyz, boxes)
    elab, epix, eu, ev = oracle_fuse(wl, xyz, boxes, Ts[0], K=K)
    assert np.array_equal(lab[0], elab) and np.array_equal(pix[0], epix)
    assert_bits_equal(uv[0, 0], eu, "u")
    assert_bits_equal(uv[0, 1], ev, "v")


def test_certified_float_projection_adversarial(ctx):
    """Labels-only calls take the certified float path (fuse_point_fast).  Aim points within a
    few ulps of every decision boundary — image edges, 32-px tile edges, box edges (integer and
    fractional) — at many depths: the float path must either agree with the FP64 reference or
    defer to it, so labels stay bit-exact."""
    wl = synth.C1
    set_camera(ctx, wl, T=False)
    rng = np.random.default_rng(77)
    boxes = synth.make_boxes(wl, n=60)
    boxes["x_max"][::4] += 0.3
    boxes["y_min"][::3] -= 0.7
    edges_u = np.unique(np.concatenate([boxes["x_min"], boxes["x_max"], np.arange(0, 417, 32), [0, 416]]))
    edges_v = np.unique(np.concatenate([boxes["y_min"], boxes["y_max"], np.arange(0, 417, 32), [0, 416]]))
    X, Y, Z = [], [], []
    for _ in range(40):
        z = np.float32(rng.uniform(0.5, 80.0))
        for eu in edges_u:
            x0 = np.float32(z * (eu - 208.0) / 208.0)
            xs = x0 + np.arange(-4, 5).astype(np.float32) * np.spacing(x0)
            vv = rng.uniform(0, 416, xs.size)
            X.append(xs); Y.append((z * (vv - 208.0) / 208.0).astype(np.float32)); Z.append(np.full(xs.size, z, np.float32))
        for ev in edges_v:
            y0 = np.float32(z * (ev - 208.0) / 208.0)
            ys = y0 + np.arange(-4, 5).astype(np.float32) * np.spacing(y0)
            uu = rng.uniform(0, 416, ys.size)
            X.append((z * (uu - 208.0) / 208.0).astype(np.float32)); Y.append(ys); Z.append(np.full(ys.size, z, np.float32))
    X, Y, Z = (np.concatenate(a) for a in (X, Y, Z))
    # corners: both coordinates on an edge at once
    cu, cv = rng.choice(edges_u, 4000), rng.choice(edges_v, 4000)
    zc = rng.uniform(0.5, 80.0, 4000).astype(np.float32)
    X = np.concatenate([X, (zc * (cu - 208.0) / 208.0).astype(np.float32)])
    Y = np.concatenate([Y, (zc * (cv - 208.0) / 208.0).astype(np.float32)])
    Z = np.concatenate([Z, zc])
    lab_fast, _, _ = ctx.fuse(X, Y, Z, boxes, want_pix=False, want_uv=False)
    lab_exact, _, _ = ctx.fuse(X, Y, Z, boxes)
    elab, _, _, _ = orc.project_label(wl.K(), 416, 416, X, Y, Z, boxes)
    assert np.array_equal(lab_exact[0], elab)
    assert np.array_equal(lab_fast[0], elab)
    assert X.size > 80000 and (elab >= 0).sum() > 20000 and (elab < 0).sum() > 20000
    # huge / tiny magnitudes must leave the fast path gracefully
    Xb = np.array([1e20, -1e20, 3e38, 1e-30, 0.0, 1e16], np.float32)
    Yb = np.array([0.0, 1e20, 0.0, 1e-30, -1e19, 0.0], np.float32)
    Zb = np.array([1e20, 1e20, 3e38, 1e-2, 1e19, 1e16], np.float32)
    lf, _, _ = ctx.fuse(Xb, Yb, Zb, boxes, want_pix=False, want_uv=False)
    le, _, _, _ = orc.project_label(wl.K(), 416, 416, Xb, Yb, Zb, boxes)
    assert np.array_equal(lf[0], le)


def test_certified_index_adversarial(ctx):
    """Beams ending within a few ulps of cell and map boundaries: the fixed-point index must
    agree with grid_map's double-precision getIndex or defer to it."""
    for nx, ny, res, px, py in [(2048, 2048, 0.1, 0.0, 0.0), (500, 200, 0.1, 16.0, 0.0), (8192, 8192, 0.05, 3.3, -7.1)]:
        g = orc.Grid.from_cells(nx, ny, res, px, py)
        ctx.grid_init_cells(nx, ny, res, px, py)
        T = np.eye(4, dtype=np.float32)
        T[0, 3], T[1, 3] = px + 0.01, py - 0.02
        ctx.set_base_transform(T)
        rng = np.random.default_rng(nx)
        i = rng.integers(-2, nx + 3, 30000)
        j = rng.integers(-2, ny + 3, 30000)
        # x of the boundary between cells i-1 and i: a = i  <=>  p = half + pos - i*res
        bx = ((0.5 * g.len_x + px) - i * res).astype(np.float32)
        by = ((0.5 * g.len_y + py) - j * res).astype(np.float32)
        k = rng.integers(-3, 4, 30000)
        bx = bx + k.astype(np.float32) * np.spacing(bx)
        by_rand = rng.uniform(py - g.len_y / 2, py + g.len_y / 2, 30000).astype(np.float32)
        bx_rand = rng.uniform(px - g.len_x / 2, px + g.len_x / 2, 30000).astype(np.float32)
        X = np.concatenate([bx, bx_rand, bx])
        Y = np.concatenate([by_rand, by + k.astype(np.float32) * np.spacing(by), by])
        Z = np.zeros_like(X)
        upd, ecells, eflags = g.accumulate(T, X, Y, Z)
        cells, flags = ctx.grid_accumulate(X, Y, Z)
        assert np.array_equal(cells, ecells), (nx, res)
        assert np.array_equal(flags, eflags)
        hit, miss = ctx.grid_counts()
        assert np.array_equal(hit, g.hit) and np.array_equal(miss, g.miss)
        assert (eflags & orc.F_CLIPPED).astype(bool).sum() > 50


def test_fuse_empty_and_no_boxes(ctx):
    wl = synth.C1
    set_camera(ctx, wl)
    lab, pix, uv = ctx.fuse(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32),
                            synth.make_boxes(wl))
    assert lab.shape == (1, 0)
    xyz = scan(small(wl))
    lab, pix, uv = ctx.fuse(*xyz, np.zeros(0, synth.BOX_DTYPE))
    assert np.all(lab == -1) and (pix >= 0).sum() > 0


def test_fuse_degenerate_boxes(ctx):
    """NaN / inverted / infinite / huge bounds behave like the double compares."""
    wl = synth.C1
    set_camera(ctx, wl)
    xyz = scan(small(wl, rings=32, azimuth=1024))
    boxes = orc.make_boxes([[np.nan, 0, 416, 416], [300, 0, 100, 416], [-np.inf, -1e300, 150.5, 1e300],
                            [150.5, 100.25, np.inf, 300.75], [0, 0, 416, 416]])
    lab, pix, _ = ctx.fuse(*xyz, boxes)
    elab, epix, _, _ = oracle_fuse(wl, xyz, boxes, synth.camera_extrinsics(1)[0])
    assert np.array_equal(lab[0], elab)
    assert set(np.unique(elab)) >= {-1, 2, 3, 4} and 0 not in elab and 1 not in elab


def test_fuse_many_boxes_multiword_masks(ctx):
    """150 boxes in one camera: the image-tile prefilter needs three 64-bit words per tile."""
    wl = small(synth.C1, rings=64, azimuth=2048)
    xyz = scan(wl)
    boxes = synth.make_boxes(wl, n=150)
    rng = np.random.default_rng(3)
    boxes = boxes[rng.permutation(150)]          # overlapping boxes in arbitrary list order
    boxes["x_max"][::7] += 0.37                  # non-integer, non-float-exact bounds too
    boxes["y_min"][::5] -= 0.11
    Ts = set_camera(ctx, wl)
    lab, pix, uv = ctx.fuse(*xyz, boxes)
    elab, epix, eu, ev = oracle_fuse(wl, xyz, boxes, Ts[0])
    assert np.array_equal(lab[0], elab) and np.array_equal(pix[0], epix)
    assert len(np.unique(elab)) > 30 and elab.max() >= 128  # all three mask words in play


def test_fuse_aos32_layout(ctx):
    wl = small(synth.C1, rings=32, azimuth=1024)
    xyz = scan(wl)
    boxes = synth.make_boxes(wl)
    Ts = set_camera(ctx, wl)
    aos = synth.points_aos32(xyz)
    aos[:, 4] = np.arange(len(aos))  # intensity lane must be ignored
    lab, pix, uv = ctx.fuse_aos32(aos, boxes)
    elab, epix, eu, ev = oracle_fuse(wl, xyz, boxes, Ts[0])
    assert np.array_equal(lab[0], elab) and np.array_equal(pix[0], epix)
    assert_bits_equal(uv[0, 0], eu)


def test_fuse_six_camera_rig_C4(ctx):
    """BASELINE config 4: 6 cameras, 300 boxes, 1M-point cloud; per-camera planes equal six
    independent reference calls."""
    wl = synth.C4
    xyz = scan(wl)
    assert xyz.shape[1] == 1048576
    Ts = set_camera(ctx, wl, ncam=6)
    per_cam = [synth.make_boxes(wl, camera=c) for c in range(6)]
    boxes = np.concatenate(per_cam)
    off = np.cumsum([0] + [len(b) for b in per_cam]).astype(np.int32)
    lab, pix, uv = ctx.fuse(*xyz, boxes, off)
    for c in range(6):
        elab, epix, eu, ev = oracle_fuse(wl, xyz, per_cam[c], Ts[c])
        assert np.array_equal(lab[c], elab), f"camera {c}"
        assert np.array_equal(pix[c], epix), f"camera {c}"
        assert_bits_equal(uv[c, 0], eu, f"u cam {c}")
        assert (elab >= 0).sum() > 1000


def test_transform_points_R1(ctx):
    wl = synth.C1
    xyz = scan(small(wl))
    Ts = set_camera(ctx, wl)
    for dense in (False, True):
        got = ctx.transform_points(0, *xyz, is_dense=dense)
        exp = orc.transform_points(Ts[0], *xyz, is_dense=dense)
        for g, e in zip(got, exp):
            assert_bits_equal(g, e, f"dense={dense}")


def _random_rigid(seed, t):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T.astype(np.float32)


def test_general_rotations_no_fma_contraction(ctx):
    """SE(3) with irrational rotation entries for BOTH extrinsics: a fused multiply-add anywhere
    in the transform would change low bits of u,v / end cells (0/+-1 matrices cannot show that)."""
    wl = small(synth.C1, rings=64, azimuth=1024)
    xyz = scan(wl)
    boxes = synth.make_boxes(wl)
    Tc = (synth.camera_extrinsics(1)[0].astype(np.float64) @ _random_rigid(1, [0.3, -0.2, 0.1]).astype(np.float64)).astype(np.float32)
    Tc[3] = [0, 0, 0, 1]
    ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], Tc.reshape(1, 16))
    lab, pix, uv = ctx.fuse(*xyz, boxes)
    elab, epix, eu, ev = oracle_fuse(wl, xyz, boxes, Tc)
    assert np.array_equal(lab[0], elab) and np.array_equal(pix[0], epix)
    assert_bits_equal(uv[0, 0], eu, "u")
    assert_bits_equal(uv[0, 1], ev, "v")
    got = ctx.transform_points(0, *xyz)
    for g_, e_ in zip(got, orc.transform_points(Tc, *xyz)):
        assert_bits_equal(g_, e_)
    lab_fast, _, _ = ctx.fuse(*xyz, boxes, want_pix=False, want_uv=False)
    assert np.array_equal(lab_fast[0], elab)
    # base transform: yaw/pitch/roll + offset
    Tb = _random_rigid(2, [1.7, -0.6, 2.4])
    Tb[:3, :3] = (0.97 * np.eye(3) + 0.03 * Tb[:3, :3]).astype(np.float32)  # keep the scan roughly level
    g = oracle_grid(wl)
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
    ctx.set_base_transform(Tb)
    upd, ecells, eflags = g.accumulate(Tb, *xyz, r_max=9.0)
    cells, flags = ctx.grid_accumulate(*xyz, None, gv.accum_params(r_max=9.0))
    assert np.array_equal(cells, ecells) and np.array_equal(flags, eflags)
    hit, miss = ctx.grid_counts()
    assert np.array_equal(hit, g.hit) and np.array_equal(miss, g.miss)
    cells2, _ = ctx.grid_accumulate(*xyz, None, gv.accum_params(r_max=9.0), want_cells=False)
    hit2, miss2 = ctx.grid_counts()
    assert np.array_equal(hit2, 2 * g.hit) and np.array_equal(miss2, 2 * g.miss)  # throughput instance


def test_project_kdtree_R4(ctx):
    wl = synth.C1
    xyz = scan(small(wl, rings=32, azimuth=2048))
    Ts = set_camera(ctx, wl)
    got = ctx.project_kdtree(0, *xyz)
    cam = orc.transform_points(Ts[0], *xyz, is_dense=False)
    exp = orc.project_kdtree(wl.K(), *cam)
    assert got.shape == exp.shape and got.shape[0] > 1000
    assert_bits_equal(got, exp)


def test_partition_by_label_stable(ctx):
    rng = np.random.default_rng(5)
    for n, nb in [(1, 1), (255, 3), (100000, 50), (300001, 300)]:
        lab = rng.integers(-1, nb, n).astype(np.int16)
        idx, off = ctx.partition_by_label(lab, nb)
        for b in range(nb):
            assert np.array_equal(idx[int(off[b]):int(off[b + 1])], np.flatnonzero(lab == b))
        assert off[nb] == (lab >= 0).sum()


# ------------------------------------------------------------------------- R6 - R9
def test_grid_reference_ctor_and_get_index(ctx):
    d = ctx.grid_init_reference(50, 20, 0.1)
    assert (d.nx, d.ny, d.pos_x, d.pos_y) == (500, 200, 16.0, 0.0)
    lo, oc = ctx.grid_download()
    assert np.all(lo == 0.0) and np.all(oc == 0.5)
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    rng = np.random.default_rng(6)
    P = np.stack([rng.uniform(-12, 44, 50000), rng.uniform(-12, 12, 50000)], 1)
    edge = np.array([[41.0, 10.0], [np.nextafter(41.0, 42.0), 0], [-9.0, 0], [np.nextafter(-9.0, 0), 0],
                     [-9 + 1e-9, -10 + 1e-9], [16, 0], [np.nan, 0], [1e300, 0], [16, -np.inf]])
    P = np.concatenate([P, edge])
    got = ctx.grid_get_index(P)
    exp = np.array([g.get_index(px, py) or (-1, -1) for px, py in P], np.int32)
    assert np.array_equal(got, exp)


def test_update_map_decay_R7(ctx):
    ctx.grid_init_reference(50, 20, 0.1)
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    rng = np.random.default_rng(7)
    lo0 = rng.uniform(-2.5, 4.0, g.nx * g.ny).astype(np.float32)
    g.log_odds[:] = lo0
    ctx.grid_upload(lo0)
    for k in range(13):
        g.update_map()
        ctx.grid_update()
        lo, oc = ctx.grid_download()
        assert_bits_equal(lo, g.log_odds, f"log_odds after {k + 1} updates")
        assert rel_close(oc, g.occupancy, OCC_RTOL)
    assert lo.min() == np.float32(-2.0) and lo.max() < np.float32(3.6)


def test_update_map_poses_R8(ctx):
    ctx.grid_init_reference(50, 20, 0.1)
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    rng = np.random.default_rng(8)
    for k in range(6):
        n = int(rng.integers(0, 40))
        poses = np.stack([rng.uniform(-10, 42, n), rng.uniform(-11, 11, n), rng.uniform(0.5, 6, n),
                          rng.uniform(0.3, 3, n)], 1)
        g.update_map_poses(poses)
        ctx.grid_update_poses(poses)
        lo, oc = ctx.grid_download()
        assert_bits_equal(lo, g.log_odds, f"frame {k}")
        assert rel_close(oc, g.occupancy, OCC_RTOL)
    assert lo.max() == np.float32(3.6) or (lo > 0).sum() > 0


def test_update_map_many_overlapping_footprints(ctx):
    """> kMaxFootCand candidates in one CTA tile exercises the overflow path."""
    ctx.grid_init_reference(50, 20, 0.1)
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    rng = np.random.default_rng(9)
    n = 400
    poses = np.stack([rng.uniform(14, 18, n), rng.uniform(-2, 2, n), rng.uniform(0.5, 3, n),
                      rng.uniform(0.3, 3, n)], 1)
    g.update_map_poses(poses)
    ctx.grid_update_poses(poses)
    lo, oc = ctx.grid_download()
    assert_bits_equal(lo, g.log_odds)


def test_update_map_points_R9(ctx):
    ctx.grid_init_reference(50, 20, 0.1)
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    rng = np.random.default_rng(10)
    n = 25
    xy = np.stack([rng.uniform(-10, 42, n), rng.uniform(-11, 11, n)], 1)
    labels = rng.integers(0, 11, n).astype(np.int32)
    for _ in range(3):
        g.update_map_points(xy, labels)
        ctx.grid_update_points(xy, labels)
    lo, oc = ctx.grid_download()
    assert_bits_equal(lo, g.log_odds)
    assert rel_close(oc, g.occupancy, OCC_RTOL)
    assert (lo > -0.6).sum() > 0


def test_to_occupancy_grid_N3(ctx):
    ctx.grid_init_reference(50, 20, 0.1)
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    rng = np.random.default_rng(12)
    occ = rng.uniform(-0.1, 1.1, g.nx * g.ny).astype(np.float32)
    occ[::1001] = np.nan
    g.occupancy[:] = occ
    ctx.grid_upload(None, occ)
    assert np.array_equal(ctx.grid_to_occupancy(), g.to_occupancy_grid())


# ------------------------------------------------------------------------- X1 - X3
def accumulate_case(ctx, wl, xyz, labels, prm_kw, gv_kw):
    g = oracle_grid(wl)
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution, wl.pos_x, wl.pos_y)
    T = synth.T_base_lidar()
    ctx.set_base_transform(T)
    upd, ecells, eflags = g.accumulate(T, *xyz, labels, **prm_kw)
    cells, flags = ctx.grid_accumulate(*xyz, labels, gv.accum_params(**gv_kw))
    assert np.array_equal(cells, ecells), "end cells"
    assert np.array_equal(flags, eflags), "beam flags"
    hit, miss = ctx.grid_counts()
    assert np.array_equal(hit, g.hit), "hit plane"
    assert np.array_equal(miss, g.miss), "miss plane (traversed-cell multiset)"
    st = ctx.stats()
    assert st["beams"] == (eflags & orc.F_VALID).astype(bool).sum()
    assert st["cells_logical"] == upd
    return g, st


@pytest.mark.parametrize("wl", [synth.C1, synth.C2], ids=["C1", "C2"])
def test_accumulate_raycast_full_config(ctx, wl):
    xyz = scan(wl)
    g, st = accumulate_case(ctx, wl, xyz, None, dict(), dict())
    assert st["cells_physical"] < st["cells_logical"]  # de-duplication did something
    poses = synth.make_footprints(wl, n=12)
    corners = orc.pose_corners(poses)
    g.finalize(1, corners)
    ctx.grid_finalize(1, corners)
    lo, oc = ctx.grid_download()
    assert_bits_equal(lo, g.log_odds, "log_odds")
    assert rel_close(oc, g.occupancy, OCC_RTOL)
    hit, miss = ctx.grid_counts()
    assert not hit.any() and not miss.any()


def test_accumulate_labelled_zgate_rangecap(ctx):
    wl = small(synth.C3, rings=32, azimuth=1024, grid_nx=512, grid_ny=384, resolution=0.2)
    xyz = scan(wl, adversarial=True)
    rng = np.random.default_rng(13)
    labels = rng.integers(-1, 5, xyz.shape[1]).astype(np.int16)
    kw = dict(occ_mode=orc.OCC_LABELLED, z_gate=(0.2, 2.5), r_max=35.0)
    g, st = accumulate_case(ctx, wl, xyz, labels, kw, kw)
    assert g.hit.sum() > 100


def test_accumulate_offcentre_map_and_origin_offmap(ctx):
    wl = small(synth.C1, rings=32, azimuth=1024, pos_x=16.0, pos_y=-3.0, grid_nx=500, grid_ny=200)
    xyz = scan(wl)
    accumulate_case(ctx, wl, xyz, None, dict(), dict())
    # sensor outside the map: every beam is dropped, nothing is written
    wl2 = wl.scaled(pos_x=500.0)
    g = oracle_grid(wl2)
    ctx.grid_init_cells(wl2.grid_nx, wl2.grid_ny, wl2.resolution, wl2.pos_x, wl2.pos_y)
    ctx.set_base_transform(synth.T_base_lidar())
    upd, ecells, eflags = g.accumulate(synth.T_base_lidar(), *xyz)
    cells, flags = ctx.grid_accumulate(*xyz)
    assert upd == -1 and np.all(cells == -1) and np.all(flags == 0)
    hit, miss = ctx.grid_counts()
    assert not hit.any() and not miss.any()


def test_accumulate_pose_change_flushes(ctx):
    """Two clouds binned under two sensor poses: rays must start at each pose's own cell."""
    wl = small(synth.C1, rings=16, azimuth=512, grid_nx=400, grid_ny=400)
    g = oracle_grid(wl)
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
    T1 = synth.T_base_lidar()
    T2 = T1.copy()
    T2[0, 3], T2[1, 3] = 3.7, -2.2
    a, b = scan(wl, frame0=0), scan(wl, frame0=1)
    for T, xyz in ((T1, a), (T2, b), (T1, b)):
        g.accumulate(T, *xyz)
        ctx.set_base_transform(T)
        ctx.grid_accumulate(*xyz, want_cells=False)
    hit, miss = ctx.grid_counts()
    assert np.array_equal(hit, g.hit) and np.array_equal(miss, g.miss)


# ------------------------------------------------------------------------- batch hot path
def batch_case(wl, nframes, ragged):
    P = wl.points_per_frame
    xyz = scan(wl, frames=nframes)
    sizes = np.full(nframes, P)
    if ragged:
        rng = np.random.default_rng(21)
        sizes = rng.integers(1, P + 1, nframes)
        sizes[1] = 0  # an empty frame
    fo = np.zeros(nframes + 1, np.uint64)
    keep = []
    for f in range(nframes):
        keep.append(np.arange(f * P, f * P + sizes[f]))
        fo[f + 1] = fo[f] + np.uint64(sizes[f])
    keep = np.concatenate(keep)
    xyz = np.ascontiguousarray(xyz[:, keep])
    per_frame = [synth.make_boxes(wl, frame=f, n=int(3 + 7 * f) % (wl.boxes_per_camera + 1)) for f in range(nframes)]
    boxes = np.concatenate(per_frame)
    bo = np.cumsum([0] + [len(b) for b in per_frame]).astype(np.int32)
    return xyz, fo, per_frame, boxes, bo


@pytest.mark.parametrize("ragged", [False, True], ids=["uniform", "ragged"])
def test_process_batch_matches_per_frame_oracle(ctx, ragged):
    wl = small(synth.C3, rings=16, azimuth=1024, grid_nx=1024, grid_ny=1024)
    nframes = 7
    xyz, fo, per_frame, boxes, bo = batch_case(wl, nframes, ragged)
    Ts = set_camera(ctx, wl)
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
    T = synth.T_base_lidar()
    ctx.set_base_transform(T)
    prm = dict(occ_mode=orc.OCC_LABELLED, r_max=wl.r_max)
    labels = ctx.process_batch(*xyz, fo, boxes, bo, gv.accum_params(**prm))
    g = oracle_grid(wl)
    for f in range(nframes):
        s, e = int(fo[f]), int(fo[f + 1])
        fx = xyz[:, s:e]
        elab, _, _, _ = oracle_fuse(wl, fx, per_frame[f], Ts[0])
        assert np.array_equal(labels[s:e], elab), f"frame {f}"
        g.accumulate(T, *fx, elab, **prm)
    hit, miss = ctx.grid_counts()
    assert np.array_equal(hit, g.hit) and np.array_equal(miss, g.miss)
    g.finalize(nframes)
    ctx.grid_finalize(nframes)
    lo, oc = ctx.grid_download()
    assert_bits_equal(lo, g.log_odds)
    assert rel_close(oc, g.occupancy, OCC_RTOL)


def test_process_batch_device_pointers_equal_host_path(ctx):
    import torch
    wl = small(synth.C3, rings=16, azimuth=1024, grid_nx=1024, grid_ny=1024)
    xyz, fo, per_frame, boxes, bo = batch_case(wl, 5, ragged=True)
    set_camera(ctx, wl)
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
    ctx.set_base_transform(synth.T_base_lidar())
    prm = gv.accum_params(r_max=wl.r_max)
    labels_h = ctx.process_batch(*xyz, fo, boxes, bo, prm)
    hit_h, miss_h = ctx.grid_counts()
    ctx.grid_reset()
    d = torch.from_numpy(xyz).cuda()
    d_boxes = torch.from_numpy(boxes.view(np.uint8).copy()).cuda()
    d_lab = torch.full((xyz.shape[1],), -7, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    ctx.process_batch(d[0], d[1], d[2], fo, d_boxes, bo, prm, d_lab)
    ctx.synchronize()
    hit_d, miss_d = ctx.grid_counts()
    assert np.array_equal(d_lab.cpu().numpy(), labels_h)
    assert np.array_equal(hit_d, hit_h) and np.array_equal(miss_d, miss_h)


# ------------------------------------------------------------------------- size-independent properties
def test_properties_at_scale(ctx):
    """64 full C3 frames (8.4M beams): too slow for the brute-force oracle in a unit test, so
    check invariants — count conservation, batch-split linearity, frame-order invariance."""
    import torch
    wl = synth.C3
    nframes = 64
    P = wl.points_per_frame
    xyz = synth.make_scans(wl, frames=nframes, device="cuda")
    boxes = np.concatenate([synth.make_boxes(wl, frame=f) for f in range(nframes)])
    d_boxes = torch.from_numpy(boxes.view(np.uint8).copy()).cuda()
    bo = (np.arange(nframes + 1) * wl.boxes_per_camera).astype(np.int32)
    fo = (np.arange(nframes + 1) * P).astype(np.uint64)
    set_camera(ctx, wl)
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
    ctx.set_base_transform(synth.T_base_lidar())
    prm = gv.accum_params(r_max=wl.r_max)
    lab = torch.empty(nframes * P, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    ctx.process_batch(xyz[0], xyz[1], xyz[2], fo, d_boxes, bo, prm, lab)
    hit, miss = ctx.grid_counts()
    st = ctx.stats()
    finite = torch.isfinite(xyz).all(dim=0).sum().item()
    assert st["beams"] == finite
    assert int(hit.sum()) + int(miss.sum()) == st["cells_logical"]     # every traversed cell counted once
    assert st["cells_physical"] < st["cells_logical"] // 4             # de-duplication
    assert hit.min() >= 0 and miss.min() >= 0
    # linearity: two half batches in reverse frame order give the same integer planes
    ctx.grid_reset()
    h = nframes // 2
    for lo_f, hi_f in ((h, nframes), (0, h)):
        sl = slice(lo_f * P, hi_f * P)
        ctx.process_batch(xyz[0, sl], xyz[1, sl], xyz[2, sl], fo[: hi_f - lo_f + 1],
                          d_boxes[lo_f * wl.boxes_per_camera * 40:], bo[: hi_f - lo_f + 1], prm, None)
        ctx.grid_raycast_flush()
    hit2, miss2 = ctx.grid_counts()
    assert np.array_equal(hit, hit2) and np.array_equal(miss, miss2)
    # labels of frame 0 agree with the oracle
    x0 = xyz[:, :P].cpu().numpy()
    elab, _, _, _ = oracle_fuse(wl, x0, boxes[: wl.boxes_per_camera], synth.camera_extrinsics(1)[0])
    assert np.array_equal(lab[:P].cpu().numpy(), elab)
    # idempotence of the clamp: finalising twice with k_decay=0 changes nothing
    ctx.grid_finalize(0)
    lo1, oc1 = ctx.grid_download()
    ctx.grid_finalize(0)
    lo2, oc2 = ctx.grid_download()
    assert np.array_equal(lo1, lo2) and np.array_equal(oc1, oc2)
    assert lo1.min() >= -2.0 and lo1.max() <= np.float32(3.6)


def test_no_silent_fallback():
    """The product must be the CUDA library: the binding exposes no alternative path."""
    from grid_vision_b200 import _lib
    lib = _lib.load()
    assert lib._name.endswith("libgridvision_b200.so")
    import grid_vision_b200.context as c
    src = open(c.__file__).read()
    assert "oracle" not in src
