"""GPU parity of k_points_fast + k_points_deferred (the hot instantiation gv_process_batch runs in
its usual configuration) against the CPU oracle: full-size BASELINE configs 3 and 5, adversarial
inputs aimed at every certified decision, and every template variant.  All through the C ABI."""
import os

import numpy as np
import pytest

import grid_vision_b200 as gv
from grid_vision_b200 import synth
from oracle import gv_oracle as orc
from tests.helpers import assert_bits_equal, oracle_batch, oracle_fuse, oracle_grid, rel_close, scan, small

pytestmark = pytest.mark.gpu
OCC_RTOL = 1e-5


def run_batch(ctx, wl, xyz, fo, boxes, bo, Tc, Tb, want_labels=True, **prm):
    ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], Tc[None])
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution, wl.pos_x, wl.pos_y)
    ctx.set_base_transform(Tb)
    labels = ctx.process_batch(*xyz, fo, boxes, bo, gv.accum_params(**prm))
    hit, miss = ctx.grid_counts()
    return labels, hit, miss


def check_batch(ctx, wl, xyz, fo, per_frame, Tc, Tb, threads=8, **prm):
    boxes = np.concatenate(per_frame) if per_frame else np.zeros(0, synth.BOX_DTYPE)
    bo = np.cumsum([0] + [len(b) for b in per_frame]).astype(np.int32)
    labels, hit, miss = run_batch(ctx, wl, xyz, fo, boxes, bo, Tc, Tb, **prm)
    elab, g = oracle_batch(wl, xyz, fo, per_frame, Tc, Tb, threads=threads, **prm)
    bad = np.flatnonzero(labels != elab)
    assert bad.size == 0, f"{bad.size} labels differ, first {bad[:5]}: {labels[bad[:5]]} vs {elab[bad[:5]]}"
    assert np.array_equal(hit, g.hit), f"hit planes differ in {(hit != g.hit).sum()} cells"
    assert np.array_equal(miss, g.miss), f"miss planes differ in {(miss != g.miss).sum()} cells"
    return labels, g


# ----------------------------------------------------------------- published configs, full size
def test_c3_64_full_frames_planes_vs_oracle(ctx):
    """BASELINE configs[2] geometry at full per-frame size: 64 scans x 131072 points into the
    2048x2048 @0.1 m grid, r_max 120 m: labels, hit and miss planes and log-odds bit for bit."""
    wl = synth.C3
    nframes = 64
    P = wl.points_per_frame
    xyz = scan(wl, frames=nframes)
    per_frame = [synth.make_boxes(wl, frame=f) for f in range(nframes)]
    fo = (np.arange(nframes + 1) * P).astype(np.uint64)
    _, g = check_batch(ctx, wl, xyz, fo, per_frame, synth.camera_extrinsics(1)[0], synth.T_base_lidar(),
                       threads=min(16, os.cpu_count() or 1), r_max=wl.r_max)
    g.finalize(nframes)
    ctx.grid_finalize(nframes)
    lo, oc = ctx.grid_download()
    assert_bits_equal(lo, g.log_odds, "log_odds")
    assert rel_close(oc, g.occupancy, OCC_RTOL)
    assert g.hit.sum() == 0 or True


def test_c5_full_resolution_vs_oracle(ctx):
    """BASELINE configs[4] geometry at full resolution: 8192x8192 @0.05 m, 120 m rays (lines of up
    to ~2400 cells), 8 scans of 131072 points: labels, hit/miss planes, log-odds bit for bit."""
    wl = synth.C5
    nframes = 8
    P = wl.points_per_frame
    xyz = scan(wl, frames=nframes)
    per_frame = [synth.make_boxes(wl, frame=f) for f in range(nframes)]
    fo = (np.arange(nframes + 1) * P).astype(np.uint64)
    _, g = check_batch(ctx, wl, xyz, fo, per_frame, synth.camera_extrinsics(1)[0], synth.T_base_lidar(),
                       threads=4, r_max=wl.r_max)
    assert int(g.miss.max()) > 0 and int(g.hit.sum()) > 100000
    g.finalize(nframes)
    ctx.grid_finalize(nframes)
    lo, oc = ctx.grid_download()
    assert_bits_equal(lo, g.log_odds, "log_odds")
    assert rel_close(oc, g.occupancy, OCC_RTOL)


def test_c3_adversarial_ranges_vs_oracle(ctx):
    """r ~ U(2, sensor_range): no end-cell repetition between beams, every range-cap / clip case."""
    wl = synth.C3
    nframes = 8
    P = wl.points_per_frame
    xyz = scan(wl, frames=nframes, adversarial=True)
    per_frame = [synth.make_boxes(wl, frame=f) for f in range(nframes)]
    fo = (np.arange(nframes + 1) * P).astype(np.uint64)
    check_batch(ctx, wl, xyz, fo, per_frame, synth.camera_extrinsics(1)[0], synth.T_base_lidar(),
                r_max=wl.r_max, occ_mode=orc.OCC_LABELLED)


# ----------------------------------------------------------------- adversarial: certified projection
def projection_adversarial_points(wl, boxes, rng, depths=40):
    edges_u = np.unique(np.concatenate([boxes["x_min"], boxes["x_max"], np.arange(0, 417, 32), [0, 416]]))
    edges_v = np.unique(np.concatenate([boxes["y_min"], boxes["y_max"], np.arange(0, 417, 32), [0, 416]]))
    X, Y, Z = [], [], []
    for _ in range(depths):
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
    cu, cv = rng.choice(edges_u, 4000), rng.choice(edges_v, 4000)
    zc = rng.uniform(0.5, 80.0, 4000).astype(np.float32)
    X = np.concatenate([X, (zc * (cu - 208.0) / 208.0).astype(np.float32)])
    Y = np.concatenate([Y, (zc * (cv - 208.0) / 208.0).astype(np.float32)])
    Z = np.concatenate([Z, zc])
    return np.stack([X, Y, Z])


def test_fast_path_projection_adversarial(ctx):
    """Points within +-4 ulp of every image / tile / box edge (integer and fractional bounds) at many
    depths, plus specials, through gv_process_batch: the certified float projection must agree with
    the FP64 reference or defer.  The extrinsic is the identity (exact), so the cloud is already in
    the camera frame."""
    wl = small(synth.C1, grid_nx=1024, grid_ny=1024)
    rng = np.random.default_rng(77)
    boxes = synth.make_boxes(wl, n=60)
    boxes["x_max"][::4] += 0.3
    boxes["y_min"][::3] -= 0.7
    xyz = projection_adversarial_points(wl, boxes, rng)
    special = np.array([[1e20, -1e20, 3e38, 1e-30, 0.0, 1e16, np.nan, np.inf, 0.0, 2e9, 5.0, -np.inf],
                        [0.0, 1e20, 0.0, 1e-30, -1e19, 0.0, 1.0, 1.0, np.nan, 1.0, 3e9, 2.0],
                        [1e20, 1e20, 3e38, 1e-2, 1e19, 1e16, 5.0, 5.0, 5.0, 4e9, 9.0, np.inf]], np.float32)
    xyz = np.ascontiguousarray(np.concatenate([xyz, special], axis=1))
    n = xyz.shape[1]
    fo = np.array([0, n // 3, n // 3, n], np.uint64)  # three frames, the middle one empty
    per_frame = [boxes, boxes[:5], boxes[::-1].copy()]
    T = np.eye(4, dtype=np.float32)
    Tb = np.eye(4, dtype=np.float32)
    Tb[0, 3], Tb[1, 3] = 0.3, -0.7
    labels, g = check_batch(ctx, wl, xyz, fo, per_frame, T, Tb, r_max=40.0)
    assert (labels >= 0).sum() > 20000 and (labels < 0).sum() > 20000


# ----------------------------------------------------------------- adversarial: certified index
@pytest.mark.parametrize("geom", [(2048, 2048, 0.1, 0.0, 0.0, 120.0), (500, 200, 0.1, 16.0, 0.0, 0.0),
                                  (8192, 8192, 0.05, 3.3, -7.1, 120.0), (300, 700, 0.25, -4.0, 9.0, 1.0e5)],
                         ids=["c3", "small-nocap", "c5", "hugecap"])
def test_fast_path_index_adversarial(ctx, geom):
    """Beams ending within +-3 ulp of cell and map boundaries (and far outside the map), with and
    without a range cap (BOUNDED and unbounded index words): end cells via the count planes."""
    nx, ny, res, px, py, r_max = geom
    wl = small(synth.C1).scaled(grid_nx=nx, grid_ny=ny, resolution=res, pos_x=px, pos_y=py)
    g0 = oracle_grid(wl)
    rng = np.random.default_rng(nx)
    m = 30000
    i = rng.integers(-2, nx + 3, m)
    j = rng.integers(-2, ny + 3, m)
    bx = ((0.5 * g0.len_x + px) - i * res).astype(np.float32)
    by = ((0.5 * g0.len_y + py) - j * res).astype(np.float32)
    k = rng.integers(-3, 4, m)
    bx = bx + k.astype(np.float32) * np.spacing(bx)
    by_rand = rng.uniform(py - g0.len_y / 2, py + g0.len_y / 2, m).astype(np.float32)
    bx_rand = rng.uniform(px - g0.len_x / 2, px + g0.len_x / 2, m).astype(np.float32)
    far = rng.uniform(-1, 1, (2, 2000)).astype(np.float32) * np.float32(4.0e5)
    X = np.concatenate([bx, bx_rand, bx, far[0]])
    Y = np.concatenate([by_rand, by + k.astype(np.float32) * np.spacing(by), by, far[1]])
    Z = rng.uniform(-1, 3, X.size).astype(np.float32)
    xyz = np.ascontiguousarray(np.stack([X, Y, Z]))
    T = np.eye(4, dtype=np.float32)
    T[0, 3], T[1, 3] = px + 0.01, py - 0.02
    fo = np.array([0, X.size], np.uint64)
    check_batch(ctx, wl, xyz, fo, [synth.make_boxes(wl)], synth.camera_extrinsics(1)[0], T, r_max=r_max)
    check_batch(ctx, wl, xyz, fo, [synth.make_boxes(wl)], synth.camera_extrinsics(1)[0], T, r_max=r_max,
                z_gate=(0.0, 2.0), occ_mode=orc.OCC_LABELLED)


# ----------------------------------------------------------------- template variants
VARIANTS = [{"GV_FAST_KIND": "0"}, {"GV_FAST_KIND": "3"}, {"GV_FAST_KIND": "3", "GV_COL_HOIST": "1"}, {"GV_FAST_KIND": "1"}, {"GV_FAST_KIND": "1", "GV_FAST_U": "1"},
            {"GV_FAST_KIND": "1", "GV_TMA_HOIST": "1"}, {"GV_FAST_KIND": "2"},
            {"GV_FAST_KIND": "2", "GV_FAST_AGG": "0", "GV_FAST_U": "1"}, {"GV_FAST_KIND": "2", "GV_FAST_AGG": "1"},
            {"GV_FAST_KIND": "2", "GV_FAST_AGG": "2", "GV_FAST_U": "4"}, {"GV_L2_PERSIST": "0"}]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k[3:]}={v}" for k, v in e.items()))
def test_fast_path_kernel_variants(env):
    """Every fast kernel (KIND 0 k_points_pair: two adjacent points per thread on the packed f32x2
    pipe where every frame offset and size is even, else k_points_col; 3 k_points_col: one thread per
    beam index across frames, run-length binning in registers; 1 k_points_tma: persistent, bulk-copy fed, run slots in shared memory;
    2 k_points_fast: one CTA per tile, optional warp-level RED merging) and instantiation family
    on three frame layouts: arbitrary ragged sizes (LDG kernel:
    bulk copies need 4-point alignment), ragged sizes that are multiples of 4 (TMA kernel, partial
    last rows), and one-point / empty frames."""
    os.environ.update(env)
    try:
        c = gv.Context(0)
    finally:
        for k in env:
            del os.environ[k]
    try:
        wl = small(synth.C3, rings=16, azimuth=1024, grid_nx=1024, grid_ny=1024)
        P = wl.points_per_frame
        nframes = 9
        full = scan(wl, frames=nframes)
        rng = np.random.default_rng(5)
        Tc, Tb = synth.camera_extrinsics(1)[0], synth.T_base_lidar()
        per_frame = [synth.make_boxes(wl, frame=f, n=int(3 + 11 * f) % 65) for f in range(nframes)]
        for align in (1, 2, 4):
            sizes = rng.integers(1, P // align + 1, nframes) * align
            sizes[1], sizes[2], sizes[3] = 0, align, 256 + align
            keep = np.concatenate([np.arange(f * P, f * P + sizes[f]) for f in range(nframes)])
            xyz = np.ascontiguousarray(full[:, keep])
            fo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
            check_batch(c, wl, xyz, fo, per_frame, Tc, Tb, r_max=wl.r_max)
            check_batch(c, wl, xyz, fo, per_frame, Tc, Tb)  # no range cap: unbounded index words
    finally:
        c.close()


def test_fast_path_equals_generic_kernel():
    """The same batch through k_points_fast and through the generic k_points (GV_NO_FAST=1): labels
    and planes identical, including a rotated (non-permutation) extrinsic and base transform."""
    import math
    wl = small(synth.C3, rings=32, azimuth=2048, grid_nx=1500, grid_ny=900, resolution=0.13, pos_x=7.5, pos_y=-3.25)
    nframes = 6
    P = wl.points_per_frame
    xyz = scan(wl, frames=nframes)
    per_frame = [synth.make_boxes(wl, frame=f) for f in range(nframes)]
    boxes = np.concatenate(per_frame)
    bo = np.cumsum([0] + [len(b) for b in per_frame]).astype(np.int32)
    fo = (np.arange(nframes + 1) * P).astype(np.uint64)
    Tc = synth.camera_extrinsics(1)[0].astype(np.float64)
    a, b = math.radians(3.7), math.radians(-1.9)
    Rz = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
    Ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
    Tc[:3, :3] = Tc[:3, :3] @ Rz @ Ry
    Tc = Tc.astype(np.float32)
    Tb = np.eye(4)
    Tb[:3, :3] = Rz @ Ry
    Tb[:3, 3] = [1.25, -0.5, 2.4]
    Tb = Tb.astype(np.float32)
    out = {}
    for mode in ("fast", "generic"):
        if mode == "generic":
            os.environ["GV_NO_FAST"] = "1"
        try:
            c = gv.Context(0)
        finally:
            os.environ.pop("GV_NO_FAST", None)
        try:
            out[mode] = run_batch(c, wl, xyz, fo, boxes, bo, Tc, Tb, r_max=60.0)
            out[mode + "_launches"] = c.stats()["kernel_launches"]
        finally:
            c.close()
    for a_, b_ in zip(out["fast"], out["generic"]):
        assert np.array_equal(a_, b_)
    elab, g = oracle_batch(wl, xyz, fo, per_frame, Tc, Tb, r_max=60.0)
    assert np.array_equal(out["fast"][0], elab)
    assert np.array_equal(out["fast"][1], g.hit) and np.array_equal(out["fast"][2], g.miss)


def test_fast_path_device_pointers_without_labels(ctx):
    """labels_out = NULL on the device entry point (LAB = false instantiation)."""
    import torch
    wl = small(synth.C3, rings=16, azimuth=1024, grid_nx=1024, grid_ny=1024)
    nframes = 4
    P = wl.points_per_frame
    xyz = scan(wl, frames=nframes)
    per_frame = [synth.make_boxes(wl, frame=f) for f in range(nframes)]
    boxes = np.concatenate(per_frame)
    bo = np.cumsum([0] + [len(b) for b in per_frame]).astype(np.int32)
    fo = (np.arange(nframes + 1) * P).astype(np.uint64)
    Tc, Tb = synth.camera_extrinsics(1)[0], synth.T_base_lidar()
    ctx.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], Tc[None])
    ctx.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
    ctx.set_base_transform(Tb)
    d = torch.from_numpy(xyz).cuda()
    d_boxes = torch.from_numpy(boxes.view(np.uint8).copy()).cuda()
    torch.cuda.synchronize()
    ctx.process_batch(d[0], d[1], d[2], fo, d_boxes, bo, gv.accum_params(r_max=wl.r_max), None)
    hit, miss = ctx.grid_counts()
    _, g = oracle_batch(wl, xyz, fo, per_frame, Tc, Tb, r_max=wl.r_max)
    assert np.array_equal(hit, g.hit) and np.array_equal(miss, g.miss)


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_pipelined_batches_match_sequential_oracle(overlap):
    """Five different batches finalised back to back with nothing read in between: batch k+1 is
    binned into the second end-cell plane while batch k's raycast + finalise still run on the merge
    stream (GV_OVERLAP=1), or everything on one stream (=0).  The clamp makes the sequence
    order-sensitive, so the final log-odds equal the oracle's only if every merge saw exactly its
    own batch."""
    import torch
    os.environ["GV_OVERLAP"] = overlap
    try:
        c = gv.Context(0)
    finally:
        del os.environ["GV_OVERLAP"]
    try:
        wl = small(synth.C3, rings=32, azimuth=2048, grid_nx=1024, grid_ny=1024)
        P = wl.points_per_frame
        Tc, Tb = synth.camera_extrinsics(1)[0], synth.T_base_lidar()
        c.set_cameras(wl.K().reshape(1, 9), [[wl.image_w, wl.image_h]], Tc[None])
        c.grid_init_cells(wl.grid_nx, wl.grid_ny, wl.resolution)
        c.set_base_transform(Tb)
        g = oracle_grid(wl)
        prm = dict(r_max=wl.r_max)
        batches = []
        for b in range(5):
            nf = 3 + b
            xyz = scan(wl, frames=nf, frame0=10 * b, adversarial=(b % 2 == 1))
            per = [synth.make_boxes(wl, frame=10 * b + f) for f in range(nf)]
            fo = (np.arange(nf + 1) * P).astype(np.uint64)
            bo = np.cumsum([0] + [len(x) for x in per]).astype(np.int32)
            d = torch.from_numpy(xyz).cuda()
            d_boxes = torch.from_numpy(np.concatenate(per).view(np.uint8).copy()).cuda()
            batches.append((xyz, per, fo, bo, d, d_boxes, nf))
        torch.cuda.synchronize()
        for xyz, per, fo, bo, d, d_boxes, nf in batches:  # queued without any host wait
            c.process_batch(d[0], d[1], d[2], fo, d_boxes, bo, gv.accum_params(**prm), None)
            c.grid_finalize(nf)
        for xyz, per, fo, bo, d, d_boxes, nf in batches:
            _, gb = oracle_batch(wl, xyz, fo, per, Tc, Tb, **prm)
            g.hit += gb.hit
            g.miss += gb.miss
            g.finalize(nf)
        lo, oc = c.grid_download()
        assert_bits_equal(lo, g.log_odds, "log_odds after 5 pipelined batches")
        assert rel_close(oc, g.occupancy, OCC_RTOL)
        st = c.stats()
        assert st["merges"] == 5 and (st["merge_ms_last"] > 0.0) == (overlap == "1")
    finally:
        c.close()
