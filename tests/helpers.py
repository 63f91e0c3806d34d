"""Shared builders for the parity tests: seeded synthetic cases + oracle-side evaluation."""
from __future__ import annotations

import numpy as np

from grid_vision_b200 import synth
from oracle import gv_oracle as orc


def bits(a: np.ndarray) -> np.ndarray:
    """float32 bit patterns with every NaN mapped to one canonical pattern."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def assert_bits_equal(a, b, what=""):
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        bad = np.flatnonzero(ba != bb)
        raise AssertionError(f"{what}: {bad.size} of {ba.size} float32 values differ, first at "
                             f"{bad[0]}: {np.ravel(a)[bad[0]]!r} vs {np.ravel(b)[bad[0]]!r}")


def scan(wl: synth.Workload, frames=1, frame0=0, adversarial=False, nan_fraction=0.01):
    xyz = synth.make_scans(wl, frame0=frame0, frames=frames, device="cpu",
                           nan_fraction=nan_fraction, adversarial=adversarial)
    return xyz.numpy()


def small(wl: synth.Workload, rings=16, azimuth=512, **kw) -> synth.Workload:
    return wl.scaled(rings=rings, azimuth=azimuth, **kw)


def oracle_fuse(wl, xyz, boxes, T_cam=None, K=None, is_dense=False):
    """R1 then R3 on the oracle: returns label, pix, u, v for one camera."""
    K = wl.K() if K is None else K
    x, y, z = xyz
    if T_cam is not None:
        x, y, z = orc.transform_points(T_cam, x, y, z, is_dense=is_dense)
    return orc.project_label(K, wl.image_w, wl.image_h, x, y, z, boxes)


def oracle_grid(wl) -> orc.Grid:
    return orc.Grid.from_cells(wl.grid_nx, wl.grid_ny, wl.resolution, wl.pos_x, wl.pos_y)


def rel_close(a, b, rtol=1e-5):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)))


def oracle_batch(wl, xyz, fo, per_frame_boxes, T_cam, T_base, threads=8, **prm):
    """Frame-parallel oracle pass over a batch (per-thread private count grids, exact integer
    merge): returns (labels int16 [n], merged Grid).  prm: occ_mode / z_gate / r_max."""
    from concurrent.futures import ThreadPoolExecutor
    nframes = len(fo) - 1
    threads = max(1, min(threads, nframes))
    grids = [oracle_grid(wl) for _ in range(threads)]
    labels = np.full(int(fo[-1]), -9, np.int16)

    def work(t):
        for f in range(t, nframes, threads):
            s, e = int(fo[f]), int(fo[f + 1])
            fx = xyz[:, s:e]
            lab, _, _, _ = oracle_fuse(wl, fx, per_frame_boxes[f], T_cam)
            labels[s:e] = lab
            grids[t].accumulate(T_base, *fx, lab, want_cells=False, **prm)

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    for g in grids[1:]:
        grids[0].hit += g.hit
        grids[0].miss += g.miss
    return labels, grids[0]
