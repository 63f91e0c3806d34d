"""Replays the committed golden fixtures (tests/golden/*.npz, made by make_golden.py):
on the CPU against the oracle (guards the oracle against regressions), on the GPU against the
CUDA path through the C ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import gv_oracle as orc
from tests.helpers import assert_bits_equal, rel_close

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def grid_from(d):
    nx, ny = (int(v) for v in d["grid"])
    return orc.Grid.from_cells(nx, ny, float(d["res"]), float(d["pos"][0]), float(d["pos"][1]))


def grid_kw(d):
    zg = None if np.isnan(d["z_gate"][0]) else (float(d["z_gate"][0]), float(d["z_gate"][1]))
    return dict(occ_mode=int(d["occ_mode"]), z_gate=zg, r_max=float(d["r_max"]))


def test_fixtures_present():
    assert len(glob.glob(os.path.join(G, "*.npz"))) >= 5


@pytest.mark.parametrize("name", ["fuse_c1_small.npz", "fuse_c2_small.npz"])
def test_oracle_fuse_golden(name):
    d = load(name)
    cam = orc.transform_points(d["T_cam"], *d["xyz"])
    lab, pix, u, v = orc.project_label(d["K"], int(d["wh"][0]), int(d["wh"][1]), *cam, d["boxes"])
    assert np.array_equal(lab, d["label"]) and np.array_equal(pix, d["pix"])
    assert_bits_equal(u, d["u"])
    assert_bits_equal(v, d["v"])
    assert_bits_equal(orc.project_kdtree(d["K"], *cam), d["uvz"])


@pytest.mark.parametrize("name", ["grid_c1_small.npz", "grid_c3_labelled.npz"])
def test_oracle_grid_golden(name):
    d = load(name)
    g = grid_from(d)
    upd, cells, flags = g.accumulate(d["T_base"], *d["xyz"], d["labels"], **grid_kw(d))
    assert upd == int(d["updates"])
    assert np.array_equal(cells, d["cells"]) and np.array_equal(flags, d["flags"])
    assert np.array_equal(g.hit, d["hit"]) and np.array_equal(g.miss, d["miss"])
    g.finalize(2, d["corners"])
    assert_bits_equal(g.log_odds, d["log_odds"])
    assert_bits_equal(g.occupancy, d["occupancy"])


def test_oracle_updates_golden():
    d = load("updates_reference_grid.npz")
    g = orc.Grid(reference_ctor=(50, 20, 0.1))
    g.log_odds[:] = d["lo0"]
    g.update_map()
    assert_bits_equal(g.log_odds, d["after_r7"])
    g.update_map_poses(d["poses"])
    assert_bits_equal(g.log_odds, d["after_r8"])
    g.update_map_points(d["xy"], d["labels"])
    assert_bits_equal(g.log_odds, d["after_r9"])
    assert_bits_equal(g.occupancy, d["occupancy"])


# ------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fuse_c1_small.npz", "fuse_c2_small.npz"])
def test_gpu_fuse_golden(ctx, name):
    d = load(name)
    ctx.set_cameras(d["K"].reshape(1, 9), [[int(d["wh"][0]), int(d["wh"][1])]], d["T_cam"].reshape(1, 16))
    lab, pix, uv = ctx.fuse(*d["xyz"], d["boxes"])
    assert np.array_equal(lab[0], d["label"]) and np.array_equal(pix[0], d["pix"])
    assert_bits_equal(uv[0, 0], d["u"])
    assert_bits_equal(uv[0, 1], d["v"])
    assert_bits_equal(ctx.project_kdtree(0, *d["xyz"]), d["uvz"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["grid_c1_small.npz", "grid_c3_labelled.npz"])
def test_gpu_grid_golden(ctx, name):
    import grid_vision_b200 as gv
    d = load(name)
    nx, ny = (int(v) for v in d["grid"])
    ctx.grid_init_cells(nx, ny, float(d["res"]), float(d["pos"][0]), float(d["pos"][1]))
    ctx.set_base_transform(d["T_base"])
    cells, flags = ctx.grid_accumulate(*d["xyz"], d["labels"], gv.accum_params(**grid_kw(d)))
    assert np.array_equal(cells, d["cells"]) and np.array_equal(flags, d["flags"])
    hit, miss = ctx.grid_counts()
    assert np.array_equal(hit, d["hit"]) and np.array_equal(miss, d["miss"])
    ctx.grid_finalize(2, d["corners"])
    lo, oc = ctx.grid_download()
    assert_bits_equal(lo, d["log_odds"])
    assert rel_close(oc, d["occupancy"], 1e-5)


@pytest.mark.gpu
def test_gpu_updates_golden(ctx):
    d = load("updates_reference_grid.npz")
    ctx.grid_init_reference(50, 20, 0.1)
    ctx.grid_upload(d["lo0"])
    ctx.grid_update()
    assert_bits_equal(ctx.grid_download()[0], d["after_r7"])
    ctx.grid_update_poses(d["poses"])
    assert_bits_equal(ctx.grid_download()[0], d["after_r8"])
    ctx.grid_update_points(d["xy"], d["labels"])
    lo, oc = ctx.grid_download()
    assert_bits_equal(lo, d["after_r9"])
    assert rel_close(oc, d["occupancy"], 1e-5)
