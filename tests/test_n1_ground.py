""""Next" row N1: RANSAC ground-plane removal (ref: src/cloud_detections.cpp:105-138).
PCL's random sample stream cannot be reproduced offline, so the oracle fixes a deterministic
hypothesis set (counter hash) and the GPU must reproduce it: same winning plane, same kept set
up to points that sit within float rounding of the distance threshold."""
import numpy as np
import pytest

from grid_vision_b200 import synth
from oracle import gv_oracle as orc

f32 = np.float32


def camera_scan(wl=synth.C1, frame=0):
    xyz = synth.make_scans(wl, frame0=frame, frames=1).numpy()
    return orc.transform_points(synth.camera_extrinsics(1)[0], *xyz)


def test_oracle_finds_the_ground_plane():
    x, y, z = camera_scan()
    keep, plane, bh, bs = orc.segment_ground(x, y, z)
    # ground is z_lidar = -1.8 -> y_cam = 1.4 (camera 0.4 m below the LiDAR origin, y down)
    assert abs(abs(plane[1]) - 1.0) < 1e-4 and abs(abs(plane[3]) - 1.4) < 1e-3
    removed = ~keep
    assert removed.sum() == pytest.approx(bs, rel=0.01) and removed.sum() > 50000
    assert np.all(np.abs(y[removed] - 1.4) < 0.05)
    assert np.isnan(x[keep]).sum() == np.isnan(x).sum()      # NaN points are not inliers: they stay
    # deterministic: same seed -> same answer; different seed -> same plane (it dominates the scan)
    k2, p2, _, _ = orc.segment_ground(x, y, z)
    assert np.array_equal(keep, k2) and np.array_equal(plane, p2)
    k3, p3, _, _ = orc.segment_ground(x, y, z, seed=7)
    assert np.allclose(np.abs(p3), np.abs(plane), atol=1e-3) and (k3 != keep).mean() < 1e-3


def degenerate_cloud():
    # every point on one line: every sampled triangle is degenerate -> no hypothesis, no model
    # (exactly: along one axis with integer coordinates, so every cross product is exactly zero;
    # a slanted float line leaves rounding-sized normals and would count as a plane)
    t = np.arange(200).astype(f32)
    return t, np.zeros(200, f32), np.zeros(200, f32)


def test_oracle_no_model_returns_none():
    keep, plane, bh, bs = orc.segment_ground(*degenerate_cloud())
    assert keep is None and bs == 0 and bh == -1
    one = np.ones(2, f32)
    assert orc.segment_ground(one, one, one)[0] is None
    # sparse scatter: the best plane is a sampled triangle plus chance inliers; still a model
    rng = np.random.default_rng(0)
    x, y, z = (rng.uniform(-50, 50, 300).astype(f32) for _ in range(3))
    keep, plane, bh, bs = orc.segment_ground(x, y, z)
    assert keep is not None and 3 <= bs < 10 and (~keep).sum() >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("frame,seed", [(0, 12345), (3, 99)])
def test_gpu_segment_ground_matches_oracle(ctx, frame, seed):
    x, y, z = camera_scan(synth.C1, frame)
    keep, plane, bh, bs = orc.segment_ground(x, y, z, seed=seed)
    out, gplane, found = ctx.segment_ground(x, y, z, seed=seed)
    assert found and np.allclose(gplane, plane, rtol=1e-5, atol=1e-6)
    exp = np.stack([x[keep], y[keep], z[keep]])
    # identical hypotheses and scores; the refined plane is reduced in a different order, so a
    # point within float rounding of the threshold may flip: allow a handful
    assert abs(out.shape[1] - exp.shape[1]) <= 8
    if out.shape[1] == exp.shape[1]:
        same = (out.view(np.uint32) == exp.view(np.uint32)).all(axis=0) | np.isnan(out).any(axis=0)
        assert same.mean() > 0.9999
    assert out.shape[1] > 50000


@pytest.mark.gpu
def test_gpu_segment_ground_no_model(ctx):
    out, plane, found = ctx.segment_ground(*degenerate_cloud())
    assert not found and out.shape[1] == 0       # ref :122-126: empty cloud
    rng = np.random.default_rng(0)
    x, y, z = (rng.uniform(-50, 50, 300).astype(f32) for _ in range(3))
    keep, eplane, _, _ = orc.segment_ground(x, y, z)
    out, plane, found = ctx.segment_ground(x, y, z)
    assert found and np.allclose(plane, eplane, rtol=1e-4, atol=1e-5) and abs(out.shape[1] - keep.sum()) <= 1


# ------------------------------------------------------------------------- independent acceptance rule
def independent_ransac(x, y, z, threshold=0.04, iters=400, seed=2024):
    """A RANSAC + least-squares plane fit written independently of the repository's (numpy float64,
    its own random sample stream, SVD refinement): the acceptance reference SURVEY 8.f asks for, since
    PCL's own sample stream cannot be reproduced.  Returns (plane[4] unit normal, inlier mask)."""
    P = np.stack([x, y, z], 1).astype(np.float64)
    ok = np.isfinite(P).all(1)
    Q = P[ok]
    rng = np.random.default_rng(seed)
    best, best_n = None, -1
    for _ in range(iters):
        a, b, c = Q[rng.choice(len(Q), 3, replace=False)]
        n = np.cross(b - a, c - a)
        ln = np.linalg.norm(n)
        if ln < 1e-9:
            continue
        n /= ln
        cnt = int((np.abs(Q @ n - a @ n) < threshold).sum())
        if cnt > best_n:
            best, best_n = (n, -a @ n), cnt
    n, d = best
    for _ in range(2):  # refine on the inliers (total least squares), like setOptimizeCoefficients(true)
        inl = np.abs(Q @ n + d) < threshold
        c = Q[inl].mean(0)
        _, _, vt = np.linalg.svd(Q[inl] - c, full_matrices=False)
        n2 = vt[-1]
        if n2 @ n < 0:
            n2 = -n2
        n, d = n2, -n2 @ c
    mask = np.zeros(len(P), bool)
    mask[ok] = np.abs(Q @ n + d) < threshold
    return np.append(n, d), mask


def tilted_scene(seed=3, n=120000):
    """A ground plane tilted ~6 degrees about two axes with 1 cm noise, plus 35 % clutter above it."""
    rng = np.random.default_rng(seed)
    nrm = np.array([0.08, 0.99, -0.07])
    nrm /= np.linalg.norm(nrm)
    u = np.cross(nrm, [0, 0, 1.0]); u /= np.linalg.norm(u)
    v = np.cross(nrm, u)
    m = int(n * 0.65)
    g = np.outer(rng.uniform(-40, 40, m), u) + np.outer(rng.uniform(-40, 40, m), v) + 1.4 * nrm
    g += rng.normal(0, 0.01, (m, 1)) * nrm
    c = rng.uniform(-40, 40, (n - m, 3)) - np.outer(rng.uniform(0.3, 3.0, n - m), nrm)
    P = np.concatenate([g, c]).astype(f32)
    P[rng.integers(0, n, n // 100)] = np.nan
    return P[:, 0].copy(), P[:, 1].copy(), P[:, 2].copy()


def jaccard(a, b):
    return (a & b).sum() / max(1, (a | b).sum())


CASES = {"C1": lambda: camera_scan(synth.C1, 0), "C2": lambda: camera_scan(synth.C2, 1), "tilted": tilted_scene}


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_agrees_with_independent_ransac(case):
    x, y, z = CASES[case]()
    keep, plane, _, _ = orc.segment_ground(x, y, z, n_hyp=256)
    iplane, imask = independent_ransac(x, y, z)
    s = 1.0 if plane[:3] @ iplane[:3] > 0 else -1.0
    assert np.allclose(s * np.asarray(plane, np.float64), iplane, atol=2e-3), (plane, iplane)
    assert jaccard(~keep & np.isfinite(x), imask) >= 0.99


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_gpu_segment_ground_agrees_with_independent_ransac(ctx, case):
    """Acceptance against an implementation that shares no code and no random numbers with the
    repository: plane coefficients within 2e-3 and removed-set Jaccard >= 0.99."""
    x, y, z = CASES[case]()
    iplane, imask = independent_ransac(x, y, z)
    out, gplane, found = ctx.segment_ground(x, y, z, n_hyp=256)
    assert found
    s = 1.0 if gplane[:3] @ iplane[:3] > 0 else -1.0
    assert np.allclose(s * gplane.astype(np.float64), iplane, atol=2e-3), (gplane, iplane)
    # the GPU returns the compacted cloud: recover the removed set from the plane it reports
    d = np.abs(x * gplane[0] + y * gplane[1] + z * gplane[2] + gplane[3])
    removed = np.isfinite(x) & np.isfinite(y) & np.isfinite(z) & (d < 0.04)
    assert abs(int((~removed).sum()) - out.shape[1]) <= 8
    assert jaccard(removed, imask) >= 0.99
