"""Synthetic inputs for the BASELINE.json configs.

Spinning-LiDAR scan model (rings x azimuth steps), a seeded scene of a ground plane
plus upright boxes, the 416x416 YOLOv4 camera, integer-valued detection boxes in the
reference's BoundingBox value domain (/root/reference src/object_detection.cpp:226-239:
bounds are static_cast<int> results stored in doubles), and the grid geometries.

torch is used only as an array library here so the same generator runs on the CPU
(tests, CPU baseline) and on the GPU (bench set-up, outside every timed region).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import torch

# The reference's BoundingBox (include/grid_vision/object_detection.hpp:27-32), 40 bytes.
BOX_DTYPE = np.dtype(
    [("x_min", "<f8"), ("y_min", "<f8"), ("x_max", "<f8"), ("y_max", "<f8"),
     ("confidence", "<f4"), ("label", "<i4")], align=True)
assert BOX_DTYPE.itemsize == 40

SEED_BASE = 20260000

# optical frame: x_cam = -y_lidar, y_cam = -z_lidar, z_cam = x_lidar
R_OPT = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])


@dataclasses.dataclass(frozen=True)
class Workload:
    """One BASELINE.json config: scan shape, boxes, camera rig and grid geometry."""
    name: str
    config_id: int
    rings: int
    azimuth: int
    frames: int
    boxes_per_camera: int
    cameras: int
    grid_nx: int
    grid_ny: int
    resolution: float
    r_max: float            # mapping range cap (metres, planar); 0 disables
    sensor_range: float = 130.0
    image_w: int = 416
    image_h: int = 416
    fx: float = 208.0
    fy: float = 208.0
    cx: float = 208.0
    cy: float = 208.0
    pos_x: float = 0.0
    pos_y: float = 0.0

    @property
    def points_per_frame(self) -> int:
        return self.rings * self.azimuth

    @property
    def points(self) -> int:
        return self.points_per_frame * self.frames

    @property
    def cells(self) -> int:
        return self.grid_nx * self.grid_ny

    def K(self) -> np.ndarray:
        return np.array([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])

    def scaled(self, **kw) -> "Workload":
        return dataclasses.replace(self, **kw)


# BASELINE.json configs[0..4]
C1 = Workload("C1 64-beam scan, 20 boxes, 200x200 @0.1 m", 1, 64, 2048, 1, 20, 1, 200, 200, 0.1, 0.0)
C2 = Workload("C2 128-beam scan, 50 boxes, 1000x1000 @0.05 m, full raycast", 2, 128, 2048, 1, 50, 1,
              1000, 1000, 0.05, 0.0)
C3 = Workload("C3 4096-scan replay into 2048x2048 @0.1 m", 3, 64, 2048, 4096, 50, 1, 2048, 2048,
              0.1, 120.0)
C4 = Workload("C4 6-camera rig, 300 boxes, 1M-point merged cloud", 4, 512, 2048, 1, 50, 6, 1000,
              1000, 0.05, 0.0)
C5 = Workload("C5 8192x8192 @0.05 m, 16M points/batch, 120 m rays", 5, 64, 2048, 128, 50, 1, 8192,
              8192, 0.05, 120.0)
CONFIGS = {1: C1, 2: C2, 3: C3, 4: C4, 5: C5}


def T_base_lidar() -> np.ndarray:
    """Sensor mounted 2.4 m above the base frame origin."""
    T = np.eye(4, dtype=np.float32)
    T[2, 3] = np.float32(2.4)
    return T


def camera_extrinsics(ncam: int = 1) -> np.ndarray:
    """T_cam<-lidar for ncam cameras at yaw 0, 360/ncam, ...; built in double, rounded once."""
    out = np.zeros((ncam, 4, 4), dtype=np.float32)
    t = np.array([0.0, -0.4, -0.2])
    for c in range(ncam):
        yaw = 2.0 * math.pi * c / ncam
        cz, sz = math.cos(yaw), math.sin(yaw)
        # rotate the lidar frame by -yaw so camera c looks along azimuth +yaw
        Rz = np.array([[cz, sz, 0.0], [-sz, cz, 0.0], [0.0, 0.0, 1.0]])
        T = np.eye(4)
        T[:3, :3] = R_OPT @ Rz
        T[:3, 3] = t
        out[c] = T.astype(np.float32)
    return out


def make_boxes(wl: Workload, frame: int = 0, camera: int = 0, n: int | None = None) -> np.ndarray:
    """Integer-valued pixel boxes, confidences descending (post-NMS order), labels 0..9."""
    n = wl.boxes_per_camera if n is None else n
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + 100 * wl.config_id + frame + 7919 * camera + 1))
    w = rng.integers(16, 129, size=n)
    h = rng.integers(16, 129, size=n)
    x0 = rng.integers(0, wl.image_w - w + 1)
    y0 = rng.integers(0, wl.image_h - h + 1)
    b = np.zeros(n, dtype=BOX_DTYPE)
    b["x_min"], b["y_min"] = x0, y0
    b["x_max"], b["y_max"] = x0 + w, y0 + h
    b["confidence"] = np.sort(rng.uniform(0.6, 1.0, size=n).astype(np.float32))[::-1]
    b["label"] = rng.integers(0, 10, size=n)
    return b


def _scene_boxes(wl: Workload, frame: int, n_obj: int = 64) -> np.ndarray:
    """n_obj upright axis-aligned boxes on the ground, LiDAR frame: rows (xmin,ymin,zmin,xmax,ymax,zmax)."""
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + 100 * wl.config_id + frame))
    c = rng.uniform(-50.0, 50.0, size=(n_obj, 2))
    # keep the sensor's own footprint clear
    near = np.hypot(c[:, 0], c[:, 1]) < 3.0
    c[near] += 6.0
    lx = rng.uniform(2.0, 5.0, size=n_obj)
    ly = rng.uniform(1.5, 2.5, size=n_obj)
    zmin = np.full(n_obj, -1.8)
    zmax = zmin + 1.5
    return np.stack([c[:, 0] - lx / 2, c[:, 1] - ly / 2, zmin, c[:, 0] + lx / 2, c[:, 1] + ly / 2,
                     zmax], axis=1)


def ray_directions(wl: Workload, device="cpu") -> torch.Tensor:
    el = torch.linspace(math.radians(-25.0), math.radians(15.0), wl.rings, dtype=torch.float64)
    az = torch.arange(wl.azimuth, dtype=torch.float64) * (2.0 * math.pi / wl.azimuth)
    ce, se = torch.cos(el)[:, None], torch.sin(el)[:, None]
    d = torch.stack([ce * torch.cos(az)[None, :], ce * torch.sin(az)[None, :],
                     se.expand(-1, wl.azimuth)], dim=-1)
    return d.reshape(-1, 3).to(torch.float32).to(device)  # ring-major, azimuth fastest


def make_scans(wl: Workload, frame0: int = 0, frames: int | None = None, device="cpu",
               nan_fraction: float = 0.01, adversarial: bool = False, chunk: int = 8,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """Returns SoA points, shape [3, frames * points_per_frame] float32 (x | y | z planes).

    Frame f occupies the slice [f*P, (f+1)*P) of each plane.  No-return beams are emitted
    at sensor_range (beyond the mapping range cap, so they are free-space-only beams) and
    nan_fraction of all slots are NaN (exercises the reference's finite test).
    """
    frames = wl.frames if frames is None else frames
    P = wl.points_per_frame
    d = ray_directions(wl, device)                                   # [P,3]
    if out is None:
        out = torch.empty((3, frames * P), dtype=torch.float32, device=device)
    inv = 1.0 / d                                                    # inf where d==0: fine for slabs
    for f0 in range(0, frames, chunk):
        f1 = min(frames, f0 + chunk)
        nf = f1 - f0
        if adversarial:
            g = torch.Generator(device="cpu")
            g.manual_seed(SEED_BASE + 100 * wl.config_id + frame0 + f0)
            r = (2.0 + (wl.sensor_range - 2.0) * torch.rand((nf, P), generator=g)).to(device)
        else:
            boxes = np.stack([_scene_boxes(wl, frame0 + f) for f in range(f0, f1)])  # [nf,64,6]
            b = torch.from_numpy(boxes).to(torch.float32).to(device)
            lo = b[:, None, :, 0:3] * inv[None, :, None, :]           # [nf,P,64,3]
            hi = b[:, None, :, 3:6] * inv[None, :, None, :]
            tmin = torch.minimum(lo, hi).amax(dim=-1)
            tmax = torch.maximum(lo, hi).amin(dim=-1)
            hit = (tmax >= tmin) & (tmin > 0.0)
            t_box = torch.where(hit, tmin, torch.full_like(tmin, float("inf"))).amin(dim=-1)
            del lo, hi, tmin, tmax, hit
            t_ground = torch.where(d[:, 2] < 0, -1.8 / d[:, 2], torch.full_like(d[:, 2], float("inf")))
            r = torch.minimum(t_box, t_ground[None, :]).clamp(max=wl.sensor_range)
        pts = r[:, :, None] * d[None, :, :]                           # [nf,P,3]
        if nan_fraction > 0:
            g = torch.Generator(device="cpu")
            g.manual_seed(SEED_BASE + 100 * wl.config_id + frame0 + f0 + 555)
            k = max(1, int(nan_fraction * nf * P))
            idx = torch.randint(0, nf * P, (k,), generator=g).to(device)
            pts.view(-1, 3)[idx] = float("nan")
        out[:, f0 * P:f1 * P] = pts.reshape(nf * P, 3).t()
    return out


def points_aos32(xyz: np.ndarray) -> np.ndarray:
    """SoA [3,n] -> pcl::PointXYZI-shaped 32-byte AoS records (w = 1, intensity = 0)."""
    n = xyz.shape[1]
    a = np.zeros((n, 8), dtype=np.float32)
    a[:, 0:3] = xyz.T
    a[:, 3] = 1.0
    return a


def make_footprints(wl: Workload, frame: int = 0, n: int = 8) -> np.ndarray:
    """n object poses (x, y, length, width) in the base frame for R8, some partly off-map."""
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + 100 * wl.config_id + frame + 31))
    hx, hy = wl.grid_nx * wl.resolution / 2, wl.grid_ny * wl.resolution / 2
    x = rng.uniform(wl.pos_x - hx * 1.05, wl.pos_x + hx * 1.05, size=n)
    y = rng.uniform(wl.pos_y - hy * 1.05, wl.pos_y + hy * 1.05, size=n)
    L = rng.uniform(1.5, 5.0, size=n)
    W = rng.uniform(0.5, 2.5, size=n)
    return np.stack([x, y, L, W], axis=1)
