"""ctypes binding for the CPU oracle (oracle/gv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgv_oracle.so")

OCC_ALL, OCC_LABELLED = 0, 1
F_VALID, F_HIT, F_CLIPPED, F_RANGECAP = 1, 2, 4, 8

BOX_DTYPE = np.dtype(
    [("x_min", "<f8"), ("y_min", "<f8"), ("x_max", "<f8"), ("y_max", "<f8"),
     ("confidence", "<f4"), ("label", "<i4")], align=True)
assert BOX_DTYPE.itemsize == 40  # include/grid_vision/object_detection.hpp:27-32
POINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"),
     ("intensity", "<f4"), ("pad1", "<f4"), ("pad2", "<f4"), ("pad3", "<f4")])
assert POINT_DTYPE.itemsize == 32  # pcl::PointXYZI


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gv_oracle.c")
    hdr = os.path.join(_HERE, "gv_oracle.h")
    stale = (not os.path.exists(_SO)) or (
        os.path.exists(src)
        and max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(_SO))
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


class _Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("res", C.c_double),
                ("len_x", C.c_double), ("len_y", C.c_double),
                ("pos_x", C.c_double), ("pos_y", C.c_double),
                ("log_odds", C.POINTER(C.c_float)), ("occupancy", C.POINTER(C.c_float)),
                ("hit", C.POINTER(C.c_int32)), ("miss", C.POINTER(C.c_int32))]


class LShape(C.Structure):
    _fields_ = [("kept", C.c_int32), ("centroid_y", C.c_float), ("mean_z", C.c_float),
                ("mean_x", C.c_float), ("major_z", C.c_float), ("major_x", C.c_float),
                ("minor_z", C.c_float), ("minor_x", C.c_float), ("length", C.c_float),
                ("width", C.c_float), ("angle_deg", C.c_float), ("qx", C.c_double),
                ("qy", C.c_double), ("qz", C.c_double), ("qw", C.c_double)]


class _AccumParams(C.Structure):
    _fields_ = [("occ_mode", C.c_int32), ("use_z_gate", C.c_int32),
                ("z_min", C.c_float), ("z_max", C.c_float), ("r_max", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.gvo_project_kdtree.restype = C.c_size_t
        _lib.gvo_accumulate.restype = C.c_int64
        _lib.gvo_estimated_depth.restype = C.c_float
        _lib.gvo_bresenham_cells.restype = C.c_int32
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _T(T):
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(-1)
    assert T.size == 16
    return T


def _K(K):
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(-1)
    assert K.size == 9
    return K


def make_boxes(xyxy, confidence=None, label=None) -> np.ndarray:
    xyxy = np.asarray(xyxy, dtype=np.float64).reshape(-1, 4)
    b = np.zeros(len(xyxy), dtype=BOX_DTYPE)
    b["x_min"], b["y_min"], b["x_max"], b["y_max"] = xyxy.T
    b["confidence"] = 1.0 if confidence is None else confidence
    b["label"] = 0 if label is None else label
    return b


def transform_points(T, x, y, z, is_dense=False):
    x, y, z = _f32(x), _f32(y), _f32(z)
    n = x.size
    ox, oy, oz = (np.empty(n, np.float32) for _ in range(3))
    T = _T(T)
    lib().gvo_transform_points(_p(T, C.c_float), _p(x, C.c_float), _p(y, C.c_float),
                               _p(z, C.c_float), C.c_size_t(n), C.c_int(int(is_dense)),
                               _p(ox, C.c_float), _p(oy, C.c_float), _p(oz, C.c_float))
    return ox, oy, oz


def project_label(K, W, H, x, y, z, boxes):
    """R3 on camera-frame points -> (label int16, pix int32, u f32, v f32)."""
    x, y, z = _f32(x), _f32(y), _f32(z)
    n = x.size
    boxes = np.ascontiguousarray(boxes, dtype=BOX_DTYPE)
    label = np.empty(n, np.int16)
    pix = np.empty(n, np.int32)
    u = np.empty(n, np.float32)
    v = np.empty(n, np.float32)
    K = _K(K)
    lib().gvo_project_label(_p(K, C.c_double), C.c_int(W), C.c_int(H), _p(x, C.c_float),
                            _p(y, C.c_float), _p(z, C.c_float), C.c_size_t(n),
                            C.c_void_p(boxes.ctypes.data), C.c_int(len(boxes)),
                            _p(label, C.c_int16), _p(pix, C.c_int32), _p(u, C.c_float),
                            _p(v, C.c_float))
    return label, pix, u, v


def extract_cloud_per_bbox_aos(points_aos, K, boxes, W, H):
    """Reference-shaped R3: AoS in, list of per-box AoS arrays out."""
    pts = np.ascontiguousarray(points_aos, dtype=POINT_DTYPE)
    boxes = np.ascontiguousarray(boxes, dtype=BOX_DTYPE)
    nb = len(boxes)
    clouds = (C.c_void_p * max(nb, 1))()
    counts = (C.c_size_t * max(nb, 1))()
    K = _K(K)
    lib().gvo_extract_cloud_per_bbox_aos(C.c_void_p(pts.ctypes.data), C.c_size_t(len(pts)),
                                         _p(K, C.c_double), C.c_void_p(boxes.ctypes.data),
                                         C.c_int(nb), C.c_int(W), C.c_int(H), clouds, counts)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    out = []
    for i in range(nb):
        m = counts[i]
        if m:
            buf = (C.c_char * (m * 32)).from_address(clouds[i])
            out.append(np.frombuffer(bytes(buf), dtype=POINT_DTYPE).copy())
            libc.free(clouds[i])
        else:
            out.append(np.zeros(0, dtype=POINT_DTYPE))
    return out


def project_kdtree(K, x, y, z):
    x, y, z = _f32(x), _f32(y), _f32(z)
    n = x.size
    uvz = np.empty((n, 3), np.float32)
    K = _K(K)
    m = lib().gvo_project_kdtree(_p(K, C.c_double), _p(x, C.c_float), _p(y, C.c_float),
                                 _p(z, C.c_float), C.c_size_t(n), _p(uvz, C.c_float))
    return uvz[:m].copy()


def box_depths(uvz, boxes, k=4):
    """N4: cloud_detections::computeDepthForBoundingBoxes on (u, v, depth) triples."""
    uvz = np.ascontiguousarray(uvz, dtype=np.float32).reshape(-1, 3)
    boxes = np.ascontiguousarray(boxes, dtype=BOX_DTYPE)
    out = np.empty(max(len(boxes), 1), np.float32)
    lib().gvo_box_depths(_p(uvz, C.c_float), C.c_size_t(len(uvz)), C.c_void_p(boxes.ctypes.data),
                         C.c_int(len(boxes)), C.c_int(k), _p(out, C.c_float))
    return out[:len(boxes)].copy()


def pixel_to_3d(K_inv, px, py, depth):
    Ki = np.ascontiguousarray(K_inv, dtype=np.float64).reshape(9)
    out = np.empty(3, np.float64)
    lib().gvo_pixel_to_3d(_p(Ki, C.c_double), C.c_float(px), C.c_float(py), C.c_float(depth), _p(out, C.c_double))
    return out


def bresenham_cells(sx, sy, ex, ey):
    n = max(abs(ex - sx), abs(ey - sy)) + 1
    out = np.empty((n, 2), np.int32)
    m = lib().gvo_bresenham_cells(C.c_int32(sx), C.c_int32(sy), C.c_int32(ex), C.c_int32(ey),
                                  _p(out, C.c_int32), C.c_int32(n))
    assert m == n
    return out


def segment_ground(x, y, z, threshold=0.04, seed=12345, n_hyp=256):
    """N1 -> (keep mask or None when no model, plane[4], best hypothesis, best score)."""
    x, y, z = _f32(x), _f32(y), _f32(z)
    keep = np.ones(x.size, np.uint8)
    plane = np.zeros(4, np.float32)
    bh, bs = C.c_int32(-1), C.c_int32(0)
    lib().gvo_segment_ground.restype = C.c_int64
    kept = lib().gvo_segment_ground(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float),
                                    C.c_size_t(x.size), C.c_float(threshold), C.c_uint32(seed),
                                    C.c_int32(n_hyp), _p(keep, C.c_uint8), _p(plane, C.c_float),
                                    C.byref(bh), C.byref(bs))
    return (None if kept < 0 else keep.astype(bool)), plane, bh.value, bs.value


def radius_outlier_keep(x, y, z, radius=0.4, min_neighbors=10):
    x, y, z = _f32(x), _f32(y), _f32(z)
    keep = np.zeros(x.size, np.uint8)
    lib().gvo_radius_outlier_keep(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float),
                                  C.c_size_t(x.size), C.c_double(radius), C.c_int32(min_neighbors),
                                  _p(keep, C.c_uint8))
    return keep.astype(bool)


def bbox_pose(x, y, z) -> LShape:
    x, y, z = _f32(x), _f32(y), _f32(z)
    out = LShape()
    lib().gvo_bbox_pose(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), C.c_size_t(x.size),
                        C.byref(out))
    return out


def estimated_depth(label: int) -> float:
    return float(lib().gvo_estimated_depth(C.c_int32(label)))


class Grid:
    """Owns a gvo_grid; layers are exposed as numpy views (column-major: lin = ix + iy*nx)."""

    def __init__(self, length_x=None, length_y=None, res=None, pos_x=0.0, pos_y=0.0,
                 reference_ctor=None):
        self._g = _Grid()
        if reference_ctor is not None:
            gx, gy, r = reference_ctor
            rc = lib().gvo_grid_init_reference(C.byref(self._g), C.c_uint8(gx), C.c_uint8(gy),
                                               C.c_double(r))
        else:
            rc = lib().gvo_grid_init(C.byref(self._g), C.c_double(length_x),
                                     C.c_double(length_y), C.c_double(res),
                                     C.c_double(pos_x), C.c_double(pos_y))
        if rc != 0:
            raise ValueError(f"gvo_grid_init failed rc={rc}")
        nc = self.nx * self.ny
        self.log_odds = np.ctypeslib.as_array(self._g.log_odds, shape=(nc,))
        self.occupancy = np.ctypeslib.as_array(self._g.occupancy, shape=(nc,))
        self.hit = np.ctypeslib.as_array(self._g.hit, shape=(nc,))
        self.miss = np.ctypeslib.as_array(self._g.miss, shape=(nc,))

    @classmethod
    def from_cells(cls, nx, ny, res, pos_x=0.0, pos_y=0.0):
        g = cls(nx * res, ny * res, res, pos_x, pos_y)
        assert (g.nx, g.ny) == (nx, ny), (g.nx, g.ny, nx, ny)
        return g

    nx = property(lambda s: s._g.nx)
    ny = property(lambda s: s._g.ny)
    res = property(lambda s: s._g.res)
    len_x = property(lambda s: s._g.len_x)
    len_y = property(lambda s: s._g.len_y)
    pos_x = property(lambda s: s._g.pos_x)
    pos_y = property(lambda s: s._g.pos_y)

    def __del__(self):
        try:
            lib().gvo_grid_free(C.byref(self._g))
        except Exception:
            pass

    def get_index(self, px, py):
        ix, iy = C.c_int32(-1), C.c_int32(-1)
        ok = lib().gvo_grid_get_index(C.byref(self._g), C.c_double(px), C.c_double(py),
                                      C.byref(ix), C.byref(iy))
        return (ix.value, iy.value) if ok else None

    def update_map(self):
        lib().gvo_update_map(C.byref(self._g))

    def update_map_poses(self, xylw):
        a = np.ascontiguousarray(xylw, dtype=np.float64).reshape(-1, 4)
        lib().gvo_update_map_poses(C.byref(self._g), _p(a, C.c_double), C.c_int(len(a)))

    def update_map_points(self, xy, labels):
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        assert len(lab) == len(a)
        lib().gvo_update_map_points(C.byref(self._g), _p(a, C.c_double), _p(lab, C.c_int32),
                                    C.c_int(len(a)))

    def update_grid_cells_fast(self, corners):
        a = np.ascontiguousarray(corners, dtype=np.float64).reshape(8)
        lib().gvo_update_grid_cells_fast(C.byref(self._g), _p(a, C.c_double))

    def accumulate(self, T_base_lidar, x, y, z, labels=None, occ_mode=OCC_ALL, z_gate=None,
                   r_max=0.0, want_cells=True):
        x, y, z = _f32(x), _f32(y), _f32(z)
        n = x.size
        T = _T(T_base_lidar)
        prm = _AccumParams(occ_mode, 0 if z_gate is None else 1,
                           0.0 if z_gate is None else z_gate[0],
                           0.0 if z_gate is None else z_gate[1], float(r_max))
        lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.int16)
        cells = np.empty(n, np.int32) if want_cells else None
        flags = np.empty(n, np.uint8) if want_cells else None
        upd = lib().gvo_accumulate(C.byref(self._g), _p(T, C.c_float), _p(x, C.c_float),
                                   _p(y, C.c_float), _p(z, C.c_float), C.c_size_t(n),
                                   _p(lab, C.c_int16), C.byref(prm), _p(cells, C.c_int32),
                                   _p(flags, C.c_uint8))
        return upd, cells, flags

    def finalize(self, k_decay=1, corners=None):
        if corners is None:
            lib().gvo_finalize(C.byref(self._g), C.c_int32(k_decay), None, C.c_int(0))
        else:
            a = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 8)
            lib().gvo_finalize(C.byref(self._g), C.c_int32(k_decay), _p(a, C.c_double),
                               C.c_int(len(a)))

    def to_occupancy_grid(self):
        out = np.empty(self.nx * self.ny, np.int8)
        lib().gvo_to_occupancy_grid(C.byref(self._g), _p(out, C.c_int8))
        return out


def pose_corners(xylw):
    """src/occupancy_grid.cpp:79-90 corner order {LB, LF, RF, RB} for n x (x,y,length,width)."""
    a = np.asarray(xylw, dtype=np.float64).reshape(-1, 4)
    x, y, L, W = a.T
    xf, xb = x + L / 2.0, x - L / 2.0
    yl, yr = y - W / 2.0, y + W / 2.0
    return np.stack([xb, yl, xf, yl, xf, yr, xb, yr], axis=1)


def point_corners(xy, labels):
    """src/occupancy_grid.cpp:107-138 corner order {LF, RF, RB, LB}; depth is float."""
    a = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    d = np.array([estimated_depth(int(l)) for l in labels], dtype=np.float32)
    dh = (d / np.float32(2)).astype(np.float64)
    d = d.astype(np.float64)
    x, y = a.T
    return np.stack([x + d, y + dh, x + d, y - dh, x, y - dh, x, y + dh], axis=1)
