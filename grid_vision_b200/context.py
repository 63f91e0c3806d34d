"""Python binding over the C ABI (include/gridvision_b200.h): one Context per GPU.

numpy arrays go through the host entry points (the library copies in/out itself and
returns synchronously); torch CUDA tensors go through the *_dev entry points on the
tensor's current stream.  Nothing here computes: every method is one C-ABI call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import AccumParams, GridDesc, GridVisionError, Stats
from .synth import BOX_DTYPE

POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"),
                        ("intensity", "<f4"), ("pad1", "<f4"), ("pad2", "<f4"), ("pad3", "<f4")])


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _is_torch(t) -> bool:
    return type(t).__module__.startswith("torch")


def _dptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def accum_params(occ_mode=_lib.OCC_ALL, z_gate=None, r_max=0.0) -> AccumParams:
    return AccumParams(int(occ_mode), 0 if z_gate is None else 1,
                       0.0 if z_gate is None else float(z_gate[0]),
                       0.0 if z_gate is None else float(z_gate[1]), float(r_max))


class Context:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.gv_create(C.byref(h), C.c_int(device))
        if rc != _lib.GV_OK:
            raise GridVisionError(rc, "gv_create", self._lib.gv_status_string(rc).decode())
        self._h = h
        self.device = device
        self.ncam = 0

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != _lib.GV_OK:
            raise GridVisionError(rc, what, self._lib.gv_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def synchronize(self):
        self._check(self._lib.gv_synchronize(self._h), "gv_synchronize")

    def use_torch_stream(self):
        """Launch on torch's current stream (ordering with the tensors' producers/consumers, and
        torch.cuda.Event brackets the kernels).  Called automatically by every method that takes
        torch CUDA tensors; re-binding happens only when torch's current stream changed."""
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        if getattr(self, "_bound_stream", None) != s:
            self._check(self._lib.gv_set_stream(self._h, C.c_void_p(s)), "gv_set_stream")
            self._bound_stream = s

    # ---- CUDA graphs: capture a sequence of device-pointer calls once, replay it with one launch
    def graph_begin(self):
        self._check(self._lib.gv_graph_begin(self._h), "gv_graph_begin")

    def graph_end(self) -> int:
        gid = C.c_int32(-1)
        self._check(self._lib.gv_graph_end(self._h, C.byref(gid)), "gv_graph_end")
        return int(gid.value)

    def graph_launch(self, gid: int):
        self._check(self._lib.gv_graph_launch(self._h, C.c_int32(gid)), "gv_graph_launch")

    def graph_destroy(self, gid: int):
        self._check(self._lib.gv_graph_destroy(self._h, C.c_int32(gid)), "gv_graph_destroy")

    def stats(self) -> dict:
        st = Stats()
        self._check(self._lib.gv_get_stats(self._h, C.byref(st)), "gv_get_stats")
        return {k: (float(getattr(st, k)) if k == "merge_ms_last" else int(getattr(st, k))) for k, _ in Stats._fields_}

    def join(self):
        """The context stream waits for the merge (finalize) still running on the internal stream."""
        self._check(self._lib.gv_join(self._h), "gv_join")

    # ------------------------------------------------------------------ fusion
    def set_cameras(self, K, wh, T_cam_lidar=None):
        K = _np(K, np.float64).reshape(-1, 9)
        ncam = K.shape[0]
        wh = _np(wh, np.int32).reshape(ncam, 2)
        T = None if T_cam_lidar is None else _np(T_cam_lidar, np.float32).reshape(ncam, 16)
        self._check(self._lib.gv_set_cameras(self._h, C.c_int(ncam), _ptr(K), _ptr(T), _ptr(wh)),
                    "gv_set_cameras")
        self.ncam = ncam

    def fuse(self, x, y, z, boxes, box_cam_offsets=None, is_dense=False, want_pix=True,
             want_uv=True):
        """-> (labels [ncam,n] int16, pix [ncam,n] int32 | None, uv [ncam,2,n] f32 | None)."""
        x, y, z = (_np(a, np.float32) for a in (x, y, z))
        n = x.size
        boxes = _np(boxes, BOX_DTYPE)
        off = None if box_cam_offsets is None else _np(box_cam_offsets, np.int32)
        labels = np.empty((self.ncam, n), np.int16)
        pix = np.empty((self.ncam, n), np.int32) if want_pix else None
        uv = np.empty((self.ncam, 2, n), np.float32) if want_uv else None
        self._check(self._lib.gv_fuse(self._h, _ptr(x), _ptr(y), _ptr(z), C.c_size_t(n),
                                      C.c_int(int(is_dense)), _ptr(boxes), C.c_int(len(boxes)),
                                      _ptr(off), _ptr(labels), _ptr(pix), _ptr(uv)), "gv_fuse")
        return labels, pix, uv

    def fuse_aos32(self, pts, boxes, box_cam_offsets=None, is_dense=False):
        pts = np.ascontiguousarray(pts)
        assert pts.nbytes % 32 == 0
        n = pts.nbytes // 32
        boxes = _np(boxes, BOX_DTYPE)
        off = None if box_cam_offsets is None else _np(box_cam_offsets, np.int32)
        labels = np.empty((self.ncam, n), np.int16)
        pix = np.empty((self.ncam, n), np.int32)
        uv = np.empty((self.ncam, 2, n), np.float32)
        self._check(self._lib.gv_fuse_aos32(self._h, _ptr(pts), C.c_size_t(n),
                                            C.c_int(int(is_dense)), _ptr(boxes),
                                            C.c_int(len(boxes)), _ptr(off), _ptr(labels),
                                            _ptr(pix), _ptr(uv)), "gv_fuse_aos32")
        return labels, pix, uv

    def fuse_dev(self, x, y, z, d_boxes, nboxes, box_cam_offsets=None, labels=None, pix=None,
                 uv=None, is_dense=False):
        """torch CUDA tensors in/out; d_boxes is a uint8 CUDA tensor holding 40-byte records."""
        off = None if box_cam_offsets is None else _np(box_cam_offsets, np.int32)
        self.use_torch_stream()
        self._check(self._lib.gv_fuse_dev(self._h, _dptr(x), _dptr(y), _dptr(z),
                                          C.c_size_t(x.numel()), C.c_int(int(is_dense)),
                                          _dptr(d_boxes), C.c_int(nboxes), _ptr(off),
                                          _dptr(labels), _dptr(pix), _dptr(uv)), "gv_fuse_dev")

    def transform_points(self, cam, x, y, z, is_dense=False):
        x, y, z = (_np(a, np.float32) for a in (x, y, z))
        n = x.size
        ox, oy, oz = (np.empty(n, np.float32) for _ in range(3))
        self._check(self._lib.gv_transform_points(self._h, C.c_int(cam), _ptr(x), _ptr(y), _ptr(z),
                                                  C.c_size_t(n), C.c_int(int(is_dense)), _ptr(ox),
                                                  _ptr(oy), _ptr(oz)), "gv_transform_points")
        return ox, oy, oz

    def project_kdtree(self, cam, x, y, z):
        x, y, z = (_np(a, np.float32) for a in (x, y, z))
        n = x.size
        uvz = np.empty((max(n, 1), 3), np.float32)
        m = C.c_size_t(0)
        self._check(self._lib.gv_project_kdtree(self._h, C.c_int(cam), _ptr(x), _ptr(y), _ptr(z),
                                                C.c_size_t(n), _ptr(uvz), C.byref(m)),
                    "gv_project_kdtree")
        return uvz[:m.value].copy()

    def partition_by_label(self, labels, nboxes):
        labels = _np(labels, np.int16)
        n = labels.size
        idx = np.empty(max(n, 1), np.uint32)
        off = np.zeros(nboxes + 1, np.uint64)
        self._check(self._lib.gv_partition_by_label(self._h, _ptr(labels), C.c_size_t(n),
                                                    C.c_int(nboxes), _ptr(idx), _ptr(off)),
                    "gv_partition_by_label")
        return idx[:int(off[nboxes])].copy(), off

    def segment_ground(self, x, y, z, threshold=0.04, seed=12345, n_hyp=256):
        """N1 -> (xyz kept [3,m] float32, plane[4], found)."""
        x, y, z = (_np(a, np.float32) for a in (x, y, z))
        n = x.size
        out = np.empty((3, max(n, 1)), np.float32)
        m, found = C.c_size_t(0), C.c_int(0)
        plane = np.zeros(4, np.float32)
        self._check(self._lib.gv_segment_ground(
            self._h, _ptr(x), _ptr(y), _ptr(z), C.c_size_t(n), C.c_float(threshold), C.c_uint32(seed),
            C.c_int(n_hyp), C.c_void_p(out[0].ctypes.data), C.c_void_p(out[1].ctypes.data),
            C.c_void_p(out[2].ctypes.data), C.byref(m), _ptr(plane), C.byref(found)), "gv_segment_ground")
        return out[:, :m.value].copy(), plane, bool(found.value)

    def bbox_pose(self, x, y, z, labels, nboxes):
        """N2: radius-outlier filter + PCA box per label -> list of _lib.LShape (kept == 0: skipped)."""
        x, y, z = (_np(a, np.float32) for a in (x, y, z))
        labels = _np(labels, np.int16)
        out = (_lib.LShape * max(nboxes, 1))()
        self._check(self._lib.gv_bbox_pose(self._h, _ptr(x), _ptr(y), _ptr(z), C.c_size_t(x.size),
                                           _ptr(labels), C.c_int(nboxes), out), "gv_bbox_pose")
        return list(out)[:nboxes]

    def box_depths(self, uvz, boxes, k=4):
        """N4: median depth of the k nearest (u, v, depth) triples of every box centre."""
        uvz = _np(uvz, np.float32).reshape(-1, 3)
        boxes = _np(boxes, BOX_DTYPE)
        out = np.empty(max(len(boxes), 1), np.float32)
        self._check(self._lib.gv_box_depths(self._h, _ptr(uvz), C.c_size_t(len(uvz)), _ptr(boxes),
                                            C.c_int(len(boxes)), C.c_int(k), _ptr(out)), "gv_box_depths")
        return out[:len(boxes)].copy()

    def pixels_to_3d(self, boxes, depths, K_inv):
        boxes = _np(boxes, BOX_DTYPE)
        depths = _np(depths, np.float32)
        K_inv = _np(K_inv, np.float64).reshape(9)
        out = np.empty((max(len(boxes), 1), 3), np.float64)
        self._check(self._lib.gv_pixels_to_3d(self._h, _ptr(boxes), _ptr(depths), C.c_int(len(boxes)),
                                              _ptr(K_inv), _ptr(out)), "gv_pixels_to_3d")
        return out[:len(boxes)].copy()

    # ------------------------------------------------------------------ grid
    def grid_init(self, length_x, length_y, resolution, pos_x=0.0, pos_y=0.0):
        self._check(self._lib.gv_grid_init(self._h, C.c_double(length_x), C.c_double(length_y),
                                           C.c_double(resolution), C.c_double(pos_x),
                                           C.c_double(pos_y)), "gv_grid_init")
        return self.grid_desc()

    def grid_init_cells(self, nx, ny, resolution, pos_x=0.0, pos_y=0.0):
        d = self.grid_init(nx * resolution, ny * resolution, resolution, pos_x, pos_y)
        assert (d.nx, d.ny) == (nx, ny), (d.nx, d.ny, nx, ny)
        return d

    def grid_init_reference(self, grid_x, grid_y, resolution):
        self._check(self._lib.gv_grid_init_reference(self._h, C.c_uint8(grid_x), C.c_uint8(grid_y),
                                                     C.c_double(resolution)),
                    "gv_grid_init_reference")
        return self.grid_desc()

    def grid_desc(self) -> GridDesc:
        d = GridDesc()
        self._check(self._lib.gv_grid_get_desc(self._h, C.byref(d)), "gv_grid_get_desc")
        return d

    def grid_reset(self):
        self._check(self._lib.gv_grid_reset(self._h), "gv_grid_reset")

    def grid_upload(self, log_odds=None, occupancy=None):
        lo = None if log_odds is None else _np(log_odds, np.float32)
        oc = None if occupancy is None else _np(occupancy, np.float32)
        self._check(self._lib.gv_grid_upload(self._h, _ptr(lo), _ptr(oc)), "gv_grid_upload")

    def grid_download(self):
        d = self.grid_desc()
        lo = np.empty(d.nx * d.ny, np.float32)
        oc = np.empty(d.nx * d.ny, np.float32)
        self._check(self._lib.gv_grid_download(self._h, _ptr(lo), _ptr(oc)), "gv_grid_download")
        return lo, oc

    def grid_counts(self):
        d = self.grid_desc()
        hit = np.empty(d.nx * d.ny, np.int32)
        miss = np.empty(d.nx * d.ny, np.int32)
        self._check(self._lib.gv_grid_counts_download(self._h, _ptr(hit), _ptr(miss)),
                    "gv_grid_counts_download")
        return hit, miss

    def grid_get_index(self, xy):
        xy = _np(xy, np.float64).reshape(-1, 2)
        out = np.empty((len(xy), 2), np.int32)
        self._check(self._lib.gv_grid_get_index(self._h, _ptr(xy), C.c_int(len(xy)), _ptr(out)),
                    "gv_grid_get_index")
        return out

    def grid_update(self):
        self._check(self._lib.gv_grid_update(self._h), "gv_grid_update")

    def grid_update_poses(self, xylw):
        a = _np(xylw, np.float64).reshape(-1, 4)
        self._check(self._lib.gv_grid_update_poses(self._h, _ptr(a), C.c_int(len(a))),
                    "gv_grid_update_poses")

    def grid_update_points(self, xy, labels):
        a = _np(xy, np.float64).reshape(-1, 2)
        lab = _np(labels, np.int32)
        assert len(lab) == len(a)
        self._check(self._lib.gv_grid_update_points(self._h, _ptr(a), _ptr(lab), C.c_int(len(a))),
                    "gv_grid_update_points")

    def grid_update_corners(self, corners):
        a = _np(corners, np.float64).reshape(-1, 8)
        self._check(self._lib.gv_grid_update_corners(self._h, _ptr(a), C.c_int(len(a))),
                    "gv_grid_update_corners")

    # ------------------------------------------------------------------ binning / raycast
    def set_base_transform(self, T_base_lidar):
        T = _np(T_base_lidar, np.float32).reshape(16)
        self._check(self._lib.gv_set_base_transform(self._h, _ptr(T)), "gv_set_base_transform")

    def grid_accumulate(self, x, y, z, labels=None, params: AccumParams | None = None,
                        want_cells=True):
        params = params or accum_params()
        if _is_torch(x):
            self.use_torch_stream()
            self._check(self._lib.gv_grid_accumulate_dev(
                self._h, _dptr(x), _dptr(y), _dptr(z), C.c_size_t(x.numel()), _dptr(labels),
                C.byref(params), None, None), "gv_grid_accumulate_dev")
            return None, None
        x, y, z = (_np(a, np.float32) for a in (x, y, z))
        n = x.size
        lab = None if labels is None else _np(labels, np.int16)
        cells = np.empty(n, np.int32) if want_cells else None
        flags = np.empty(n, np.uint8) if want_cells else None
        self._check(self._lib.gv_grid_accumulate(self._h, _ptr(x), _ptr(y), _ptr(z), C.c_size_t(n),
                                                 _ptr(lab), C.byref(params), _ptr(cells),
                                                 _ptr(flags)), "gv_grid_accumulate")
        return cells, flags

    def grid_raycast_flush(self):
        self._check(self._lib.gv_grid_raycast_flush(self._h), "gv_grid_raycast_flush")

    def grid_finalize(self, k_decay=1, corners=None, multi=False):
        a = None if corners is None else _np(corners, np.float64).reshape(-1, 8)
        fn = self._lib.gv_grid_finalize_multi if multi else self._lib.gv_grid_finalize
        self._check(fn(self._h, C.c_int32(k_decay), _ptr(a), C.c_int(0 if a is None else len(a))),
                    "gv_grid_finalize_multi" if multi else "gv_grid_finalize")

    def process_batch(self, x, y, z, frame_offsets, boxes, box_frame_offsets,
                      params: AccumParams | None = None, labels_out=None):
        """Whole hot path for a batch.  numpy planes -> host entry point (returns labels);
        torch CUDA planes -> device entry point (boxes: uint8 CUDA tensor of 40-byte records,
        labels_out: int16 CUDA tensor or None)."""
        params = params or accum_params()
        fo = _np(frame_offsets, np.uint64)
        bo = _np(box_frame_offsets, np.int32)
        nf = len(fo) - 1
        assert len(bo) == nf + 1
        if _is_torch(x):
            self.use_torch_stream()
            self._check(self._lib.gv_process_batch_dev(
                self._h, _dptr(x), _dptr(y), _dptr(z), _ptr(fo), C.c_int(nf), _dptr(boxes),
                _ptr(bo), C.byref(params), _dptr(labels_out)), "gv_process_batch_dev")
            return labels_out
        boxes = _np(boxes, BOX_DTYPE)
        if labels_out is None:
            labels_out = np.empty(int(fo[-1]), np.int16)
        self._check(self._lib.gv_process_batch(
            self._h, _ptr(x), _ptr(y), _ptr(z), _ptr(fo), C.c_int(nf), _ptr(boxes), _ptr(bo),
            C.byref(params), _ptr(labels_out)), "gv_process_batch")
        return labels_out

    def process_batch_ptrs(self, px, py, pz, frame_offsets, boxes_ptr, box_frame_offsets, params,
                           labels_ptr):
        """Raw host-pointer form (pinned torch tensors): no numpy conversion on the hot path."""
        fo = _np(frame_offsets, np.uint64)
        bo = _np(box_frame_offsets, np.int32)
        self._check(self._lib.gv_process_batch(
            self._h, C.c_void_p(px), C.c_void_p(py), C.c_void_p(pz), _ptr(fo), C.c_int(len(fo) - 1),
            C.c_void_p(boxes_ptr), _ptr(bo), C.byref(params),
            None if not labels_ptr else C.c_void_p(labels_ptr)), "gv_process_batch")

    def grid_to_occupancy(self):
        d = self.grid_desc()
        out = np.empty(d.nx * d.ny, np.int8)
        self._check(self._lib.gv_grid_to_occupancy(self._h, _ptr(out)), "gv_grid_to_occupancy")
        return out

    # ------------------------------------------------------------------ multi-GPU
    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        rc = lib.gv_nccl_unique_id(buf)
        if rc != _lib.GV_OK:
            raise GridVisionError(rc, "gv_nccl_unique_id")
        return buf.raw

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(448)
        self._check(self._lib.gv_ipc_export(self._h, buf), "gv_ipc_export")
        return buf.raw

    def ipc_import(self, blobs: bytes, world: int, rank: int):
        assert len(blobs) == 448 * world
        self._check(self._lib.gv_ipc_import(self._h, C.c_char_p(blobs), C.c_int(world), C.c_int(rank)),
                    "gv_ipc_import")

    def ipc_close(self):
        self._check(self._lib.gv_ipc_close(self._h), "gv_ipc_close")

    def nccl_init(self, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == 128
        self._check(self._lib.gv_nccl_init(self._h, C.c_char_p(unique_id), C.c_int(rank),
                                           C.c_int(world)), "gv_nccl_init")
