"""Loader for libgridvision_b200.so (the C ABI in include/gridvision_b200.h).

There is no fallback: if the CUDA library is missing or fails to load this raises, and
every product call raises with it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgridvision_b200.so")

GV_OK = 0
OCC_ALL, OCC_LABELLED = 0, 1
F_VALID, F_HIT, F_CLIPPED, F_RANGECAP = 1, 2, 4, 8


class GridVisionError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        super().__init__(f"{what}: status {status} ({detail})" if detail else f"{what}: status {status}")


class AccumParams(C.Structure):
    _fields_ = [("occ_mode", C.c_int32), ("use_z_gate", C.c_int32),
                ("z_min", C.c_float), ("z_max", C.c_float), ("r_max", C.c_double)]


class GridDesc(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("resolution", C.c_double),
                ("length_x", C.c_double), ("length_y", C.c_double),
                ("pos_x", C.c_double), ("pos_y", C.c_double)]


class LShape(C.Structure):
    _fields_ = [("kept", C.c_int32), ("centroid_y", C.c_float), ("mean_z", C.c_float),
                ("mean_x", C.c_float), ("major_z", C.c_float), ("major_x", C.c_float),
                ("minor_z", C.c_float), ("minor_x", C.c_float), ("length", C.c_float),
                ("width", C.c_float), ("angle_deg", C.c_float), ("qx", C.c_double),
                ("qy", C.c_double), ("qz", C.c_double), ("qw", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("beams", C.c_uint64), ("cells_logical", C.c_uint64),
                ("cells_physical", C.c_uint64), ("distinct_ends", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("merges", C.c_uint64), ("merge_ms_last", C.c_double),
                ("deferred_points", C.c_uint64)]


# every symbol include/gridvision_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "gv_version", "gv_status_string", "gv_create", "gv_destroy", "gv_last_error",
    "gv_synchronize", "gv_join", "gv_stream", "gv_set_stream", "gv_get_stats",
    "gv_graph_begin", "gv_graph_end", "gv_graph_launch", "gv_graph_destroy", "gv_debug_pair_thresholds",
    "gv_set_cameras", "gv_fuse", "gv_fuse_aos32", "gv_fuse_dev", "gv_transform_points",
    "gv_project_kdtree", "gv_partition_by_label", "gv_segment_ground", "gv_bbox_pose",
    "gv_box_depths", "gv_pixels_to_3d",
    "gv_grid_init_reference", "gv_grid_init", "gv_grid_get_desc", "gv_grid_reset",
    "gv_grid_upload", "gv_grid_download", "gv_grid_counts_download", "gv_grid_layers_dev",
    "gv_grid_get_index",
    "gv_grid_update", "gv_grid_update_poses", "gv_grid_update_points", "gv_grid_update_corners",
    "gv_set_base_transform", "gv_grid_accumulate", "gv_grid_accumulate_dev",
    "gv_grid_raycast_flush", "gv_grid_finalize",
    "gv_process_batch", "gv_process_batch_dev",
    "gv_grid_to_occupancy",
    "gv_nccl_unique_id", "gv_nccl_init", "gv_nccl_world", "gv_grid_finalize_multi",
    "gv_ipc_export", "gv_ipc_import", "gv_ipc_close",
    "gv_microbench_atomics",
]

_lib = None


def load() -> C.CDLL:
    """dlopen the library (torch first, so its bundled libnccl/libcudart win the SONAME race)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GridVisionError(-1, "libgridvision_b200.so is not built",
                              f"expected {LIB_PATH}; run `make lib` or __graft_entry__.build()")
    import torch  # noqa: F401  (loads libnccl.so.2 / CUDA runtime the library binds to)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.gv_status_string.restype = C.c_char_p
    lib.gv_last_error.restype = C.c_char_p
    lib.gv_last_error.argtypes = [C.c_void_p]
    lib.gv_stream.restype = C.c_void_p
    lib.gv_stream.argtypes = [C.c_void_p]
    lib.gv_destroy.restype = None
    lib.gv_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib
