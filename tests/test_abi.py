"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/gridvision_b200.h declares; without a GPU it refuses to create a context (no CPU
fallback); the reference's struct layouts are binary-compatible."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gridvision_b200.h")).read()
    return sorted(set(re.findall(r"GV_API\s+[\w\s\*]+?\b(gv_\w+)\s*\(", hdr)))


def test_header_symbols_are_exported():
    from grid_vision_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == syms, set(_lib.EXPORTS) ^ set(syms)
    assert lib.gv_version() == 100
    assert lib.gv_status_string(0) == b"ok"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        return  # exercised by the -m gpu tests
    import grid_vision_b200 as gv
    try:
        gv.Context(0)
    except gv.GridVisionError as e:
        assert e.status == 6  # GV_ERR_NO_DEVICE
    else:
        raise AssertionError("Context() must fail without a B200")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: the product may mention it in comments (each kernel
    names the oracle function it mirrors) but never imports, includes or links it."""
    pkg = os.path.join(ROOT, "grid_vision_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dp, f)
            if f.endswith(".py"):
                src = open(path).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), path
                assert "gv_oracle" not in src, path
            elif f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(path).read()
                assert not re.search(r"#include\s*[<\"][^>\"]*oracle", src), path
                assert "dlopen" not in src, path
    mk = open(os.path.join(ROOT, "Makefile")).read()
    lib_rule = mk[mk.index("$(LIB):"):mk.index("oracle:")]
    assert "oracle" not in lib_rule


def test_reference_struct_layouts():
    from grid_vision_b200.context import POINT_DTYPE
    from grid_vision_b200.synth import BOX_DTYPE
    # include/grid_vision/object_detection.hpp:27-32: 4 doubles, float, enum -> 40 bytes
    assert BOX_DTYPE.itemsize == 40
    assert [BOX_DTYPE.fields[k][1] for k in ("x_min", "y_min", "x_max", "y_max", "confidence", "label")] == \
        [0, 8, 16, 24, 32, 36]
    assert POINT_DTYPE.itemsize == 32  # pcl::PointXYZI
    from grid_vision_b200._lib import AccumParams, GridDesc, Stats
    assert C.sizeof(AccumParams) == 24 and C.sizeof(GridDesc) == 48 and C.sizeof(Stats) == 64


def test_synthetic_workloads_match_baseline_configs():
    from grid_vision_b200 import synth
    assert synth.C1.points == 131072 and synth.C1.boxes_per_camera == 20 and synth.C1.cells == 200 * 200
    assert synth.C2.points == 262144 and synth.C2.boxes_per_camera == 50 and synth.C2.cells == 1000 * 1000
    assert synth.C3.frames == 4096 and synth.C3.points == 4096 * 131072 and synth.C3.cells == 2048 * 2048
    assert synth.C4.points == 1048576 and synth.C4.cameras * synth.C4.boxes_per_camera == 300
    assert synth.C5.points == 16777216 and synth.C5.cells == 8192 * 8192 and synth.C5.r_max == 120.0
    b = synth.make_boxes(synth.C2)
    assert np.all(b["x_min"] == np.floor(b["x_min"])) and np.all(np.diff(b["confidence"]) <= 0)
    T = synth.camera_extrinsics(6)
    assert T.shape == (6, 4, 4) and T.dtype == np.float32
    # deterministic generator
    a = synth.make_scans(synth.C1.scaled(rings=4, azimuth=64), frames=2).numpy()
    c = synth.make_scans(synth.C1.scaled(rings=4, azimuth=64), frames=2).numpy()
    assert np.array_equal(a, c, equal_nan=True) and np.isnan(a).any()
