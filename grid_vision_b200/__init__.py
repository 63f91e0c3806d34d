"""grid_vision_b200 — B200 (sm_100a) implementation of grid-vision's point-cloud ->
occupancy-grid hot path behind a C ABI (include/gridvision_b200.h).

Layout: csrc/ (CUDA kernels + C ABI), lib/ (built .so, git-ignored), context.py (ctypes
binding), reference_api.py (Python mirror of the reference's C++ entry points),
synth.py (synthetic BASELINE.json workloads), sharding.py (frame sharding across GPUs).
"""
from ._lib import (F_CLIPPED, F_HIT, F_RANGECAP, F_VALID, OCC_ALL, OCC_LABELLED,  # noqa: F401
                   GridVisionError)
from .context import Context, accum_params  # noqa: F401

__all__ = ["Context", "accum_params", "GridVisionError", "OCC_ALL", "OCC_LABELLED",
           "F_VALID", "F_HIT", "F_CLIPPED", "F_RANGECAP"]
