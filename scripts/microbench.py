#!/usr/bin/env python
"""Atomic-throughput ceilings of this GPU (gv_microbench_atomics): prints one JSON object."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(device=0, reps=256):
    from grid_vision_b200 import _lib
    lib = _lib.load()
    out = {}
    v = C.c_double()
    cases = [("red64_spread_4Mcells", 0, 0, 1 << 22, 0), ("red64_spread_64Mcells", 0, 0, 1 << 26, 0),
             ("red64_runs4_4Mcells", 0, 3, 1 << 22, 0), ("red64_contig_4Mcells", 0, 1, 1 << 22, 0),
             ("red64_same_4Mcells", 0, 4, 1 << 22, 0),
             ("red32_spread_4Mcells", 1, 0, 1 << 22, 0), ("red32_contig_4Mcells", 1, 1, 1 << 22, 0),
             ("red32_stride2048_4Mcells", 1, 2, 1 << 22, 2048), ("red32_stride8192_64Mcells", 1, 2, 1 << 26, 8192),
             ("red32_contig_64Mcells", 1, 1, 1 << 26, 0), ("red32_spread_64Mcells", 1, 0, 1 << 26, 0),
             ("atoms32_shared_conflict_free", 2, 0, 1 << 20, 0)]
    for name, kind, pat, nc, stride in cases:
        rc = lib.gv_microbench_atomics(C.c_int(device), C.c_int(kind), C.c_int(pat), C.c_size_t(nc),
                                       C.c_uint(stride), C.c_int(reps), C.byref(v))
        out[name] = v.value if rc == 0 else f"rc={rc}"
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
