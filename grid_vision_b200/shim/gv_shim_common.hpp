// gv_shim_common.hpp — shared helpers of the drop-in translation units.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "gridvision_b200.h"

namespace gv_shim {

// One library context per process for the free functions of namespace cloud_detections
// (the reference node is single-threaded, ref: src/grid_vision_node.cpp:533-540).
// Device from $GV_DEVICE (default 0).  There is no CPU fallback: if the context cannot be
// created the shim reports it once and the calls produce the reference's own "nothing found"
// results (empty clouds / empty vector), mirroring its silent error conventions.
inline gv_ctx *context()
{
  static gv_ctx *ctx = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char *d = std::getenv("GV_DEVICE");
    const int rc = gv_create(&ctx, d ? std::atoi(d) : 0);
    if (rc != GV_OK) {
      std::fprintf(stderr, "[grid_vision_b200] gv_create failed: %s\n", gv_status_string(rc));
      ctx = nullptr;
    }
  }
  return ctx;
}

inline bool ok(gv_ctx *ctx, int rc, const char *what)
{
  if (rc == GV_OK) return true;
  std::fprintf(stderr, "[grid_vision_b200] %s failed: %s (%s)\n", what, gv_status_string(rc),
               ctx ? gv_last_error(ctx) : "");
  return false;
}

}  // namespace gv_shim
