// transform_lidar_to_camera_b200.hpp — drop-in for the point-cloud transform inside
// GridVision::transformLidarToCamera (ref: src/grid_vision_node.cpp:280-307; decl
// include/grid_vision/grid_vision_node.hpp:95-97).  The member itself needs the node's tf buffer,
// so what is replaced is its one compute call, `pcl_ros::transformPointCloud(lidar_cloud,
// *transformed_cloud, tf_transform)` at :304 -> `gv_shim::transformPointCloud(...)` (same
// arguments, see INTEGRATION.md for the guarded two-line edit).
#pragma once
#include <vector>

#include "gv_shim_common.hpp"

namespace gv_shim {

// pcl_ros semantics: tf2::Transform -> float 4x4 (doubles narrowed), every field of every point
// copied, xyz through PCL's Transformer<float>::se3 on the GPU (gv_transform_points); with
// is_dense == false non-finite points pass through untouched.
template <typename CloudT, typename TransformT>
inline void transformPointCloud(const CloudT &in, CloudT &out, const TransformT &t)
{
  out = in;
  const size_t n = in.points.size();
  gv_ctx *ctx = context();
  if(!ctx || n == 0)
    return;
  float T[16] = {0};
  for(int r = 0; r < 3; ++r)
  {
    T[4 * r + 0] = static_cast<float>(t.getBasis()[r].x());
    T[4 * r + 1] = static_cast<float>(t.getBasis()[r].y());
    T[4 * r + 2] = static_cast<float>(t.getBasis()[r].z());
  }
  T[3] = static_cast<float>(t.getOrigin().x());
  T[7] = static_cast<float>(t.getOrigin().y());
  T[11] = static_cast<float>(t.getOrigin().z());
  T[15] = 1.0f;
  const double K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const int32_t wh[2] = {1, 1};
  if(!ok(ctx, gv_set_cameras(ctx, 1, K, T, wh), "gv_set_cameras"))
    return;
  std::vector<float> x(n), y(n), z(n), ox(n), oy(n), oz(n);
  for(size_t i = 0; i < n; ++i)
  {
    x[i] = in.points[i].x;
    y[i] = in.points[i].y;
    z[i] = in.points[i].z;
  }
  if(!ok(ctx, gv_transform_points(ctx, 0, x.data(), y.data(), z.data(), n, in.is_dense ? 1 : 0, ox.data(),
                                  oy.data(), oz.data()),
         "gv_transform_points"))
    return;
  for(size_t i = 0; i < n; ++i)
  {
    out.points[i].x = ox[i];
    out.points[i].y = oy[i];
    out.points[i].z = oz[i];
  }
}

}  // namespace gv_shim
