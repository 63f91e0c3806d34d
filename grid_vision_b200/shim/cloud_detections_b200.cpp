// cloud_detections_b200.cpp — drop-in definitions of the four hot-path functions of namespace
// cloud_detections, with the reference's exact signatures
// (ref: include/grid_vision/cloud_detections.hpp:29-30,46-48,50-52), over the C ABI.
//
// Build it INSTEAD of the bodies at ref: src/cloud_detections.cpp:8-40, 43-87, 250-298, 300-321 (guard
// those with `#ifndef GRID_VISION_B200`); the rest of that file
// (segmentGroundPlane, bboxPoseEstimation, computePCABoundingBox, ...) keeps running unchanged.
#include "grid_vision/cloud_detections.hpp"

#include <cstring>

#include "gv_shim_common.hpp"

static_assert(sizeof(BoundingBox) == sizeof(gv_box), "BoundingBox must stay 40 bytes");
static_assert(sizeof(pcl::PointXYZI) == sizeof(gv_point_xyzi), "pcl::PointXYZI must stay 32 bytes");

namespace cloud_detections
{
  namespace
  {
    bool set_camera(gv_ctx *ctx, const Eigen::Matrix3d &K, int image_width, int image_height)
    {
      double Kd[9];
      for(int r = 0; r < 3; ++r)
        for(int c = 0; c < 3; ++c)
          Kd[3 * r + c] = K(r, c);
      const int32_t wh[2] = {image_width, image_height};
      // no extrinsic: both callers hand over a cloud that is already in the camera frame
      return gv_shim::ok(ctx, gv_set_cameras(ctx, 1, Kd, nullptr, wh), "gv_set_cameras");
    }
  }

  // ref: src/cloud_detections.cpp:8-40 — the per-point projection loop runs on the GPU (order
  // preserving compaction); the FLANN tree itself is not on the hot path and stays as it was.
  void buildKDTree(pcl::KdTreeFLANN<pcl::PointXYZ> &kdtree,
                   pcl::PointCloud<pcl::PointXYZ>::Ptr image_points,
                   const pcl::PointCloud<pcl::PointXYZI>::Ptr lidar_points,
                   const Eigen::Matrix3d &K)
  {
    gv_ctx *ctx = gv_shim::context();
    const size_t n = lidar_points->points.size();
    if(ctx && n > 0 && set_camera(ctx, K, 1, 1))
    {
      std::vector<float> x(n), y(n), z(n), uvz(3 * n);
      for(size_t i = 0; i < n; ++i)
      {
        x[i] = lidar_points->points[i].x;
        y[i] = lidar_points->points[i].y;
        z[i] = lidar_points->points[i].z;
      }
      size_t m = 0;
      if(gv_shim::ok(ctx, gv_project_kdtree(ctx, 0, x.data(), y.data(), z.data(), n, uvz.data(), &m),
                     "gv_project_kdtree"))
      {
        for(size_t i = 0; i < m; ++i)
        {
          pcl::PointXYZ pt;
          pt.x = uvz[3 * i + 0];
          pt.y = uvz[3 * i + 1];
          pt.z = uvz[3 * i + 2];
          image_points->push_back(pt);
        }
      }
    }
    if(!image_points->empty())
    {
      kdtree.setInputCloud(image_points);
    }
  }

  // ref: src/cloud_detections.cpp:43-87 — the k-NN median depth of every box on the GPU (exact
  // brute-force search over the (u, v, depth) triples; the FLANN tree argument is not consulted).
  std::vector<float>
  computeDepthForBoundingBoxes(pcl::KdTreeFLANN<pcl::PointXYZ> &,
                               pcl::PointCloud<pcl::PointXYZ>::Ptr image_points,
                               const std::vector<BoundingBox> &bboxes, uint16_t k)
  {
    std::vector<float> depths(bboxes.size(), -1.0f);
    gv_ctx *ctx = gv_shim::context();
    const size_t m = image_points ? image_points->points.size() : 0;
    if(!ctx || bboxes.empty() || m == 0 || k == 0)
      return depths;
    std::vector<float> uvz(3 * m);
    for(size_t i = 0; i < m; ++i)
    {
      uvz[3 * i + 0] = image_points->points[i].x;
      uvz[3 * i + 1] = image_points->points[i].y;
      uvz[3 * i + 2] = image_points->points[i].z;
    }
    gv_shim::ok(ctx,
                gv_box_depths(ctx, uvz.data(), m, reinterpret_cast<const gv_box *>(bboxes.data()),
                              static_cast<int>(bboxes.size()), k > 64 ? 64 : k, depths.data()),
                "gv_box_depths");
    return depths;
  }

  // ref: src/cloud_detections.cpp:250-298
  void
  extractCloudPerBBox(const pcl::PointCloud<pcl::PointXYZI> &cloud,
                      const Eigen::Matrix3d &K, const std::vector<BoundingBox> &bboxes,
                      std::vector<pcl::PointCloud<pcl::PointXYZI>> &output_clouds,
                      int image_width, int image_height)
  {
    output_clouds.clear();
    output_clouds.resize(bboxes.size());
    for(auto &c : output_clouds)
    {
      c.width = 0;
      c.height = 1;
      c.is_dense = true;
    }
    gv_ctx *ctx = gv_shim::context();
    const size_t n = cloud.points.size();
    const int nb = static_cast<int>(bboxes.size());
    if(!ctx || n == 0 || nb == 0 || !set_camera(ctx, K, image_width, image_height))
      return;

    std::vector<int16_t> labels(n);
    if(!gv_shim::ok(ctx,
                    gv_fuse_aos32(ctx, reinterpret_cast<const gv_point_xyzi *>(cloud.points.data()), n,
                                  1, reinterpret_cast<const gv_box *>(bboxes.data()), nb, nullptr,
                                  labels.data(), nullptr, nullptr),
                    "gv_fuse_aos32"))
      return;

    // per-box clouds in input order (the push_back order of :285)
    std::vector<uint32_t> indices(n);
    std::vector<uint64_t> offsets(static_cast<size_t>(nb) + 1, 0);
    if(nb <= 1024)
    {
      if(!gv_shim::ok(ctx,
                      gv_partition_by_label(ctx, labels.data(), n, nb, indices.data(), offsets.data()),
                      "gv_partition_by_label"))
        return;
    }
    else
    {
      for(size_t i = 0; i < n; ++i)
        if(labels[i] >= 0)
          offsets[labels[i] + 1]++;
      for(int b = 0; b < nb; ++b)
        offsets[b + 1] += offsets[b];
      std::vector<uint64_t> cur(offsets.begin(), offsets.end() - 1);
      for(size_t i = 0; i < n; ++i)
        if(labels[i] >= 0)
          indices[cur[labels[i]]++] = static_cast<uint32_t>(i);
    }
    for(int b = 0; b < nb; ++b)
    {
      auto &c = output_clouds[b];
      c.points.resize(offsets[b + 1] - offsets[b]);
      for(uint64_t k = offsets[b]; k < offsets[b + 1]; ++k)
        c.points[k - offsets[b]] = cloud.points[indices[k]];
      c.width = c.points.size(); // :292-297
      c.height = 1;
      c.is_dense = true;
    }
  }

  // ref: src/cloud_detections.cpp:300-321 — same orchestration, same error convention
  std::vector<LShapePose>
  computeBBoxPose(const pcl::PointCloud<pcl::PointXYZI>::Ptr &input_cloud,
                  const Eigen::Matrix3d &K, const std::vector<BoundingBox> &bboxes,
                  int image_height, int image_width)
  {
    pcl::PointCloud<pcl::PointXYZI> segmented_cloud = segmentGroundPlane(input_cloud);

    if(segmented_cloud.empty())
      return {};

    std::vector<pcl::PointCloud<pcl::PointXYZI>> output_clouds;
    extractCloudPerBBox(segmented_cloud, K, bboxes, output_clouds, image_width, image_height);

    std::vector<LShapePose> bboxes_pose;
    bboxPoseEstimation(output_clouds, bboxes_pose);

    return bboxes_pose;
  }
}
