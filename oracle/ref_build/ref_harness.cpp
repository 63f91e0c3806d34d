// ref_harness.cpp — C entry points around the REFERENCE'S OWN functions, compiled from
// /root/reference/src/{occupancy_grid,cloud_detections}.cpp against oracle/ref_build/stubs.
// TEST INFRASTRUCTURE: used by tests/test_oracle_vs_ref.py to pin oracle/gv_oracle.c.
#include "grid_vision/cloud_detections.hpp"
#include "grid_vision/occupancy_grid.hpp"

#include <cstring>
#include <optional>

#ifdef GV_SHIM_BUILD
#include "transform_lidar_to_camera_b200.hpp"
#define GV_TRANSFORM_CLOUD gv_shim::transformPointCloud
#else
#define GV_TRANSFORM_CLOUD pcl_ros::transformPointCloud
#endif

static_assert(sizeof(BoundingBox) == 40, "BoundingBox layout (object_detection.hpp:27-32)");

extern "C" {

// cloud_detections::extractCloudPerBBox (src/cloud_detections.cpp:250-298) on camera-frame
// points.  The point index rides in the intensity lane; labels[i] = box whose output cloud
// received point i (-1 = none).  Returns 1 iff every output cloud kept input order and got
// width = size, height = 1, is_dense = true (:292-297).
int ref_extract_cloud_per_bbox(const float *x, const float *y, const float *z, size_t n,
                               const double *K, const BoundingBox *boxes, int nb, int W, int H,
                               int16_t *labels)
{
  pcl::PointCloud<pcl::PointXYZI> cloud;
  cloud.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    cloud.points[i].x = x[i];
    cloud.points[i].y = y[i];
    cloud.points[i].z = z[i];
    cloud.points[i].intensity = (float)i;
  }
  Eigen::Matrix3d Km;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Km(r, c) = K[3 * r + c];
  std::vector<BoundingBox> bb(boxes, boxes + nb);
  std::vector<pcl::PointCloud<pcl::PointXYZI>> out(3);  // must be cleared + resized by the callee
  cloud_detections::extractCloudPerBBox(cloud, Km, bb, out, W, H);
  int ok = (int)out.size() == nb;
  for (size_t i = 0; i < n; ++i) labels[i] = -1;
  for (int b = 0; b < nb && ok; ++b) {
    float prev = -1.0f;
    for (const auto &p : out[b].points) {
      if (!(p.intensity > prev)) ok = 0;
      prev = p.intensity;
      if (labels[(size_t)p.intensity] != -1) ok = 0;  // a point may land in one box only
      labels[(size_t)p.intensity] = (int16_t)b;
    }
    if (out[b].width != out[b].points.size() || out[b].height != 1 || !out[b].is_dense) ok = 0;
  }
  return ok;
}

// cloud_detections::buildKDTree projection loop (src/cloud_detections.cpp:8-40)
size_t ref_build_kdtree(const float *x, const float *y, const float *z, size_t n, const double *K,
                        float *uvz)
{
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>);
  cloud->points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    cloud->points[i].x = x[i];
    cloud->points[i].y = y[i];
    cloud->points[i].z = z[i];
  }
  Eigen::Matrix3d Km;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Km(r, c) = K[3 * r + c];
  pcl::PointCloud<pcl::PointXYZ>::Ptr img(new pcl::PointCloud<pcl::PointXYZ>);
  pcl::KdTreeFLANN<pcl::PointXYZ> tree;
  cloud_detections::buildKDTree(tree, img, cloud, Km);
  for (size_t i = 0; i < img->points.size(); ++i) {
    uvz[3 * i + 0] = img->points[i].x;
    uvz[3 * i + 1] = img->points[i].y;
    uvz[3 * i + 2] = img->points[i].z;
  }
  return img->points.size();
}

// The compute call of GridVision::transformLidarToCamera (src/grid_vision_node.cpp:296-304):
// tf2::Transform -> pcl_ros::transformPointCloud.  R: row-major 3x3 doubles, t: origin.
void ref_transform_cloud(const float *x, const float *y, const float *z, size_t n, int is_dense, const double *R,
                         const double *t, float *ox, float *oy, float *oz, float *intensity_out)
{
  pcl::PointCloud<pcl::PointXYZI> in, out;
  in.points.resize(n);
  in.is_dense = is_dense != 0;
  for (size_t i = 0; i < n; ++i) {
    in.points[i].x = x[i];
    in.points[i].y = y[i];
    in.points[i].z = z[i];
    in.points[i].intensity = (float)i;
  }
  tf2::Transform tf;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) tf.basis[r].v[c] = R[3 * r + c];
    tf.origin.v[r] = t[r];
  }
  GV_TRANSFORM_CLOUD(in, out, tf);
  for (size_t i = 0; i < n; ++i) {
    ox[i] = out.points[i].x;
    oy[i] = out.points[i].y;
    oz[i] = out.points[i].z;
    intensity_out[i] = out.points[i].intensity;
  }
}

// cloud_detections::computeDepthForBoundingBoxes (src/cloud_detections.cpp:43-87) on (u, v, depth)
// triples as buildKDTree leaves them (tree over the same cloud), k as the node passes k_near_
void ref_box_depths(const float *uvz, size_t m, const BoundingBox *boxes, int nb, int k, float *depths)
{
  pcl::PointCloud<pcl::PointXYZ>::Ptr img(new pcl::PointCloud<pcl::PointXYZ>);
  img->points.resize(m);
  for (size_t i = 0; i < m; ++i) {
    img->points[i].x = uvz[3 * i + 0];
    img->points[i].y = uvz[3 * i + 1];
    img->points[i].z = uvz[3 * i + 2];
  }
  pcl::KdTreeFLANN<pcl::PointXYZ> tree;
  if (!img->empty()) tree.setInputCloud(img);
  std::vector<BoundingBox> bb(boxes, boxes + nb);
  const std::vector<float> d = cloud_detections::computeDepthForBoundingBoxes(tree, img, bb, (uint16_t)k);
  for (int i = 0; i < nb; ++i) depths[i] = d[i];
}

// cloud_detections::pixelTo3D (src/cloud_detections.cpp:89-103)
void ref_pixel_to_3d(const double *Kinv, float px, float py, float depth, double *out)
{
  Eigen::Matrix3d Ki;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Ki(r, c) = Kinv[3 * r + c];
  const geometry_msgs::msg::Point p = cloud_detections::pixelTo3D(cv::Point2f(px, py), depth, Ki);
  out[0] = p.x;
  out[1] = p.y;
  out[2] = p.z;
}

// cloud_detections::computeBBoxPose error convention (src/cloud_detections.cpp:308-309):
// an empty ground-removed cloud yields an empty result
int ref_compute_bbox_pose_empty()
{
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>);
  Eigen::Matrix3d Km;
  std::vector<BoundingBox> bb(2);
  return (int)cloud_detections::computeBBoxPose(cloud, Km, bb, 416, 416).size();
}

// OccupancyGridMap (src/occupancy_grid.cpp:4-196)
void *ref_grid_new(uint8_t gx, uint8_t gy, double res) { return new OccupancyGridMap("base", gx, gy, res); }
void ref_grid_free(void *g) { delete static_cast<OccupancyGridMap *>(g); }

// The node's own construction pattern (src/grid_vision_node.cpp:35 with the member declared at
// include/grid_vision/grid_vision_node.hpp:64): a temporary assigned into a std::optional, i.e.
// the object every later updateMap sees is a moved copy, not the one the constructor ran on.
void *ref_grid_new_like_node(uint8_t gx, uint8_t gy, double res)
{
  std::optional<OccupancyGridMap> occ_grid_;
  occ_grid_ = OccupancyGridMap("base", gx, gy, res);
  return new OccupancyGridMap(std::move(*occ_grid_));
}

void ref_grid_desc(void *g, int *nx, int *ny, double *len, double *pos, double *res)
{
  auto &m = static_cast<OccupancyGridMap *>(g)->grid_map_;
  *nx = m.getSize()(0);
  *ny = m.getSize()(1);
  len[0] = m.getLength()(0);
  len[1] = m.getLength()(1);
  pos[0] = m.getPosition()(0);
  pos[1] = m.getPosition()(1);
  *res = m.getResolution();
}

void ref_grid_read(void *g, float *lo, float *occ)
{
  auto &m = static_cast<OccupancyGridMap *>(g)->grid_map_;
  std::memcpy(lo, m["log_odds"].d.data(), m["log_odds"].d.size() * sizeof(float));
  std::memcpy(occ, m["occupancy"].d.data(), m["occupancy"].d.size() * sizeof(float));
}

void ref_grid_write_log_odds(void *g, const float *lo)
{
  auto &m = static_cast<OccupancyGridMap *>(g)->grid_map_;
  std::memcpy(m["log_odds"].d.data(), lo, m["log_odds"].d.size() * sizeof(float));
}

int ref_grid_get_index(void *g, double x, double y, int *ix, int *iy)
{
  auto &m = static_cast<OccupancyGridMap *>(g)->grid_map_;
  grid_map::Index idx;
  const bool ok = m.getIndex(grid_map::Position(x, y), idx);
  *ix = idx.x();
  *iy = idx.y();
  return ok ? 1 : 0;
}

// the three overloads, called the way the node calls them (src/grid_vision_node.cpp:145,208,230,235)
void ref_update_map(void *g)
{
  auto *o = static_cast<OccupancyGridMap *>(g);
  o->updateMap(o->grid_map_);
}

void ref_update_map_poses(void *g, const double *xylw, int n)
{
  auto *o = static_cast<OccupancyGridMap *>(g);
  std::vector<LShapePose> poses(n);
  for (int i = 0; i < n; ++i) {
    poses[i].pose.position.x = xylw[4 * i + 0];
    poses[i].pose.position.y = xylw[4 * i + 1];
    poses[i].length = xylw[4 * i + 2];
    poses[i].width = xylw[4 * i + 3];
    poses[i].height = 1.0;
  }
  o->updateMap(o->grid_map_, poses);
}

void ref_update_map_points(void *g, const double *xy, const int32_t *labels, int n)
{
  auto *o = static_cast<OccupancyGridMap *>(g);
  std::vector<geometry_msgs::msg::Point> pts(n);
  std::vector<BoundingBox> bb(n);
  for (int i = 0; i < n; ++i) {
    pts[i].x = xy[2 * i + 0];
    pts[i].y = xy[2 * i + 1];
    bb[i].label = static_cast<ObjectClass>(labels[i]);
  }
  o->updateMap(o->grid_map_, pts, bb);
}

}  // extern "C"
