// occupancy_grid_b200.cpp — drop-in definitions of class OccupancyGridMap
// (ref: include/grid_vision/occupancy_grid.hpp:13-40, src/occupancy_grid.cpp:4-196) over the
// C ABI.  Build it INSTEAD of src/occupancy_grid.cpp.  The header is untouched, so the library
// context of each map lives in a side table keyed by the object's grid_map_ member.
#include "grid_vision/occupancy_grid.hpp"

#include <cstring>
#include <map>

#include "gv_shim_common.hpp"

namespace
{
  struct MapState
  {
    gv_ctx *ctx = nullptr;
    int nx = 0, ny = 0;
  };
  std::map<const grid_map::GridMap *, MapState> &states()
  {
    static std::map<const grid_map::GridMap *, MapState> s;
    return s;
  }

  // The grid the node passes is its own public member grid_map_ (ref:
  // src/grid_vision_node.cpp:145,208,230,235).  Host-authoritative drop-in: the log_odds layer
  // is uploaded, updated on the GPU, and both layers are written back, so any host-side edit of
  // the map between calls is honoured exactly like in the reference.
  template <typename F> void run_update(grid_map::GridMap &grid_map, F &&update)
  {
    auto it = states().find(&grid_map);
    if(it == states().end() || !it->second.ctx)
      return;
    gv_ctx *ctx = it->second.ctx;
    Eigen::MatrixXf &lo = grid_map["log_odds"];
    Eigen::MatrixXf &oc = grid_map["occupancy"];
    if(!gv_shim::ok(ctx, gv_grid_upload(ctx, lo.data(), nullptr), "gv_grid_upload"))
      return;
    if(!gv_shim::ok(ctx, update(ctx), "gv_grid_update"))
      return;
    gv_shim::ok(ctx, gv_grid_download(ctx, lo.data(), oc.data()), "gv_grid_download");
  }
}

OccupancyGridMap::OccupancyGridMap(const std::string &base_link, uint8_t grid_x,
                                   uint8_t grid_y, double resolution)
{
  // host-side container exactly as the reference builds it (:8-13) ...
  grid_map_ = grid_map::GridMap({"log_odds", "occupancy"});
  grid_map_.setFrameId(base_link);
  grid_map_.setGeometry(grid_map::Length(grid_x, grid_y), resolution);
  grid_map_.setPosition(grid_map::Position(grid_x / 3, 0.0));
  grid_map_["log_odds"].setConstant(log_odds_prior_);
  grid_map_["occupancy"].setConstant(init_probability_);
  // ... and its device twin with the same geometry
  MapState st;
  const char *d = std::getenv("GV_DEVICE");
  if(gv_create(&st.ctx, d ? std::atoi(d) : 0) == GV_OK
     && gv_shim::ok(st.ctx, gv_grid_init_reference(st.ctx, grid_x, grid_y, resolution),
                    "gv_grid_init_reference"))
  {
    gv_grid_desc desc;
    gv_grid_get_desc(st.ctx, &desc);
    st.nx = desc.nx;
    st.ny = desc.ny;
  }
  else
  {
    std::fprintf(stderr, "[grid_vision_b200] no B200 context: OccupancyGridMap updates are disabled\n");
    if(st.ctx)
      gv_destroy(st.ctx);
    st.ctx = nullptr;
  }
  states()[&grid_map_] = st;
}

// ref: src/occupancy_grid.cpp:16-31
void OccupancyGridMap::updateMap(grid_map::GridMap &grid_map)
{
  run_update(grid_map, [](gv_ctx *ctx) { return gv_grid_update(ctx); });
}

// ref: src/occupancy_grid.cpp:33-63 (+ :107-138 corners and :185-196 depths, evaluated on the device)
void OccupancyGridMap::updateMap(grid_map::GridMap &grid_map,
                                 const std::vector<geometry_msgs::msg::Point> &base_points,
                                 const std::vector<BoundingBox> &bboxes)
{
  const int n = static_cast<int>(base_points.size());
  std::vector<double> xy(2 * static_cast<size_t>(n));
  std::vector<int32_t> labels(n);
  for(int i = 0; i < n; ++i)
  {
    xy[2 * i + 0] = base_points[i].x;
    xy[2 * i + 1] = base_points[i].y;
    labels[i] = static_cast<int32_t>(bboxes[i].label);
  }
  run_update(grid_map, [&](gv_ctx *ctx) { return gv_grid_update_points(ctx, xy.data(), labels.data(), n); });
}

// ref: src/occupancy_grid.cpp:65-105 (+ :140-183)
void OccupancyGridMap::updateMap(grid_map::GridMap &grid_map,
                                 const std::vector<LShapePose> &bboxes_pose)
{
  const int n = static_cast<int>(bboxes_pose.size());
  std::vector<double> xylw(4 * static_cast<size_t>(n));
  for(int i = 0; i < n; ++i)
  {
    xylw[4 * i + 0] = bboxes_pose[i].pose.position.x;
    xylw[4 * i + 1] = bboxes_pose[i].pose.position.y;
    xylw[4 * i + 2] = bboxes_pose[i].length;
    xylw[4 * i + 3] = bboxes_pose[i].width;
  }
  run_update(grid_map, [&](gv_ctx *ctx) { return gv_grid_update_poses(ctx, xylw.data(), n); });
}

// The three private helpers are part of the class, so they stay defined; the update paths above
// evaluate their arithmetic on the device (k_footprint_rects), these host versions serve any
// other caller.  ref: src/occupancy_grid.cpp:107-138
std::array<geometry_msgs::msg::Point, 4>
OccupancyGridMap::computeBoundingBox3D(const geometry_msgs::msg::Point &base_center,
                                       ObjectClass label)
{
  std::array<geometry_msgs::msg::Point, 4> corners;
  const float d = getEstimatedDepth(label);
  const double sx[4] = {1.0, 1.0, 0.0, 0.0}, sy[4] = {0.5, -0.5, -0.5, 0.5};
  for(int i = 0; i < 4; ++i)
  {
    corners[i].x = base_center.x + sx[i] * d;
    corners[i].y = base_center.y + sy[i] * d;
    corners[i].z = base_center.z;
  }
  return corners;
}

// ref: src/occupancy_grid.cpp:140-183.  Private and not used by the update paths above; note
// that routing it through gv_grid_finalize(k_decay = 0) also clamps the layer, which the
// reference's helper alone does not do (its callers clamp right after).
void OccupancyGridMap::updateGridCellsFast(
  grid_map::GridMap &grid_map, const std::array<geometry_msgs::msg::Point, 4> &bbox_corners)
{
  auto it = states().find(&grid_map);
  if(it == states().end() || !it->second.ctx)
    return;
  gv_ctx *ctx = it->second.ctx;
  double c[8];
  for(int i = 0; i < 4; ++i)
  {
    c[2 * i + 0] = bbox_corners[i].x;
    c[2 * i + 1] = bbox_corners[i].y;
  }
  // one footprint, no decay: finalize with k_decay = 0 leaves every other cell's log-odds as is
  Eigen::MatrixXf &lo = grid_map["log_odds"];
  if(gv_shim::ok(ctx, gv_grid_upload(ctx, lo.data(), nullptr), "gv_grid_upload")
     && gv_shim::ok(ctx, gv_grid_finalize(ctx, 0, c, 1), "gv_grid_finalize"))
    gv_shim::ok(ctx, gv_grid_download(ctx, lo.data(), nullptr), "gv_grid_download");
}

// ref: src/occupancy_grid.cpp:185-196
float OccupancyGridMap::getEstimatedDepth(ObjectClass class_label)
{
  switch(class_label)
  {
  case ObjectClass::VEHICLE: return 3.5f;
  case ObjectClass::PERSON: return 0.6f;
  case ObjectClass::BIKE: return 2.5f;
  case ObjectClass::MOTORBIKE: return 2.5f;
  default: return -1.0f;
  }
}
