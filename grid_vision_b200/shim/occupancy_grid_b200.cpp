// occupancy_grid_b200.cpp — drop-in definitions of class OccupancyGridMap
// (ref: include/grid_vision/occupancy_grid.hpp:13-40, src/occupancy_grid.cpp:4-196) over the
// C ABI.  Build it INSTEAD of src/occupancy_grid.cpp.  The header is untouched, so there is no
// member to hang a library context on, and the object is copied/moved by the node
// (`occ_grid_ = OccupancyGridMap(...)` into a std::optional, ref: src/grid_vision_node.cpp:35,
// include/grid_vision/grid_vision_node.hpp:64).  The device twin is therefore found from the
// GEOMETRY of the grid_map the caller passes (size, resolution, position) and created on first
// use, never from an object address.
#include "grid_vision/occupancy_grid.hpp"

#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "gv_shim_common.hpp"

namespace
{
  struct MapState
  {
    gv_ctx *ctx = nullptr;
    // log_odds as the device last left it on the host: when the caller's layer still equals it
    // the device copy is current and the upload is skipped (host edits are detected, not assumed)
    std::vector<float> shadow;
    bool shadow_valid = false;
  };
  using GeomKey = std::tuple<int, int, double, double, double>;

  struct Registry
  {
    std::map<GeomKey, MapState> states;
    ~Registry()
    {
      for(auto &kv : states)
        if(kv.second.ctx)
          gv_destroy(kv.second.ctx);
    }
  };
  Registry &registry()
  {
    static Registry r;
    return r;
  }

  // One context per distinct map geometry (the node has exactly one map).  nullptr when no B200
  // context can be had: the update is then skipped, like the reference's silent error paths.
  MapState *state_for(const grid_map::GridMap &grid_map)
  {
    const auto size = grid_map.getSize();
    const auto pos = grid_map.getPosition();
    const double res = grid_map.getResolution();
    if(size(0) <= 0 || size(1) <= 0 || !(res > 0.0))
      return nullptr;
    const GeomKey key(size(0), size(1), res, pos(0), pos(1));
    auto &states = registry().states;
    auto it = states.find(key);
    if(it != states.end())
      return it->second.ctx ? &it->second : nullptr;
    MapState st;
    const char *d = std::getenv("GV_DEVICE");
    // gv_grid_init takes lengths: size * resolution reproduces the size exactly (setGeometry's
    // own length_ = size * resolution, grid_map::GridMap::setGeometry)
    if(gv_create(&st.ctx, d ? std::atoi(d) : 0) != GV_OK
       || !gv_shim::ok(st.ctx, gv_grid_init(st.ctx, size(0) * res, size(1) * res, res, pos(0), pos(1)),
                       "gv_grid_init"))
    {
      std::fprintf(stderr, "[grid_vision_b200] no B200 context: OccupancyGridMap updates are disabled\n");
      if(st.ctx)
        gv_destroy(st.ctx);
      st.ctx = nullptr;
    }
    else
    {
      gv_grid_desc desc;
      gv_grid_get_desc(st.ctx, &desc);
      if(desc.nx != size(0) || desc.ny != size(1))
      {
        std::fprintf(stderr, "[grid_vision_b200] device grid %dx%d != host grid %dx%d\n", desc.nx, desc.ny,
                     (int)size(0), (int)size(1));
        gv_destroy(st.ctx);
        st.ctx = nullptr;
      }
    }
    auto &slot = states[key];
    slot = std::move(st);
    return slot.ctx ? &slot : nullptr;
  }

  // The grid the node passes is its own public member grid_map_ (ref:
  // src/grid_vision_node.cpp:145,208,230,235).  Host-authoritative drop-in: both layers are
  // written back after every update, and a host-side edit of log_odds between calls is honoured
  // exactly like in the reference (the layer is re-uploaded whenever it differs from what the
  // device last wrote).
  template <typename F> void run_update(grid_map::GridMap &grid_map, F &&update, bool with_occupancy = true)
  {
    MapState *st = state_for(grid_map);
    if(!st)
      return;
    gv_ctx *ctx = st->ctx;
    Eigen::MatrixXf &lo = grid_map["log_odds"];
    Eigen::MatrixXf &oc = grid_map["occupancy"];
    const size_t n = static_cast<size_t>(grid_map.getSize()(0)) * static_cast<size_t>(grid_map.getSize()(1));
    const bool current = st->shadow_valid && st->shadow.size() == n
                         && std::memcmp(st->shadow.data(), lo.data(), n * sizeof(float)) == 0;
    st->shadow_valid = false;
    if(!current && !gv_shim::ok(ctx, gv_grid_upload(ctx, lo.data(), nullptr), "gv_grid_upload"))
      return;
    if(!gv_shim::ok(ctx, update(ctx), "gv_grid_update"))
      return;
    if(!gv_shim::ok(ctx, gv_grid_download(ctx, lo.data(), with_occupancy ? oc.data() : nullptr), "gv_grid_download"))
      return;
    st->shadow.assign(lo.data(), lo.data() + n);
    st->shadow_valid = true;
  }
}

OccupancyGridMap::OccupancyGridMap(const std::string &base_link, uint8_t grid_x,
                                   uint8_t grid_y, double resolution)
{
  // host-side container exactly as the reference builds it (:8-13); the device twin is created
  // by the first update from this geometry (state_for), so copies and moves of *this are harmless
  grid_map_ = grid_map::GridMap({"log_odds", "occupancy"});
  grid_map_.setFrameId(base_link);
  grid_map_.setGeometry(grid_map::Length(grid_x, grid_y), resolution);
  grid_map_.setPosition(grid_map::Position(grid_x / 3, 0.0));
  grid_map_["log_odds"].setConstant(log_odds_prior_);
  grid_map_["occupancy"].setConstant(init_probability_);
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
  double c[8];
  for(int i = 0; i < 4; ++i)
  {
    c[2 * i + 0] = bbox_corners[i].x;
    c[2 * i + 1] = bbox_corners[i].y;
  }
  // one footprint, no decay: finalize with k_decay = 0 leaves every other cell's log-odds as is
  run_update(grid_map, [&](gv_ctx *ctx) { return gv_grid_finalize(ctx, 0, c, 1); }, false);
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
