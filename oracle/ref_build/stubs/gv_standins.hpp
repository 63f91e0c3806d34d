// gv_standins.hpp — minimal stand-ins for the third-party headers the reference's hot-path
// sources include (Eigen, PCL, grid_map, OpenCV, tf2, geometry_msgs, onnxruntime, spdlog).
//
// TEST INFRASTRUCTURE.  Purpose: compile /root/reference/src/occupancy_grid.cpp and
// /root/reference/src/cloud_detections.cpp UNMODIFIED, where they lie, so that the
// reference's OWN code (predicates, constants, control flow, operation order in
// extractCloudPerBBox / buildKDTree / OccupancyGridMap::updateMap / updateGridCellsFast /
// computeBoundingBox3D / getEstimatedDepth) runs for real and pins oracle/gv_oracle.c.
// None of the real libraries exist in the build container, so the third-party pieces are
// restated here from their published behaviour, only as far as those two files use them:
//   Eigen      Matrix3d * Vector3d accumulates (a0*b0 + a1*b1) + a2*b2 per row in double;
//              MatrixXf is column-major; array() += s, cwiseMax/cwiseMin, block().array() += s
//   grid_map   GridMap geometry / getIndex / at / GridMapIterator (grid_map_core, start index 0)
//   PCL        PointXYZI / PointXYZ / PointCloud / isFinite
// The pieces the exercised functions never reach (RANSAC, radius outlier removal, FLANN,
// cv::PCA) are declared with inert bodies so the translation units link.
#pragma once

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------ Eigen
namespace Eigen {

struct Vector3d {
  double v[3];
  Vector3d() : v{0, 0, 0} {}
  Vector3d(double a, double b, double c) : v{a, b, c} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
inline Vector3d operator*(double s, const Vector3d &a) { return Vector3d(s * a.v[0], s * a.v[1], s * a.v[2]); }

struct Matrix3d {
  double m[3][3];
  Matrix3d() : m{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}} {}
  double &operator()(int r, int c) { return m[r][c]; }
  double operator()(int r, int c) const { return m[r][c]; }
  // fixed-size 3x3 * 3x1: coefficient-wise inner products, accumulated left to right
  Vector3d operator*(const Vector3d &b) const
  {
    Vector3d r;
    for (int i = 0; i < 3; ++i) r.v[i] = (m[i][0] * b.v[0] + m[i][1] * b.v[1]) + m[i][2] * b.v[2];
    return r;
  }
  struct CommaInit {
    Matrix3d *M;
    int k;
    CommaInit &operator,(double x)
    {
      M->m[k / 3][k % 3] = x;
      ++k;
      return *this;
    }
  };
  CommaInit operator<<(double x)
  {
    m[0][0] = x;
    return CommaInit{this, 1};
  }
  Matrix3d inverse() const { return Matrix3d(); }  // not reached by the exercised functions
};

struct Vector4f {
  float v[4] = {0, 0, 0, 0};
  float &operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
};
struct VectorXf {
  std::vector<float> d;
};

template <typename T>
struct Pair2 {
  T a[2];
  Pair2() : a{T(), T()} {}
  template <typename A, typename B>
  Pair2(A x, B y) : a{static_cast<T>(x), static_cast<T>(y)} {}
  T &x() { return a[0]; }
  T &y() { return a[1]; }
  T x() const { return a[0]; }
  T y() const { return a[1]; }
  T &operator()(int i) { return a[i]; }
  T operator()(int i) const { return a[i]; }
  T &operator[](int i) { return a[i]; }
  T operator[](int i) const { return a[i]; }
};
using Array2i = Pair2<int>;
using Array2d = Pair2<double>;
using Vector2d = Pair2<double>;

// Column-major dynamic float matrix with exactly the expression surface occupancy_grid.cpp uses.
struct MatrixXf {
  int r = 0, c = 0;
  std::vector<float> d;
  void resize(int rows, int cols)
  {
    r = rows;
    c = cols;
    d.assign((size_t)rows * cols, 0.0f);
  }
  int rows() const { return r; }
  int cols() const { return c; }
  float *data() { return d.data(); }
  const float *data() const { return d.data(); }
  float &operator()(int i, int j) { return d[(size_t)i + (size_t)j * r]; }
  float operator()(int i, int j) const { return d[(size_t)i + (size_t)j * r]; }
  void setConstant(float v) { std::fill(d.begin(), d.end(), v); }

  struct ArrayRef {  // m.array() and m.block(...).array()
    MatrixXf *M;
    int i0, j0, nr, nc;
    void operator+=(float s)
    {
      for (int j = j0; j < j0 + nc; ++j)
        for (int i = i0; i < i0 + nr; ++i) (*M)(i, j) = (*M)(i, j) + s;
    }
  };
  ArrayRef array() { return ArrayRef{this, 0, 0, r, c}; }
  struct Block {
    MatrixXf *M;
    int i0, j0, nr, nc;
    ArrayRef array() { return ArrayRef{M, i0, j0, nr, nc}; }
  };
  Block block(int i0, int j0, int nr, int nc) { return Block{this, i0, j0, nr, nc}; }

  // cwiseMax(a).cwiseMin(b): scalar_max_op(x,a) = (x < a) ? a : x ; scalar_min_op(x,b) = (b < x) ? b : x
  struct Clamped {
    const MatrixXf *M;
    bool has_lo = false, has_hi = false;
    float lo = 0, hi = 0;
    Clamped cwiseMin(float b) const
    {
      Clamped e = *this;
      e.has_hi = true;
      e.hi = b;
      return e;
    }
    Clamped cwiseMax(float a) const
    {
      Clamped e = *this;
      e.has_lo = true;
      e.lo = a;
      return e;
    }
  };
  Clamped cwiseMax(float a) const
  {
    Clamped e{this};
    e.has_lo = true;
    e.lo = a;
    return e;
  }
  Clamped cwiseMin(float b) const
  {
    Clamped e{this};
    e.has_hi = true;
    e.hi = b;
    return e;
  }
  MatrixXf &operator=(const Clamped &e)
  {
    for (size_t k = 0; k < d.size(); ++k) {
      float x = e.M->d[k];
      if (e.has_lo) x = (x < e.lo) ? e.lo : x;
      if (e.has_hi) x = (e.hi < x) ? e.hi : x;
      d[k] = x;
    }
    return *this;
  }
};

}  // namespace Eigen

// ------------------------------------------------------------------------- geometry_msgs
namespace geometry_msgs {
namespace msg {
struct Point {
  double x = 0, y = 0, z = 0;
};
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
};
struct Pose {
  Point position;
  Quaternion orientation;
};
}  // namespace msg
}  // namespace geometry_msgs

// ------------------------------------------------------------------------------ tf2
namespace tf2 {
struct Quaternion {
  double q[4] = {0, 0, 0, 1};
  void setRPY(double roll, double pitch, double yaw)
  {
    const double hr = roll * 0.5, hp = pitch * 0.5, hy = yaw * 0.5;
    const double cr = std::cos(hr), sr = std::sin(hr), cp = std::cos(hp), sp = std::sin(hp),
                 cy = std::cos(hy), sy = std::sin(hy);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
  }
  double x() const { return q[0]; }
  double y() const { return q[1]; }
  double z() const { return q[2]; }
  double w() const { return q[3]; }
};
// tf2::Transform as pcl_ros consumes it: a 3x3 double basis (rows addressable, elements by
// x()/y()/z()) and a double origin
struct Vector3 {
  double v[3] = {0, 0, 0};
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};
struct Matrix3x3 {
  Vector3 r[3];
  const Vector3 &operator[](int i) const { return r[i]; }
  Vector3 &operator[](int i) { return r[i]; }
};
struct Transform {
  Matrix3x3 basis;
  Vector3 origin;
  const Matrix3x3 &getBasis() const { return basis; }
  const Vector3 &getOrigin() const { return origin; }
};
}  // namespace tf2

// ------------------------------------------------------------------------------ OpenCV
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795
namespace cv {
struct Point2f {
  float x = 0, y = 0;
  Point2f() {}
  Point2f(float a, float b) : x(a), y(b) {}
  float dot(const Point2f &o) const { return x * o.x + y * o.y; }
};
inline Point2f operator-(const Point2f &a, const Point2f &b) { return Point2f(a.x - b.x, a.y - b.y); }
inline Point2f operator+(const Point2f &a, const Point2f &b) { return Point2f(a.x + b.x, a.y + b.y); }
inline Point2f operator*(float s, const Point2f &a) { return Point2f(s * a.x, s * a.y); }
struct Scalar {
  Scalar(double = 0, double = 0, double = 0, double = 0) {}
};
struct Mat {
  int rows = 0, cols = 0;
  std::vector<float> d;
  Mat() {}
  Mat(int r, int c, int) : rows(r), cols(c), d((size_t)r * c, 0.0f) {}
  bool empty() const { return d.empty(); }
  template <typename T>
  T &at(int i, int j = 0) { return d[(size_t)i * (cols ? cols : 1) + j]; }
  template <typename T>
  const T &at(int i, int j = 0) const { return d[(size_t)i * (cols ? cols : 1) + j]; }
};
struct PCA {  // inert: computePCABoundingBox is outside the exercised path ("next" row N2)
  enum { DATA_AS_ROW = 0 };
  Mat mean, eigenvectors;
  PCA(const Mat &, const Mat &, int) : mean(1, 2, CV_32F), eigenvectors(2, 2, CV_32F)
  {
    eigenvectors.at<float>(0, 0) = 1.0f;
    eigenvectors.at<float>(1, 1) = 1.0f;
  }
};
}  // namespace cv

// ------------------------------------------------------------------------ onnxruntime / spdlog
namespace Ort {
struct Session {};
struct Env {};
struct SessionOptions {};
struct Value {};
}  // namespace Ort

// ------------------------------------------------------------------------------ PCL
#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0, y = 0, z = 0, w = 1.0f;
};
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, w = 1.0f;
  float intensity = 0, pad1 = 0, pad2 = 0, pad3 = 0;
};
static_assert(sizeof(PointXYZI) == 32 && sizeof(PointXYZ) == 16, "PCL point layouts");

template <typename P>
inline bool isFinite(const P &p)
{
  return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z);
}

template <typename P>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<P>>;
  using ConstPtr = std::shared_ptr<const PointCloud<P>>;
  std::vector<P> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  void push_back(const P &p)
  {
    points.push_back(p);
    width = (uint32_t)points.size();
    height = 1;
  }
  bool empty() const { return points.empty(); }
  size_t size() const { return points.size(); }
};

struct PointIndices {
  using Ptr = std::shared_ptr<PointIndices>;
  std::vector<int> indices;
};
struct ModelCoefficients {
  using Ptr = std::shared_ptr<ModelCoefficients>;
  std::vector<float> values;
};
enum { SACMODEL_PLANE = 0 };
enum { SAC_RANSAC = 0 };

// inert: ground removal is a "next" row, never reached by the harness
template <typename P>
struct SACSegmentation {
  void setOptimizeCoefficients(bool) {}
  void setModelType(int) {}
  void setMethodType(int) {}
  void setDistanceThreshold(double) {}
  void setInputCloud(const typename PointCloud<P>::Ptr &) {}
  void segment(PointIndices &, ModelCoefficients &) {}
};
template <typename P>
struct ExtractIndices {
  typename PointCloud<P>::Ptr in;
  void setInputCloud(const typename PointCloud<P>::Ptr &c) { in = c; }
  void setIndices(const PointIndices::Ptr &) {}
  void setNegative(bool) {}
  void filter(PointCloud<P> &out) { out = *in; }
};
template <typename P>
struct RadiusOutlierRemoval {
  typename PointCloud<P>::Ptr in;
  void setInputCloud(const typename PointCloud<P>::Ptr &c) { in = c; }
  void setRadiusSearch(double) {}
  void setMinNeighborsInRadius(int) {}
  void filter(PointCloud<P> &out) { out = *in; }
};
template <typename P>
inline unsigned compute3DCentroid(const PointCloud<P> &c, Eigen::Vector4f &o)
{
  for (const auto &p : c.points) {
    o[0] += p.x;
    o[1] += p.y;
    o[2] += p.z;
  }
  const float n = c.points.empty() ? 1.0f : (float)c.points.size();
  o[0] /= n;
  o[1] /= n;
  o[2] /= n;
  return (unsigned)c.points.size();
}
// pcl::KdTreeFLANN stand-in: an EXACT k-nearest-neighbour search (what FLANN's single KD-tree
// returns with PCL's default exact parameters), brute force.  Distance = flann::L2_Simple<float>
// (one `result += diff*diff` per coordinate, float); non-finite points are not indexed (PCL's
// convertCloudToArray drops them); k is clamped to the number of indexed points; results are
// ascending in distance, ties by lower index (upstream leaves ties unspecified).
template <typename P>
struct KdTreeFLANN {
  typename PointCloud<P>::Ptr in;
  void setInputCloud(const typename PointCloud<P>::Ptr &c) { in = c; }
  int nearestKSearch(const P &q, int k, std::vector<int> &idx, std::vector<float> &dist)
  {
    idx.clear();
    dist.clear();
    if (!in || k <= 0) return 0;
    for (size_t i = 0; i < in->points.size(); ++i) {
      const P &p = in->points[i];
      if (!(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
      float d = 0.0f, diff;
      diff = p.x - q.x; d += diff * diff;
      diff = p.y - q.y; d += diff * diff;
      diff = p.z - q.z; d += diff * diff;
      if ((int)idx.size() == k && !(d < dist.back())) continue;
      size_t pos = idx.size();
      if ((int)idx.size() < k) { idx.push_back(0); dist.push_back(0.0f); } else pos = (size_t)k - 1;
      while (pos > 0 && d < dist[pos - 1]) { dist[pos] = dist[pos - 1]; idx[pos] = idx[pos - 1]; --pos; }
      dist[pos] = d;
      idx[pos] = (int)i;
    }
    return (int)idx.size();
  }
};
}  // namespace pcl

// pcl_ros::transformPointCloud(in, out, tf2::Transform) as the node calls it
// (src/grid_vision_node.cpp:304).  RECALLED FROM UPSTREAM (pcl_ros transforms.hpp + pcl
// common/impl/transforms.hpp), unverified offline: the tf2 transform becomes an Eigen::Matrix4f
// (doubles narrowed to float), the cloud is copied (every field), and each point's xyz becomes
// Transformer<float>::se3: c0*x + (c1*y + (c2*z + c3)) per row; with is_dense == false a point
// with a non-finite coordinate is left as it is.
namespace pcl_ros {
template <typename P>
inline void transformPointCloud(const pcl::PointCloud<P> &in, pcl::PointCloud<P> &out, const tf2::Transform &t)
{
  float m[3][4];
  for (int r = 0; r < 3; ++r) {
    m[r][0] = (float)t.getBasis()[r].x();
    m[r][1] = (float)t.getBasis()[r].y();
    m[r][2] = (float)t.getBasis()[r].z();
  }
  m[0][3] = (float)t.getOrigin().x();
  m[1][3] = (float)t.getOrigin().y();
  m[2][3] = (float)t.getOrigin().z();
  out = in;
  for (auto &p : out.points) {
    if (!in.is_dense && !(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
    const float x = p.x, y = p.y, z = p.z;
    p.x = m[0][0] * x + (m[0][1] * y + (m[0][2] * z + m[0][3]));
    p.y = m[1][0] * x + (m[1][1] * y + (m[1][2] * z + m[1][3]));
    p.z = m[2][0] * x + (m[2][1] * y + (m[2][2] * z + m[2][3]));
  }
}
}  // namespace pcl_ros

// ------------------------------------------------------------------------------ grid_map
// grid_map_core GridMap, restated (RECALLED FROM UPSTREAM, unverified offline) for a map that
// is never moved (start index 0): setGeometry, setPosition, operator[], at, getIndex
// (GridMapMath getIndexFromPosition + checkIfPositionWithinMap + checkIfIndexInRange) and
// GridMapIterator.  This is the same statement as oracle/gv_oracle.c gvo_grid_get_index; what
// the _ref build adds is the reference's own code around it.
namespace grid_map {
using Matrix = Eigen::MatrixXf;
using Index = Eigen::Array2i;
using Size = Eigen::Array2i;
using Length = Eigen::Array2d;
using Position = Eigen::Vector2d;

class GridMap {
public:
  GridMap() {}
  GridMap(const std::vector<std::string> &layers)
  {
    for (const auto &l : layers) data_[l] = Matrix();
    layers_ = layers;
  }
  void setFrameId(const std::string &f) { frame_ = f; }
  void setGeometry(const Length &length, double resolution, const Position &position = Position(0.0, 0.0))
  {
    size_(0) = static_cast<int>(std::round(length(0) / resolution));
    size_(1) = static_cast<int>(std::round(length(1) / resolution));
    resolution_ = resolution;
    length_(0) = (double)size_(0) * resolution_;
    length_(1) = (double)size_(1) * resolution_;
    position_ = position;
    for (auto &kv : data_) kv.second.resize(size_(0), size_(1));
  }
  void setPosition(const Position &p) { position_ = p; }
  Matrix &operator[](const std::string &layer) { return data_.at(layer); }
  const Matrix &operator[](const std::string &layer) const { return data_.at(layer); }
  float &at(const std::string &layer, const Index &i) { return data_.at(layer)(i(0), i(1)); }
  float at(const std::string &layer, const Index &i) const { return data_.at(layer)(i(0), i(1)); }
  const Size &getSize() const { return size_; }
  const Length &getLength() const { return length_; }
  const Position &getPosition() const { return position_; }
  double getResolution() const { return resolution_; }

  bool getIndex(const Position &position, Index &index) const
  {
    // getIndexFromPosition
    double iv[2];
    for (int a = 0; a < 2; ++a) {
      const double offset = 0.5 * length_(a);
      iv[a] = ((position(a) - offset) - position_(a)) / resolution_;
    }
    // checkIfPositionWithinMap (evaluated first here so the int cast below is defined; the
    // conjunction is what upstream returns)
    bool within = true;
    for (int a = 0; a < 2; ++a) {
      const double offset = 0.5 * length_(a);
      const double q = -((position(a) - position_(a)) - offset);
      within = within && (q >= 0.0) && (q < length_(a));
    }
    if (!within) return false;
    index(0) = static_cast<int>(-iv[0]);
    index(1) = static_cast<int>(-iv[1]);
    // checkIfIndexInRange
    return index(0) >= 0 && index(1) >= 0 && index(0) < size_(0) && index(1) < size_(1);
  }

private:
  std::map<std::string, Matrix> data_;
  std::vector<std::string> layers_;
  std::string frame_;
  Size size_;
  Length length_;
  Position position_;
  double resolution_ = 0.0;
};

class GridMapIterator {
public:
  explicit GridMapIterator(const GridMap &m) : size_(m.getSize()), lin_(0) {}
  bool isPastEnd() const { return lin_ >= (size_t)size_(0) * (size_t)size_(1); }
  GridMapIterator &operator++()
  {
    ++lin_;
    return *this;
  }
  Index operator*() const { return Index((int)(lin_ % (size_t)size_(0)), (int)(lin_ / (size_t)size_(0))); }

private:
  Size size_;
  size_t lin_;
};
}  // namespace grid_map
