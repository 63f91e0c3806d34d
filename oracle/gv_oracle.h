/*
 * gv_oracle.h — CPU ORACLE for the grid-vision point-cloud -> occupancy-grid hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (grid_vision_b200/) never links, imports or calls anything in oracle/.
 *
 * What it is: a dependency-free plain-C restatement of the reference's CPU arithmetic
 * for the path, function by function (citations are relative to /root/reference):
 *   R1  src/grid_vision_node.cpp:280-307   LiDAR->camera extrinsic transform
 *                                          (pcl_ros::transformPointCloud, float SE(3))
 *   R3  src/cloud_detections.cpp:250-298   extractCloudPerBBox (project + label)
 *   R4  src/cloud_detections.cpp:8-40      buildKDTree projection loop
 *   R6  src/occupancy_grid.cpp:4-14        grid geometry / initial state
 *   R7  src/occupancy_grid.cpp:16-31       updateMap(grid)          decay/clamp/sigmoid
 *   R8  src/occupancy_grid.cpp:65-105      updateMap(grid, poses)   + :140-183 block add
 *   R9  src/occupancy_grid.cpp:33-63       updateMap(grid, points, boxes) + :107-138,185-196
 * plus the north-star extensions that have NO reference code:
 *   X1  base-frame transform + cell binning (grid_map getIndex convention)
 *   X2  per-beam Bresenham raycast (grid_map LineIterator convention)
 *   X3  integer counts -> float log-odds in one canonical order
 *
 * PINNING STATUS.  The reference ships no tests, golden vectors or fixtures and none
 * of its dependencies (ROS 2, PCL, Eigen, grid_map, OpenCV) exist in the build
 * container, so the third-party arithmetic underneath R1 (PCL Transformer<float>::se3,
 * SSE2 add nesting), R6-R9 (grid_map getIndexFromPosition / checkIfPositionWithinMap,
 * column-major MatrixXf) and X2 (grid_map LineIterator) is restated FROM THE PUBLISHED
 * UPSTREAM ALGORITHMS AND IS UNVERIFIED OFFLINE: for those pieces parity is
 * "unpinned by upstream".  What *is* pinned:
 *   - the reference's own code (R3/R4/R5 and R6-R9 control flow, predicates, constants)
 *     is compiled unmodified from /root/reference against minimal stand-in headers by
 *     oracle/ref_build/ (output oracle/_ref/) and this restatement is checked against
 *     it bit-for-bit in tests/test_oracle_vs_ref.py;
 *   - hand-derived known-answer vectors (tests/golden/, SURVEY.md §8.c list).
 * The three recalled upstream pieces live in exactly one place each (gvo_se3,
 * gvo_grid_get_index, gvo_bresenham_*) so they can be corrected in one edit.
 *
 * Float semantics: build with -O2 -ffp-contract=off and NO -march / -ffast-math: the
 * reference's CMakeLists.txt:1-8 sets no arch or optimisation flags, i.e. baseline
 * x86-64 SSE2, separate multiply and add, no FMA.
 */
#ifndef GV_ORACLE_H_
#define GV_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pcl::PointXYZI layout stand-in: 32 bytes, 16-aligned (x,y,z,pad | intensity,pad*3). */
typedef struct {
  float x, y, z, w;
  float intensity, pad1, pad2, pad3;
} gvo_point_xyzi;

/* include/grid_vision/object_detection.hpp:27-32  struct BoundingBox  (sizeof 40). */
typedef struct {
  double x_min, y_min, x_max, y_max;
  float confidence;
  int32_t label; /* enum class ObjectClass, object_detection.hpp:12-25 */
} gvo_box;

/* ObjectClass values used by R9 (object_detection.hpp:12-25). */
enum {
  GVO_BIKE = 0, GVO_MOTORBIKE = 1, GVO_PERSON = 2, GVO_VEHICLE = 9, GVO_UNKNOWN = 10
};

/* ---------------------------------------------------------------- R1 --- */
/* T is a row-major 4x4 float matrix (last row ignored).  out = c0*x + (c1*y + (c2*z + c3))
 * per PCL's SSE2 Transformer<float>::se3.  is_dense==0: non-finite points are copied
 * through untouched (pcl::transformPointCloud, !is_dense branch). */
void gvo_transform_points(const float T[16], const float *x, const float *y, const float *z,
                          size_t n, int is_dense, float *ox, float *oy, float *oz);

/* ---------------------------------------------------------------- R3 --- */
/* Points are in the CAMERA frame.  K row-major 3x3 double.  Outputs (any may be NULL):
 *  label[i] = index of first box (list order) containing (u,v), else -1
 *  pix[i]   = (int)v * W + (int)u for points that pass the image test, else -1
 *  u[i],v[i]= the float pixel coordinates (NaN for points rejected before projection) */
void gvo_project_label(const double K[9], int image_width, int image_height, const float *x,
                       const float *y, const float *z, size_t n, const gvo_box *boxes,
                       int nboxes, int16_t *label, int32_t *pix, float *u, float *v);

/* Reference-shaped R3: AoS 32-byte points in, one growable AoS cloud per box out
 * (std::vector::push_back growth policy).  counts[nboxes] receives sizes; clouds[i] is
 * malloc'd (caller frees each, may be NULL when empty). */
void gvo_extract_cloud_per_bbox_aos(const gvo_point_xyzi *pts, size_t n, const double K[9],
                                    const gvo_box *boxes, int nboxes, int image_width,
                                    int image_height, gvo_point_xyzi **clouds, size_t *counts);

/* ---------------------------------------------------------------- R4 --- */
/* buildKDTree projection loop: skip z <= 0, emit (u, v, z) compacted in input order.
 * uvz has room for 3*n floats; returns the number of points emitted. */
size_t gvo_project_kdtree(const double K[9], const float *x, const float *y, const float *z,
                          size_t n, float *uvz);

/* ------------------------------------------------------------ R6 .. R9 --- */
typedef struct {
  int32_t nx, ny;        /* size(0) (x cells, matrix rows), size(1) (y cells, matrix cols) */
  double res;            /* resolution                                                 */
  double len_x, len_y;   /* length = size * resolution                                 */
  double pos_x, pos_y;   /* map centre                                                 */
  float *log_odds;       /* nx*ny, column-major: lin = ix + iy*nx                      */
  float *occupancy;      /* nx*ny                                                      */
  int32_t *hit, *miss;   /* nx*ny each (X1/X2 integer planes)                          */
} gvo_grid;

/* grid_map setGeometry: size = (int)round(length/res); length = size*res.  Layers are
 * malloc'd and initialised to log_odds=0.0f, occupancy=0.5f, hit=miss=0. */
int gvo_grid_init(gvo_grid *g, double length_x, double length_y, double res, double pos_x,
                  double pos_y);
/* Reference constructor geometry: lengths are uint8 metres, centre = (grid_x / 3, 0)
 * with INTEGER division (src/occupancy_grid.cpp:10-11). */
int gvo_grid_init_reference(gvo_grid *g, uint8_t grid_x, uint8_t grid_y, double res);
void gvo_grid_free(gvo_grid *g);

/* grid_map::GridMap::getIndex restatement.  Returns 1 and writes (ix,iy) when inside. */
int gvo_grid_get_index(const gvo_grid *g, double px, double py, int32_t *ix, int32_t *iy);

void gvo_update_map(gvo_grid *g);                                        /* R7 */
/* R8: poses as n x (pos_x, pos_y, length, width) doubles. */
void gvo_update_map_poses(gvo_grid *g, const double *xylw, int n);
/* R9: points as n x (x, y) doubles + class label per point. */
void gvo_update_map_points(gvo_grid *g, const double *xy, const int32_t *labels, int n);
/* :140-183 on explicit corners, n x 4 x (x,y) doubles, reference corner order. */
void gvo_update_grid_cells_fast(gvo_grid *g, const double corners[8]);
float gvo_estimated_depth(int32_t label);                                /* :185-196 */

/* ------------------------------------------------------------ X1 .. X3 --- */
enum { GVO_OCC_ALL = 0, GVO_OCC_LABELLED = 1 };
enum { GVO_F_VALID = 1, GVO_F_HIT = 2, GVO_F_CLIPPED = 4, GVO_F_RANGECAP = 8 };

typedef struct {
  int32_t occ_mode;   /* GVO_OCC_*                                         */
  int32_t use_z_gate; /* apply z_min <= z_base <= z_max to hits            */
  float z_min, z_max;
  double r_max;       /* planar range cap in metres; <= 0 disables         */
} gvo_accum_params;

/* Per-beam bin + raycast (no de-duplication: this is the brute-force ground truth).
 * T_base_lidar row-major 4x4 float; sensor origin = its translation column.
 * labels may be NULL unless occ_mode==GVO_OCC_LABELLED.  cell_out / flags_out nullable.
 * Returns the number of traversed-cell updates performed (misses + hits), or -1 when the
 * sensor origin is off-map (every beam dropped). */
int64_t gvo_accumulate(gvo_grid *g, const float T_base_lidar[16], const float *x,
                       const float *y, const float *z, size_t n, const int16_t *labels,
                       const gvo_accum_params *prm, int32_t *cell_out, uint8_t *flags_out);

/* grid_map LineIterator restatement: writes up to cap cells (ix,iy pairs) of the line
 * from (sx,sy) to (ex,ey); returns nCells = max(|dx|,|dy|)+1. */
int32_t gvo_bresenham_cells(int32_t sx, int32_t sy, int32_t ex, int32_t ey, int32_t *cells_xy,
                            int32_t cap);

/* X3 (+R8 footprints): l += k_decay*(-0.2f); l += miss*(-0.4f); l += hit*1.2f;
 * each valid footprint covering the cell, in list order, l += 0.85f; clamp; sigmoid;
 * hit = miss = 0.  corners: nfoot x 4 x (x,y) doubles (may be NULL when nfoot==0). */
void gvo_finalize(gvo_grid *g, int32_t k_decay, const double *corners, int nfoot);

/* N1 (next row): cloud_detections::segmentGroundPlane (src/cloud_detections.cpp:105-138):
 * pcl::SACSegmentation(SACMODEL_PLANE, SAC_RANSAC, distance threshold 0.04, optimise
 * coefficients) then ExtractIndices(negative) removes the plane's inliers, order kept; no model
 * -> EMPTY cloud (:122-126).  PCL draws its samples from a boost::mt19937 stream that cannot be
 * reproduced offline, so parity with PCL is statistical by nature; THIS restatement fixes a
 * deterministic hypothesis set instead, which the GPU shares bit for bit:
 *   - hypothesis h in [0, n_hyp): three point indices mix32(seed, 3h+j) % n; invalid if a sample
 *     is non-finite or the triangle is degenerate; plane = normalised (p1-p0) x (p2-p0), float
 *   - score = #{ i : |a x + b y + c z + d| < threshold }   (float, left-to-right, strict <)
 *   - best = highest score, lowest h on ties; no valid hypothesis (score < 3) -> no model;
 *     exactly 3 inliers -> the hypothesis itself is the plane (PCL refines only beyond the sample size)
 *   - refinement (optimizeModelCoefficients): centroid + covariance of the best plane's inliers
 *     in double, normal = eigenvector of the smallest eigenvalue (cyclic Jacobi), d = -n.c
 *   - removed = points within threshold of the refined plane (selectWithinDistance)
 * keep[i] = 1 for points that stay.  Returns the number kept, or -1 when no model was found
 * (the reference then returns an empty cloud). */
int64_t gvo_segment_ground(const float *x, const float *y, const float *z, size_t n, float threshold,
                           uint32_t seed, int32_t n_hyp, uint8_t *keep, float plane[4],
                           int32_t *best_h, int32_t *best_score);

/* N2 (next row): cloud_detections::bboxPoseEstimation + computePCABoundingBox
 * (src/cloud_detections.cpp:140-247) for ONE per-box cloud in the camera frame.
 *   1. pcl::RadiusOutlierRemoval(r = 0.4, min neighbours 10): keep point i iff the number of
 *      points j (itself included) with |p_j - p_i|^2 < (float)(0.4*0.4) exceeds 10; order kept.
 *      (PCL / FLANN conventions RECALLED FROM UPSTREAM: k counts the query point, k <= min_pts
 *      removes, FLANN's radius test is strict on the squared float distance.)
 *   2. pcl::compute3DCentroid -> only centroid.y is used (:209)
 *   3. cv::PCA on the N x 2 float rows (z, x): mean, 2x2 covariance, eigenvectors by descending
 *      eigenvalue (closed form here; OpenCV's Jacobi agrees to rounding, eigenvector signs are
 *      arbitrary in both)
 *   4. extents of (pt - mean) along the two axes -> length, width (:224-237)
 *   5. pose.position = (mean_x, centroid_y, mean_z); angle = atan2(major.y, major.x) in DEGREES,
 *      fed to setRPY(0, -angle, 0) as if it were radians (:246,:254-255: a reference quirk, kept).
 * Returns the number of points kept; 0 means the box is skipped (:201-202 data.empty()). */
typedef struct {
  int32_t kept;
  float centroid_y;            /* pose.position.y            */
  float mean_z, mean_x;        /* PCA mean (center.x, center.y): pose.position.z / .x */
  float major_z, major_x;      /* first eigenvector (row 0)  */
  float minor_z, minor_x;      /* second eigenvector (row 1) */
  float length, width;
  float angle_deg;
  double qx, qy, qz, qw;       /* tf2::Quaternion::setRPY(0, -angle_deg, 0) */
} gvo_lshape;

int32_t gvo_radius_outlier_keep(const float *x, const float *y, const float *z, size_t n,
                                double radius, int32_t min_neighbors, uint8_t *keep);
/* N4 (src/cloud_detections.cpp:43-103) */
void gvo_box_depths(const float *uvz, size_t m, const void *boxes40, int nb, int k, float *depths);
void gvo_pixel_to_3d(const double Kinv[9], float px, float py, float depth, double out[3]);
void gvo_bbox_pose(const float *x, const float *y, const float *z, size_t n, gvo_lshape *out);

/* N3 (next row): grid_map_ros toOccupancyGrid(layer "occupancy", 0, 1) cell conversion
 * into nav_msgs/OccupancyGrid data order. */
void gvo_to_occupancy_grid(const gvo_grid *g, int8_t *data);

#ifdef __cplusplus
}
#endif
#endif /* GV_ORACLE_H_ */
