/*
 * gridvision_b200.h — C ABI of libgridvision_b200.so
 *
 * B200 (sm_100a) implementation of grid-vision's point-cloud -> occupancy-grid hot path.
 * Plain C: pointers, sizes and PODs only (no PCL / Eigen / ROS / torch types), so the
 * reference node (C++), a ctypes binding or any other FFI can bind it directly.
 * Citations "ref:" are relative to the upstream repository rohankhaire-work/grid-vision.
 *
 * Conventions
 *  - every entry point returns a gv_status (0 = GV_OK) and never throws; the message of
 *    the last failure is kept per context (gv_last_error).  This mirrors the reference,
 *    where no exception crosses these calls (ref: src/cloud_detections.cpp:308-309,
 *    src/occupancy_grid.cpp:171-172: failures are empty results / silent skips).
 *  - a gv_ctx is NOT re-entrant (the reference's caller is the single-threaded
 *    rclcpp::spin executor, ref: src/grid_vision_node.cpp:533-540); use one per thread/GPU.
 *  - functions without a suffix take HOST pointers, copy in/out themselves and are
 *    synchronous on return (drop-in for the reference's blocking calls);
 *    functions ending in _dev take DEVICE pointers for the bulk arrays (points, labels and
 *    gv_box records; per-frame offset tables stay host arrays) and are asynchronous on the
 *    context stream (gv_stream / gv_set_stream).  A fresh context owns a private NON-BLOCKING
 *    stream, which is not ordered against any other stream (not even the legacy default
 *    stream): a caller whose device buffers are produced or consumed on another stream must
 *    either hand that stream over with gv_set_stream, or order the two with events /
 *    gv_synchronize.  The Python binding binds torch's current stream automatically.
 *  - point clouds are SoA float planes (x[], y[], z[]); the reference's
 *    pcl::PointCloud<pcl::PointXYZI> 32-byte AoS records are accepted by the *_aos32 entry
 *    points and de-interleaved on the device.
 *  - matrices are ROW-MAJOR: K is 3x3 double, extrinsics are 4x4 float (last row ignored).
 *  - grid layers are column-major float planes, lin = ix + iy*nx, identical to
 *    grid_map's Eigen::MatrixXf storage (rows = x cells), so a shim can memcpy them.
 *  - there is no CPU fallback: without a usable sm_100 device gv_create fails.
 */
#ifndef GRIDVISION_B200_H_
#define GRIDVISION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GV_API __attribute__((visibility("default")))
#else
#define GV_API
#endif

#define GV_VERSION 100 /* 0.1.0 */

typedef enum {
  GV_OK = 0,
  GV_ERR_INVALID = 1,  /* bad argument                                   */
  GV_ERR_CUDA = 2,     /* CUDA runtime failure (see gv_last_error)       */
  GV_ERR_STATE = 3,    /* call order: grid/camera/base transform not set */
  GV_ERR_NCCL = 4,
  GV_ERR_OVERFLOW = 5, /* more than 2^31-1 beams since the last finalize */
  GV_ERR_NO_DEVICE = 6
} gv_status;

typedef struct gv_ctx gv_ctx;

/* ref: include/grid_vision/object_detection.hpp:27-32  struct BoundingBox (sizeof 40).
 * Binary-compatible: a std::vector<BoundingBox>::data() can be passed as const gv_box*. */
typedef struct {
  double x_min, y_min, x_max, y_max;
  float confidence;
  int32_t label; /* enum class ObjectClass, ref: object_detection.hpp:12-25 */
} gv_box;

/* ref: pcl::PointXYZI as used for GridVision::cloud_ (grid_vision_node.hpp:61): 32 bytes. */
typedef struct {
  float x, y, z, w;
  float intensity, pad1, pad2, pad3;
} gv_point_xyzi;

typedef struct {
  int32_t nx, ny;      /* cells along x (matrix rows) and y (matrix cols) */
  double resolution;
  double length_x, length_y;
  double pos_x, pos_y; /* map centre */
} gv_grid_desc;

enum { GV_OCC_ALL = 0, GV_OCC_LABELLED = 1 };
/* per-beam flags written by gv_grid_accumulate (parity output) */
enum { GV_F_VALID = 1, GV_F_HIT = 2, GV_F_CLIPPED = 4, GV_F_RANGECAP = 8 };

typedef struct {
  int32_t occ_mode;   /* GV_OCC_ALL: every in-map return is a hit; GV_OCC_LABELLED: only label >= 0 */
  int32_t use_z_gate; /* hits additionally need z_min <= z_base <= z_max              */
  float z_min, z_max;
  double r_max;       /* planar range cap in metres (beam truncated, no hit); <= 0 disables */
} gv_accum_params;

typedef struct {
  uint64_t beams;             /* valid beams binned since the last finalize               */
  uint64_t cells_logical;     /* traversed-cell updates those beams stand for (sum of n)  */
  uint64_t cells_physical;    /* atomic cell updates actually issued after de-duplication */
  uint64_t distinct_ends;     /* (start,end) pairs walked                                 */
  uint64_t kernel_launches;   /* kernels launched by this context so far                  */
  uint64_t merges;            /* gv_grid_finalize / _multi calls so far                    */
  double merge_ms_last;       /* device time of the last one (raycast [+ exchange] + finalise) */
  uint64_t deferred_points;   /* points the certified kernels handed to the exact FP64 pass (since reset) */
} gv_stats;

/* ------------------------------------------------------------------ lifecycle --- */
GV_API int gv_version(void);
GV_API const char *gv_status_string(int status);
GV_API int gv_create(gv_ctx **out, int device);
GV_API void gv_destroy(gv_ctx *ctx);
GV_API const char *gv_last_error(const gv_ctx *ctx);
GV_API int gv_synchronize(gv_ctx *ctx);
/* gv_grid_finalize and gv_grid_finalize_multi return as soon as their work is queued: the merge
 * of the batch binned so far (raycast, [exchange between ranks,] finalise) runs on an internal
 * stream, ordered after the binning, while the caller's stream is free to bin the NEXT batch
 * (gv_process_batch*, gv_grid_accumulate*) into a second end-cell plane.  Every entry point
 * that reads or writes grid state waits for it by itself; gv_join makes the caller's stream
 * wait explicitly (e.g. before recording a timing event).  $GV_OVERLAP=0 disables the overlap. */
GV_API int gv_join(gv_ctx *ctx);
GV_API void *gv_stream(gv_ctx *ctx);               /* cudaStream_t */
/* launch on the caller's stream from now on; the handle is used as given (NULL = the legacy
 * default stream).  A fresh context owns a private non-blocking stream. */
GV_API int gv_set_stream(gv_ctx *ctx, void *stream);
GV_API int gv_get_stats(gv_ctx *ctx, gv_stats *out);

/* CUDA graphs for the node's per-scan call pattern (ref: src/grid_vision_node.cpp:139-237 runs the
 * same fusion + grid update for every scan at 20 Hz): the device-pointer calls made between
 * gv_graph_begin and gv_graph_end (gv_process_batch_dev, gv_grid_accumulate_dev, gv_grid_finalize,
 * gv_grid_update*, ...) are captured from the context stream instead of executed, and gv_graph_launch
 * replays them with ONE launch.  The captured calls read and write the same device buffers at every
 * replay (refill them in place), with the arguments they were captured with.  Run the sequence
 * once, un-captured, first: a captured call may not allocate, synchronise or copy from pageable
 * memory (GV_ERR_STATE / GV_ERR_CUDA otherwise).  Single-GPU contexts on a non-default stream. */
GV_API int gv_graph_begin(gv_ctx *ctx);
GV_API int gv_graph_end(gv_ctx *ctx, int32_t *graph_id);
GV_API int gv_graph_launch(gv_ctx *ctx, int32_t graph_id);
GV_API int gv_graph_destroy(gv_ctx *ctx, int32_t graph_id);

/* Host-only probe (needs no device): the thresholds of the certified image test of the fused point
 * kernel for an image axis of `size` pixels with principal point c (ref intrinsics:
 * src/object_detection.cpp:241-247): out[6] = {size/2, E, a_in, a_out, e6, e0} with
 * |RN(q - size/2)| < a_in => certainly 0 <= u < size, > a_out => certainly outside, E the interval of
 * the box tests.  For the CPU tests of the error analysis. */
GV_API int gv_debug_pair_thresholds(double size, float c, float *out);

/* ------------------------------------------------------------ fusion (R1-R5) --- */
/* Camera rig.  K: ncam x 9 (ref: src/object_detection.cpp:241-247 setIntrinsicMatrix);
 * T_cam_lidar: ncam x 16, the float matrix pcl_ros::transformPointCloud applies
 * (ref: src/grid_vision_node.cpp:300-304), or NULL when clouds arrive already in the
 * camera frame (then gv_fuse is exactly extractCloudPerBBox); wh: ncam x (width,height)
 * (ref: src/cloud_detections.cpp:254, note the (rows, cols) swap at :303,:313-314). */
GV_API int gv_set_cameras(gv_ctx *ctx, int ncam, const double *K, const float *T_cam_lidar,
                          const int32_t *wh);

/* Replaces the per-point loop of cloud_detections::extractCloudPerBBox
 * (ref: src/cloud_detections.cpp:250-298; decl include/grid_vision/cloud_detections.hpp:46-48)
 * fused with GridVision::transformLidarToCamera (ref: src/grid_vision_node.cpp:280-307).
 * boxes: all cameras' lists concatenated; box_cam_offsets[ncam+1] delimits them (NULL for
 * one camera = all nboxes).  Outputs are cam-major planes of n entries each:
 *   labels_out[c*n+i] first box (index local to camera c's list) containing point i, else -1
 *   pix_out   [c*n+i] (int)v*W+(int)u or -1            (nullable)
 *   uv_out    [(2c)*n+i] = u, [(2c+1)*n+i] = v, NaN when rejected before projection (nullable) */
GV_API int gv_fuse(gv_ctx *ctx, const float *x, const float *y, const float *z, size_t n,
                   int is_dense, const gv_box *boxes, int nboxes, const int32_t *box_cam_offsets,
                   int16_t *labels_out, int32_t *pix_out, float *uv_out);
GV_API int gv_fuse_aos32(gv_ctx *ctx, const gv_point_xyzi *pts, size_t n, int is_dense,
                         const gv_box *boxes, int nboxes, const int32_t *box_cam_offsets,
                         int16_t *labels_out, int32_t *pix_out, float *uv_out);
GV_API int gv_fuse_dev(gv_ctx *ctx, const float *d_x, const float *d_y, const float *d_z, size_t n,
                       int is_dense, const gv_box *d_boxes, int nboxes,
                       const int32_t *box_cam_offsets, int16_t *d_labels_out, int32_t *d_pix_out,
                       float *d_uv_out);

/* Replaces GridVision::transformLidarToCamera alone (ref: src/grid_vision_node.cpp:280-307):
 * out = T_cam_lidar[cam] * p with PCL's float op order. */
GV_API int gv_transform_points(gv_ctx *ctx, int cam, const float *x, const float *y, const float *z,
                               size_t n, int is_dense, float *ox, float *oy, float *oz);

/* Replaces the projection loop of cloud_detections::buildKDTree
 * (ref: src/cloud_detections.cpp:13-33; decl cloud_detections.hpp:29-30): camera-frame points
 * in (extrinsic of `cam` applied first when set), (u, v, depth) triples out, compacted in
 * input order; *m_out = number emitted (predicate z <= 0 only, no bounds test). */
GV_API int gv_project_kdtree(gv_ctx *ctx, int cam, const float *x, const float *y, const float *z,
                             size_t n, float *uvz_out, size_t *m_out);

/* Stable per-box partition of point indices from labels (the push_back order of
 * ref: src/cloud_detections.cpp:280-297): offsets_out[nboxes+1], indices_out[n] (only the
 * first offsets_out[nboxes] entries are meaningful). Host pointers. */
GV_API int gv_partition_by_label(gv_ctx *ctx, const int16_t *labels, size_t n, int nboxes,
                                 uint32_t *indices_out, uint64_t *offsets_out);

/* "Next" row N1: the step that precedes extractCloudPerBBox inside computeBBoxPose —
 * cloud_detections::segmentGroundPlane (ref: src/cloud_detections.cpp:105-138; decl
 * cloud_detections.hpp:40-41): RANSAC plane (distance threshold 0.04 m, coefficients refined by
 * least squares) and removal of its inliers, order preserved.  PCL's random sample stream cannot
 * be reproduced, so the hypotheses are a deterministic counter-hashed set (seed, n_hyp <= 1024)
 * scored in one pass; parity is against the repository's oracle, which fixes the same set.
 * Outputs: compacted cloud (ox,oy,oz need room for n floats), *m_out points kept, plane_out
 * (a,b,c,d).  *found_out = 0 when no model exists: the reference then returns an EMPTY cloud
 * (:122-126) and computeBBoxPose returns {} (:308-309); *m_out is 0 in that case. */
GV_API int gv_segment_ground(gv_ctx *ctx, const float *x, const float *y, const float *z, size_t n,
                             float threshold, uint32_t seed, int n_hyp, float *ox, float *oy,
                             float *oz, size_t *m_out, float *plane_out, int *found_out);

/* "Next" row N2: the step that follows extractCloudPerBBox inside computeBBoxPose —
 * cloud_detections::bboxPoseEstimation + computePCABoundingBox
 * (ref: src/cloud_detections.cpp:140-247; decls cloud_detections.hpp:43-44,54).
 * Per box: pcl::RadiusOutlierRemoval(0.4 m, 10 neighbours), centroid, cv::PCA on (z, x), extents
 * along the principal axes.  Input: the camera-frame cloud + the labels gv_fuse produced.
 * out[b].kept == 0 means the reference would emit no pose for box b (:201-202).  The eigenvector
 * sign is normalised (major_z >= 0); OpenCV's is arbitrary. */
typedef struct {
  int32_t kept;                /* points that survived the radius-outlier filter           */
  float centroid_y;            /* pose.position.y (:209)                                   */
  float mean_z, mean_x;        /* PCA mean: pose.position.z, pose.position.x (:235-237)     */
  float major_z, major_x, minor_z, minor_x; /* PCA eigenvectors, rows (z, x)               */
  float length, width;         /* extents along major / minor (:224-225)                   */
  float angle_deg;             /* atan2(major.y, major.x) * 180 / pi (:232)                */
  double qx, qy, qz, qw;       /* tf2::Quaternion::setRPY(0, -angle_deg, 0) (:246-247)     */
} gv_lshape;
GV_API int gv_bbox_pose(gv_ctx *ctx, const float *x, const float *y, const float *z, size_t n,
                        const int16_t *labels, int nboxes, gv_lshape *out);

/* "Next" row N4: depth of the static detections — cloud_detections::computeDepthForBoundingBoxes
 * (ref: src/cloud_detections.cpp:43-87; decl cloud_detections.hpp:32-35; call site
 * src/grid_vision_node.cpp:179 with k = k_near, 4 in config/grid_vision_cfg.yaml:20).
 * uvz: the m (u, v, depth) triples of gv_project_kdtree.  Per box: the k nearest triples of
 * (centre_x, centre_y, 0) in the tree's 3-D float metric (depth is a coordinate), the median
 * (element size/2 of the ascending depths), -1 when none.  Exact search; ties in distance go to
 * the lower index (FLANN leaves them unspecified).  k <= 64. */
GV_API int gv_box_depths(gv_ctx *ctx, const float *uvz, size_t m, const gv_box *boxes, int nboxes,
                         int k, float *depths_out);
/* cloud_detections::pixelTo3D (ref: src/cloud_detections.cpp:89-103) for every box centre, as
 * GridVision::convertPixelsTo3D loops it (ref: src/grid_vision_node.cpp:309-335) before the TF
 * hop: xyz_out[3i..] = depths[i] * (K_inv * (centre_x, centre_y, 1)), camera frame, doubles. */
GV_API int gv_pixels_to_3d(gv_ctx *ctx, const gv_box *boxes, const float *depths, int nboxes,
                           const double *K_inv, double *xyz_out);

/* --------------------------------------------------------- grid state (R6) --- */
/* ref: OccupancyGridMap::OccupancyGridMap, src/occupancy_grid.cpp:4-14
 * (decl include/grid_vision/occupancy_grid.hpp:16): size = round(length/res), centre
 * (grid_x/3, 0) with uint8 integer division, log_odds = 0, occupancy = 0.5. */
GV_API int gv_grid_init_reference(gv_ctx *ctx, uint8_t grid_x, uint8_t grid_y, double resolution);
/* general geometry (the uint8-metre constructor cannot express BASELINE configs 3 and 5) */
GV_API int gv_grid_init(gv_ctx *ctx, double length_x, double length_y, double resolution,
                        double pos_x, double pos_y);
GV_API int gv_grid_get_desc(gv_ctx *ctx, gv_grid_desc *out);
GV_API int gv_grid_reset(gv_ctx *ctx); /* back to log_odds 0 / occupancy 0.5 / counts 0 */
/* the public member grid_map_ (ref: occupancy_grid.hpp:22) lives on the device; these move
 * its two layers to/from grid_map["log_odds"] / ["occupancy"] MatrixXf storage. NULL skips. */
GV_API int gv_grid_upload(gv_ctx *ctx, const float *log_odds, const float *occupancy);
GV_API int gv_grid_download(gv_ctx *ctx, float *log_odds, float *occupancy);
GV_API int gv_grid_counts_download(gv_ctx *ctx, int32_t *hit, int32_t *miss);
GV_API int gv_grid_layers_dev(gv_ctx *ctx, float **d_log_odds, float **d_occupancy,
                              int32_t **d_hit, int32_t **d_miss);
/* grid_map::GridMap::getIndex for host positions (ref call site: src/occupancy_grid.cpp:152):
 * ixy_out[2i..2i+1] = cell or (-1,-1) when outside the map. */
GV_API int gv_grid_get_index(gv_ctx *ctx, const double *xy, int n, int32_t *ixy_out);

/* ------------------------------------------------ per-frame updates (R7-R9) --- */
/* ref: void OccupancyGridMap::updateMap(grid_map::GridMap&), src/occupancy_grid.cpp:16-31 */
GV_API int gv_grid_update(gv_ctx *ctx);
/* ref: updateMap(grid_map::GridMap&, const std::vector<LShapePose>&), :65-105 + :140-183.
 * xylw: n x (pose.position.x, pose.position.y, length, width). */
GV_API int gv_grid_update_poses(gv_ctx *ctx, const double *xylw, int n);
/* ref: updateMap(grid_map::GridMap&, const std::vector<geometry_msgs::msg::Point>&,
 * const std::vector<BoundingBox>&), :33-63 + :107-138 + :185-196.
 * xy: n x (x, y) base-frame points; labels: n ObjectClass values. */
GV_API int gv_grid_update_points(gv_ctx *ctx, const double *xy, const int32_t *labels, int n);
/* updateGridCellsFast on explicit corners (ref: :140-183), n x 4 x (x, y). */
GV_API int gv_grid_update_corners(gv_ctx *ctx, const double *corners, int n);

/* ----------------------------- binning + raycast + finalise (north-star X1-X3) --- */
/* T_base_lidar: 4x4 float; the sensor origin is its translation column. */
GV_API int gv_set_base_transform(gv_ctx *ctx, const float *T_base_lidar);
/* X1+X2 for one cloud in the LiDAR frame.  labels (camera-0 plane of gv_fuse) may be NULL
 * unless occ_mode == GV_OCC_LABELLED.  cell_out[i] = end cell (lin) or -1 for dropped
 * beams, flags_out[i] = GV_F_* (both nullable, parity outputs).  Beams are binned into a
 * per-end-cell (total, hit) plane; identical (start,end) pairs are walked once by
 * gv_grid_raycast_flush (called implicitly by finalize / a base-transform change). */
GV_API int gv_grid_accumulate(gv_ctx *ctx, const float *x, const float *y, const float *z,
                              size_t n, const int16_t *labels, const gv_accum_params *prm,
                              int32_t *cell_out, uint8_t *flags_out);
GV_API int gv_grid_accumulate_dev(gv_ctx *ctx, const float *d_x, const float *d_y,
                                  const float *d_z, size_t n, const int16_t *d_labels,
                                  const gv_accum_params *prm, int32_t *d_cell_out,
                                  uint8_t *d_flags_out);
GV_API int gv_grid_raycast_flush(gv_ctx *ctx);
/* X3 (+ R8 footprints): l += k_decay*(-0.2f); l += miss*(-0.4f); l += hit*1.2f; +0.85f per
 * covering valid footprint; clamp [-2, 3.6]; sigmoid; counts cleared.
 * corners: nfoot x 4 x (x,y) or NULL. */
GV_API int gv_grid_finalize(gv_ctx *ctx, int32_t k_decay, const double *corners, int nfoot);

/* The whole hot path for a batch of scans that share one sensor pose: per point
 * transform -> project -> label (camera 0, per-frame box lists) -> base transform -> cell
 * -> bin, in ONE kernel that reads each point once.  frame_offsets[nframes+1] delimit the
 * frames in the point planes, box_frame_offsets[nframes+1] delimit each frame's boxes.
 * labels_out (n int16) is nullable.  Follow with gv_grid_finalize.
 * Performance notes (results are identical either way):
 *  - one sensor pose per raycast: all beams binned between two sweeps share the start cell, so a
 *    moving ego needs gv_set_base_transform (which sweeps what was binned) per pose: batch scans
 *    that share a pose, a pose change costs one sweep;
 *  - the fastest kernel (two points per thread) needs every frame offset and size to be even, the
 *    planes 8-byte and labels_out 4-byte aligned, one camera with an extrinsic, canonical K and at
 *    most 64 boxes per frame; other layouts take the one-point-per-thread kernel, other
 *    configurations the generic one;
 *  - the _dev entry point runs asynchronously on the context stream (a private non-blocking stream
 *    unless gv_set_stream was called): order it against the producers of its inputs and the
 *    consumers of its outputs with that stream. */
GV_API int gv_process_batch(gv_ctx *ctx, const float *x, const float *y, const float *z,
                            const uint64_t *frame_offsets, int nframes, const gv_box *boxes,
                            const int32_t *box_frame_offsets, const gv_accum_params *prm,
                            int16_t *labels_out);
GV_API int gv_process_batch_dev(gv_ctx *ctx, const float *d_x, const float *d_y, const float *d_z,
                                const uint64_t *frame_offsets, int nframes, const gv_box *d_boxes,
                                const int32_t *box_frame_offsets, const gv_accum_params *prm,
                                int16_t *d_labels_out);

/* ---------------------------------------------------------------- consumers --- */
/* nav_msgs/OccupancyGrid cells of grid_map::GridMapRosConverter::toOccupancyGrid(grid,
 * "occupancy", 0.0, 1.0, msg) (ref: src/grid_vision_node.cpp:265-278): nx*ny int8. */
GV_API int gv_grid_to_occupancy(gv_ctx *ctx, int8_t *data_out);

/* ---------------------------------------------------------------- multi-GPU --- */
/* One context per rank/GPU, frames sharded across ranks.  id128 comes from rank 0
 * (gv_nccl_unique_id) and is distributed by the caller (e.g. torch.distributed). */
GV_API int gv_nccl_unique_id(void *id128_out);
GV_API int gv_nccl_init(gv_ctx *ctx, const void *id128, int rank, int world);
GV_API int gv_nccl_world(gv_ctx *ctx, int *rank_out, int *world_out);
/* Collective finalize.  Every rank must have initialised the SAME grid geometry and set the SAME
 * base transform (verified collectively the first time after either changed: GV_ERR_STATE on a
 * mismatch).  All ranks' binned beams are summed (exact int sums over NVLink), the
 * de-duplicated raycast is split across ranks, partial miss planes are reduce-scattered,
 * every rank finalises its slab of the grid and the slabs are all-gathered, so each rank
 * ends with the identical full grid.  Bit-identical to a single-context run. */
GV_API int gv_grid_finalize_multi(gv_ctx *ctx, int32_t k_decay, const double *corners, int nfoot);
/* Optional: map every rank's grid planes into every other rank (cudaIpc over NVLink P2P) so
 * that gv_grid_finalize_multi runs FUSED over peer memory: the raycast sweep sums and clears the
 * binned-beam planes of all ranks itself (no all-reduce), and the finalise kernel sums all
 * ranks' count planes for its slab and writes the finished slab into every rank's grid (no
 * reduce-scatter / all-gather); the three barriers of that path are flag words in peer memory
 * (no NCCL call at all).
 * Call gv_ipc_export on every rank after gv_grid_init*, exchange the GV_IPC_BLOB_BYTES blobs
 * (rank order), call gv_ipc_import on every rank.  Re-do after re-initialising the grid. */
#define GV_IPC_BLOB_BYTES 448
GV_API int gv_ipc_export(gv_ctx *ctx, void *blob_out);
GV_API int gv_ipc_import(gv_ctx *ctx, const void *blobs, int world, int rank);
/* unmap the peers again: gv_grid_finalize_multi goes back to the NCCL collectives */
GV_API int gv_ipc_close(gv_ctx *ctx);

/* -------------------------------------------------------------- measurement --- */
/* Atomic-throughput ceilings of this device for the two atomic-bound kernels (the
 * raycast's cells/s is reported beside "RED.global to L2-resident lines" and "atomicAdd int32 to
 * shared, conflict-free").  Not part of the hot path and has no reference counterpart; bench.py
 * calls it to express k_sweep_walk / k_points atomic rates as fractions of a measured peak.
 * kind 0: RED.64 global, 1: RED.32 global, 2: shared int32 atomicAdd (conflict-free).
 * pattern (global kinds): 0 spread, 1 32 consecutive cells, 2 32 cells `stride` apart,
 * 3 runs of 4 lanes per cell, 4 one cell per warp.  ncells_pow2 cells of 8 bytes are allocated. */
GV_API int gv_microbench_atomics(int device, int kind, int pattern, size_t ncells_pow2,
                                 unsigned stride, int reps, double *ops_per_s_out);

#ifdef __cplusplus
}
#endif
#endif /* GRIDVISION_B200_H_ */
