/*
 * gv_oracle.c — CPU ORACLE (test infrastructure; see gv_oracle.h for scope and pinning).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared gv_oracle.c -lm
 * No -march, no -ffast-math: baseline x86-64 SSE2 float semantics like the reference
 * build (/root/reference/CMakeLists.txt:1-8 sets no arch/optimisation flags).
 * All citations are relative to /root/reference.
 */
#include "gv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* R1  LiDAR -> camera extrinsic transform                                   */
/* ------------------------------------------------------------------------- */

/* PCL common/impl/transforms.hpp, detail::Transformer<float>::se3 (SSE2 build):
 *   p0 = x*c0; p1 = y*c1; p2 = z*c2; out = p0 + (p1 + (p2 + c3))
 * with c_j the j-th COLUMN of the 4x4 matrix.  Separate multiply and add roundings.
 * RECALLED FROM UPSTREAM, UNVERIFIED OFFLINE: this function is the single place that
 * pins the add nesting.  Call site: src/grid_vision_node.cpp:304. */
static inline void gvo_se3(const float T[16], float x, float y, float z, float out[3])
{
  for (int r = 0; r < 3; ++r) {
    const float p0 = T[4 * r + 0] * x;
    const float p1 = T[4 * r + 1] * y;
    const float p2 = T[4 * r + 2] * z;
    const float s2 = p2 + T[4 * r + 3];
    const float s1 = p1 + s2;
    out[r] = p0 + s1;
  }
}

static inline int gvo_finite3(float x, float y, float z)
{
  return isfinite(x) && isfinite(y) && isfinite(z);
}

void gvo_transform_points(const float T[16], const float *x, const float *y, const float *z,
                          size_t n, int is_dense, float *ox, float *oy, float *oz)
{
  for (size_t i = 0; i < n; ++i) {
    /* pcl::transformPointCloud: dense clouds transform every point; otherwise points
     * with a non-finite coordinate are left as copied. */
    if (!is_dense && !gvo_finite3(x[i], y[i], z[i])) {
      ox[i] = x[i];
      oy[i] = y[i];
      oz[i] = z[i];
      continue;
    }
    float o[3];
    gvo_se3(T, x[i], y[i], z[i], o);
    ox[i] = o[0];
    oy[i] = o[1];
    oz[i] = o[2];
  }
}

/* ------------------------------------------------------------------------- */
/* R3 / R4  pinhole projection and box labelling                             */
/* ------------------------------------------------------------------------- */

/* src/cloud_detections.cpp:268-273 (and :19-24):
 *   Eigen::Vector3d img = K * Vector3d(pt.x, pt.y, pt.z);
 *   float u = img.x() / img.z();  float v = img.y() / img.z();
 * Eigen's fixed-size 3x3 * 3x1 product accumulates left to right, (a0*b0 + a1*b1) + a2*b2,
 * in double with separate mul/add.  For the reference's K (zero skew, float-valued
 * entries: src/object_detection.cpp:241-247, include/grid_vision/vision_orientation.hpp:18-25)
 * every product is exact and each row has <= 2 non-zero terms, so the value does not
 * depend on that order. */
static inline void gvo_project(const double K[9], float xf, float yf, float zf, float *u, float *v)
{
  const double X = (double)xf, Y = (double)yf, Z = (double)zf;
  const double ix = (K[0] * X + K[1] * Y) + K[2] * Z;
  const double iy = (K[3] * X + K[4] * Y) + K[5] * Z;
  const double iz = (K[6] * X + K[7] * Y) + K[8] * Z;
  *u = (float)(ix / iz);
  *v = (float)(iy / iz);
}

/* src/cloud_detections.cpp:262-289, one point.  Returns label (-1 = none). */
static inline int gvo_label_point(const double K[9], int W, int H, float x, float y, float z,
                                  const gvo_box *boxes, int nboxes, int32_t *pix, float *uo,
                                  float *vo)
{
  *pix = -1;
  *uo = NAN;
  *vo = NAN;
  /* :264  if(!pcl::isFinite(pt) || pt.z <= 0.001f) continue; */
  if (!gvo_finite3(x, y, z) || z <= 0.001f) return -1;
  float u, v;
  gvo_project(K, x, y, z, &u, &v);
  *uo = u;
  *vo = v;
  /* :276  if(u < 0 || u >= image_width || v < 0 || v >= image_height) continue;
   * float-vs-int comparisons: the int converts to float. */
  if (u < 0 || u >= (float)W || v < 0 || v >= (float)H) return -1;
  *pix = (int32_t)v * W + (int32_t)u;
  /* :280-288  first box in list order wins; float promoted to double; inclusive. */
  for (int i = 0; i < nboxes; ++i) {
    if ((double)u >= boxes[i].x_min && (double)u <= boxes[i].x_max &&
        (double)v >= boxes[i].y_min && (double)v <= boxes[i].y_max)
      return i;
  }
  return -1;
}

void gvo_project_label(const double K[9], int image_width, int image_height, const float *x,
                       const float *y, const float *z, size_t n, const gvo_box *boxes,
                       int nboxes, int16_t *label, int32_t *pix, float *u, float *v)
{
  for (size_t i = 0; i < n; ++i) {
    int32_t p;
    float uu, vv;
    const int l = gvo_label_point(K, image_width, image_height, x[i], y[i], z[i], boxes, nboxes,
                                  &p, &uu, &vv);
    if (label) label[i] = (int16_t)l;
    if (pix) pix[i] = p;
    if (u) u[i] = uu;
    if (v) v[i] = vv;
  }
}

void gvo_extract_cloud_per_bbox_aos(const gvo_point_xyzi *pts, size_t n, const double K[9],
                                    const gvo_box *boxes, int nboxes, int image_width,
                                    int image_height, gvo_point_xyzi **clouds, size_t *counts)
{
  /* :256-259 output_clouds.clear(); resize(bboxes.size()); */
  size_t *cap = (size_t *)calloc((size_t)(nboxes > 0 ? nboxes : 1), sizeof(size_t));
  for (int i = 0; i < nboxes; ++i) {
    clouds[i] = NULL;
    counts[i] = 0;
  }
  for (size_t k = 0; k < n; ++k) {
    int32_t p;
    float uu, vv;
    const int l = gvo_label_point(K, image_width, image_height, pts[k].x, pts[k].y, pts[k].z,
                                  boxes, nboxes, &p, &uu, &vv);
    if (l < 0) continue;
    /* :285 output_clouds[i].points.push_back(pt)  — doubling growth like std::vector. */
    if (counts[l] == cap[l]) {
      cap[l] = cap[l] ? 2 * cap[l] : 1;
      void *np = NULL;
      if (posix_memalign(&np, 16, cap[l] * sizeof(gvo_point_xyzi)) != 0) abort();
      if (clouds[l]) {
        memcpy(np, clouds[l], counts[l] * sizeof(gvo_point_xyzi));
        free(clouds[l]);
      }
      clouds[l] = (gvo_point_xyzi *)np;
    }
    clouds[l][counts[l]++] = pts[k];
  }
  free(cap);
}

size_t gvo_project_kdtree(const double K[9], const float *x, const float *y, const float *z,
                          size_t n, float *uvz)
{
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    /* src/cloud_detections.cpp:16  if(p.z <= 0) continue;   (NaN z passes) */
    if (z[i] <= 0) continue;
    float u, v;
    gvo_project(K, x[i], y[i], z[i], &u, &v);
    uvz[3 * m + 0] = u; /* :28-30 */
    uvz[3 * m + 1] = v;
    uvz[3 * m + 2] = z[i];
    ++m;
  }
  return m;
}

/* N4: cloud_detections::computeDepthForBoundingBoxes (src/cloud_detections.cpp:43-87) on the
 * (u, v, depth) triples buildKDTree emitted.  The tree is 3-D: depth is its third coordinate
 * (:28-30) and the query's is 0 (:60), so "nearest" minimises (u-cx)^2 + (v-cy)^2 + depth^2 in
 * FLANN's float accumulation order (pcl::KdTreeFLANN's default flann::L2_Simple<float>: one
 * `result += diff*diff` per coordinate).  Non-finite points are not in the tree (PCL drops them
 * when it converts the cloud).  Median = element size/2 of the ascending depths (:78-82);
 * -1 when nothing was found (:51).  FLANN's order among equal distances is unspecified: lower
 * index first here. */
void gvo_box_depths(const float *uvz, size_t m, const void *boxes40, int nb, int k, float *depths)
{
  const gvo_box *B = (const gvo_box *)boxes40;
  if (k > 64) k = 64;
  for (int b = 0; b < nb; ++b) {
    depths[b] = -1.0f;
    /* :57-60 centre in double (x_min + (x_max - x_min) / 2.0f), narrowed to the float fields */
    const float qx = (float)(B[b].x_min + ((B[b].x_max - B[b].x_min) / 2.0f));
    const float qy = (float)(B[b].y_min + ((B[b].y_max - B[b].y_min) / 2.0f));
    float bd[64];
    size_t bi[64];
    int found = 0;
    for (size_t i = 0; i < m; ++i) {
      const float u = uvz[3 * i], v = uvz[3 * i + 1], z = uvz[3 * i + 2];
      if (!(isfinite(u) && isfinite(v) && isfinite(z))) continue;
      float d = 0.0f, diff;
      diff = u - qx; d += diff * diff;
      diff = v - qy; d += diff * diff;
      diff = z - 0.0f; d += diff * diff;
      if (found == k && !(d < bd[k - 1])) continue; /* ties keep the earlier index */
      int pos = found < k ? found : k - 1;
      while (pos > 0 && d < bd[pos - 1]) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bd[pos] = d;
      bi[pos] = i;
      if (found < k) ++found;
    }
    if (found == 0 || k <= 0) continue;
    float dv[64];
    for (int j = 0; j < found; ++j) dv[j] = uvz[3 * bi[j] + 2];
    for (int j = 1; j < found; ++j) { /* ascending */
      const float t = dv[j];
      int q = j;
      while (q > 0 && t < dv[q - 1]) { dv[q] = dv[q - 1]; --q; }
      dv[q] = t;
    }
    depths[b] = dv[found / 2];
  }
}

/* cloud_detections::pixelTo3D (src/cloud_detections.cpp:89-103): depth * (K_inv * (u, v, 1)),
 * Eigen's 3x3 * 3x1 accumulated left to right in double. */
void gvo_pixel_to_3d(const double Kinv[9], float px, float py, float depth, double out[3])
{
  const double h[3] = {(double)px, (double)py, 1.0};
  for (int r = 0; r < 3; ++r) {
    const double t = (Kinv[3 * r] * h[0] + Kinv[3 * r + 1] * h[1]) + Kinv[3 * r + 2] * h[2];
    out[r] = (double)depth * t;
  }
}

/* ------------------------------------------------------------------------- */
/* R6  grid geometry (grid_map_core GridMap::setGeometry / setPosition)      */
/* ------------------------------------------------------------------------- */

int gvo_grid_init(gvo_grid *g, double length_x, double length_y, double res, double pos_x,
                  double pos_y)
{
  memset(g, 0, sizeof(*g));
  /* grid_map::GridMap::setGeometry: size(i) = static_cast<int>(round(length(i)/resolution));
   * length = size.cast<double>() * resolution.  RECALLED FROM UPSTREAM. */
  g->nx = (int32_t)round(length_x / res);
  g->ny = (int32_t)round(length_y / res);
  if (g->nx <= 0 || g->ny <= 0) return -1;
  g->res = res;
  g->len_x = (double)g->nx * res;
  g->len_y = (double)g->ny * res;
  g->pos_x = pos_x;
  g->pos_y = pos_y;
  const size_t nc = (size_t)g->nx * (size_t)g->ny;
  g->log_odds = (float *)malloc(nc * sizeof(float));
  g->occupancy = (float *)malloc(nc * sizeof(float));
  g->hit = (int32_t *)calloc(nc, sizeof(int32_t));
  g->miss = (int32_t *)calloc(nc, sizeof(int32_t));
  if (!g->log_odds || !g->occupancy || !g->hit || !g->miss) return -2;
  /* src/occupancy_grid.cpp:12-13 with include/grid_vision/occupancy_grid.hpp:27-28 */
  for (size_t i = 0; i < nc; ++i) {
    g->log_odds[i] = 0.0f;
    g->occupancy[i] = 0.5f;
  }
  return 0;
}

int gvo_grid_init_reference(gvo_grid *g, uint8_t grid_x, uint8_t grid_y, double res)
{
  /* src/occupancy_grid.cpp:10-11: Length(grid_x, grid_y); Position(grid_x / 3, 0.0) —
   * grid_x is uint8_t so grid_x / 3 is INTEGER division (50/3 -> 16). */
  return gvo_grid_init(g, (double)grid_x, (double)grid_y, res, (double)(grid_x / 3), 0.0);
}

void gvo_grid_free(gvo_grid *g)
{
  free(g->log_odds);
  free(g->occupancy);
  free(g->hit);
  free(g->miss);
  memset(g, 0, sizeof(*g));
}

/* grid_map_core GridMapMath.cpp getIndexFromPosition + checkIfPositionWithinMap +
 * checkIfIndexInRange, startIndex = 0 (the reference never moves the map).
 * RECALLED FROM UPSTREAM, UNVERIFIED OFFLINE — the single place that pins it:
 *   offset      = 0.5 * length
 *   indexVector = ((position - offset) - mapPosition) / resolution     (true division)
 *   index       = (int)(-indexVector)                                  (trunc toward 0)
 *   within      : q = -((position - mapPosition) - offset); 0 <= q < length on both axes
 *   in range    : 0 <= index < size
 * Call site in the reference: src/occupancy_grid.cpp:152. */
static inline double gvo_index_coord(double p, double len, double pos, double res)
{
  const double offset = 0.5 * len;
  return -(((p - offset) - pos) / res);
}

int gvo_grid_get_index(const gvo_grid *g, double px, double py, int32_t *ix, int32_t *iy)
{
  const double qx = -((px - g->pos_x) - 0.5 * g->len_x);
  const double qy = -((py - g->pos_y) - 0.5 * g->len_y);
  /* NaN compares false -> outside. Evaluated before the int cast so the cast is defined. */
  if (!(qx >= 0.0 && qy >= 0.0 && qx < g->len_x && qy < g->len_y)) return 0;
  const double ax = gvo_index_coord(px, g->len_x, g->pos_x, g->res);
  const double ay = gvo_index_coord(py, g->len_y, g->pos_y, g->res);
  const int32_t i = (int32_t)ax;
  const int32_t j = (int32_t)ay;
  if (!(i >= 0 && j >= 0 && i < g->nx && j < g->ny)) return 0;
  *ix = i;
  *iy = j;
  return 1;
}

/* ------------------------------------------------------------------------- */
/* R7 / R8 / R9  per-frame updates, exactly as the reference structures them  */
/* ------------------------------------------------------------------------- */

/* include/grid_vision/occupancy_grid.hpp:25-31 */
static const float kLogOddsFree = -0.4f;     /* declared, unused by the reference */
static const float kLogOddsOccupied = 1.2f;  /* declared, unused by the reference */
static const float kLogOddsDecay = -0.2f;
static const float kMinLogOdds = -2.0f;
static const float kMaxLogOdds = 3.6f;
static const float kBlockIncrement = 0.85f;  /* literal at src/occupancy_grid.cpp:182 */

static void gvo_decay(gvo_grid *g)
{
  /* src/occupancy_grid.cpp:19,38,69  grid["log_odds"].array() += log_odds_decay_; */
  const size_t nc = (size_t)g->nx * (size_t)g->ny;
  for (size_t i = 0; i < nc; ++i) g->log_odds[i] = g->log_odds[i] + kLogOddsDecay;
}

static void gvo_clamp_sigmoid(gvo_grid *g)
{
  const size_t nc = (size_t)g->nx * (size_t)g->ny;
  /* :21-22  cwiseMax(min_log_odds_).cwiseMin(max_log_odds_) */
  for (size_t i = 0; i < nc; ++i) {
    float l = g->log_odds[i];
    l = l < kMinLogOdds ? kMinLogOdds : l;
    l = l > kMaxLogOdds ? kMaxLogOdds : l;
    g->log_odds[i] = l;
  }
  /* :25-30  probability = 1.0f / (1.0f + std::exp(-log_odds_value)) in float */
  for (size_t i = 0; i < nc; ++i) {
    const float l = g->log_odds[i];
    g->occupancy[i] = 1.0f / (1.0f + expf(-l));
  }
}

/* src/occupancy_grid.cpp:140-183.  Returns 1 and the inclusive rectangle when every
 * corner indexes validly, 0 (whole footprint skipped, :171-172) otherwise. */
static int gvo_footprint_rect(const gvo_grid *g, const double c[8], int32_t r[4])
{
  int32_t minx = 0, miny = 0, maxx = 0, maxy = 0;
  for (int i = 0; i < 4; ++i) {
    int32_t ix, iy;
    if (!gvo_grid_get_index(g, c[2 * i], c[2 * i + 1], &ix, &iy)) return 0; /* :152-156 */
    if (i == 0) {
      minx = maxx = ix;
      miny = maxy = iy;
    } else {
      if (ix < minx) minx = ix;
      if (iy < miny) miny = iy;
      if (ix > maxx) maxx = ix;
      if (iy > maxy) maxy = iy;
    }
  }
  r[0] = minx;
  r[1] = miny;
  r[2] = maxx;
  r[3] = maxy;
  return 1;
}

void gvo_update_grid_cells_fast(gvo_grid *g, const double corners[8])
{
  int32_t r[4];
  if (!gvo_footprint_rect(g, corners, r)) return;
  /* :178-182  grid_data.block(min.x, min.y, nx, ny).array() += 0.85f  (column-major) */
  for (int32_t j = r[1]; j <= r[3]; ++j)
    for (int32_t i = r[0]; i <= r[2]; ++i) {
      const size_t lin = (size_t)i + (size_t)j * (size_t)g->nx;
      g->log_odds[lin] = g->log_odds[lin] + kBlockIncrement;
    }
}

void gvo_update_map(gvo_grid *g)
{
  gvo_decay(g);         /* :19 */
  gvo_clamp_sigmoid(g); /* :21-30 */
}

/* src/occupancy_grid.cpp:79-90: corners {left_back, left_front, right_front, right_back}. */
static void gvo_pose_corners(const double p[4], double c[8])
{
  const double x = p[0], y = p[1], length = p[2], width = p[3];
  const double xf = x + length / 2.0, xb = x - length / 2.0;
  const double yl = y - width / 2.0, yr = y + width / 2.0;
  c[0] = xb; c[1] = yl; /* left_back   */
  c[2] = xf; c[3] = yl; /* left_front  */
  c[4] = xf; c[5] = yr; /* right_front */
  c[6] = xb; c[7] = yr; /* right_back  */
}

void gvo_update_map_poses(gvo_grid *g, const double *xylw, int n)
{
  gvo_decay(g); /* :69 */
  for (int k = 0; k < n; ++k) {
    double c[8];
    gvo_pose_corners(xylw + 4 * k, c);
    gvo_update_grid_cells_fast(g, c); /* :93 */
  }
  gvo_clamp_sigmoid(g); /* :96-104 */
}

float gvo_estimated_depth(int32_t label)
{
  /* src/occupancy_grid.cpp:185-196 */
  switch (label) {
  case GVO_VEHICLE: return 3.5f;
  case GVO_PERSON: return 0.6f;
  case GVO_BIKE: return 2.5f;
  case GVO_MOTORBIKE: return 2.5f;
  default: return -1.0f;
  }
}

/* src/occupancy_grid.cpp:107-138: corners {LF, RF, RB, LB}; estimated_depth is float,
 * promoted to double when added to the double point coordinates. */
static void gvo_point_corners(const double xy[2], int32_t label, double c[8])
{
  const float d = gvo_estimated_depth(label);
  c[0] = xy[0] + d; c[1] = xy[1] + (d / 2); /* :118-119 */
  c[2] = xy[0] + d; c[3] = xy[1] - (d / 2); /* :123-124 */
  c[4] = xy[0];     c[5] = xy[1] - (d / 2); /* :128-129 */
  c[6] = xy[0];     c[7] = xy[1] + (d / 2); /* :133-134 */
}

void gvo_update_map_points(gvo_grid *g, const double *xy, const int32_t *labels, int n)
{
  gvo_decay(g); /* :38 */
  for (int k = 0; k < n; ++k) {
    double c[8];
    gvo_point_corners(xy + 2 * k, labels[k], c); /* :48-49 */
    gvo_update_grid_cells_fast(g, c);            /* :52 */
  }
  gvo_clamp_sigmoid(g); /* :54-62 */
}

/* ------------------------------------------------------------------------- */
/* X2  Bresenham (grid_map_core LineIterator)                                */
/* ------------------------------------------------------------------------- */

/* grid_map_core/src/iterators/LineIterator.cpp initializeIterationParameters + operator++.
 * RECALLED FROM UPSTREAM, UNVERIFIED OFFLINE — the single place that pins it. */
typedef struct {
  int32_t x, y, inc1x, inc1y, inc2x, inc2y, den, num, add, n;
} gvo_line;

static inline void gvo_line_init(gvo_line *L, int32_t sx, int32_t sy, int32_t ex, int32_t ey)
{
  const int32_t dx = ex >= sx ? ex - sx : sx - ex;
  const int32_t dy = ey >= sy ? ey - sy : sy - ey;
  L->x = sx;
  L->y = sy;
  L->inc1x = L->inc2x = ex >= sx ? 1 : -1;
  L->inc1y = L->inc2y = ey >= sy ? 1 : -1;
  if (dx >= dy) {
    L->inc1x = 0;
    L->inc2y = 0;
    L->den = dx;
    L->num = dx / 2;
    L->add = dy;
    L->n = dx + 1;
  } else {
    L->inc2x = 0;
    L->inc1y = 0;
    L->den = dy;
    L->num = dy / 2;
    L->add = dx;
    L->n = dy + 1;
  }
}

static inline void gvo_line_step(gvo_line *L)
{
  L->num += L->add;
  if (L->num >= L->den) {
    L->num -= L->den;
    L->x += L->inc1x;
    L->y += L->inc1y;
  }
  L->x += L->inc2x;
  L->y += L->inc2y;
}

int32_t gvo_bresenham_cells(int32_t sx, int32_t sy, int32_t ex, int32_t ey, int32_t *cells_xy,
                            int32_t cap)
{
  gvo_line L;
  gvo_line_init(&L, sx, sy, ex, ey);
  for (int32_t k = 0; k < L.n; ++k) {
    if (k < cap) {
      cells_xy[2 * k] = L.x;
      cells_xy[2 * k + 1] = L.y;
    }
    if (k + 1 < L.n) gvo_line_step(&L);
  }
  return L.n;
}

/* ------------------------------------------------------------------------- */
/* X1 + X2  bin + raycast, one beam at a time (brute force, no de-duplication) */
/* ------------------------------------------------------------------------- */

/* Geometry of free-space-only beams (range-capped and/or off-map endpoints) — a
 * specification authored here (the reference has no raycast), chosen so that
 * it is cheap on the GPU and bit-reproducible everywhere: every step is ONE separately rounded
 * IEEE binary32 operation (add, sub, mul, div, sqrt — no FMA), in exactly this order.
 *   range cap   d = p - o;  r2 = dx*dx + dy*dy;  capped iff r2 > rmax*rmax;
 *               p' = o + (rmax / sqrt(r2)) * d
 *   index coord a(p) = (c0 - p) * inv_res,  c0 = (float)(0.5*len + pos), inv_res = (float)(1/res)
 *   clip        one parametric clip of o_a -> e_a against [0,nx) x [0,ny), end cell = the
 *               truncated clipped coordinate, clamped into the map (NaN clamps to 0).
 * Only WHETHER an endpoint is inside the map is decided by the exact grid_map getIndex. */
typedef struct {
  float oxf, oyf;       /* sensor origin (translation column of T_base<-lidar) */
  float c0xf, c0yf, inv_resf;
  float oaxf, oayf;     /* a(origin) */
  float rmaxf, rmax2f;
} gvo_beam_geom;

static inline void gvo_beam_geom_init(gvo_beam_geom *b, const gvo_grid *g, const float T[16],
                                      double r_max)
{
  b->oxf = T[3];
  b->oyf = T[7];
  b->c0xf = (float)(0.5 * g->len_x + g->pos_x);
  b->c0yf = (float)(0.5 * g->len_y + g->pos_y);
  b->inv_resf = (float)(1.0 / g->res);
  b->oaxf = (b->c0xf - b->oxf) * b->inv_resf;
  b->oayf = (b->c0yf - b->oyf) * b->inv_resf;
  b->rmaxf = (float)r_max;
  b->rmax2f = b->rmaxf * b->rmaxf;
}

static inline int32_t gvo_clamp_cell(float c, int32_t n)
{
  /* written so that NaN falls to 0 and the int cast is always defined */
  return c >= 0.0f ? (c >= (float)n ? n - 1 : (int32_t)c) : 0;
}

static inline void gvo_clip_end(const gvo_beam_geom *b, float pxf, float pyf, int32_t nx,
                                int32_t ny, int32_t *ex, int32_t *ey)
{
  const float nxf = (float)nx, nyf = (float)ny;
  const float eax = (b->c0xf - pxf) * b->inv_resf;
  const float eay = (b->c0yf - pyf) * b->inv_resf;
  const float dax = eax - b->oaxf, day = eay - b->oayf;
  float t = 1.0f;
  if (eax < 0.0f) {
    const float tt = (0.0f - b->oaxf) / dax;
    if (tt < t) t = tt;
  } else if (eax >= nxf) {
    const float tt = (nxf - b->oaxf) / dax;
    if (tt < t) t = tt;
  }
  if (eay < 0.0f) {
    const float tt = (0.0f - b->oayf) / day;
    if (tt < t) t = tt;
  } else if (eay >= nyf) {
    const float tt = (nyf - b->oayf) / day;
    if (tt < t) t = tt;
  }
  const float mx = t * dax, my = t * day;
  *ex = gvo_clamp_cell(b->oaxf + mx, nx);
  *ey = gvo_clamp_cell(b->oayf + my, ny);
}

int64_t gvo_accumulate(gvo_grid *g, const float T[16], const float *x, const float *y,
                       const float *z, size_t n, const int16_t *labels,
                       const gvo_accum_params *prm, int32_t *cell_out, uint8_t *flags_out)
{
  /* sensor origin in the base frame = translation column of T (float), promoted */
  const double ox = (double)T[3], oy = (double)T[7];
  int32_t sx, sy;
  const int origin_ok = gvo_grid_get_index(g, ox, oy, &sx, &sy);
  const int cap = prm->r_max > 0.0;
  gvo_beam_geom bg;
  gvo_beam_geom_init(&bg, g, T, prm->r_max);
  int64_t updates = 0;

  for (size_t i = 0; i < n; ++i) {
    if (cell_out) cell_out[i] = -1;
    if (flags_out) flags_out[i] = 0;
    if (!origin_ok) continue;
    if (!gvo_finite3(x[i], y[i], z[i])) continue; /* no direction: beam dropped */
    float pb[3];
    gvo_se3(T, x[i], y[i], z[i], pb); /* X1: p_base = T_base<-lidar * p, R1 op order */
    if (!gvo_finite3(pb[0], pb[1], pb[2])) continue;
    float pxf = pb[0], pyf = pb[1];
    int hit_ok = 1;
    uint8_t flags = GVO_F_VALID;
    if (cap) {
      const float dx = pxf - bg.oxf, dy = pyf - bg.oyf;
      const float dx2 = dx * dx, dy2 = dy * dy;
      const float r2 = dx2 + dy2;
      if (r2 > bg.rmax2f) {
        const float sf = bg.rmaxf / sqrtf(r2);
        const float mx = sf * dx, my = sf * dy;
        pxf = bg.oxf + mx;
        pyf = bg.oyf + my;
        hit_ok = 0;
        flags |= GVO_F_RANGECAP;
      }
    }
    int32_t ex, ey;
    /* X1: cell = grid_map getIndex of the (float) base-frame position, promoted to double */
    if (!gvo_grid_get_index(g, (double)pxf, (double)pyf, &ex, &ey)) {
      gvo_clip_end(&bg, pxf, pyf, g->nx, g->ny, &ex, &ey);
      hit_ok = 0;
      flags |= GVO_F_CLIPPED;
    }
    if (hit_ok && prm->use_z_gate && !(pb[2] >= prm->z_min && pb[2] <= prm->z_max)) hit_ok = 0;
    if (hit_ok && prm->occ_mode == GVO_OCC_LABELLED && !(labels && labels[i] >= 0)) hit_ok = 0;
    if (hit_ok) flags |= GVO_F_HIT;
    if (cell_out) cell_out[i] = ex + ey * g->nx;
    if (flags_out) flags_out[i] = flags;

    /* X2: cells 0..n-2 get a miss; the end cell gets the hit (or a miss for no-hit beams) */
    gvo_line L;
    gvo_line_init(&L, sx, sy, ex, ey);
    for (int32_t k = 0; k + 1 < L.n; ++k) {
      g->miss[(size_t)L.x + (size_t)L.y * (size_t)g->nx] += 1;
      gvo_line_step(&L);
    }
    const size_t lin = (size_t)L.x + (size_t)L.y * (size_t)g->nx;
    if (hit_ok) g->hit[lin] += 1;
    else g->miss[lin] += 1;
    updates += L.n;
  }
  return origin_ok ? updates : -1;
}

/* ------------------------------------------------------------------------- */
/* X3  counts -> log-odds (+ R8 footprints), clamp, sigmoid                  */
/* ------------------------------------------------------------------------- */

void gvo_finalize(gvo_grid *g, int32_t k_decay, const double *corners, int nfoot)
{
  const size_t nc = (size_t)g->nx * (size_t)g->ny;
  /* canonical order; each product and each sum separately rounded (no FMA).
   * k_decay == 1 with zero counts reproduces R7/R8 bit for bit: 1.0f*(-0.2f) == -0.2f and
   * adding 0.0f*c == -0.0f / +0.0f leaves any l unchanged. */
  for (size_t i = 0; i < nc; ++i) {
    float l = g->log_odds[i];
    const float d = (float)k_decay * kLogOddsDecay;
    l = l + d;
    const float m = (float)g->miss[i] * kLogOddsFree;
    l = l + m;
    const float h = (float)g->hit[i] * kLogOddsOccupied;
    l = l + h;
    g->log_odds[i] = l;
    g->hit[i] = 0;
    g->miss[i] = 0;
  }
  for (int k = 0; k < nfoot; ++k) gvo_update_grid_cells_fast(g, corners + 8 * k);
  gvo_clamp_sigmoid(g);
}

/* ------------------------------------------------------------------------- */
/* N1  ground-plane removal (src/cloud_detections.cpp:105-138)                */
/* ------------------------------------------------------------------------- */

static inline uint32_t gvo_mix32(uint32_t seed, uint32_t k)
{
  /* counter-based integer hash (two rounds of a multiply-xorshift mixer), identical on the GPU */
  uint32_t v = seed ^ (k * 0x9E3779B9u);
  v ^= v >> 16; v *= 0x85EBCA6Bu;
  v ^= v >> 13; v *= 0xC2B2AE35u;
  v ^= v >> 16;
  return v;
}

static inline float gvo_plane_dist(const float p[4], float x, float y, float z)
{
  const float a = p[0] * x, b = p[1] * y, c = p[2] * z;
  const float s1 = a + b;
  const float s2 = s1 + c;
  return fabsf(s2 + p[3]);
}

/* plane through three points, float; returns 0 for a degenerate / non-finite sample */
static int gvo_plane_from3(const float *x, const float *y, const float *z, size_t i0, size_t i1,
                           size_t i2, float p[4])
{
  if (!gvo_finite3(x[i0], y[i0], z[i0]) || !gvo_finite3(x[i1], y[i1], z[i1]) ||
      !gvo_finite3(x[i2], y[i2], z[i2]))
    return 0;
  const float ux = x[i1] - x[i0], uy = y[i1] - y[i0], uz = z[i1] - z[i0];
  const float vx = x[i2] - x[i0], vy = y[i2] - y[i0], vz = z[i2] - z[i0];
  const float cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
  const float n2 = (cx * cx + cy * cy) + cz * cz;
  if (!(n2 > 0.0f) || !isfinite(n2)) return 0;
  const float inv = 1.0f / sqrtf(n2);
  p[0] = cx * inv; p[1] = cy * inv; p[2] = cz * inv;
  p[3] = -((p[0] * x[i0] + p[1] * y[i0]) + p[2] * z[i0]);
  return 1;
}

/* eigenvector of the smallest eigenvalue of a symmetric 3x3 matrix (cyclic Jacobi, double) */
static void gvo_smallest_eigvec3(const double A[6] /* xx xy xz yy yz zz */, double v[3])
{
  double a[3][3] = {{A[0], A[1], A[2]}, {A[1], A[3], A[4]}, {A[2], A[4], A[5]}};
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(a[p][q]) < 1e-300) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  if (a[1][1] < a[m][m]) m = 1;
  if (a[2][2] < a[m][m]) m = 2;
  v[0] = V[0][m]; v[1] = V[1][m]; v[2] = V[2][m];
}

int64_t gvo_segment_ground(const float *x, const float *y, const float *z, size_t n, float threshold,
                           uint32_t seed, int32_t n_hyp, uint8_t *keep, float plane[4],
                           int32_t *best_h, int32_t *best_score)
{
  for (size_t i = 0; i < n; ++i) keep[i] = 1;
  *best_h = -1;
  *best_score = 0;
  plane[0] = plane[1] = plane[2] = plane[3] = 0.0f;
  if (n < 3) return -1;
  float bp[4] = {0, 0, 0, 0};
  for (int32_t h = 0; h < n_hyp; ++h) {
    float p[4];
    const size_t i0 = gvo_mix32(seed, 3u * (uint32_t)h) % n, i1 = gvo_mix32(seed, 3u * (uint32_t)h + 1u) % n,
                 i2 = gvo_mix32(seed, 3u * (uint32_t)h + 2u) % n;
    if (!gvo_plane_from3(x, y, z, i0, i1, i2, p)) continue;
    int32_t score = 0;
    for (size_t i = 0; i < n; ++i) score += gvo_plane_dist(p, x[i], y[i], z[i]) < threshold;
    if (score > *best_score) {
      *best_score = score;
      *best_h = h;
      memcpy(bp, p, sizeof(bp));
    }
  }
  if (*best_score < 3) return -1; /* :122-126 no planar model -> the reference returns an empty cloud */
  if (*best_score == 3) {         /* PCL refines only with more inliers than the sample size */
    memcpy(plane, bp, sizeof(bp));
    int64_t kept3 = 0;
    for (size_t i = 0; i < n; ++i) {
      keep[i] = !(gvo_plane_dist(plane, x[i], y[i], z[i]) < threshold);
      kept3 += keep[i];
    }
    return kept3;
  }
  /* setOptimizeCoefficients(true): least-squares plane of the inliers */
  double c[3] = {0, 0, 0}, m = 0;
  for (size_t i = 0; i < n; ++i)
    if (gvo_plane_dist(bp, x[i], y[i], z[i]) < threshold) {
      c[0] += x[i]; c[1] += y[i]; c[2] += z[i];
      m += 1;
    }
  c[0] /= m; c[1] /= m; c[2] /= m;
  double A[6] = {0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < n; ++i)
    if (gvo_plane_dist(bp, x[i], y[i], z[i]) < threshold) {
      const double dx = x[i] - c[0], dy = y[i] - c[1], dz = z[i] - c[2];
      A[0] += dx * dx; A[1] += dx * dy; A[2] += dx * dz;
      A[3] += dy * dy; A[4] += dy * dz; A[5] += dz * dz;
    }
  double v[3];
  gvo_smallest_eigvec3(A, v);
  /* orient like the RANSAC hypothesis so the refinement never flips the normal */
  if (v[0] * bp[0] + v[1] * bp[1] + v[2] * bp[2] < 0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }
  plane[0] = (float)v[0]; plane[1] = (float)v[1]; plane[2] = (float)v[2];
  plane[3] = (float)(-(v[0] * c[0] + v[1] * c[1] + v[2] * c[2]));
  int64_t kept = 0;
  for (size_t i = 0; i < n; ++i) {
    keep[i] = !(gvo_plane_dist(plane, x[i], y[i], z[i]) < threshold); /* :128-135 setNegative(true) */
    kept += keep[i];
  }
  return kept;
}

/* ------------------------------------------------------------------------- */
/* N2  radius outlier removal + PCA box (src/cloud_detections.cpp:140-247)    */
/* ------------------------------------------------------------------------- */

int32_t gvo_radius_outlier_keep(const float *x, const float *y, const float *z, size_t n,
                                double radius, int32_t min_neighbors, uint8_t *keep)
{
  /* :150-154 RadiusOutlierRemoval: setRadiusSearch(0.4), setMinNeighborsInRadius(10).
   * pcl::KdTreeFLANN::radiusSearch hands FLANN static_cast<float>(radius * radius); FLANN's
   * RadiusResultSet keeps dist < radius2 (strict); k includes the query; k <= min_pts removes. */
  const float r2 = (float)(radius * radius);
  int32_t kept = 0;
  for (size_t i = 0; i < n; ++i) {
    int32_t k = 0;
    for (size_t j = 0; j < n; ++j) {
      const float dx = x[j] - x[i], dy = y[j] - y[i], dz = z[j] - z[i];
      const float dxx = dx * dx, dyy = dy * dy, dzz = dz * dz;
      const float s = dxx + dyy;
      const float d2 = s + dzz;
      if (d2 < r2) ++k;
    }
    keep[i] = k > min_neighbors;
    kept += keep[i];
  }
  return kept;
}

void gvo_bbox_pose(const float *x, const float *y, const float *z, size_t n, gvo_lshape *out)
{
  memset(out, 0, sizeof(*out));
  out->qw = 1.0;
  if (n == 0) return;
  uint8_t *keep = (uint8_t *)malloc(n);
  const int32_t m = gvo_radius_outlier_keep(x, y, z, n, 0.4, 10, keep);
  out->kept = m;
  if (m == 0) { /* :201-202 if(data.empty()) continue; */
    free(keep);
    return;
  }
  /* :157-158 compute3DCentroid (float accumulation in input order) */
  float cy = 0.0f;
  double sz = 0, sx = 0;
  for (size_t i = 0; i < n; ++i)
    if (keep[i]) {
      cy = cy + y[i];
      sz += (double)z[i];
      sx += (double)x[i];
    }
  out->centroid_y = cy / (float)m;
  /* :189-191 cv::PCA(data, Mat(), DATA_AS_ROW) on rows (z, x) */
  const double mz = sz / m, mx = sx / m;
  double czz = 0, czx = 0, cxx = 0;
  for (size_t i = 0; i < n; ++i)
    if (keep[i]) {
      const double dz = (double)z[i] - mz, dx = (double)x[i] - mx;
      czz += dz * dz;
      czx += dz * dx;
      cxx += dx * dx;
    }
  czz /= m;
  czx /= m;
  cxx /= m;
  /* symmetric 2x2 eigen decomposition, larger eigenvalue first */
  const double tr = czz + cxx, df = czz - cxx;
  const double rad = sqrt(df * df + 4.0 * czx * czx);
  const double l1 = 0.5 * (tr + rad);
  double vz, vx;
  if (fabs(czx) > 1e-300) {
    vz = l1 - cxx;
    vx = czx;
  } else if (czz >= cxx) {
    vz = 1.0;
    vx = 0.0;
  } else {
    vz = 0.0;
    vx = 1.0;
  }
  const double nv = sqrt(vz * vz + vx * vx);
  vz /= nv;
  vx /= nv;
  if (vz < 0 || (vz == 0 && vx < 0)) { /* sign convention of THIS restatement: major.z >= 0 */
    vz = -vz;
    vx = -vx;
  }
  out->mean_z = (float)mz;
  out->mean_x = (float)mx;
  out->major_z = (float)vz;
  out->major_x = (float)vx;
  out->minor_z = (float)(-vx); /* orthogonal complement */
  out->minor_x = (float)vz;
  /* :204-222 extents along the axes, float like the reference's cv::Point2f arithmetic */
  float minL = 3.402823466e+38f, maxL = -3.402823466e+38f, minW = minL, maxW = maxL;
  for (size_t i = 0; i < n; ++i)
    if (keep[i]) {
      const float dz = z[i] - out->mean_z, dx = x[i] - out->mean_x;
      const float pL = dz * out->major_z + dx * out->major_x;
      const float pW = dz * out->minor_z + dx * out->minor_x;
      if (pL < minL) minL = pL;
      if (pL > maxL) maxL = pL;
      if (pW < minW) minW = pW;
      if (pW > maxW) maxW = pW;
    }
  out->length = maxL - minL;
  out->width = maxW - minW;
  /* :232 angle in degrees; :246-247 q.setRPY(0, -angle, 0) (degrees passed as radians: kept) */
  out->angle_deg = atan2f(out->major_x, out->major_z) * 180.0f / 3.14159265358979323846f;
  const double hp = -(double)out->angle_deg * 0.5;
  out->qx = 0.0;
  out->qy = sin(hp);
  out->qz = 0.0;
  out->qw = cos(hp);
  free(keep);
}

/* ------------------------------------------------------------------------- */
/* N3  nav_msgs/OccupancyGrid cell conversion (grid_map_ros toOccupancyGrid)  */
/* ------------------------------------------------------------------------- */

void gvo_to_occupancy_grid(const gvo_grid *g, int8_t *data)
{
  /* call site src/grid_vision_node.cpp:270: toOccupancyGrid(grid, "occupancy", 0.0, 1.0, msg).
   * RECALLED FROM UPSTREAM grid_map_ros GridMapRosConverter::toOccupancyGrid:
   *   value = (cell - dataMin) / (dataMax - dataMin); NaN -> -1 else 0 + clamp01(value)*100;
   *   data[nCells - lin - 1] = value  (float -> int8 truncation), lin column-major. */
  const size_t nc = (size_t)g->nx * (size_t)g->ny;
  const float dataMin = 0.0f, dataMax = 1.0f;
  for (size_t lin = 0; lin < nc; ++lin) {
    float value = (g->occupancy[lin] - dataMin) / (dataMax - dataMin);
    if (isnan(value)) value = -1.0f;
    else {
      float c = value < 0.0f ? 0.0f : value;
      c = c > 1.0f ? 1.0f : c;
      value = 0.0f + c * 100.0f;
    }
    data[nc - lin - 1] = (int8_t)value;
  }
}
