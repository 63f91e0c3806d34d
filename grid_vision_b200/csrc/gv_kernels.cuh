// gv_kernels.cuh — sm_100a kernels for the grid-vision point-cloud -> occupancy-grid path.
//
// Every parity-critical floating-point operation is written with an explicit
// round-to-nearest intrinsic (__fmul_rn, __dadd_rn, __ddiv_rn, ...) so nvcc can never
// contract a multiply-add into an FMA: the reference CPU build has no FMA
// (baseline x86-64, /root/reference CMakeLists.txt:1-8) and the summation order is part
// of the contract.  __fma_rn is used only where both products are exact in double, so
// that fma(a,b,c) == RN(a*b + c) by construction (see project_point).
//
// No tensor cores anywhere: nothing on this path is a dense contraction.  The kernels are
// HBM-streaming (points, grid planes) or atomic-bound (binning, raycast).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gv {

constexpr int kMaxCam = 8;
constexpr int kThreads = 256;
constexpr int kPtsPerThread = 4;
constexpr int kTilePts = kThreads * kPtsPerThread;  // 1024 points per CTA tile
constexpr int kMaxFootCand = 256;

// ----------------------------------------------------------------------------------
// parameter blocks (passed by value as __grid_constant__, no __constant__ state so that
// several contexts/streams can share a device)
// ----------------------------------------------------------------------------------
struct CamDev {
  double K[9];   // row-major intrinsics
  float T[12];   // row-major 3x4 T_cam<-lidar
  float Wf, Hf;  // (float)image_width/height: the reference compares float u against int W
  int W, H;
  int has_T;     // 0: cloud already in the camera frame
  int canon;     // K == [fx 0 cx; 0 fy cy; 0 0 1] with float-representable entries
  int box_begin, box_end;
  int t_small;   // every |T entry| < 1e6 (with small3 inputs the transform cannot overflow)
  // certified float projection (canonical K only): exact float copies of fx, fy, cx, cy and the
  // constants of the error bound E(q) = e6*|q| + e0 (see fuse_point_fast)
  float fxf, fyf, cxf, cyf;
  float e6, e0u, e0v;
};

struct GridGeom {
  int nx, ny;
  double res, len_x, len_y, pos_x, pos_y, half_x, half_y;  // half = 0.5*len (exact)
};

struct BinDev {
  float T[12];  // row-major 3x4 T_base<-lidar
  GridGeom g;
  int sx, sy;  // origin cell (grid_map getIndex of the sensor origin)
  int origin_ok;
  int occ_mode, use_z_gate;
  float z_min, z_max;
  // float geometry of free-space-only beams (oracle gvo_beam_geom)
  float oxf, oyf, c0xf, c0yf, inv_resf, oaxf, oayf, rmaxf, rmax2f;
  int cap;
  // certified fixed-point index: k = trunc((c0 - p) * (2^16 / res)) holds cell and fraction
  double c0xd, c0yd, mres;
  int klim_x, klim_y;  // size << 16
  int fast_index_ok;
  int t_small;
};

struct PointArgs {
  const float *x, *y, *z;
  unsigned long long n;
  int is_dense;
  int vec_ok;  // x,y,z base pointers are 16-byte aligned
  // fusion
  int ncam;
  CamDev cam[kMaxCam];
  const float4 *boxes;  // (x_min, y_min, x_max, y_max) pre-rounded to float, see k_round_boxes
  // image-tile prefilter (k_box_masks): per box set (camera, or frame in batch mode) one
  // bit mask per tile of 2^mask_shift x 2^mask_shift pixels naming the boxes that overlap it
  const unsigned long long *masks;  // [nsets][mask_stride]
  int mask_shift[kMaxCam], mask_tx[kMaxCam];
  int mask_words;   // 64-bit words per tile
  int mask_stride;  // words per set
  int smem_boxes;   // float4 slots reserved for boxes in dynamic shared memory
  int16_t *labels;  // cam-major planes, stride n (nullable)
  int32_t *pix;     // nullable
  float *uv;        // nullable
  // batch mode (per-frame box lists, camera 0 only); nframes == 0 -> single cloud
  int nframes;
  unsigned tile0;                        // first tile of this launch (chunked batches)
  int tile_pts;                          // points per CTA tile (multiple of kTilePts)
  const unsigned long long *tile_start;  // [ntiles] first point of the tile
  const unsigned long long *tile_end;    // [ntiles] end of the tile's frame (exclusive)
  const int4 *tile_boxes;                // [ntiles] (box_begin, box_end, frame, 0)
  // binning
  BinDev bin;
  const int16_t *labels_in;  // labels for GV_OCC_LABELLED when not fusing in the same pass
  unsigned long long *ends;  // per cell: low 32 = beams ending here, high 32 = of which hits
  int32_t *cell_out;         // nullable
  uint8_t *flags_out;        // nullable
};

// ----------------------------------------------------------------------------------
// device helpers — each mirrors one oracle function (oracle/gv_oracle.c) op for op
// ----------------------------------------------------------------------------------
__device__ __forceinline__ bool finitef(float v)
{
  return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u;
}
__device__ __forceinline__ bool finite3(float x, float y, float z)
{
  return finitef(x) && finitef(y) && finitef(z);
}

// all three |v| < 1e9 (hence finite): one integer max over the sign-stripped bit patterns, so
// NaN and Inf (larger patterns) fail.  With |T entries| < 1e6 (t_small, checked on the host) the
// SE(3) outputs are then below 3.1e15 in magnitude: finite, no separate test needed.
__device__ __forceinline__ bool small3(float x, float y, float z)
{
  const unsigned ux = __float_as_uint(x) & 0x7fffffffu, uy = __float_as_uint(y) & 0x7fffffffu,
                 uz = __float_as_uint(z) & 0x7fffffffu;
  return max(max(ux, uy), uz) < 0x4e6e6b28u;  // bits of 1.0e9f
}

// R1: PCL Transformer<float>::se3 (SSE2): out = c0*x + (c1*y + (c2*z + c3)); call site
// ref: src/grid_vision_node.cpp:304.
__device__ __forceinline__ void se3(const float *T, float x, float y, float z, float &ox,
                                    float &oy, float &oz)
{
  ox = __fadd_rn(__fmul_rn(T[0], x),
                 __fadd_rn(__fmul_rn(T[1], y), __fadd_rn(__fmul_rn(T[2], z), T[3])));
  oy = __fadd_rn(__fmul_rn(T[4], x),
                 __fadd_rn(__fmul_rn(T[5], y), __fadd_rn(__fmul_rn(T[6], z), T[7])));
  oz = __fadd_rn(__fmul_rn(T[8], x),
                 __fadd_rn(__fmul_rn(T[9], y), __fadd_rn(__fmul_rn(T[10], z), T[11])));
}

// R3/R4 projection, ref: src/cloud_detections.cpp:268-273 and :19-24.
//   img = K * (double)(x,y,z), accumulated (a0*b0 + a1*b1) + a2*b2;  u = (float)(img.x/img.z)
// Canonical K (zero skew, bottom row 0 0 1, float-valued entries — the only K the
// reference can build, ref: src/object_detection.cpp:241-247 fed from float CAMParams):
// fx*X and cx*Z are exact in double, the 0*Y term is a signed zero, so
// img.x = RN(fx*X + cx*Z) = fma(fx, X, cx*Z) unless the sum is zero (sign-of-zero games),
// in which case the generic path is taken.  img.z = Z exactly for finite X, Y.
__device__ __forceinline__ void project_point(const CamDev &c, float xf, float yf, float zf,
                                              float &u, float &v)
{
  const double X = (double)xf, Y = (double)yf, Z = (double)zf;
  double ix, iy, iz;
  bool generic = !c.canon;
  if (!generic) {
    ix = __fma_rn(c.K[0], X, __dmul_rn(c.K[2], Z));
    iy = __fma_rn(c.K[4], Y, __dmul_rn(c.K[5], Z));
    iz = Z;
    generic = (ix == 0.0) || (iy == 0.0) || !finitef(xf) || !finitef(yf);
  }
  if (generic) {
    ix = __dadd_rn(__dadd_rn(__dmul_rn(c.K[0], X), __dmul_rn(c.K[1], Y)), __dmul_rn(c.K[2], Z));
    iy = __dadd_rn(__dadd_rn(__dmul_rn(c.K[3], X), __dmul_rn(c.K[4], Y)), __dmul_rn(c.K[5], Z));
    iz = __dadd_rn(__dadd_rn(__dmul_rn(c.K[6], X), __dmul_rn(c.K[7], Y)), __dmul_rn(c.K[8], Z));
  }
  u = __double2float_rn(__ddiv_rn(ix, iz));
  v = __double2float_rn(__ddiv_rn(iy, iz));
}

// grid_map getIndexFromPosition restatement (oracle gvo_index_coord / gvo_grid_get_index).
__device__ __forceinline__ double index_coord(double p, double half, double pos, double res)
{
  return -__ddiv_rn(__dsub_rn(__dsub_rn(p, half), pos), res);
}

__device__ __forceinline__ bool grid_get_index(const GridGeom &g, double px, double py, int &ix,
                                               int &iy)
{
  const double qx = -__dsub_rn(__dsub_rn(px, g.pos_x), g.half_x);
  const double qy = -__dsub_rn(__dsub_rn(py, g.pos_y), g.half_y);
  if (!(qx >= 0.0 && qy >= 0.0 && qx < g.len_x && qy < g.len_y)) return false;
  const double ax = index_coord(px, g.half_x, g.pos_x, g.res);
  const double ay = index_coord(py, g.half_y, g.pos_y, g.res);
  const int i = __double2int_rz(ax);
  const int j = __double2int_rz(ay);
  if (!(i >= 0 && j >= 0 && i < g.nx && j < g.ny)) return false;
  ix = i;
  iy = j;
  return true;
}

// Certified fast version of grid_get_index for float positions (X1).  The contract is
//   inside  <=>  0 <= -((p - pos) - half) < len  (per axis, double)      [checkIfPositionWithinMap]
//   index   =   trunc(-(((p - half) - pos) / res))                       [getIndexFromPosition]
// Fast path: a' = (c0 - p) * (2^16/res) in double with c0 = RN(half + pos); k = trunc(a') is a
// 16.16 fixed-point image of the exact index coordinate a (the conversion saturates, NaN -> 0).
// a' / 2^16 differs from the reference's rounded a (and q/res from a) by at most
// 8 * 2^-53 * (|p| + half + |pos|) / res, which the host guarantees is below 2^-20 cells
// (fast_index_ok, which also requires size <= 16384 so k fits), so when the 16-bit fraction of
// k lies in [8, 2^16 - 8] both formulas truncate to the same cell, and when
// 8 <= k < (size << 16) - 8 the point is certainly inside.  Anything else (a point within
// 1.2e-4 cells of a cell or map boundary, NaN, a huge coordinate) takes the exact path.
// Returns true with (ix,iy) when inside; false when outside.  Bit-identical by construction.
template <bool COMMON>
__device__ __forceinline__ bool grid_get_index_cert(const BinDev &b, float pxf, float pyf, int &ix,
                                                    int &iy)
{
  if (COMMON || b.fast_index_ok) {
    const int kx = __double2int_rz(__dmul_rn(__dsub_rn(b.c0xd, (double)pxf), b.mres));
    const int ky = __double2int_rz(__dmul_rn(__dsub_rn(b.c0yd, (double)pyf), b.mres));
    const unsigned fx = (unsigned)kx & 0xffffu, fy = (unsigned)ky & 0xffffu;
    const bool frac_ok = (fx - 8u <= 0xffffu - 16u) && (fy - 8u <= 0xffffu - 16u);
    if (frac_ok && kx >= 8 && ky >= 8 && kx < b.klim_x - 8 && ky < b.klim_y - 8) {
      ix = kx >> 16;
      iy = ky >> 16;
      return true;
    }
    // certainly outside: more than 8/2^16 cells beyond an edge on some axis
    if (kx < -8 || ky < -8 || kx >= b.klim_x + 8 || ky >= b.klim_y + 8) return false;
  }
  return grid_get_index(b.g, (double)pxf, (double)pyf, ix, iy);
}

__device__ __forceinline__ int clamp_cell(float c, int n)
{
  return c >= 0.0f ? (c >= (float)n ? n - 1 : (int)c) : 0;  // NaN -> 0
}

// oracle gvo_clip_end: all-float parametric clip in index space, one rounded op per step.
__device__ __forceinline__ void clip_end(const BinDev &b, float pxf, float pyf, int &ex, int &ey)
{
  const float nxf = (float)b.g.nx, nyf = (float)b.g.ny;
  const float eax = __fmul_rn(__fsub_rn(b.c0xf, pxf), b.inv_resf);
  const float eay = __fmul_rn(__fsub_rn(b.c0yf, pyf), b.inv_resf);
  const float dax = __fsub_rn(eax, b.oaxf), day = __fsub_rn(eay, b.oayf);
  float t = 1.0f;
  if (eax < 0.0f) {
    const float tt = __fdiv_rn(__fsub_rn(0.0f, b.oaxf), dax);
    if (tt < t) t = tt;
  } else if (eax >= nxf) {
    const float tt = __fdiv_rn(__fsub_rn(nxf, b.oaxf), dax);
    if (tt < t) t = tt;
  }
  if (eay < 0.0f) {
    const float tt = __fdiv_rn(__fsub_rn(0.0f, b.oayf), day);
    if (tt < t) t = tt;
  } else if (eay >= nyf) {
    const float tt = __fdiv_rn(__fsub_rn(nyf, b.oayf), day);
    if (tt < t) t = tt;
  }
  ex = clamp_cell(__fadd_rn(b.oaxf, __fmul_rn(t, dax)), b.g.nx);
  ey = clamp_cell(__fadd_rn(b.oayf, __fmul_rn(t, day)), b.g.ny);
}

// X1: one beam -> (end cell, flags).  Mirrors the per-point body of oracle gvo_accumulate.
// COMMON: t_small and fast_index_ok are known true at compile time (host-checked).
template <bool COMMON>
__device__ __forceinline__ void bin_point(const BinDev &b, float x, float y, float z, int label,
                                          int &cell, unsigned &flags)
{
  cell = -1;
  flags = 0;
  float bx, by, bz;
  if ((COMMON || b.t_small) && small3(x, y, z)) {
    se3(b.T, x, y, z, bx, by, bz);  // cannot overflow: finite
  } else {
    if (!finite3(x, y, z)) return;
    se3(b.T, x, y, z, bx, by, bz);
    if (!finite3(bx, by, bz)) return;
  }
  bool hit_ok = true;
  flags = 1u;  // GV_F_VALID
  if (b.cap) {
    const float dx = __fsub_rn(bx, b.oxf), dy = __fsub_rn(by, b.oyf);
    const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (r2 > b.rmax2f) {
      const float sf = __fdiv_rn(b.rmaxf, __fsqrt_rn(r2));
      bx = __fadd_rn(b.oxf, __fmul_rn(sf, dx));
      by = __fadd_rn(b.oyf, __fmul_rn(sf, dy));
      hit_ok = false;
      flags |= 8u;  // GV_F_RANGECAP
    }
  }
  int ex, ey;
  if (!grid_get_index_cert<COMMON>(b, bx, by, ex, ey)) {
    clip_end(b, bx, by, ex, ey);
    hit_ok = false;
    flags |= 4u;  // GV_F_CLIPPED
  }
  // branch-free hit predicate (z gate, labelled-only mode)
  hit_ok = hit_ok & (!b.use_z_gate | ((bz >= b.z_min) & (bz <= b.z_max))) &
           ((b.occ_mode != 1) | (label >= 0));
  flags |= hit_ok ? 2u : 0u;  // GV_F_HIT
  cell = ex + ey * b.g.nx;
}

// ----------------------------------------------------------------------------------
// K1/K2: fused transform + project + box-label (+ base transform + cell + bin).
// One CTA = one tile of tile_pts consecutive points of one frame, 256 points per pass (one per
// thread, coalesced 128-byte rows per warp, next pass prefetched).  The frame's
// detection boxes and its image-tile box masks are staged once per CTA in shared memory.
// No barrier after the staging one: warps retire independently (projection and box tests
// make per-warp work very uneven).
// ----------------------------------------------------------------------------------
// Certified float version of R3's decisions for one camera-frame point (canonical K only).
// q = (fx*X + cx*Z)/Z evaluated in binary32 (one multiply, one FMA, one 2-ulp division) is within
//   |q - u_ref| <= 1.01 * 2^-24 * (|cx| + 6|u|)
// of the reference's u_ref = (float)((double)(fx*X + cx*Z) / Z): the product cx*Z and the FMA
// each contribute one half-ulp relative error (the first scaled by |cx*Z|/|Z| = |cx|, the second
// by |u|), the division 4 * 2^-24 |u|, and the reference's own narrowing to float 2^-24 |u|.
// E(q) = e6*|q| + e0 with e6 = 6 * 2^-22, e0 = 2^-22 (|cx| + 1) is four times that bound.  A
// decision (u < 0, u >= W, u >= x_min, u <= x_max, which 32-px tile) is taken here only if it
// is the same for every value in [q - E, q + E]; otherwise the function returns false and the
// caller evaluates the exact FP64 path for this point.  Returns true with the label set.
__device__ __forceinline__ bool fuse_point_fast(const CamDev &cam, float X, float Y, float Z,
                                                const float4 *s_box, int nb,
                                                const unsigned long long *mset, int shift, int mtx,
                                                int mwords, int &lab)
{
  lab = -1;
  const float q = __fdividef(fmaf(cam.fxf, X, cam.cxf * Z), Z);
  const float r = __fdividef(fmaf(cam.fyf, Y, cam.cyf * Z), Z);
  const float Eu = fmaf(fabsf(q), cam.e6, cam.e0u), Ev = fmaf(fabsf(r), cam.e6, cam.e0v);
  const float ql = q - Eu, qh = q + Eu, rl = r - Ev, rh = r + Ev;
  // ref: src/cloud_detections.cpp:276 image test
  if (!(ql >= 0.0f && qh < cam.Wf && rl >= 0.0f && rh < cam.Hf)) {
    // certainly outside the image -> no label; else too close to an image edge to call
    return qh < 0.0f || ql >= cam.Wf || rh < 0.0f || rl >= cam.Hf;
  }
  if (nb <= 0) return true;
  const int tu = (int)ql >> shift, tv = (int)rl >> shift;
  if (tu != ((int)qh >> shift) || tv != ((int)rh >> shift)) return false;  // straddles a tile edge
  const unsigned long long *mrow = mset + (tv * mtx + tu) * mwords;
#pragma unroll 1
  for (int w = 0; w < mwords; ++w) {
    unsigned long long m = mrow[w];
    while (m) {
      const int b = w * 64 + __ffsll((long long)m) - 1;
      m &= m - 1;
      const float4 B = s_box[b];
      if (ql >= B.x && qh <= B.z && rl >= B.y && rh <= B.w) {  // certainly inside: first match
        lab = b;
        return true;
      }
      if (!(qh < B.x || ql > B.z || rh < B.y || rl > B.w)) return false;  // not certainly outside
    }
  }
  return true;
}

// R1 + R3 for one point against one camera: returns the label (index local to the camera's /
// frame's box list, -1 = none); pix/u/v are the parity outputs (EXACT_UV: always the FP64 path).
// COMMON: has_T, t_small, canon and mwords == 1 are known true at compile time (host-checked).
template <bool EXACT_UV, bool COMMON>
__device__ __forceinline__ int fuse_point(const CamDev &cam, int is_dense, float px, float py,
                                          float pz, const float4 *s_box, int nb,
                                          const unsigned long long *mset, int shift, int mtx,
                                          int mwords, int &pix, float &u, float &v)
{
  int lab = -1;
  pix = -1;
  u = v = __int_as_float(0x7fc00000);
  float X = px, Y = py, Z = pz;
  bool ok;
  const bool small = (COMMON || cam.t_small) && small3(px, py, pz);
  if (small) {
    // finite in, |T| < 1e6: finite out (< 3.1e15), so ref :264 reduces to the depth test
    if (COMMON || cam.has_T) se3(cam.T, px, py, pz, X, Y, Z);
    ok = Z > 0.001f;
  } else {
    // R1 (pcl::transformPointCloud: non-finite points pass through when !is_dense)
    if (cam.has_T && (is_dense || finite3(X, Y, Z))) se3(cam.T, px, py, pz, X, Y, Z);
    // ref: src/cloud_detections.cpp:264
    ok = finite3(X, Y, Z) && !(Z <= 0.001f);
  }
  if (ok) {
    if (!EXACT_UV && (COMMON || cam.canon) && small) {
      // magnitudes are inside what the error analysis (and __fdividef) covers
      if (fuse_point_fast(cam, X, Y, Z, s_box, nb, mset, shift, mtx, COMMON ? 1 : mwords, lab))
        return lab;
      lab = -1;
    }
    project_point(cam, X, Y, Z, u, v);
    // ref: :276  (float vs int -> the int is converted to float)
    if (!(u < 0.0f || u >= cam.Wf || v < 0.0f || v >= cam.Hf)) {
      const int iu = (int)u, iv = (int)v;
      pix = iv * cam.W + iu;
      // ref: :280-288 first box in list order wins, inclusive bounds.  Only boxes whose
      // rectangle overlaps this point's image tile can contain it (k_box_masks); they are
      // visited in ascending index order, so the first hit is the reference's.  The double
      // bounds were rounded to float on the device (k_round_boxes) so these float compares
      // decide exactly like the reference's float-vs-double compares.
      if (nb > 0) {
        const unsigned long long *mrow = mset + ((iv >> shift) * mtx + (iu >> shift)) * mwords;
#pragma unroll 1
        for (int w = 0; w < mwords && lab < 0; ++w) {
          unsigned long long m = mrow[w];
          while (m) {
            const int b = w * 64 + __ffsll((long long)m) - 1;
            m &= m - 1;
            const float4 B = s_box[b];
            if (u >= B.x && u <= B.z && v >= B.y && v <= B.w) {
              lab = b;
              break;
            }
          }
        }
      }
    }
  }
  return lab;
}

// MULTI = false: one camera (index 0, every parameter a compile-time constant-bank operand);
// MULTI = true: loop over a.ncam cameras of a rig (BASELINE config 4).
// EXACT_UV = true: a parity output (u,v / pixel / end cell / beam flags) is requested; the
// projection then always runs in FP64 and the nullable parity pointers are honoured.  The
// throughput instantiations (EXACT_UV = false) write labels and bin beams only.
// COMMON = true: the usual configuration (extrinsic present, sane magnitudes, canonical K,
// certified index usable, <= 64 boxes per set) is fixed at compile time, which removes a dozen
// uniform-flag branches from the hot loop; the host selects it, anything else runs COMMON = false.
template <bool FUSE, bool BIN, bool MULTI, bool EXACT_UV, bool COMMON>
__global__ void __launch_bounds__(kThreads) k_points(const __grid_constant__ PointArgs a)
{
  extern __shared__ float4 s_dyn[];
  float4 *s_box = s_dyn;
  const unsigned long long *s_mask = reinterpret_cast<const unsigned long long *>(s_dyn + a.smem_boxes);

  unsigned long long start, end;
  int bb = 0, be = 0, set0 = 0;
  const unsigned tile = blockIdx.x + a.tile0;
  if (a.nframes > 0) {
    start = a.tile_start[tile];
    end = a.tile_end[tile];
    const int4 br = a.tile_boxes[tile];
    bb = br.x;
    be = br.y;
    set0 = br.z;
  } else {
    start = (unsigned long long)tile * (unsigned)a.tile_pts;
    end = a.n;
  }
  if (end > start + (unsigned)a.tile_pts) end = start + (unsigned)a.tile_pts;
  if (end <= start) return;
  const unsigned cnt = (unsigned)(end - start);  // <= tile_pts: 32-bit tile-local indexing below

  if (FUSE) {
    // stage boxes + tile masks: batch mode -> this frame's set; otherwise every camera's set
    const int b0 = a.nframes > 0 ? bb : 0;
    const int b1 = a.nframes > 0 ? be : a.cam[a.ncam - 1].box_end;
    for (int i = b0 + threadIdx.x; i < b1; i += kThreads) s_box[i - b0] = a.boxes[i];
    const int nm = (a.nframes > 0 ? 1 : a.ncam) * a.mask_stride;
    const unsigned long long *gm = a.masks + (size_t)set0 * a.mask_stride;
    unsigned long long *sm = const_cast<unsigned long long *>(s_mask);
    for (int i = threadIdx.x; i < nm; i += kThreads) sm[i] = gm[i];
    __syncthreads();
  }

  // per-thread pointers that advance by one pass (256 points) per iteration
  const unsigned long long t0 = start + threadIdx.x;
  const float *xp = a.x + t0, *yp = a.y + t0, *zp = a.z + t0;
  int16_t *lab_p = a.labels ? a.labels + t0 : nullptr;
  int32_t *pix_p = (EXACT_UV && a.pix) ? a.pix + t0 : nullptr;
  float *uv_p = (EXACT_UV && a.uv) ? a.uv + t0 : nullptr;
  const int16_t *labin_p = (!FUSE && a.labels_in) ? a.labels_in + t0 : nullptr;
  int32_t *cell_p = (EXACT_UV && a.cell_out) ? a.cell_out + t0 : nullptr;
  uint8_t *flag_p = (EXACT_UV && a.flags_out) ? a.flags_out + t0 : nullptr;
  if (BIN && !a.bin.origin_ok) {
    // sensor outside the map: every beam is dropped (parity outputs say so)
    if (EXACT_UV)
      for (unsigned k = threadIdx.x; k < cnt; k += kThreads) {
        if (cell_p) cell_p[k - threadIdx.x] = -1;
        if (flag_p) flag_p[k - threadIdx.x] = 0;
      }
    if (!FUSE) return;
  }
  const bool do_bin = BIN && a.bin.origin_ok;
  const int nb0 = a.nframes > 0 ? be - bb : a.cam[0].box_end - a.cam[0].box_begin;
  const float4 *box0 = s_box + (a.nframes > 0 ? 0 : a.cam[0].box_begin);

  // One point per thread per pass, next pass prefetched.  (A 4-points-per-thread version with
  // 128-bit loads ran out of instruction cache: 3.4k SASS instructions, 41% no-instruction
  // stalls in ncu; the path is issue-bound, not load-bound, so scalar coalesced loads win.)
  int left = (int)cnt - (int)threadIdx.x;  // points this thread still has to do: every kThreads-th
  float nx = 0.0f, ny = 0.0f, nz = 0.0f;
  if (left > 0) {
    nx = __ldg(xp);
    ny = __ldg(yp);
    nz = __ldg(zp);
  }
#pragma unroll 1
  while (left > 0) {
    const float px = nx, py = ny, pz = nz;
    left -= kThreads;
    if (left > 0) {
      nx = __ldg(xp + kThreads);
      ny = __ldg(yp + kThreads);
      nz = __ldg(zp + kThreads);
    }
    xp += kThreads;
    yp += kThreads;
    zp += kThreads;
    int lab0 = -1;

    if (FUSE) {
      if (!MULTI) {
        int pix;
        float u, v;
        lab0 = fuse_point<EXACT_UV, COMMON>(a.cam[0], a.is_dense, px, py, pz, box0, nb0, s_mask,
                                            a.mask_shift[0], a.mask_tx[0], a.mask_words, pix, u, v);
        if (lab_p) {
          *lab_p = (int16_t)lab0;
          lab_p += kThreads;
        }
        if (EXACT_UV) {
          if (pix_p) {
            *pix_p = pix;
            pix_p += kThreads;
          }
          if (uv_p) {
            uv_p[0] = u;
            uv_p[a.n] = v;
            uv_p += kThreads;
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < a.ncam; ++c) {
          const CamDev &cam = a.cam[c];
          int pix;
          float u, v;
          const int lab = fuse_point<EXACT_UV, false>(cam, a.is_dense, px, py, pz, s_box + cam.box_begin,
                                                      cam.box_end - cam.box_begin,
                                                      s_mask + c * a.mask_stride, a.mask_shift[c],
                                                      a.mask_tx[c], a.mask_words, pix, u, v);
          if (c == 0) lab0 = lab;
          const unsigned long long plane = (unsigned long long)c * a.n;
          if (lab_p) lab_p[plane] = (int16_t)lab;
          if (EXACT_UV) {
            if (pix_p) pix_p[plane] = pix;
            if (uv_p) {
              uv_p[2 * plane] = u;
              uv_p[2 * plane + a.n] = v;
            }
          }
        }
        if (lab_p) lab_p += kThreads;
        if (EXACT_UV) {
          if (pix_p) pix_p += kThreads;
          if (uv_p) uv_p += kThreads;
        }
      }
    }

    if (BIN && do_bin) {
      int label = lab0;
      if (!FUSE && labin_p) {
        label = *labin_p;
        labin_p += kThreads;
      }
      int cell;
      unsigned flags;
      bin_point<COMMON>(a.bin, px, py, pz, label, cell, flags);
      // one 64-bit RED per beam: low word counts beams ending in the cell, high word hits
      if (cell >= 0) atomicAdd(a.ends + cell, (flags & 2u) ? 0x100000001ull : 1ull);
      if (EXACT_UV) {
        if (cell_p) {
          *cell_p = cell;
          cell_p += kThreads;
        }
        if (flag_p) {
          *flag_p = (uint8_t)flags;
          flag_p += kThreads;
        }
      }
    }
  }
}

struct BoxRaw {  // ref: include/grid_vision/object_detection.hpp:27-32 (40 bytes)
  double x_min, y_min, x_max, y_max;
  float confidence;
  int label;
};

// double bounds -> float so that float compares decide exactly like the reference's float-vs-double
// compares: x_min, y_min round up, x_max, y_max round down (k_round_boxes, k_box_masks)
__device__ __forceinline__ float4 round_box(const BoxRaw &b)
{
  return make_float4(__double2float_ru(b.x_min), __double2float_ru(b.y_min), __double2float_rd(b.x_max),
                     __double2float_rd(b.y_max));
}

// Image-tile prefilter for the box test.  One CTA per box set (a camera, or a frame in batch
// mode); bit b of masks[set][tile*words + b/64] is set iff box b (index local to the set)
// overlaps the tile's pixel square [tx*S, (tx+1)*S) x [ty*S, (ty+1)*S), S = 2^shift.
// A point with float pixel (u,v) lies in tile ((int)u >> shift, (int)v >> shift), and a box
// contains it only if x_min <= u <= x_max, so boxes outside the mask can never match; the
// test is conservative, never lossy.  NaN bounds compare false everywhere -> bit clear.
// rev32: bit (31 - b%32) of 32-bit half (b%64)/32 instead of bit b%64, so that a count-leading-
// zeros walk visits boxes in ascending order (k_points_fast).
__global__ void __launch_bounds__(kThreads) k_box_masks(const float4 *__restrict__ boxes,
                                                        const int *__restrict__ set_offsets,
                                                        int shift, int tiles_x, int tiles_y,
                                                        int words, int stride, int rev32,
                                                        unsigned long long *__restrict__ masks,
                                                        const BoxRaw *__restrict__ raw, float4 *boxes_out)
{
  const int set = blockIdx.x;
  const int b0 = set_offsets[set], b1 = set_offsets[set + 1];
  const float S = (float)(1 << shift), D = rev32 ? 1.0f : 0.0f;
  // the usual set (<= 256 boxes) is staged in shared memory once: every tile tests every box.
  // raw != NULL (sets of <= 256 boxes only, host-checked): this CTA also does k_round_boxes' job
  // for its set, one launch less per batch.
  __shared__ float4 s_b[kThreads];
  const bool staged = b1 - b0 <= kThreads;
  if (staged && b0 + (int)threadIdx.x < b1) {
    if (raw) {
      const float4 B = round_box(raw[b0 + threadIdx.x]);
      s_b[threadIdx.x] = B;
      boxes_out[b0 + threadIdx.x] = B;
    } else {
      s_b[threadIdx.x] = boxes[b0 + threadIdx.x];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < tiles_x * tiles_y * words; t += kThreads) {
    const int w = t % words, tile = t / words;
    const float x0 = (float)(tile % tiles_x) * S, y0 = (float)(tile / tiles_x) * S;
    unsigned long long m = 0ull;
    const int bw0 = b0 + w * 64;
    const int bw1 = bw0 + 64 < b1 ? bw0 + 64 : b1;
    for (int b = bw0; b < bw1; ++b) {
      const float4 B = staged ? s_b[b - b0] : boxes[b];
      const int bit = rev32 ? ((b - bw0) & 32) + 31 - ((b - bw0) & 31) : b - bw0;
      // rev32 (the certified kernels): the tile is widened by D = 1 px, because k_points_pair takes
      // the tile from its approximate pixel (within 1e-3 px of the reference's) without a straddle test
      if (B.z >= x0 - D && B.x < x0 + S + D && B.w >= y0 - D && B.y < y0 + S + D) m |= 1ull << bit;
    }
    masks[(size_t)set * stride + t] = m;
  }
}

// N4: pcl::PointXYZI 32-byte AoS -> SoA planes (ref: pcl::fromROSMsg output consumed at
// src/grid_vision_node.cpp:103-106,157)
__global__ void __launch_bounds__(kThreads) k_aos32_to_soa(const float4 *__restrict__ pts,
                                                           unsigned long long n,
                                                           float *__restrict__ x,
                                                           float *__restrict__ y,
                                                           float *__restrict__ z)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(pts + 2 * i);  // first 16 bytes of the record: x, y, z, w
  x[i] = p.x;
  y[i] = p.y;
  z[i] = p.z;
}

// R1 alone: GridVision::transformLidarToCamera (ref: src/grid_vision_node.cpp:280-307)
__global__ void __launch_bounds__(kThreads) k_transform(const float *__restrict__ x,
                                                        const float *__restrict__ y,
                                                        const float *__restrict__ z,
                                                        unsigned long long n, int is_dense,
                                                        const __grid_constant__ CamDev cam,
                                                        float *__restrict__ ox,
                                                        float *__restrict__ oy,
                                                        float *__restrict__ oz)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float X = x[i], Y = y[i], Z = z[i];
  if (is_dense || finite3(X, Y, Z)) se3(cam.T, x[i], y[i], z[i], X, Y, Z);
  ox[i] = X;
  oy[i] = Y;
  oz[i] = Z;
}

// ----------------------------------------------------------------------------------
// R4: buildKDTree projection loop (ref: src/cloud_detections.cpp:13-33) with an
// order-preserving compaction: count pass -> exclusive scan of CTA counts -> scatter pass.
// ----------------------------------------------------------------------------------
__device__ __forceinline__ bool kd_keep(const CamDev &cam, int is_dense, float x, float y, float z,
                                        float &X, float &Y, float &Z)
{
  X = x; Y = y; Z = z;
  if (cam.has_T && (is_dense || finite3(x, y, z))) se3(cam.T, x, y, z, X, Y, Z);
  return !(Z <= 0.0f);  // ref: :16  if(p.z <= 0) continue;   (NaN depth passes)
}

__global__ void __launch_bounds__(kThreads) k_kdtree_count(const float *__restrict__ x,
                                                           const float *__restrict__ y,
                                                           const float *__restrict__ z,
                                                           unsigned long long n, int is_dense,
                                                           const __grid_constant__ CamDev cam,
                                                           unsigned *__restrict__ cta_count)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  float X, Y, Z;
  const bool keep = i < n && kd_keep(cam, is_dense, x[i], y[i], z[i], X, Y, Z);
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) cta_count[blockIdx.x] = (unsigned)c;
}

__global__ void __launch_bounds__(kThreads) k_kdtree_scatter(const float *__restrict__ x,
                                                             const float *__restrict__ y,
                                                             const float *__restrict__ z,
                                                             unsigned long long n, int is_dense,
                                                             const __grid_constant__ CamDev cam,
                                                             const unsigned *__restrict__ cta_base,
                                                             float *__restrict__ uvz)
{
  __shared__ unsigned s_warp[kThreads / 32];
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  float X, Y, Z;
  const bool keep = i < n && kd_keep(cam, is_dense, x[i], y[i], z[i], X, Y, Z);
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  unsigned base = cta_base[blockIdx.x];
  for (unsigned w = 0; w < warp; ++w) base += s_warp[w];
  if (keep) {
    const unsigned long long o = (unsigned long long)base + __popc(m & ((1u << lane) - 1u));
    float u, v;
    project_point(cam, X, Y, Z, u, v);
    uvz[3 * o + 0] = u;
    uvz[3 * o + 1] = v;
    uvz[3 * o + 2] = Z;
  }
}

// in-place exclusive scan of up to 1024*blockDim chunks: level kernel (CTA scans 1024
// values, emits its total) — composed recursively on the host (scan_u32)
__global__ void __launch_bounds__(1024) k_scan_block(unsigned *data, unsigned long long n,
                                                     unsigned *block_sums)
{
  __shared__ unsigned s_w[32];
  const unsigned long long i = (unsigned long long)blockIdx.x * 1024 + threadIdx.x;
  const unsigned v = i < n ? data[i] : 0u;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned w = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= (unsigned)o) w += t;
    }
    s_w[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const unsigned warp_base = warp ? s_w[warp - 1] : 0u;
  if (i < n) data[i] = warp_base + incl - v;
  if (threadIdx.x == 1023 && block_sums) block_sums[blockIdx.x] = warp_base + incl;
}

__global__ void __launch_bounds__(1024) k_scan_add(unsigned *data, unsigned long long n,
                                                   const unsigned *block_base)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * 1024 + threadIdx.x;
  if (i < n) data[i] += block_base[blockIdx.x];
}

// ----------------------------------------------------------------------------------
// stable partition of point indices by label (R3's per-box push_back order,
// ref: src/cloud_detections.cpp:280-297): histogram -> scan -> ranked scatter
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_label_hist(const int16_t *__restrict__ labels,
                                                         unsigned long long n, int nboxes,
                                                         unsigned nblocks,
                                                         unsigned *__restrict__ hist)
{
  extern __shared__ unsigned s_h[];
  for (int b = threadIdx.x; b < nboxes; b += kThreads) s_h[b] = 0;
  __syncthreads();
  const unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
  if (i < n) {
    const int l = labels[i];
    if (l >= 0 && l < nboxes) atomicAdd(&s_h[l], 1u);
  }
  __syncthreads();
  // label-major layout so one exclusive scan yields box-concatenated output offsets
  for (int b = threadIdx.x; b < nboxes; b += kThreads)
    hist[(unsigned long long)b * nblocks + blockIdx.x] = s_h[b];
}

__global__ void __launch_bounds__(kThreads) k_label_scatter(const int16_t *__restrict__ labels,
                                                            unsigned long long n, int nboxes,
                                                            unsigned nblocks,
                                                            const unsigned *__restrict__ base,
                                                            unsigned *__restrict__ indices)
{
  extern __shared__ unsigned s_cnt[];  // [8 warps][nboxes] -> exclusive over warps
  constexpr int kWarps = kThreads / 32;
  for (int b = threadIdx.x; b < nboxes * kWarps; b += kThreads) s_cnt[b] = 0;
  __syncthreads();
  const unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int l = -1;
  if (i < n) {
    l = labels[i];
    if (l >= nboxes) l = -1;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, l);
  const unsigned rank = __popc(peers & ((1u << lane) - 1u));
  if (l >= 0 && rank == 0) s_cnt[warp * nboxes + l] = __popc(peers);
  __syncthreads();
  for (int b = threadIdx.x; b < nboxes; b += kThreads) {
    unsigned run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const unsigned c = s_cnt[w * nboxes + b];
      s_cnt[w * nboxes + b] = run;
      run += c;
    }
  }
  __syncthreads();
  if (l >= 0) {
    const unsigned o =
      base[(unsigned long long)l * nblocks + blockIdx.x] + s_cnt[warp * nboxes + l] + rank;
    indices[o] = (unsigned)i;
  }
}

// ----------------------------------------------------------------------------------
// K3: de-duplicated raycast as a distance-ordered sweep.
// All beams binned since the last flush share one start cell s, and a Bresenham line is a pure
// function of (start, end), so the traversed-cell multiset of the whole batch is
//   sum over distinct end cells e of  w_e * line(s, e)           (exact: integer sums commute).
// End cells are organised by (direction, major-axis distance D from s): pieces of the column
// x = sx +- D (x-major lines, |dy| <= D) and of the row y = sy +- D (y-major lines, |dx| < D) —
// the grid_map LineIterator's own case split.  32 end cells of one such piece, taken in order,
// give 32 lines that one warp can walk in lock step:
//   * all 32 lines have exactly D + 1 cells: no divergence, no tail;
//   * at step k every lane is at the same major coordinate and the minor coordinates are
//     monotone in the lane index, so equal cells form contiguous lane runs: one shuffle + one
//     ballot find the runs, and the run's weight is a difference of the warp prefix sum of w_e
//     (computed once, the weights do not change along the walk): ONE RED per distinct cell.
// The end cells themselves are settled too (hit += high word, miss += low - high) and the ends
// plane is cleared.  Multi-GPU: span i belongs to rank (i - i / world) % world (the table is identical on
// every rank).  LineIterator restatement: oracle gvo_line_init / gvo_line_step.
// ----------------------------------------------------------------------------------
// Peer-memory views of one plane on every rank (cudaIpc-mapped, NVLink P2P); p[rank] is local.
constexpr int kMaxPeers = 16;
template <typename T>
struct Peers {
  T *p[kMaxPeers];
};

struct SweepEntry {
  int dir;  // 0: x = sx + D, 1: x = sx - D (x-major); 2: y = sy + D, 3: y = sy - D (y-major); 4: origin
  int D;
  int m0, m1;  // inclusive range of the minor coordinate (absolute index), clipped to the map
};

// Scheduling: the span list is sorted by decreasing D; rank r owns one span of every round of `world`
// spans (rotating, see k_sweep_compact) and
// warps pull work through an atomic counter, longest first.  (Measured: static round-robin is
// 20 % slower — most spans carry no beams and the busy ones cluster — and cutting lines into
// segments for more parallelism costs more in set-up than it gains.)
__device__ __forceinline__ int sweep_find_entry(const unsigned *__restrict__ prefix, int n_entries,
                                                unsigned unit)
{
  int lo = 0, hi = n_entries;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (prefix[mid] <= unit) lo = mid;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ size_t sweep_end_cell(const SweepEntry &E, int mc, int sx, int sy, int nx,
                                                 int &ex, int &ey)
{
  if (E.dir == 4) { ex = sx; ey = sy; }
  else if (E.dir < 2) { ex = E.dir == 0 ? sx + E.D : sx - E.D; ey = mc; }
  else { ey = E.dir == 2 ? sy + E.D : sy - E.D; ex = mc; }
  return (size_t)ex + (size_t)ey * (size_t)nx;
}

// Two kernels.  k_sweep_compact: item = kSpanCells consecutive cells of one entry; a warp compacts
// the span's non-empty cells, in order, into dense BATCHES of 32 lines appended to a global list
// (a subsequence of a sorted sequence is sorted, so the monotone-run argument holds), settles the
// end cells and clears the plane.  k_sweep_walk: one batch per warp task.  With 13-25 % of the
// cells carrying beams this keeps ~3x more lanes busy than walking 32 raw cells at a time, and
// a dense span (the map border, where every clipped beam ends) is spread over many warps
// instead of being walked batch after batch by one.
// CHUNKS (32-cell chunks per span, template parameter): 32 for maps up to 2048 cells a side; larger,
// sparser maps use 128, so that a span's few end cells do not each drag a mostly-padded batch of 32
// lanes through thousands of steps (BASELINE config 5: 5 lines per 1024-cell span).

// One CTA per span: 8 warps x 4 chunks, four independent loads per thread, a block prefix over
// the per-chunk ballots, one atomic per span to reserve its batches.  (A warp-per-span version
// spent ~0.1 ms in 64 dependent loads per warp whatever the rank's share of the spans.)
// P2P (multi-GPU over peer memory): the entries of a cell are summed over EVERY rank's plane with
// NVLink loads and every rank's cell is cleared with NVLink stores, right here: span ownership
// partitions the cells, so this CTA is the only reader of these cells on any rank.  This is the
// all-reduce of the end-cell planes, restricted to what this rank walks, fused into the sweep.
template <bool P2P, int kSpanChunks>
__global__ void __launch_bounds__(kThreads, kSpanChunks <= 32 ? 4 : 2) k_sweep_compact(
  unsigned long long *__restrict__ ends, int32_t *__restrict__ hit, int32_t *__restrict__ miss,
  const SweepEntry *__restrict__ entries, const unsigned *__restrict__ item_prefix,
  const int *__restrict__ item_entry, int n_entries,
  unsigned n_items, unsigned item0, int reach, int sx, int sy, int nx, unsigned rank, unsigned world, int clear_ends,
  unsigned *__restrict__ counters /* [0] item counter, [1] batch count */,
  int *__restrict__ batch_entry, int *__restrict__ batch_mi, unsigned *__restrict__ batch_w,
  unsigned long long *__restrict__ stats, const __grid_constant__ Peers<unsigned long long> peer_ends)
{
  constexpr int kWarps = kThreads / 32;
  constexpr int kSpanCells = 32 * kSpanChunks;
  constexpr int kPer = kSpanChunks / kWarps;  // chunks per warp
  static_assert(kSpanChunks % kWarps == 0, "span must split evenly over the warps");
  __shared__ unsigned s_base;
  __shared__ int s_cnt[kSpanChunks], s_last[kWarps];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long st_beams = 0, st_logical = 0, st_lines = 0;
  // one span per CTA (grid-stride if the grid is smaller): CTAs are dispatched in index order, so
  // the batch list still fills longest lines first
  for (unsigned s_item = blockIdx.x;; s_item += gridDim.x) {
    __syncthreads();  // s_cnt / s_base of the previous span are no longer read
    // round t hands item t * world + (rank + t) % world to this rank: a rotation, so that a rank
    // does not keep the same (direction, half-column) of every distance (with 8 spans per distance
    // and 8 ranks, plain i % world gave the x-major directions, whose REDs are strided, to four
    // ranks and the cheaper y-major ones to the other four: 0.6 ms of barrier wait per step)
    // item0 / reach (single GPU with a range cap): no beam binned since the last sweep ends farther than
    // `reach` cells from the start cell on either axis, so the items of longer distances (the first
    // item0 of the table) are not launched at all and a span wholly beyond reach on the minor axis
    // returns before it reads anything (a 8192^2 map with 120 m rays: 2/3 of the end-cell plane)
    const unsigned long long item64 = item0 + (unsigned long long)s_item * world + (rank + s_item) % world;
    if (item64 >= n_items) break;
    const unsigned item = (unsigned)item64;
    const int ei = item_entry[item];  // (a binary search of item_prefix here cost ~13 dependent loads per CTA)
    const SweepEntry E = entries[ei];
    const int span0 = E.m0 + (int)(item - item_prefix[ei]) * kSpanCells;
    {
      const int sm = E.dir < 2 ? sy : sx;  // the start cell's minor coordinate
      if (E.dir < 4 && (span0 > sm + reach || span0 + kSpanCells - 1 < sm - reach)) continue;  // block-uniform
    }
    // four independent loads per thread
    unsigned long long e[kPer];
    size_t elin[kPer];
    int mi[kPer];
    int last = -1;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      mi[j] = span0 + ((int)wid * kPer + j) * 32 + (int)lane;
      int ex, ey;
      elin[j] = sweep_end_cell(E, mi[j] <= E.m1 ? mi[j] : E.m1, sx, sy, nx, ex, ey);
      e[j] = 0ull;
      if (mi[j] <= E.m1) {
        if (P2P) {
#pragma unroll
          for (int r = 0; r < kMaxPeers; ++r)
            if (r < (int)world) e[j] += peer_ends.p[r][elin[j]];
        } else {
          e[j] = ends[elin[j]];
        }
      }
    }
    unsigned nz[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      nz[j] = __ballot_sync(0xffffffffu, e[j] != 0ull);
      if (lane == 0) s_cnt[wid * kPer + j] = __popc(nz[j]);
      if (nz[j]) last = span0 + ((int)wid * kPer + j) * 32 + (31 - __clz((int)nz[j]));
    }
    if (lane == 0) s_last[wid] = last;
    __syncthreads();
    int count = 0, before = 0, last_mi = E.m0;
    for (int c = 0; c < kSpanChunks; ++c) {
      if (c == (int)wid * kPer) before = count;
      count += s_cnt[c];
    }
    for (int w = 0; w < kWarps; ++w)
      if (s_last[w] >= 0) last_mi = s_last[w];  // warps hold ascending cells: the last hit wins
    if (count == 0) continue;  // block-uniform
    const int nbatch = E.D > 0 ? (count + 31) >> 5 : 0;  // the start cell itself has no line to walk
    if (threadIdx.x == 0 && nbatch) s_base = atomicAdd(counters + 1, (unsigned)nbatch);
    __syncthreads();
    const size_t base = nbatch ? (size_t)s_base * 32 : 0;
    // settle + clear the end cells, append them in order
    int pos = before;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if (e[j] != 0ull) {
        const unsigned w = (unsigned)(e[j] & 0xffffffffull), hits = (unsigned)(e[j] >> 32);
        if (nbatch) {
          const size_t o = base + (size_t)(pos + __popc(nz[j] & ((1u << lane) - 1u)));
          batch_mi[o] = mi[j];
          batch_w[o] = w;
        }
        if (P2P) {
#pragma unroll
          for (int r = 0; r < kMaxPeers; ++r)
            if (r < (int)world) peer_ends.p[r][elin[j]] = 0ull;  // this thread is the cell's only reader anywhere
        } else if (clear_ends) {
          ends[elin[j]] = 0ull;  // single GPU: this thread is the only reader of the cell
        }
        if (hits) hit[elin[j]] += (int32_t)hits;  // only this thread ever writes hit[elin] in this kernel
        if (w - hits) atomicAdd(miss + elin[j], (int32_t)(w - hits));  // other lines pass through it
        st_beams += w;
        st_logical += (unsigned long long)w * (unsigned)(E.D + 1);
        st_lines += 1;
      }
      pos += __popc(nz[j]);
    }
    if (nbatch) {
      // pad the last batch with weight-0 shadows of the last line, tag every batch with its entry
      const int padded = nbatch * 32;
      for (int i = count + (int)threadIdx.x; i < padded; i += kThreads) {
        batch_mi[base + i] = last_mi;
        batch_w[base + i] = 0u;
      }
      for (int b = (int)threadIdx.x; b < nbatch; b += kThreads) batch_entry[s_base + b] = ei;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    st_beams += __shfl_xor_sync(0xffffffffu, st_beams, o);
    st_logical += __shfl_xor_sync(0xffffffffu, st_logical, o);
    st_lines += __shfl_xor_sync(0xffffffffu, st_lines, o);
  }
  if (lane == 0 && st_lines) {
    atomicAdd(stats + 0, st_beams);
    atomicAdd(stats + 1, st_logical);
    atomicAdd(stats + 3, st_lines);
  }
}

// x-major batches (the 32 lanes share x and differ in y) accumulate into missT, a TRANSPOSED miss
// plane (lin = y + x * ny): in the grid's own layout their REDs are nx * 4 bytes apart, one L2 line
// per lane and step, while the y-major batches' are neighbours.  k_miss_fold adds missT back.
__global__ void __launch_bounds__(kThreads) k_sweep_walk(
  int32_t *__restrict__ miss, int32_t *__restrict__ missT, const SweepEntry *__restrict__ entries,
  unsigned *__restrict__ counters /* [1] batch count, [2] walk counter */,
  const int *__restrict__ batch_entry, const int *__restrict__ batch_mi,
  const unsigned *__restrict__ batch_w, int sx, int sy, int nx, int ny, unsigned nseg, int seg,
  unsigned long long *__restrict__ stats)
{
  // nseg > 1 (small sweeps: one scan): a batch is cut into nseg pieces of `seg` steps, each its own
  // warp task, so that the longest lines are not walked by one warp from end to end (the critical
  // path of a single 1000^2 scan); the state of the LineIterator after k0 steps has a closed form.
  const unsigned lane = threadIdx.x & 31;
  const unsigned nbatch = counters[1];
  unsigned long long st_physical = 0;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(counters + 2, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    const unsigned b = nseg > 1u ? t / nseg : t;
    if (b >= nbatch) break;
    const int k0 = nseg > 1u ? (int)(t - b * nseg) * seg : 0;
    const SweepEntry E = entries[batch_entry[b]];
    if (k0 >= E.D) continue;  // this batch's lines are shorter than the piece's start (warp-uniform)
    const int mi = batch_mi[(size_t)b * 32 + lane];
    const unsigned w = batch_w[(size_t)b * 32 + lane];
    const bool xmajor = E.dir < 2;
    // inclusive warp prefix sum of the weights (loop invariant)
    unsigned P = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned tt = __shfl_up_sync(0xffffffffu, P, o);
      if (lane >= (unsigned)o) P += tt;
    }
    // LineIterator state: den = D, num = D/2, add = minor delta; the major coordinate advances
    // every step, the minor one when num >= den.
    const int den = E.D;
    int num = den / 2;
    const int dminor = mi - (xmajor ? sy : sx);
    const int add = dminor >= 0 ? dminor : -dminor;
    const int smajor = (E.dir == 0 || E.dir == 2) ? 1 : -1;
    int32_t *const plane = xmajor ? missT : miss;
    // both planes put the minor coordinate in the fast index: cell = major * stride + minor
    const int stride_major = xmajor ? ny : nx;
    int lin = (xmajor ? sx : sy) * stride_major + (xmajor ? sy : sx);
    const int dlin_major = smajor * stride_major, dlin_minor = dminor >= 0 ? 1 : -1;
    const unsigned mask_le = 0xffffffffu >> (31u - lane);
    const unsigned Pex = P - w;  // exclusive prefix
    unsigned st_phys32 = 0u;
    int k1 = den;
    if (nseg > 1u) {
      // after k0 steps: num = (den/2 + k0*add) mod den, and the minor coordinate has advanced
      // (den/2 + k0*add) div den times (add <= den: at most one minor step per major step)
      const unsigned acc = (unsigned)num + (unsigned)k0 * (unsigned)add;  // < 2^31: k0, add <= 16384
      const unsigned q = acc / (unsigned)den;
      num = (int)(acc - q * (unsigned)den);
      lin += k0 * dlin_major + (int)q * dlin_minor;
      k1 = k0 + seg < den ? k0 + seg : den;
    }
#pragma unroll 1
    for (int k = k0; k < k1; ++k) {
      // every lane is at the same major coordinate and the minor coordinates are monotone in the
      // lane index, so equal cells form contiguous lane runs: one RED per run (by its last lane),
      // the run's weight a difference of the warp prefix sums
      const int prev = __shfl_up_sync(0xffffffffu, lin, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, (lane == 0u) | (lin != prev));
      const int start = 31 - __clz((int)(heads & mask_le));
      const unsigned sum = P - __shfl_sync(0xffffffffu, Pex, start);
      const bool tail = (((heads >> 1) | 0x80000000u) >> lane) & 1u;
      if (tail && sum) {
        atomicAdd(plane + lin, (int32_t)sum);
        ++st_phys32;
      }
      num += add;
      const bool carry = num >= den;
      num -= carry ? den : 0;
      lin += dlin_major + (carry ? dlin_minor : 0);
    }
    st_physical += st_phys32;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) st_physical += __shfl_xor_sync(0xffffffffu, st_physical, o);
  if (lane == 0 && st_physical) atomicAdd(stats + 2, st_physical);
}

// miss[x + y * nx] += missT[y + x * ny] over the cell rectangle [x0, x1] x [y0, y1] the x-major
// lines can reach, and missT is zero again.  32 x 32 tiles through shared memory: both planes are
// read and written in 128-byte rows.  Runs after k_sweep_walk on the same stream: plain adds.
__global__ void __launch_bounds__(256) k_miss_fold(int32_t *__restrict__ miss, int32_t *__restrict__ missT, int nx,
                                                   int ny, int x0, int y0, int x1, int y1, unsigned *__restrict__ counters)
{
  __shared__ int t[32][33];
  // last kernel of a sweep: leave the sweep's work counters zero for the next one
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0u;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int bx = x0 + (int)blockIdx.x * 32, by = y0 + (int)blockIdx.y * 32;
  bool any = false;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int x = bx + j, y = by + tx;
    int v = 0;
    if (x <= x1 && y <= y1) {
      const size_t o = (size_t)x * (unsigned)ny + (unsigned)y;
      v = missT[o];
      if (v) missT[o] = 0;
    }
    t[j][tx] = v;
    any |= v != 0;
  }
  if (!__syncthreads_or(any)) return;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int x = bx + tx, y = by + j;
    const int v = t[tx][j];
    if (v) miss[(size_t)y * (unsigned)nx + (unsigned)x] += v;
  }
}

// Stream-ordered barrier between the ranks of one node over peer memory: every rank owns a flag
// word per peer (cudaIpc-mapped everywhere).  Rank r publishes `epoch` into slot r of every rank's
// array, then waits until every slot of its own array has reached `epoch`.  One kernel per rank
// per GPU (never two ranks on one GPU: a spinning kernel must not depend on a kernel that may not
// be resident).  A rank that never arrives would hang the others, so the spin gives up after
// ~2 s and raises *timeout_flag, which the host reports at the next synchronising call.
__global__ void k_peer_barrier(const __grid_constant__ Peers<unsigned> flags, unsigned rank, unsigned world,
                               unsigned epoch, unsigned *timeout_flag)
{
  const unsigned r = threadIdx.x;
  if (r >= world) return;
  __threadfence_system();  // this rank's earlier kernels' writes are visible before the arrival is
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[r] + rank), "r"(epoch) : "memory");
  const unsigned *mine = flags.p[rank] + r;
  const long long t0 = clock64();
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if (clock64() - t0 > 4000000000ll) {
      *timeout_flag = 1u;
      break;
    }
    __nanosleep(64);
  }
  __threadfence_system();
}

// ----------------------------------------------------------------------------------
// R8/R9 footprints -> index rectangles (ref: src/occupancy_grid.cpp:72-93,107-138,140-172)
// mode 0: explicit corners n x 8; 1: poses n x (x,y,length,width) (:79-90);
// mode 2: points n x (x,y) + class label (:107-138 with :185-196)
// rect = (min_ix, min_iy, max_ix, max_iy); skipped footprints get an empty rect.
// ----------------------------------------------------------------------------------
__device__ __forceinline__ float estimated_depth(int label)
{
  switch (label) {  // ref: src/occupancy_grid.cpp:185-196
  case 9: return 3.5f;  // VEHICLE
  case 2: return 0.6f;  // PERSON
  case 0: return 2.5f;  // BIKE
  case 1: return 2.5f;  // MOTORBIKE
  default: return -1.0f;
  }
}

__global__ void k_footprint_rects(const double *__restrict__ in, const int32_t *__restrict__ labels,
                                  int n, int mode, const __grid_constant__ GridGeom g,
                                  int4 *__restrict__ rects)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double c[8];
  if (mode == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = in[8 * k + i];
  } else if (mode == 1) {
    const double x = in[4 * k], y = in[4 * k + 1], L = in[4 * k + 2], W = in[4 * k + 3];
    const double hl = __ddiv_rn(L, 2.0), hw = __ddiv_rn(W, 2.0);
    const double xf = __dadd_rn(x, hl), xb = __dsub_rn(x, hl);
    const double yl = __dsub_rn(y, hw), yr = __dadd_rn(y, hw);
    c[0] = xb; c[1] = yl; c[2] = xf; c[3] = yl; c[4] = xf; c[5] = yr; c[6] = xb; c[7] = yr;
  } else {
    const double x = in[2 * k], y = in[2 * k + 1];
    const float d = estimated_depth(labels[k]);
    const double dd = (double)d, dh = (double)__fdiv_rn(d, 2.0f);
    c[0] = __dadd_rn(x, dd); c[1] = __dadd_rn(y, dh);
    c[2] = __dadd_rn(x, dd); c[3] = __dsub_rn(y, dh);
    c[4] = x;                c[5] = __dsub_rn(y, dh);
    c[6] = x;                c[7] = __dadd_rn(y, dh);
  }
  int minx = 0, miny = 0, maxx = -1, maxy = -1;
  bool valid = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ix, iy;
    if (!grid_get_index(g, c[2 * i], c[2 * i + 1], ix, iy)) {  // ref: :152-156
      valid = false;
      break;
    }
    if (i == 0) {
      minx = maxx = ix;
      miny = maxy = iy;
    } else {
      minx = min(minx, ix); miny = min(miny, iy);
      maxx = max(maxx, ix); maxy = max(maxy, iy);
    }
  }
  rects[k] = valid ? make_int4(minx, miny, maxx, maxy) : make_int4(1, 1, 0, 0);
}

// grid_map getIndex for host-supplied positions (parity hook)
__global__ void k_get_index(const double *__restrict__ xy, int n, const __grid_constant__ GridGeom g,
                            int32_t *__restrict__ out)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int ix = -1, iy = -1;
  if (!grid_get_index(g, xy[2 * k], xy[2 * k + 1], ix, iy)) ix = iy = -1;
  out[2 * k] = ix;
  out[2 * k + 1] = iy;
}

// ----------------------------------------------------------------------------------
// K4: finalise — one streaming pass over the grid.
//   l += k_decay*(-0.2f); l += miss*(-0.4f); l += hit*1.2f; +0.85f per covering footprint;
//   clamp [-2, 3.6]; occupancy = 1/(1+exp(-l)); counts cleared.
// R7 (ref: src/occupancy_grid.cpp:16-31) is COUNTS=false, nfoot=0, k_decay=1;
// R8/R9 (ref: :65-105, :33-63) add the footprint rectangles.  Adding the same constant
// 0.85f once per covering footprint is order-independent, so only the cover count is
// needed to reproduce the reference's sequential block adds bit for bit.
// cell0/ncell select a slab (multi-GPU: each rank finalises its own slab).
// ----------------------------------------------------------------------------------
struct FinalizeArgs {
  float *log_odds, *occupancy;
  int32_t *hit, *miss;
  // P2P: every rank's planes; the slab owner sums all ranks' counts (fused reduce-scatter),
  // writes log-odds/occupancy into every rank's grid (fused all-gather) and clears the counts
  int world;
  Peers<float> peer_lo, peer_occ;
  Peers<int32_t> peer_hit, peer_miss;
  unsigned long long cell0, ncell;  // slab [cell0, cell0+ncell), cell0 % 4 == 0
  int nx;
  float decay;  // (float)k_decay * -0.2f, computed on the host in float
  const int4 *rects;
  int nfoot;
};

template <bool COUNTS, bool P2P>
__global__ void __launch_bounds__(kThreads) k_finalize(const __grid_constant__ FinalizeArgs a)
{
  __shared__ int4 s_rect[kMaxFootCand];
  __shared__ int s_nc;
  const unsigned long long blk0 = a.cell0 + (unsigned long long)blockIdx.x * (kThreads * 4);
  if (a.nfoot > 0) {
    if (threadIdx.x == 0) s_nc = 0;
    __syncthreads();
    unsigned long long blk1 = blk0 + kThreads * 4 - 1;
    const unsigned long long last = a.cell0 + a.ncell - 1;
    if (blk1 > last) blk1 = last;
    const int iy0 = (int)(blk0 / (unsigned)a.nx), iy1 = (int)(blk1 / (unsigned)a.nx);
    for (int r = threadIdx.x; r < a.nfoot; r += kThreads) {
      const int4 R = a.rects[r];
      if (R.z >= R.x && R.w >= iy0 && R.y <= iy1) {
        const int slot = atomicAdd(&s_nc, 1);
        if (slot < kMaxFootCand) s_rect[slot] = R;
      }
    }
    __syncthreads();
  }
  const int ncand = a.nfoot > 0 ? s_nc : 0;

  const unsigned long long i0 = blk0 + (unsigned long long)threadIdx.x * 4;
  const unsigned long long endc = a.cell0 + a.ncell;
  if (i0 >= endc) return;
  const bool full = i0 + 4 <= endc;
  float l[4];
  int h[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0};
  if (full) {
    const float4 v = *reinterpret_cast<const float4 *>(a.log_odds + i0);
    l[0] = v.x; l[1] = v.y; l[2] = v.z; l[3] = v.w;
    if (COUNTS && !P2P) {
      const int4 hv = *reinterpret_cast<const int4 *>(a.hit + i0);
      const int4 mv = *reinterpret_cast<const int4 *>(a.miss + i0);
      h[0] = hv.x; h[1] = hv.y; h[2] = hv.z; h[3] = hv.w;
      m[0] = mv.x; m[1] = mv.y; m[2] = mv.z; m[3] = mv.w;
    }
    if (COUNTS && P2P) {
      for (int r = 0; r < a.world; ++r) {
        const int4 hv = *reinterpret_cast<const int4 *>(a.peer_hit.p[r] + i0);
        const int4 mv = *reinterpret_cast<const int4 *>(a.peer_miss.p[r] + i0);
        h[0] += hv.x; h[1] += hv.y; h[2] += hv.z; h[3] += hv.w;
        m[0] += mv.x; m[1] += mv.y; m[2] += mv.z; m[3] += mv.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = i0 + j < endc;
      l[j] = ok ? a.log_odds[i0 + j] : 0.0f;
      if (COUNTS && ok && !P2P) {
        h[j] = a.hit[i0 + j];
        m[j] = a.miss[i0 + j];
      }
      if (COUNTS && ok && P2P)
        for (int r = 0; r < a.world; ++r) {
          h[j] += a.peer_hit.p[r][i0 + j];
          m[j] += a.peer_miss.p[r][i0 + j];
        }
    }
  }
  int ix = (int)(i0 % (unsigned)a.nx), iy = (int)(i0 / (unsigned)a.nx);
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = __fadd_rn(l[j], a.decay);
    if (COUNTS) {
      v = __fadd_rn(v, __fmul_rn((float)m[j], -0.4f));  // log_odds_free_  (occupancy_grid.hpp:25)
      v = __fadd_rn(v, __fmul_rn((float)h[j], 1.2f));   // log_odds_occupied_ (:26)
    }
    if (ncand) {
      int cover = 0;
      if (ncand <= kMaxFootCand) {
        for (int r = 0; r < ncand; ++r) {
          const int4 R = s_rect[r];
          cover += (ix >= R.x && ix <= R.z && iy >= R.y && iy <= R.w) ? 1 : 0;
        }
      } else {  // candidate list overflowed: scan every footprint
        for (int r = 0; r < a.nfoot; ++r) {
          const int4 R = a.rects[r];
          cover += (ix >= R.x && ix <= R.z && iy >= R.y && iy <= R.w) ? 1 : 0;
        }
      }
      for (int r = 0; r < cover; ++r) v = __fadd_rn(v, 0.85f);  // ref: occupancy_grid.cpp:182
    }
    v = v < -2.0f ? -2.0f : v;  // cwiseMax(min_log_odds_)  ref: :21-22
    v = v > 3.6f ? 3.6f : v;    // cwiseMin(max_log_odds_)
    l[j] = v;
    o[j] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));  // ref: :28
    if (++ix == a.nx) {
      ix = 0;
      ++iy;
    }
  }
  if (full && P2P) {
    for (int r = 0; r < a.world; ++r) {
      *reinterpret_cast<float4 *>(a.peer_lo.p[r] + i0) = make_float4(l[0], l[1], l[2], l[3]);
      *reinterpret_cast<float4 *>(a.peer_occ.p[r] + i0) = make_float4(o[0], o[1], o[2], o[3]);
    }  // the count planes are cleared by each rank locally after the closing barrier
  } else if (!full && P2P) {
    for (int r = 0; r < a.world; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + j < endc) {
          a.peer_lo.p[r][i0 + j] = l[j];
          a.peer_occ.p[r][i0 + j] = o[j];
        }
  } else if (full) {
    *reinterpret_cast<float4 *>(a.log_odds + i0) = make_float4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<float4 *>(a.occupancy + i0) = make_float4(o[0], o[1], o[2], o[3]);
    if (COUNTS) {
      *reinterpret_cast<int4 *>(a.hit + i0) = make_int4(0, 0, 0, 0);
      *reinterpret_cast<int4 *>(a.miss + i0) = make_int4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j < endc) {
        a.log_odds[i0 + j] = l[j];
        a.occupancy[i0 + j] = o[j];
        if (COUNTS) {
          a.hit[i0 + j] = 0;
          a.miss[i0 + j] = 0;
        }
      }
  }
}

__global__ void __launch_bounds__(kThreads) k_fill_f32(float *p, unsigned long long n, float v)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride)
    p[i] = v;
}

// N3: nav_msgs/OccupancyGrid cells (grid_map_ros toOccupancyGrid(grid,"occupancy",0,1,msg),
// call site ref: src/grid_vision_node.cpp:270): reversed linear order, float -> int8.
__global__ void __launch_bounds__(kThreads) k_to_occupancy(const float *__restrict__ occ,
                                                           unsigned long long nc,
                                                           int8_t *__restrict__ data)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  float value = __fdiv_rn(__fsub_rn(occ[i], 0.0f), __fsub_rn(1.0f, 0.0f));
  if (value != value) value = -1.0f;
  else {
    float c = value < 0.0f ? 0.0f : value;
    c = c > 1.0f ? 1.0f : c;
    value = __fadd_rn(0.0f, __fmul_rn(c, 100.0f));
  }
  data[nc - i - 1] = (int8_t)value;
}

// ----------------------------------------------------------------------------------
// N1: ground-plane removal (ref: src/cloud_detections.cpp:105-138; oracle gvo_segment_ground).
// Deterministic batched RANSAC: all hypotheses are scored in ONE pass over the cloud (planes in
// shared memory, per-warp ballots), the best one is refined by a least-squares fit of its
// inliers (double moments + cyclic Jacobi, one thread) and the points within the threshold of
// the refined plane are dropped by an order-preserving compaction.
// ----------------------------------------------------------------------------------
constexpr int kMaxHyp = 1024;

__device__ __forceinline__ double block_sum(double v, double *s_red)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
  return t;
}


__host__ __device__ __forceinline__ unsigned mix32(unsigned seed, unsigned k)
{
  unsigned v = seed ^ (k * 0x9E3779B9u);
  v ^= v >> 16; v *= 0x85EBCA6Bu;
  v ^= v >> 13; v *= 0xC2B2AE35u;
  v ^= v >> 16;
  return v;
}

__device__ __forceinline__ float plane_dist(const float4 p, float x, float y, float z)
{
  return fabsf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, x), __fmul_rn(p.y, y)), __fmul_rn(p.z, z)), p.w));
}

struct GroundState {
  float4 plane;        // refined plane (a, b, c, d)
  float4 best_plane;   // winning hypothesis
  int best_h, best_score, found, pad;
  double mom[10];      // count, sum x y z, then xx xy xz yy yz zz about the centroid
};

__global__ void k_n1_hypotheses(const float *__restrict__ x, const float *__restrict__ y,
                                const float *__restrict__ z, unsigned n, unsigned seed, int n_hyp,
                                float4 *__restrict__ planes, int *__restrict__ valid)
{
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_hyp) return;
  const unsigned i0 = mix32(seed, 3u * h) % n, i1 = mix32(seed, 3u * h + 1u) % n, i2 = mix32(seed, 3u * h + 2u) % n;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  int ok = finite3(x[i0], y[i0], z[i0]) && finite3(x[i1], y[i1], z[i1]) && finite3(x[i2], y[i2], z[i2]);
  if (ok) {
    const float ux = __fsub_rn(x[i1], x[i0]), uy = __fsub_rn(y[i1], y[i0]), uz = __fsub_rn(z[i1], z[i0]);
    const float vx = __fsub_rn(x[i2], x[i0]), vy = __fsub_rn(y[i2], y[i0]), vz = __fsub_rn(z[i2], z[i0]);
    const float cx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
    const float cy = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz));
    const float cz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
    ok = n2 > 0.0f && finitef(n2);
    if (ok) {
      const float inv = __fdiv_rn(1.0f, __fsqrt_rn(n2));
      p.x = __fmul_rn(cx, inv); p.y = __fmul_rn(cy, inv); p.z = __fmul_rn(cz, inv);
      p.w = -__fadd_rn(__fadd_rn(__fmul_rn(p.x, x[i0]), __fmul_rn(p.y, y[i0])), __fmul_rn(p.z, z[i0]));
    }
  }
  planes[h] = p;
  valid[h] = ok;
}

__global__ void __launch_bounds__(kThreads) k_n1_score(const float *__restrict__ x, const float *__restrict__ y,
                                                       const float *__restrict__ z, unsigned long long n,
                                                       const float4 *__restrict__ planes,
                                                       const int *__restrict__ valid, int n_hyp,
                                                       float threshold, int *__restrict__ scores)
{
  __shared__ float4 s_p[kMaxHyp];
  __shared__ int s_cnt[kMaxHyp];
  for (int h = threadIdx.x; h < n_hyp; h += kThreads) {
    s_p[h] = valid[h] ? planes[h] : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fc00000));  // NaN: never an inlier
    s_cnt[h] = 0;
  }
  __syncthreads();
  const unsigned long long stride = (unsigned long long)gridDim.x * kThreads;
  const unsigned long long nround = (n + 31ull) & ~31ull;
  for (unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; i < nround; i += stride) {
    const bool live = i < n;
    const float px = live ? x[i] : 0.f, py = live ? y[i] : 0.f, pz = live ? z[i] : 0.f;
    for (int h = 0; h < n_hyp; ++h) {
      const unsigned m = __ballot_sync(0xffffffffu, live && plane_dist(s_p[h], px, py, pz) < threshold);
      if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt[h], __popc(m));
    }
  }
  __syncthreads();
  for (int h = threadIdx.x; h < n_hyp; h += kThreads)
    if (s_cnt[h]) atomicAdd(&scores[h], s_cnt[h]);
}

__global__ void k_n1_best(const float4 *__restrict__ planes, const int *__restrict__ scores, int n_hyp,
                          GroundState *__restrict__ st)
{
  // single warp: highest score, lowest hypothesis index on ties
  int best = 0, bh = -1;
  for (int h = threadIdx.x; h < n_hyp; h += 32)
    if (scores[h] > best) { best = scores[h]; bh = h; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ob = __shfl_xor_sync(0xffffffffu, best, o), oh = __shfl_xor_sync(0xffffffffu, bh, o);
    if (ob > best || (ob == best && oh >= 0 && (bh < 0 || oh < bh))) { best = ob; bh = oh; }
  }
  if (threadIdx.x == 0) {
    st->best_h = bh;
    st->best_score = best;
    st->found = best >= 3;
    st->best_plane = bh >= 0 ? planes[bh] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < 10; ++i) st->mom[i] = 0.0;
  }
}

// pass = 0: count + sums of the winning plane's inliers; pass = 1: second moments about the centroid
__global__ void __launch_bounds__(kThreads) k_n1_moments(const float *__restrict__ x, const float *__restrict__ y,
                                                         const float *__restrict__ z, unsigned long long n,
                                                         float threshold, int pass, GroundState *__restrict__ st)
{
  __shared__ double s_red[kThreads / 32];
  if (!st->found) return;
  const float4 bp = st->best_plane;
  double a[6] = {0, 0, 0, 0, 0, 0};
  const double cnt = st->mom[0];
  const double cx = pass ? st->mom[1] / cnt : 0.0, cy = pass ? st->mom[2] / cnt : 0.0, cz = pass ? st->mom[3] / cnt : 0.0;
  const unsigned long long stride = (unsigned long long)gridDim.x * kThreads;
  for (unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const float px = x[i], py = y[i], pz = z[i];
    if (plane_dist(bp, px, py, pz) < threshold) {
      if (!pass) { a[0] += 1.0; a[1] += px; a[2] += py; a[3] += pz; }
      else {
        const double dx = px - cx, dy = py - cy, dz = pz - cz;
        a[0] += dx * dx; a[1] += dx * dy; a[2] += dx * dz; a[3] += dy * dy; a[4] += dy * dz; a[5] += dz * dz;
      }
    }
  }
  const int nv = pass ? 6 : 4, base = pass ? 4 : 0;
  for (int k = 0; k < nv; ++k) {
    const double t = block_sum(a[k], s_red);
    if (threadIdx.x == 0 && t != 0.0) atomicAdd(&st->mom[base + k], t);
  }
}

__global__ void k_n1_refine(GroundState *__restrict__ st)
{
  if (!st->found) return;
  if (st->best_score <= 3) {  // PCL refines only with more inliers than the sample size
    st->plane = st->best_plane;
    return;
  }
  // eigenvector of the smallest eigenvalue (cyclic Jacobi, double) — oracle gvo_smallest_eigvec3
  const double *A = st->mom + 4;
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
        for (int k = 0; k < 3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int m = 0;
  if (a[1][1] < a[m][m]) m = 1;
  if (a[2][2] < a[m][m]) m = 2;
  double v[3] = {V[0][m], V[1][m], V[2][m]};
  const float4 bp = st->best_plane;
  if (v[0] * bp.x + v[1] * bp.y + v[2] * bp.z < 0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }
  const double cnt = st->mom[0], cx = st->mom[1] / cnt, cy = st->mom[2] / cnt, cz = st->mom[3] / cnt;
  st->plane = make_float4((float)v[0], (float)v[1], (float)v[2], (float)(-(v[0] * cx + v[1] * cy + v[2] * cz)));
}

// order-preserving removal of the refined plane's inliers: count pass / scatter pass
__global__ void __launch_bounds__(kThreads) k_n1_count(const float *__restrict__ x, const float *__restrict__ y,
                                                       const float *__restrict__ z, unsigned long long n,
                                                       float threshold, const GroundState *__restrict__ st,
                                                       unsigned *__restrict__ cta_count, uint8_t *__restrict__ keep)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
  bool k = false;
  if (i < n) {
    k = st->found ? !(plane_dist(st->plane, x[i], y[i], z[i]) < threshold) : true;
    if (keep) keep[i] = k;
  }
  const int c = __syncthreads_count(k);
  if (threadIdx.x == 0) cta_count[blockIdx.x] = (unsigned)c;
}

__global__ void __launch_bounds__(kThreads) k_n1_scatter(const float *__restrict__ x, const float *__restrict__ y,
                                                         const float *__restrict__ z, unsigned long long n,
                                                         float threshold, const GroundState *__restrict__ st,
                                                         const unsigned *__restrict__ cta_base,
                                                         float *__restrict__ ox, float *__restrict__ oy,
                                                         float *__restrict__ oz)
{
  __shared__ unsigned s_warp[kThreads / 32];
  const unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
  const bool k = i < n && (st->found ? !(plane_dist(st->plane, x[i], y[i], z[i]) < threshold) : true);
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, k);
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  unsigned base = cta_base[blockIdx.x];
  for (unsigned w = 0; w < warp; ++w) base += s_warp[w];
  if (k) {
    const unsigned long long o = (unsigned long long)base + __popc(m & ((1u << lane) - 1u));
    ox[o] = x[i]; oy[o] = y[i]; oz[o] = z[i];
  }
}

// ----------------------------------------------------------------------------------
// N2: per-box radius-outlier filter + PCA box (ref: src/cloud_detections.cpp:140-247;
// oracle gvo_radius_outlier_keep / gvo_bbox_pose).  Points arrive partitioned by box
// (k_label_hist/scatter): box b owns positions [off[b], off[b+1]) of the index list.
// ----------------------------------------------------------------------------------
struct LShapeDev {  // binary twin of gv_lshape / oracle gvo_lshape
  int kept;
  float centroid_y, mean_z, mean_x, major_z, major_x, minor_z, minor_x, length, width, angle_deg;
  double qx, qy, qz, qw;
};

// ref: :150-154 RadiusOutlierRemoval(0.4, 10): keep iff #{j : |p_j - p_i|^2 < r2} (self included)
// > min_nb.  One CTA = 256 query points of one box; the box's points stream through shared
// memory.  d2 is evaluated exactly like the oracle: (dx*dx + dy*dy) + dz*dz, single roundings.
__global__ void __launch_bounds__(kThreads) k_n2_neighbors(
  const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
  const unsigned *__restrict__ indices, const int2 *__restrict__ block_tab /* (box, first query) */,
  const unsigned long long *__restrict__ offsets, float r2, int min_nb, uint8_t *__restrict__ keep)
{
  __shared__ float sx[kThreads], sy[kThreads], sz[kThreads];
  const int2 bt = block_tab[blockIdx.x];
  const unsigned long long b0 = offsets[bt.x], b1 = offsets[bt.x + 1];
  const unsigned long long q = b0 + (unsigned)bt.y + threadIdx.x;
  const bool live = q < b1;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) {
    const unsigned i = indices[q];
    qx = x[i]; qy = y[i]; qz = z[i];
  }
  int k = 0;
  for (unsigned long long t = b0; t < b1; t += kThreads) {
    const unsigned long long j = t + threadIdx.x;
    __syncthreads();
    if (j < b1) {
      const unsigned i = indices[j];
      sx[threadIdx.x] = x[i]; sy[threadIdx.x] = y[i]; sz[threadIdx.x] = z[i];
    }
    __syncthreads();
    const int m = (int)(b1 - t < (unsigned long long)kThreads ? b1 - t : (unsigned long long)kThreads);
    if (live)
      for (int c = 0; c < m; ++c) {
        const float dx = __fsub_rn(sx[c], qx), dy = __fsub_rn(sy[c], qy), dz = __fsub_rn(sz[c], qz);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        k += d2 < r2 ? 1 : 0;
      }
  }
  if (live) keep[q] = k > min_nb ? 1 : 0;
}

// ref: :157-158 centroid, :189-191 cv::PCA on rows (z, x), :204-237 extents, :232-247 pose.
// One CTA per box.  Moments are reduced in double (the reference accumulates in float in
// input order: results agree to float rounding, compared with a tolerance in the tests); the
// extents use the oracle's float expression on the float mean/axes.
__global__ void __launch_bounds__(kThreads) k_n2_box(
  const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
  const unsigned *__restrict__ indices, const unsigned long long *__restrict__ offsets,
  const uint8_t *__restrict__ keep, LShapeDev *__restrict__ out)
{
  __shared__ double s_red[kThreads / 32];
  __shared__ float s_ax[6];  // mean_z, mean_x, major_z, major_x, minor_z, minor_x
  __shared__ float s_mm[4][kThreads / 32];
  const int b = blockIdx.x;
  const unsigned long long b0 = offsets[b], b1 = offsets[b + 1];
  double cnt = 0, sy = 0, sz = 0, sx = 0;
  for (unsigned long long j = b0 + threadIdx.x; j < b1; j += kThreads)
    if (keep[j]) {
      const unsigned i = indices[j];
      cnt += 1.0; sy += (double)y[i]; sz += (double)z[i]; sx += (double)x[i];
    }
  cnt = block_sum(cnt, s_red);
  sy = block_sum(sy, s_red);
  sz = block_sum(sz, s_red);
  sx = block_sum(sx, s_red);
  LShapeDev r;
  memset(&r, 0, sizeof(r));
  r.qw = 1.0;
  r.kept = (int)cnt;
  if (r.kept == 0) {  // ref: :201-202 data.empty() -> box skipped
    if (threadIdx.x == 0) out[b] = r;
    return;
  }
  const double mz = sz / cnt, mx = sx / cnt;
  double czz = 0, czx = 0, cxx = 0;
  for (unsigned long long j = b0 + threadIdx.x; j < b1; j += kThreads)
    if (keep[j]) {
      const unsigned i = indices[j];
      const double dz = (double)z[i] - mz, dx = (double)x[i] - mx;
      czz += dz * dz; czx += dz * dx; cxx += dx * dx;
    }
  czz = block_sum(czz, s_red) / cnt;
  czx = block_sum(czx, s_red) / cnt;
  cxx = block_sum(cxx, s_red) / cnt;
  if (threadIdx.x == 0) {
    // symmetric 2x2 eigen decomposition, larger eigenvalue first (oracle gvo_bbox_pose)
    const double tr = czz + cxx, df = czz - cxx;
    const double rad = sqrt(df * df + 4.0 * czx * czx);
    const double l1 = 0.5 * (tr + rad);
    double vz, vx;
    if (fabs(czx) > 1e-300) { vz = l1 - cxx; vx = czx; }
    else if (czz >= cxx) { vz = 1.0; vx = 0.0; }
    else { vz = 0.0; vx = 1.0; }
    const double nv = sqrt(vz * vz + vx * vx);
    vz /= nv; vx /= nv;
    if (vz < 0 || (vz == 0 && vx < 0)) { vz = -vz; vx = -vx; }
    s_ax[0] = (float)mz; s_ax[1] = (float)mx;
    s_ax[2] = (float)vz; s_ax[3] = (float)vx;
    s_ax[4] = (float)(-vx); s_ax[5] = (float)vz;
  }
  __syncthreads();
  const float fmz = s_ax[0], fmx = s_ax[1], az = s_ax[2], ax = s_ax[3], wz = s_ax[4], wx = s_ax[5];
  float minL = 3.402823466e+38f, maxL = -3.402823466e+38f, minW = minL, maxW = maxL;
  for (unsigned long long j = b0 + threadIdx.x; j < b1; j += kThreads)
    if (keep[j]) {
      const unsigned i = indices[j];
      const float dz = __fsub_rn(z[i], fmz), dx = __fsub_rn(x[i], fmx);
      const float pL = __fadd_rn(__fmul_rn(dz, az), __fmul_rn(dx, ax));
      const float pW = __fadd_rn(__fmul_rn(dz, wz), __fmul_rn(dx, wx));
      minL = fminf(minL, pL); maxL = fmaxf(maxL, pL);
      minW = fminf(minW, pW); maxW = fmaxf(maxW, pW);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    minL = fminf(minL, __shfl_xor_sync(0xffffffffu, minL, o));
    maxL = fmaxf(maxL, __shfl_xor_sync(0xffffffffu, maxL, o));
    minW = fminf(minW, __shfl_xor_sync(0xffffffffu, minW, o));
    maxW = fmaxf(maxW, __shfl_xor_sync(0xffffffffu, maxW, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s_mm[0][threadIdx.x >> 5] = minL; s_mm[1][threadIdx.x >> 5] = maxL;
    s_mm[2][threadIdx.x >> 5] = minW; s_mm[3][threadIdx.x >> 5] = maxW;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      s_mm[0][0] = fminf(s_mm[0][0], s_mm[0][w]); s_mm[1][0] = fmaxf(s_mm[1][0], s_mm[1][w]);
      s_mm[2][0] = fminf(s_mm[2][0], s_mm[2][w]); s_mm[3][0] = fmaxf(s_mm[3][0], s_mm[3][w]);
    }
    r.centroid_y = (float)(sy / cnt);
    r.mean_z = fmz; r.mean_x = fmx;
    r.major_z = az; r.major_x = ax; r.minor_z = wz; r.minor_x = wx;
    r.length = __fsub_rn(s_mm[1][0], s_mm[0][0]);
    r.width = __fsub_rn(s_mm[3][0], s_mm[2][0]);
    // ref: :232 degrees; :246-247 setRPY(0, -angle, 0) takes the DEGREE value as radians (kept)
    r.angle_deg = __fdiv_rn(__fmul_rn(atan2f(ax, az), 180.0f), 3.14159265358979323846f);
    const double hp = -(double)r.angle_deg * 0.5;
    r.qx = 0.0; r.qy = sin(hp); r.qz = 0.0; r.qw = cos(hp);
    out[b] = r;
  }
}

// ref: src/cloud_detections.cpp:282-283 compares float u,v (promoted) against the double box
// bounds, inclusive.  u >= x_min  <=>  u >= RU_f32(x_min)  and  u <= x_max  <=>  u <= RD_f32(x_max)
// for every float u, so rounding the bounds once (toward +inf for mins, -inf for maxes) turns
// the per-point double compares into exact float compares.  NaN bounds stay NaN (never match).
__global__ void k_round_boxes(const BoxRaw *__restrict__ in, int n, float4 *__restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = round_box(in[i]);
}

// ----------------------------------------------------------------------------------
// N4: cloud_detections::computeDepthForBoundingBoxes (ref: src/cloud_detections.cpp:43-87;
// oracle gvo_box_depths).  One CTA per box; the k nearest (u, v, depth) triples of the box centre
// (cx, cy, 0) are found by k selection rounds over the whole list: round r takes the smallest
// (distance, index) key greater than round r-1's.  Distance = FLANN's float accumulation
// ((du*du) + dv*dv) + depth*depth, non-finite triples are not candidates (PCL does not put them in
// the tree).  Brute force on purpose: at the reference's sizes (1e5 points, <= 50 boxes, k = 4)
// this is tens of microseconds and has no approximate-search parameter to agree on.
// ----------------------------------------------------------------------------------
constexpr int kMaxKnn = 64;

__global__ void __launch_bounds__(kThreads) k_box_knn_depth(const float *__restrict__ uvz, unsigned m,
                                                            const BoxRaw *__restrict__ boxes, int k,
                                                            float *__restrict__ depths)
{
  __shared__ unsigned long long s_best;
  __shared__ float s_depth[kMaxKnn];
  const BoxRaw B = boxes[blockIdx.x];
  // ref :57-60: centre in double, narrowed to pcl::PointXYZ's floats
  const float qx = __double2float_rn(__dadd_rn(B.x_min, __ddiv_rn(__dsub_rn(B.x_max, B.x_min), 2.0)));
  const float qy = __double2float_rn(__dadd_rn(B.y_min, __ddiv_rn(__dsub_rn(B.y_max, B.y_min), 2.0)));
  unsigned long long last = 0ull;
  int found = 0;
  for (int r = 0; r < k; ++r) {
    if (threadIdx.x == 0) s_best = ~0ull;
    __syncthreads();
    unsigned long long best = ~0ull;
    for (unsigned i = threadIdx.x; i < m; i += kThreads) {
      const float u = uvz[3 * i], v = uvz[3 * i + 1], z = uvz[3 * i + 2];
      if (!finite3(u, v, z)) continue;
      const float du = __fsub_rn(u, qx), dv = __fsub_rn(v, qy);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)), __fmul_rn(z, z));
      // d >= 0: its bit pattern orders like the value, so (bits, index) is the lexicographic key
      const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | i;
      if ((r == 0 || key > last) && key < best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
      best = t < best ? t : best;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&s_best, best);
    __syncthreads();
    const unsigned long long sel = s_best;
    __syncthreads();
    if (sel == ~0ull) break;  // fewer than k finite triples
    last = sel;
    if (threadIdx.x == 0) s_depth[r] = uvz[3 * (size_t)(unsigned)sel + 2];
    ++found;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // ref :78-82: std::nth_element at size/2 == element size/2 of the ascending depths
    for (int j = 1; j < found; ++j) {
      const float t = s_depth[j];
      int q = j;
      while (q > 0 && t < s_depth[q - 1]) {
        s_depth[q] = s_depth[q - 1];
        --q;
      }
      s_depth[q] = t;
    }
    depths[blockIdx.x] = found ? s_depth[found / 2] : -1.0f;  // ref :51 default
  }
}

// cloud_detections::pixelTo3D for every box centre (ref: src/cloud_detections.cpp:89-103 called
// from src/grid_vision_node.cpp:318-325): depth * (K_inv * (cx, cy, 1)), Eigen's left-to-right
// double accumulation; the centre is a cv::Point2f (floats).
__global__ void k_pixels_to_3d(const BoxRaw *__restrict__ boxes, const float *__restrict__ depths, int n,
                               const double *__restrict__ Kinv, double *__restrict__ xyz)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const BoxRaw B = boxes[i];
  const double px = (double)__double2float_rn(__dadd_rn(B.x_min, __ddiv_rn(__dsub_rn(B.x_max, B.x_min), 2.0)));
  const double py = (double)__double2float_rn(__dadd_rn(B.y_min, __ddiv_rn(__dsub_rn(B.y_max, B.y_min), 2.0)));
  const double d = (double)depths[i];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double t = __dadd_rn(__dadd_rn(__dmul_rn(Kinv[3 * r], px), __dmul_rn(Kinv[3 * r + 1], py)), Kinv[3 * r + 2]);
    xyz[3 * i + r] = __dmul_rn(d, t);
  }
}

// sensor origin -> start cell and continuous index coordinates (single thread)
struct OriginOut {
  int sx, sy, ok, pad;
};

__global__ void k_origin_setup(double ox, double oy, const __grid_constant__ GridGeom g,
                               OriginOut *out)
{
  OriginOut o;
  o.sx = o.sy = -1;
  o.pad = 0;
  o.ok = grid_get_index(g, ox, oy, o.sx, o.sy) ? 1 : 0;
  *out = o;
}

}  // namespace gv
