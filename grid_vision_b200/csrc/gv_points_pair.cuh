// gv_points_pair.cuh — k_points_pair: the hot instantiation of K1/K2 with TWO points per thread on
// Blackwell's packed binary32 pipe (FMUL2 / FFMA2 / FADD2: one issue slot, two lanes).
//
// What it computes is IDENTICAL to k_points_col (gv_points_fast.cuh) and so to k_points<FUSE,BIN>:
// per point the label of R3 (ref: src/cloud_detections.cpp:250-298 after the extrinsic of
// src/grid_vision_node.cpp:280-307) and one contribution to the end-cell plane (X1).  The kernel
// is instruction-issue bound (ncu, round 2: 83-91 % of issue slots at 21 % of DRAM), so the
// design goal is thread-instructions per point:
//   * thread t owns the ADJACENT points (2t, 2t+1) of every frame of its frame group: one
//     64-bit load per plane, one 32-bit label store, and every SE(3) row, the projection, the range
//     test and the IEEE sqrt / division of the range cap run on both points in one packed
//     instruction.  Adjacent azimuth steps of a ring take the same side of nearly every branch,
//     so the pair shares its control flow;
//   * the parity contract forbids contracting a multiply and an add (the reference build has no
//     FMA), and ptxas DOES contract mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (checked in SASS, with
//     or without -fmad=false).  Every add that consumes a product is therefore written as
//     fma(p, one, c) with `one` = 1.0f passed as a kernel parameter: p * 1 is exact, so the FMA
//     rounds p + c exactly like the separate add, and ptxas cannot fold what it cannot see;
//   * decisions are certified exactly as in fast_point (binary32 projection with an error
//     interval, the double-FMA 16.16 cell index); a point whose decision is not provably the
//     reference's sets its bit in the deferral bitmap and k_points_deferred re-runs it with the exact
//     FP64 code.  The image-tile masks are built 1 px wide (k_box_masks, rev32 mode), so the tile
//     is taken from the approximate pixel without a straddle test;
//   * control flow is structured (no early exits): a non-finite or deferred lane keeps flowing
//     through the arithmetic with its outputs masked, so the warp reconverges after every block.
// Needs 8-byte aligned planes, 4-byte aligned labels, even frame offsets and sizes (host-checked;
// anything else runs k_points_col).
#pragma once

#include "gv_points_fast.cuh"

namespace gv {

typedef unsigned long long f32x2;  // two binary32 lanes: low word = even point, high word = odd point

__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 bc2(float v) { return pk2(v, v); }
__device__ __forceinline__ float lo2(f32x2 v)
{
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo;
}
__device__ __forceinline__ float hi2(f32x2 v)
{
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return hi;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// only for sums whose operands are NOT products (ptxas would contract those): see mad2
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// RN(RN(a * b) + c): product and sum rounded separately, == __fadd_rn(__fmul_rn(a, b), c)
__device__ __forceinline__ f32x2 mad2(f32x2 a, f32x2 b, f32x2 c, f32x2 one) { return fma2(mul2(a, b), one, c); }

// one row of PCL's se3 on two points: c0*x + (c1*y + (c2*z + c3)), every operation rounded separately
__device__ __forceinline__ f32x2 se3_row2(const float *T, f32x2 x, f32x2 y, f32x2 z, f32x2 one)
{
  return mad2(bc2(T[0]), x, mad2(bc2(T[1]), y, mad2(bc2(T[2]), z, bc2(T[3]), one), one), one);
}

// sqrt(x) correctly rounded on both lanes (sqrt_rn_inrange, packed)
__device__ __forceinline__ f32x2 sqrt_rn_inrange2(f32x2 x)
{
  const f32x2 y = pk2(rsqrt_approx(lo2(x)), rsqrt_approx(hi2(x)));
  const f32x2 g = mul2(x, y), h = mul2(y, bc2(0.5f));
  const f32x2 r = fma2(mul2(g, bc2(-1.0f)), g, x);
  return fma2(r, h, g);
}

// a / b correctly rounded on both lanes (div_rn_inrange, packed)
__device__ __forceinline__ f32x2 div_rn_inrange2(f32x2 a, f32x2 b)
{
  const f32x2 nb = mul2(b, bc2(-1.0f));
  const f32x2 y0 = pk2(rcp_approx(lo2(b)), rcp_approx(hi2(b)));
  const f32x2 e = fma2(nb, y0, bc2(1.0f));
  const f32x2 y1 = fma2(y0, e, y0);
  const f32x2 q0 = mul2(a, y1);
  const f32x2 r = fma2(nb, q0, a);
  return fma2(r, y1, q0);
}

// Parameters of k_points_pair beyond FastArgs (filled by fill_pair_args)
struct PairArgs {
  float one;                  // 1.0f, opaque to ptxas (see mad2)
  float half_w, half_h;       // image centre: the image test is |q - W/2| against a threshold
  float ain_u, ain_v;         // |q - W/2| <  ain  => certainly 0 <= u < W
  float aout_u, aout_v;       // |q - W/2| >  aout => certainly u < 0 or u >= W
  float eu, ev;               // E(q) for any |q| <= W + 1: the interval of the box tests
  float inv_tile;             // 2^-mask_shift
  unsigned mask_bias;         // 0x4340 * (mask_tx + 1): see tile lookup
};

// Labels of both lanes.  mu = the candidate boxes of the lanes' image tiles (bit-reversed halves:
// leading one = lowest box index), visited in list order; pend = the lane is certainly inside the
// image and still unlabelled.  A candidate from the other lane's tile does not overlap this lane's
// (1-px dilated) tile, so it tests "certainly outside" like any other miss.  The first candidate that
// is not certainly outside ends the lane: certainly inside -> its label (ref :280-288, first box
// wins), else the point is deferred.
template <typename Boxes>
__device__ __forceinline__ void label_pair(const Boxes &bsrc, unsigned long long mu, bool pend0, bool pend1, f32x2 q,
                                           f32x2 r, float eu, float ev, int &lab0, int &lab1, unsigned &def)
{
  const f32x2 ql = add2(q, bc2(-eu)), qh = add2(q, bc2(eu)), rl = add2(r, bc2(-ev)), rh = add2(r, bc2(ev));
  const float ql0 = lo2(ql), ql1 = hi2(ql), qh0 = lo2(qh), qh1 = hi2(qh);
  const float rl0 = lo2(rl), rl1 = hi2(rl), rh0 = lo2(rh), rh1 = hi2(rh);
  unsigned mw = (unsigned)mu;
  unsigned base = 0;
#pragma unroll 1
  for (;;) {
    if (mw == 0u) {
      if (base) break;
      base = 32u * 16u;
      mw = (unsigned)(mu >> 32);
      if (mw == 0u) break;
    }
    const unsigned p = (unsigned)__clz((int)mw);
    mw &= ~(0x80000000u >> p);
    const float4 B = bsrc.box_at(base + 16u * p);
    const bool o0 = !pend0 | (qh0 < B.x) | (ql0 > B.z) | (rh0 < B.y) | (rl0 > B.w);  // certainly outside this box
    const bool o1 = !pend1 | (qh1 < B.x) | (ql1 > B.z) | (rh1 < B.y) | (rl1 > B.w);
    if (!(o0 & o1)) {
      const int id = (int)((base >> 4) + p);
      if (!o0) {
        if ((ql0 >= B.x) & (qh0 <= B.z) & (rl0 >= B.y) & (rh0 <= B.w)) lab0 = id;
        else def |= 1u;
        pend0 = false;
      }
      if (!o1) {
        if ((ql1 >= B.x) & (qh1 <= B.z) & (rl1 >= B.y) & (rh1 <= B.w)) lab1 = id;
        else def |= 2u;
        pend1 = false;
      }
      if (!(pend0 | pend1)) break;
    }
  }
}

// end cell of one lane whose certified index did not say "inside" (same code as fast_point)
__device__ __forceinline__ void offmap_lane(const FastArgs &a, float bx, float by, unsigned tx, unsigned ty,
                                            bool word_ok, int &lin, unsigned &def, unsigned bit)
{
  const int sx = (int)tx, sy = (int)ty;  // k - 8, signed view
  const bool out = !word_ok | (sx < -16) | (sy < -16) | (sx >= (int)a.klim_x) | (sy >= (int)a.klim_y);
  if (!out) {
    def |= bit;  // within 2^-13 cells of a cell or map boundary
  } else {
    // off-map endpoint: clip the free-space-only beam to the map (oracle gvo_clip_end, all float)
    const float eax = __fmul_rn(__fsub_rn(a.c0xf, bx), a.inv_resf);
    const float eay = __fmul_rn(__fsub_rn(a.c0yf, by), a.inv_resf);
    const float dax = __fsub_rn(eax, a.oaxf), day = __fsub_rn(eay, a.oayf);
    float t = 1.0f;
    if (eax < 0.0f || eax >= a.nxf) {
      const float tt = div_rn_inrange(eax < 0.0f ? a.noaxf : a.paxf, dax);
      if (tt < t) t = tt;
    }
    if (eay < 0.0f || eay >= a.nyf) {
      const float tt = div_rn_inrange(eay < 0.0f ? a.noayf : a.payf, day);
      if (tt < t) t = tt;
    }
    const int ex = (int)fminf(fmaxf(__fadd_rn(a.oaxf, __fmul_rn(t, dax)), 0.0f), a.nxm1f);
    const int ey = (int)fminf(fmaxf(__fadd_rn(a.oayf, __fmul_rn(t, day)), 0.0f), a.nym1f);
    lin = ex + ey * a.hot.nx;
  }
}

// run-length binning of one lane (see k_points_col): (cell, beams | hits << 16) in two registers
__device__ __forceinline__ void run_bin_reg(unsigned long long *ends, int &run_cell, unsigned &run, int lin, bool valid,
                                            bool hit)
{
  const bool changed = valid & (lin != run_cell);
  if (changed & (run != 0u)) atomicAdd(ends + run_cell, ((unsigned long long)(run >> 16) << 32) | (run & 0xffffu));
  if (changed) {
    run_cell = lin;
    run = 0u;
  }
  if (valid) run += hit ? 0x10001u : 1u;
}

#ifndef GV_PAIR_MINB
#define GV_PAIR_MINB 5  // default CTAs per SM the register allocation aims at ($GV_PAIR_MINB at run time: 3..6)
#endif

// grid = (pair-column blocks, frame groups); a.frames_per_cta <= min(kColFrames, 32767) keeps the
// 16-bit run counters exact
template <bool BOUNDED, bool LAB, bool ZGATE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_points_pair(const __grid_constant__ FastArgs a,
                                                                        const __grid_constant__ PairArgs pa)
{
  __shared__ uint4 s_rec[kColFrames + 1];  // {element offset relative to the group's first frame, points, first box, -}
  const unsigned pidx = blockIdx.x * kThreads + threadIdx.x;
  const unsigned idx = 2u * pidx;  // this thread's even point; idx + 1 is its odd point
  const int f0 = a.frame0 + (int)blockIdx.y * a.frames_per_cta;
  int nf = a.frame0 + a.nframes - f0;
  if (nf > a.frames_per_cta) nf = a.frames_per_cta;
  if (nf <= 0) return;
  const uint4 first = __ldg(a.frames + f0);
  const unsigned long long off0 = ((unsigned long long)first.y << 32) | first.x;
  for (int k = threadIdx.x; k <= nf; k += kThreads) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (k < nf) {
      const uint4 g = __ldg(a.frames + f0 + k);
      r.x = (unsigned)((((unsigned long long)g.y << 32) | g.x) - off0);  // < 2^32: host-checked
      r.y = g.z;
      r.z = g.w;
    }
    s_rec[k] = r;  // entry nf: an empty frame (ends the prefetch chain)
  }
  __syncthreads();
  const FastHot &h = a.hot;
  const FastWarm &w = a.warm;

  const float *xb = a.x + off0 + idx, *yb = a.y + off0 + idx, *zb = a.z + off0 + idx;
  int16_t *lb = LAB ? a.labels + off0 + idx : nullptr;
  unsigned fcur = (unsigned)f0;
  unsigned idx_r = idx;
  unsigned sa_k = (unsigned)__cvta_generic_to_shared(s_rec);  // this frame's record
  const unsigned sa_end = sa_k + 16u * (unsigned)nf;
  // opaque to the optimiser: otherwise these loop invariants are re-derived from the parameter
  // block and the special registers inside the loop (a dozen instructions per iteration)
  asm volatile("" : "+l"(xb), "+l"(yb), "+l"(zb), "+l"(lb), "+r"(idx_r), "+r"(sa_k));
  const f32x2 one = bc2(pa.one);
  f32x2 nx = 0ull, ny = 0ull, nz = 0ull;
  if (idx_r < first.z) {
    nx = __ldcs(reinterpret_cast<const unsigned long long *>(xb));
    ny = __ldcs(reinterpret_cast<const unsigned long long *>(yb));
    nz = __ldcs(reinterpret_cast<const unsigned long long *>(zb));
  }
  int cell0 = -1, cell1 = -1;
  unsigned run0 = 0u, run1 = 0u;
#pragma unroll 1
  for (; sa_k != sa_end; sa_k += 16u, ++fcur) {
    uint4 rcur;
    uint2 rnxt;  // the next frame's {offset, points} (entry nf: an empty frame)
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rcur.x), "=r"(rcur.y), "=r"(rcur.z), "=r"(rcur.w) : "r"(sa_k));
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+16];" : "=r"(rnxt.x), "=r"(rnxt.y) : "r"(sa_k));
    const f32x2 px = nx, py = ny, pz = nz;
    if (idx_r < rnxt.y) {  // the next frame's pair: in flight while this one is processed
      nx = __ldcs(reinterpret_cast<const unsigned long long *>(xb + rnxt.x));
      ny = __ldcs(reinterpret_cast<const unsigned long long *>(yb + rnxt.x));
      nz = __ldcs(reinterpret_cast<const unsigned long long *>(zb + rnxt.x));
    }
    if (!(idx_r < rcur.y)) continue;

    const float x0 = lo2(px), x1 = hi2(px), y0 = lo2(py), y1 = hi2(py), z0 = lo2(pz), z1 = hi2(pz);
    // all three |v| < 1e9?  max.NaN propagates NaN, so NaN and Inf fail the compare.  A lane that
    // fails keeps flowing (its values are garbage, every use below is masked by ok / def).
    const float m0 = fmax3_nan_abs(x0, y0, z0), m1 = fmax3_nan_abs(x1, y1, z1);
    const bool ok0 = m0 < 1.0e9f, ok1 = m1 < 1.0e9f;
    unsigned def = 0u;  // bit 0 / 1: the even / odd point is deferred to k_points_deferred
    if (!(ok0 & ok1)) {
      // finite but huge: nothing is certified.  non-finite: no label (ref :264), beam dropped (X1)
      if (!ok0 & (m0 < __int_as_float(0x7f800000))) def |= 1u;
      if (!ok1 & (m1 < __int_as_float(0x7f800000))) def |= 2u;
    }
    // With |T| < 1e6 (host-checked) every transformed coordinate of an ok lane is finite (< 3.1e15).

    // ---------------- camera: depth row first.  Ends with the image-tile mask loads in flight; the
    // box tests that consume them run after the base-frame block (label_pair below).
    int lab0 = -1, lab1 = -1;
    bool in0 = false, in1 = false;
    f32x2 q = 0ull, r = 0ull;
    unsigned long long cm0 = 0ull, cm1 = 0ull;
    {
      const f32x2 Z = se3_row2(h.Tcz, px, py, pz, one);
      const float Z0 = lo2(Z), Z1 = hi2(Z);
      const bool fr0 = Z0 > 0.001f, fr1 = Z1 > 0.001f;  // ref :264 (NaN: false)
      if (fr0 | fr1) {
        const f32x2 X = se3_row2(w.Tcxy, px, py, pz, one), Y = se3_row2(w.Tcxy + 4, px, py, pz, one);
        // certified projection: q = fx*(X/Z) + cx in binary32 with rcp.approx (1 ulp):
        //   |q - u_ref| <= 2^-24 (5|u| + 3|cx|), and E(q) = 2^-22 (6|q| + 1.5|cx| + 1) is at least
        //   twice that (fast_point).  Inside or near the image |q| <= W + 1, so the constant
        //   pa.eu >= E(q) there; the image test itself uses thresholds derived from E(q) on the host.
        const f32x2 rz = pk2(rcp_approx(Z0), rcp_approx(Z1));
        q = fma2(bc2(w.fx), mul2(X, rz), bc2(w.cx));
        r = fma2(bc2(w.fy), mul2(Y, rz), bc2(w.cy));
        const f32x2 dq = add2(q, bc2(-pa.half_w)), dr = add2(r, bc2(-pa.half_h));
        const float aq0 = fabsf(lo2(dq)), aq1 = fabsf(hi2(dq)), ar0 = fabsf(lo2(dr)), ar1 = fabsf(hi2(dr));
        in0 = fr0 & (aq0 < pa.ain_u) & (ar0 < pa.ain_v);  // certainly inside the image (:276)
        in1 = fr1 & (aq1 < pa.ain_u) & (ar1 < pa.ain_v);
        // neither certainly inside nor certainly outside: too close to an image edge to call
        if (fr0 & !in0 & !((aq0 > pa.aout_u) | (ar0 > pa.aout_v))) def |= 1u;
        if (fr1 & !in1 & !((aq1 > pa.aout_u) | (ar1 > pa.aout_v))) def |= 2u;
        if (in0 | in1) {
          // image tile: (q / S + 192) has ulp 2^-16, so bits >> 16 = 0x4340 + floor(q / S) up to
          // a rounding of 2^-16 tiles, which the 1-px dilation of the tile masks covers
          const f32x2 tu = fma2(q, bc2(pa.inv_tile), bc2(192.0f)), tv = fma2(r, bc2(pa.inv_tile), bc2(192.0f));
          const unsigned long long *mrow = a.masks + (size_t)fcur * a.mask_stride;
          if (in0)
            cm0 = __ldg(mrow + ((__float_as_uint(lo2(tv)) >> 16) * (unsigned)a.mask_tx +
                                (__float_as_uint(lo2(tu)) >> 16) - pa.mask_bias));
          if (in1)
            cm1 = __ldg(mrow + ((__float_as_uint(hi2(tv)) >> 16) * (unsigned)a.mask_tx +
                                (__float_as_uint(hi2(tu)) >> 16) - pa.mask_bias));
        }
      }
    }

    // ---------------- base frame: end cell (oracle gvo_accumulate, per-point body)
    f32x2 bx = se3_row2(h.Tb, px, py, pz, one), by = se3_row2(h.Tb + 4, px, py, pz, one);
    bool cap0, cap1;
    {
      const f32x2 dx = add2(bx, bc2(-h.oxf)), dy = add2(by, bc2(-h.oyf));
      const f32x2 r2 = fma2(mul2(dx, dx), one, mul2(dy, dy));
      cap0 = lo2(r2) > h.rmax2f;
      cap1 = hi2(r2) > h.rmax2f;
      if (cap0 | cap1) {  // beyond the mapping range: free-space-only beam shortened to r_max
        const f32x2 sf = div_rn_inrange2(bc2(w.rmaxf), sqrt_rn_inrange2(r2));
        const f32x2 cx = mad2(sf, dx, bc2(h.oxf), one), cy = mad2(sf, dy, bc2(h.oyf), one);
        bx = pk2(cap0 ? lo2(cx) : lo2(bx), cap1 ? hi2(cx) : hi2(bx));
        by = pk2(cap0 ? lo2(cy) : lo2(by), cap1 ? hi2(cy) : hi2(by));
      }
    }
    // certified index (fast_point): low word of fma((double)b, -1/res, C) = 16.16 index coordinate
    int lin0, lin1;
    bool ins0, ins1;
    {
      const float bx0 = lo2(bx), bx1 = hi2(bx), by0 = lo2(by), by1 = hi2(by);
      const double rx0 = fma((double)bx0, h.nires, h.Cx), ry0 = fma((double)by0, h.nires, h.Cy);
      const double rx1 = fma((double)bx1, h.nires, h.Cx), ry1 = fma((double)by1, h.nires, h.Cy);
      const unsigned tx0 = (unsigned)__double2loint(rx0) - h.kb8, ty0 = (unsigned)__double2loint(ry0) - h.kb8;
      const unsigned tx1 = (unsigned)__double2loint(rx1) - h.kb8, ty1 = (unsigned)__double2loint(ry1) - h.kb8;
      bool wok0 = true, wok1 = true;
      if (!BOUNDED) {
        wok0 = ((unsigned)__double2hiint(rx0) == a.hi0) & ((unsigned)__double2hiint(ry0) == a.hi0);
        wok1 = ((unsigned)__double2hiint(rx1) == a.hi0) & ((unsigned)__double2hiint(ry1) == a.hi0);
      }
      ins0 = wok0 & (tx0 < h.klim_x16) & (ty0 < h.klim_y16) & (max(tx0 & 0xffffu, ty0 & 0xffffu) < 0xfff0u);
      ins1 = wok1 & (tx1 < h.klim_x16) & (ty1 < h.klim_y16) & (max(tx1 & 0xffffu, ty1 & 0xffffu) < 0xfff0u);
      lin0 = (int)(tx0 >> 16) + (int)(ty0 >> 16) * h.nx;
      lin1 = (int)(tx1 >> 16) + (int)(ty1 >> 16) * h.nx;
      // a lane without a usable point must not drag the warp into the off-map code
      if (!((ins0 | !ok0) & (ins1 | !ok1))) {
        if (!ins0 & ok0) offmap_lane(a, bx0, by0, tx0, ty0, wok0, lin0, def, 1u);
        if (!ins1 & ok1) offmap_lane(a, bx1, by1, tx1, ty1, wok1, lin1, def, 2u);
      }
    }
    // ---------------- labels (the masks have had the whole base-frame block to arrive)
    if (in0 | in1)
      label_pair(GmemBoxes{a.boxes + rcur.z, nullptr}, cm0 | cm1, in0, in1, q, r, pa.eu, pa.ev, lab0, lab1, def);
    bool hit0 = ins0 & !cap0 & (lab0 >= h.lab_min), hit1 = ins1 & !cap1 & (lab1 >= h.lab_min);
    if (ZGATE) {
      const f32x2 bz = se3_row2(a.Tbz, px, py, pz, one);
      hit0 &= (lo2(bz) >= a.z_min) & (lo2(bz) <= a.z_max);
      hit1 &= (hi2(bz) >= a.z_min) & (hi2(bz) <= a.z_max);
    }
    // labels of the pair in one streaming store (a deferred lane's half is rewritten by k_points_deferred)
    if (LAB) __stcs(reinterpret_cast<unsigned *>(lb + rcur.x), ((unsigned)lab0 & 0xffffu) | ((unsigned)lab1 << 16));
    run_bin_reg(a.ends, cell0, run0, lin0, ok0 & !(def & 1u), hit0);
    run_bin_reg(a.ends, cell1, run1, lin1, ok1 & !(def & 2u), hit1);
    if (def) atomicOr(a.defer_bits + (size_t)fcur * a.defer_stride + (idx_r >> 5), def << (idx_r & 31u));
  }
  if (run0) atomicAdd(a.ends + cell0, ((unsigned long long)(run0 >> 16) << 32) | (run0 & 0xffffu));
  if (run1) atomicAdd(a.ends + cell1, ((unsigned long long)(run1 >> 16) << 32) | (run1 & 0xffffu));
}

}  // namespace gv
