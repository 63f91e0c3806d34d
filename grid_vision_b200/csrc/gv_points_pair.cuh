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

// Hot parameters of k_points_pair, in 16-byte rows in the order the loop reads them (so that one
// 128-bit constant load brings a whole row into uniform registers); filled by fill_pair_args
struct __align__(16) PairArgs {
  float Tcz[4];                               // camera extrinsic, depth row (R1)
  float Tcx[4], Tcy[4];                       // camera extrinsic rows x, y
  float fx, fy, cx, cy;                       // certified R3
  float half_w, half_h, ain_u, ain_v;         // |q - W/2| <  ain  => certainly 0 <= u < W
  float aout_u, aout_v, eu, ev;               // |q - W/2| >  aout => certainly outside; eu >= E(q) for |q| <= W + 1
  float inv_tile;                             // 2^-mask_shift
  unsigned mask_bias;                         // 0x4340 * (mask_tx + 1): see the tile lookup
  unsigned mask_tx, mask_stride;
  float Tbx[4], Tby[4];                       // base transform rows x, y (X1)
  float noxf, noyf, rmax2f, rmaxf;            // -origin; rmax2f = +inf when the range cap is disabled
  float oxf, oyf, one, pad0;                  // one = 1.0f, opaque to ptxas (see mad2)
  double nires, Cx;                           // index FMA (fast_point)
  double Cy;
  unsigned kbm;
  int nx;
  unsigned klim_xm, klim_ym;
  int lab_min;
  unsigned defer_stride;
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
  // next candidate in list order: byte offset of its box, or false when the masks are exhausted
  auto next = [&](unsigned &off) -> bool {
    if (mw == 0u) {
      if (base) return false;
      base = 32u * 16u;
      mw = (unsigned)(mu >> 32);
      if (mw == 0u) return false;
    }
    const unsigned p = (unsigned)__clz((int)mw);
    mw &= ~(0x80000000u >> p);
    off = base + 16u * p;
    return true;
  };
  auto test = [&](const float4 B, unsigned off) -> bool {  // true: both lanes are settled
    const bool o0 = !pend0 | (qh0 < B.x) | (ql0 > B.z) | (rh0 < B.y) | (rl0 > B.w);  // certainly outside this box
    const bool o1 = !pend1 | (qh1 < B.x) | (ql1 > B.z) | (rh1 < B.y) | (rl1 > B.w);
    if (!(o0 & o1)) {
      const int id = (int)(off >> 4);
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
      return !(pend0 | pend1);
    }
    return false;
  };
  // (Prefetching the next candidate's box while this one is tested was measured: the five extra live
  // registers spill in the main loop under the 48-register budget, 2.81 -> 3.24 ms.)
  unsigned off;
#pragma unroll 1
  while (next(off))
    if (test(bsrc.box_at(off), off)) break;
}

// end cell of one lane whose certified index did not say "inside" (same code as fast_point)
__device__ __forceinline__ void offmap_lane(const FastArgs &a, float bx, float by, unsigned tx, unsigned ty,
                                            bool word_ok, int &lin, unsigned &def, unsigned bit)
{
  const int sx = (int)tx, sy = (int)ty;  // k - margin, signed view
  const bool out = !word_ok | (sx < -16) | (sy < -16) | (sx >= (int)a.klim_x) | (sy >= (int)a.klim_y);
  if (!out) {
    def |= bit;  // within a few 2^-16 cells of a cell or map boundary
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

#ifndef GV_PAIR_MINB
#define GV_PAIR_MINB 5  // default CTAs per SM the register allocation aims at ($GV_PAIR_MINB at run time: 3..6)
#endif

struct PairRuns {  // run-length binning state of the two lanes: (cell, beams | hits << 16)
  int cell0, cell1;
  unsigned run0, run1;
};

// the beam's end cell differs from the open run: flush the run, open a new one
__device__ __forceinline__ void run_flush_reg(unsigned long long *ends, int &run_cell, unsigned &run, int lin, bool valid)
{
  if (valid & (lin != run_cell)) {
    if (run) atomicAdd(ends + run_cell, ((unsigned long long)(run >> 16) << 32) | (run & 0xffffu));
    run_cell = lin;
    run = 0u;
  }
}

// One pair of points: label store, deferral bits, run-length bins.  rcur = the frame's record.
template <bool BOUNDED, bool LAB, bool ZGATE>
__device__ __forceinline__ void pair_body(const FastArgs &a, const PairArgs &pa, const f32x2 one, const unsigned idx,
                                          const unsigned fcur, const uint4 rcur, const f32x2 px, const f32x2 py,
                                          const f32x2 pz, PairRuns &st)
{
  // A point with a non-finite or huge coordinate is not tested up front: it keeps flowing, and
  // IEEE arithmetic flags it.  0 * NaN = 0 * Inf = NaN, so a NaN or Inf coordinate makes EVERY
  // SE(3) row non-finite whatever the matrix: the depth test is false (no label, ref :264), r2
  // below is NaN or +Inf.  r2 = NaN drops the beam (X1); r2 >= 1e30 (Inf, or a finite coordinate
  // large enough to threaten overflow: |T| < 1e6 is host-checked, so r2 < 1e30 bounds |x|, |y| of
  // the base frame) defers the point.  A huge z alone leaves r2 small; it can only make the
  // camera rows or the projection overflow: the rows are the reference's own arithmetic (same
  // NaN / Inf), an infinite q or r is certainly outside the image as it is in the reference, and a
  // NaN q or r fails both image tests: deferred.
  unsigned def = 0u;  // bit 0 / 1: the even / odd point is deferred to k_points_deferred

  // ---------------- camera: depth row first.  Ends with the image-tile mask loads in flight; the
  // box tests that consume them run after the base-frame block (label_pair below).
  int lab0 = -1, lab1 = -1;
  bool in0 = false, in1 = false;
  f32x2 q = 0ull, r = 0ull;
  unsigned long long cm0 = 0ull, cm1 = 0ull;
  {
    const f32x2 Z = se3_row2(pa.Tcz, px, py, pz, one);
    const float Z0 = lo2(Z), Z1 = hi2(Z);
    const bool fr0 = Z0 > 0.001f, fr1 = Z1 > 0.001f;  // ref :264 (NaN: false)
    if (fr0 | fr1) {
      const f32x2 X = se3_row2(pa.Tcx, px, py, pz, one), Y = se3_row2(pa.Tcy, px, py, pz, one);
      // certified projection: q = fx*(X/Z) + cx in binary32 with rcp.approx (1 ulp):
      //   |q - u_ref| <= 2^-24 (5|u| + 3|cx|), and E(q) = 2^-22 (6|q| + 1.5|cx| + 1) is at least
      //   twice that (fast_point).  Inside or near the image |q| <= W + 1, so the constant
      //   pa.eu >= E(q) there; the image test itself uses thresholds derived from E(q) on the host.
      const f32x2 rz = pk2(rcp_approx(Z0), rcp_approx(Z1));
      q = fma2(bc2(pa.fx), mul2(X, rz), bc2(pa.cx));
      r = fma2(bc2(pa.fy), mul2(Y, rz), bc2(pa.cy));
      const f32x2 dq = add2(q, bc2(-pa.half_w)), dr = add2(r, bc2(-pa.half_h));
      const float aq0 = fabsf(lo2(dq)), aq1 = fabsf(hi2(dq)), ar0 = fabsf(lo2(dr)), ar1 = fabsf(hi2(dr));
      in0 = fr0 & (aq0 < pa.ain_u) & (ar0 < pa.ain_v);  // certainly inside the image (:276)
      in1 = fr1 & (aq1 < pa.ain_u) & (ar1 < pa.ain_v);
      // neither certainly inside nor certainly outside: too close to an image edge to call
      const bool e0 = fr0 & !in0 & !((aq0 > pa.aout_u) | (ar0 > pa.aout_v));
      const bool e1 = fr1 & !in1 & !((aq1 > pa.aout_u) | (ar1 > pa.aout_v));
      def |= (e0 ? 1u : 0u) | (e1 ? 2u : 0u);
      if (in0 | in1) {
        // image tile: (q / S + 192) has ulp 2^-16, so bits >> 16 = 0x4340 + floor(q / S) up to
        // a rounding of 2^-16 tiles, which the 1-px dilation of the tile masks covers
        const f32x2 tu = fma2(q, bc2(pa.inv_tile), bc2(192.0f)), tv = fma2(r, bc2(pa.inv_tile), bc2(192.0f));
        if (in0)
          cm0 = __ldg(a.masks + ((__float_as_uint(lo2(tv)) >> 16) * pa.mask_tx + (__float_as_uint(lo2(tu)) >> 16) + rcur.w));
        if (in1)
          cm1 = __ldg(a.masks + ((__float_as_uint(hi2(tv)) >> 16) * pa.mask_tx + (__float_as_uint(hi2(tu)) >> 16) + rcur.w));
      }
    }
  }

  // ---------------- base frame: end cell (oracle gvo_accumulate, per-point body)
  f32x2 bx = se3_row2(pa.Tbx, px, py, pz, one), by = se3_row2(pa.Tby, px, py, pz, one);
  bool cap0, cap1, ok0, ok1;
  {
    const f32x2 dx = add2(bx, bc2(pa.noxf)), dy = add2(by, bc2(pa.noyf));
    const f32x2 r2 = fma2(mul2(dx, dx), one, mul2(dy, dy));
    ok0 = lo2(r2) < 1.0e30f;
    ok1 = hi2(r2) < 1.0e30f;
    cap0 = lo2(r2) > pa.rmax2f;
    cap1 = hi2(r2) > pa.rmax2f;
    // r2 >= 1e30 defers.  With a range cap (BOUNDED) such a lane is also capped: tested in there.
    if (!BOUNDED) def |= (lo2(r2) >= 1.0e30f ? 1u : 0u) | (hi2(r2) >= 1.0e30f ? 2u : 0u);
    if (cap0 | cap1) {  // beyond the mapping range: free-space-only beam shortened to r_max
      if (BOUNDED) def |= (lo2(r2) >= 1.0e30f ? 1u : 0u) | (hi2(r2) >= 1.0e30f ? 2u : 0u);
      const f32x2 sf = div_rn_inrange2(bc2(pa.rmaxf), sqrt_rn_inrange2(r2));
      const f32x2 cx = mad2(sf, dx, bc2(pa.oxf), one), cy = mad2(sf, dy, bc2(pa.oyf), one);
      bx = pk2(cap0 ? lo2(cx) : lo2(bx), cap1 ? hi2(cx) : hi2(bx));
      by = pk2(cap0 ? lo2(cy) : lo2(by), cap1 ? hi2(cy) : hi2(by));
    }
  }
  // certified index (fast_point): low word of fma((double)b, -1/res, C) = 16.16 index coordinate
  int lin0, lin1;
  bool ins0, ins1;
  {
    const float bx0 = lo2(bx), bx1 = hi2(bx), by0 = lo2(by), by1 = hi2(by);
    const double rx0 = fma((double)bx0, pa.nires, pa.Cx), ry0 = fma((double)by0, pa.nires, pa.Cy);
    const double rx1 = fma((double)bx1, pa.nires, pa.Cx), ry1 = fma((double)by1, pa.nires, pa.Cy);
    const unsigned tx0 = (unsigned)__double2loint(rx0) - pa.kbm, ty0 = (unsigned)__double2loint(ry0) - pa.kbm;
    const unsigned tx1 = (unsigned)__double2loint(rx1) - pa.kbm, ty1 = (unsigned)__double2loint(ry1) - pa.kbm;
    bool wok0 = true, wok1 = true;
    if (!BOUNDED) {
      wok0 = ((unsigned)__double2hiint(rx0) == a.hi0) & ((unsigned)__double2hiint(ry0) == a.hi0);
      wok1 = ((unsigned)__double2hiint(rx1) == a.hi0) & ((unsigned)__double2hiint(ry1) == a.hi0);
    }
    ins0 = wok0 & (tx0 < pa.klim_xm) & (ty0 < pa.klim_ym) & (max(tx0 & 0xffffu, ty0 & 0xffffu) < kFracLim);
    ins1 = wok1 & (tx1 < pa.klim_xm) & (ty1 < pa.klim_ym) & (max(tx1 & 0xffffu, ty1 & 0xffffu) < kFracLim);
    lin0 = (int)(tx0 >> 16) + (int)(ty0 >> 16) * pa.nx;
    lin1 = (int)(tx1 >> 16) + (int)(ty1 >> 16) * pa.nx;
    // a lane without a usable point must not drag the warp into the off-map code
    if (!((ins0 | !ok0) & (ins1 | !ok1))) {
      if (!ins0 & ok0) offmap_lane(a, bx0, by0, tx0, ty0, wok0, lin0, def, 1u);
      if (!ins1 & ok1) offmap_lane(a, bx1, by1, tx1, ty1, wok1, lin1, def, 2u);
    }
  }
  // ---------------- labels (the masks have had the whole base-frame block to arrive)
  if (in0 | in1)
    label_pair(GmemBoxes{a.boxes + rcur.z, nullptr}, cm0 | cm1, in0, in1, q, r, pa.eu, pa.ev, lab0, lab1, def);
  bool hit0 = ins0 & !cap0 & (lab0 >= pa.lab_min), hit1 = ins1 & !cap1 & (lab1 >= pa.lab_min);
  if (ZGATE) {
    const f32x2 bz = se3_row2(a.Tbz, px, py, pz, one);
    hit0 &= (lo2(bz) >= a.z_min) & (lo2(bz) <= a.z_max);
    hit1 &= (hi2(bz) >= a.z_min) & (hi2(bz) <= a.z_max);
  }
  // labels of the pair in one streaming store (a deferred lane's half is rewritten by k_points_deferred)
  if (LAB) __stcs(reinterpret_cast<unsigned *>(a.labels + (rcur.x + idx)), ((unsigned)lab0 & 0xffffu) | ((unsigned)lab1 << 16));
  {
    const bool v0 = ok0 & !(def & 1u), v1 = ok1 & !(def & 2u);
    // common case first: both beams end where this thread's previous beams ended
    if ((v0 & (lin0 != st.cell0)) | (v1 & (lin1 != st.cell1))) {
      run_flush_reg(a.ends, st.cell0, st.run0, lin0, v0);
      run_flush_reg(a.ends, st.cell1, st.run1, lin1, v1);
    }
    st.run0 += v0 ? (hit0 ? 0x10001u : 1u) : 0u;
    st.run1 += v1 ? (hit1 ? 0x10001u : 1u) : 0u;
  }
  if (def) atomicOr(a.defer_bits + (fcur * pa.defer_stride + (idx >> 5)), def << (idx & 31u));
}

// grid = (pair-column blocks, frame groups); a.frames_per_cta <= min(kColFrames, 32767) keeps the
// 16-bit run counters exact.  Every point index of the launch is below 2^32 (host-checked): the
// loop carries ONE 32-bit element offset and takes the plane bases from the parameter block.
template <bool BOUNDED, bool LAB, bool ZGATE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_points_pair(const __grid_constant__ FastArgs a,
                                                                const __grid_constant__ PairArgs pa)
{
  __shared__ uint4 s_rec[kColFrames + 1];  // {element offset of the frame, points, first box, frame * mask_stride - mask_bias}
  const int f0 = a.frame0 + (int)blockIdx.y * a.frames_per_cta;
  int nf = a.frame0 + a.nframes - f0;
  if (nf > a.frames_per_cta) nf = a.frames_per_cta;
  if (nf <= 0) return;
  for (int k = threadIdx.x; k <= nf; k += kThreads) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (k < nf) {
      const uint4 g = __ldg(a.frames + f0 + k);
      r.x = g.x;  // g.y == 0: host-checked
      r.y = g.z;
      r.z = g.w;
      r.w = (unsigned)(f0 + k) * pa.mask_stride - pa.mask_bias;
    }
    s_rec[k] = r;  // entry nf: an empty frame (ends the prefetch chain)
  }
  __syncthreads();

  unsigned idx = 2u * (blockIdx.x * kThreads + threadIdx.x);  // this thread's even point; idx + 1 is its odd point
  unsigned fcur = (unsigned)f0;
  unsigned sa_k = (unsigned)__cvta_generic_to_shared(s_rec);  // this frame's record
  const unsigned sa_end = sa_k + 16u * (unsigned)nf;
  // opaque to the optimiser: otherwise these loop invariants are re-derived from the special
  // registers inside the loop
  asm volatile("" : "+r"(idx), "+r"(sa_k));
  const f32x2 one = bc2(pa.one);
  PairRuns st{-1, -1, 0u, 0u};
  f32x2 nx = 0ull, ny = 0ull, nz = 0ull;
  {
    uint2 r0;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r0.x), "=r"(r0.y) : "r"(sa_k));
    if (idx < r0.y) {
      const unsigned o = r0.x + idx;
      nx = __ldcs(reinterpret_cast<const unsigned long long *>(a.x + o));
      ny = __ldcs(reinterpret_cast<const unsigned long long *>(a.y + o));
      nz = __ldcs(reinterpret_cast<const unsigned long long *>(a.z + o));
    }
  }
  // (Two frames per trip with ping-pong point registers, to drop the six register copies of the
  // software pipeline, was measured slower: under the 48-register budget the in-flight loads spill.)
#pragma unroll 1
  for (; sa_k != sa_end; sa_k += 16u, ++fcur) {
    uint4 rcur;
    uint2 rnxt;  // the next frame's {offset, points} (entry nf: an empty frame)
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rcur.x), "=r"(rcur.y), "=r"(rcur.z), "=r"(rcur.w) : "r"(sa_k));
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+16];" : "=r"(rnxt.x), "=r"(rnxt.y) : "r"(sa_k));
    const f32x2 px = nx, py = ny, pz = nz;
    if (idx < rnxt.y) {  // the next frame's pair: in flight while this one is processed
      const unsigned o = rnxt.x + idx;
      nx = __ldcs(reinterpret_cast<const unsigned long long *>(a.x + o));
      ny = __ldcs(reinterpret_cast<const unsigned long long *>(a.y + o));
      nz = __ldcs(reinterpret_cast<const unsigned long long *>(a.z + o));
    }
    if (idx < rcur.y) pair_body<BOUNDED, LAB, ZGATE>(a, pa, one, idx, fcur, rcur, px, py, pz, st);
  }
  if (st.run0) atomicAdd(a.ends + st.cell0, ((unsigned long long)(st.run0 >> 16) << 32) | (st.run0 & 0xffffu));
  if (st.run1) atomicAdd(a.ends + st.cell1, ((unsigned long long)(st.run1 >> 16) << 32) | (st.run1 & 0xffffu));
}

}  // namespace gv
