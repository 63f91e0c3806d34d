// gv_api.cu — host side of the C ABI declared in include/gridvision_b200.h.
//
// Owns device memory, streams and (optionally) an NCCL communicator; every compute step is
// a kernel from gv_kernels.cuh.  There is deliberately no CPU implementation of any step
// here: if no sm_100 device is usable gv_create fails and nothing else can be called.
#include "gridvision_b200.h"
#include "gv_kernels.cuh"
#include "gv_points_fast.cuh"
#include "gv_points_pair.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#ifdef GV_WITH_NCCL
#include <nccl.h>
#endif

using namespace gv;

namespace {

struct Scratch {
  void *p = nullptr;
  size_t cap = 0;
};

enum {
  S_X, S_Y, S_Z, S_AOS, S_LAB, S_PIX, S_UV, S_BOX_RAW, S_BOX_F4, S_FRAME_OFF, S_BOX_OFF,
  S_TILE_PREFIX, S_TILE_START, S_TILE_END, S_TILE_BOX, S_CELL, S_FLAGS, S_FOOT_IN, S_FOOT_LAB,
  S_RECT, S_SMALL, S_OX, S_OY, S_OZ, S_SCAN0, S_SCAN1, S_SCAN2, S_SCAN3, S_UVZ, S_INDICES,
  S_LABELS_IN, S_MASKS, S_SET_OFF, S_SWEEP, S_SWEEP_PREFIX, S_SWEEP_ITEM, S_BATCH_ENTRY, S_BATCH_MI, S_BATCH_W, S_N2_TAB, S_N2_KEEP, S_N2_OUT, S_N2_OFF, S_N1_PLANES, S_N1_VALID, S_N1_SCORES, S_N1_STATE, S_DEFER, S_DEFER_LIST, S_DEFER_COUNT, S_FRAMES, S_KNN_DEPTH, S_KNN_XYZ, S_COUNT
};

constexpr size_t kPlanePad = 4096;  // slack cells so multi-GPU slabs can be equal-sized

}  // namespace

struct gv_ctx {
  int device = 0;
  int num_sms = 148;
  size_t max_smem_optin = 48 * 1024;  // cudaDevAttrMaxSharedMemoryPerBlockOptin
  size_t persist_bytes = 0;           // persisting-L2 carve-out this context asked for
  size_t max_window = 0;              // cudaDeviceProp::accessPolicyMaxWindowSize
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> events;
  std::string err;

  int ncam = 0;
  CamDev cam[kMaxCam];

  bool has_grid = false;
  GridGeom g{};
  size_t ncells = 0;
  float *d_lo = nullptr, *d_occ = nullptr;
  int32_t *d_hit = nullptr, *d_miss = nullptr;
  int32_t *d_missT = nullptr;  // transposed miss plane of the x-major lines (k_sweep_walk), all-zero between sweeps
  double sweep_reach = 0.0;    // cells a beam binned since the last sweep can extend from the origin (inf: no range cap)
  // Two end-cell planes: beams are binned into d_ends (= d_ends_buf[ends_cur]) on the caller's
  // stream while the previous batch's plane is merged (raycast + [exchange] + finalise) on
  // merge_stream.  ev_merge_done[i] / merge_busy[i]: the merge that consumed plane i.
  unsigned long long *d_ends = nullptr;
  unsigned long long *d_ends_buf[2] = {nullptr, nullptr};
  int ends_cur = 0;
  cudaStream_t merge_stream = nullptr;
  cudaEvent_t ev_bin_done = nullptr, ev_merge_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_merge_t0 = nullptr, ev_merge_t1 = nullptr;  // timing of the last merge
  bool merge_busy[2] = {false, false}, merge_timed = false;
  int overlap = -1;     // $GV_OVERLAP: 1 finalize on merge_stream, 0 on the caller's stream; default (-1): on
                        // merge_stream when there are several ranks (there the merge is latency-bound and
                        // hides under the next batch's binning; on one GPU both are issue-bound and it
                        // only slows the binning kernel down by what it gains)
  unsigned long long merges = 0;
  unsigned *d_list_count = nullptr;  // work-item counter of the raycast sweep
  bool list_count_dirty = true;      // the counters are not known to be zero
  SweepEntry *d_sweep = nullptr;     // sweep table for the current start cell
  unsigned *d_sweep_prefix = nullptr;  // [n_sweep+1] first work item of each entry
  int *d_sweep_item = nullptr;         // [n_sweep_items] entry of each work item
  std::vector<int> h_sweep_D;          // host copies: distance of each entry (non-increasing) ...
  std::vector<unsigned> h_sweep_prefix;  // ... and its first work item
  int n_sweep = 0;
  unsigned n_sweep_items = 0;
  bool counts_dirty = false, ends_dirty = false;
  bool multi_checked = false;  // ranks verified to share geometry + pose (gv_grid_finalize_multi)
  bool use_fast = true;  // $GV_NO_FAST=1 forces the generic k_points (A/B measurements)
  int fast_unroll = 2;   // $GV_FAST_U: points per thread per iteration of k_points_fast
  // CUDA graphs of captured call sequences (gv_graph_*)
  struct GraphSlot {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    size_t nodes = 0;
  };
  std::vector<GraphSlot> graphs;
  bool capturing = false;
  int walk_seg = 64;             // $GV_WALK_SEG: steps per piece when a small sweep's lines are cut (0: never)
  long long walk_seg_max_beams = 4ll << 20;  // sweeps of more beams than this walk whole lines (throughput mode)
  int span_chunks = 32;          // 32-cell chunks per sweep span (32 / 64 / 128), chosen by build_sweep_table
  int span_chunks_env = 0;       // $GV_SPAN_CHUNKS overrides the choice
  int pair_waves = 32;           // $GV_PAIR_WAVES: CTA waves k_points_pair's grid aims at (fewer = longer run-length merging)
  int pair_minb = GV_PAIR_MINB;  // $GV_PAIR_MINB: CTAs per SM k_points_pair's register allocation aims at (3..6)
  bool col_hoist = false; // $GV_COL_HOIST=1: k_points_col keeps FastHot in registers (fewer instructions, 3 CTAs/SM)
  int fast_kind = 0;     // $GV_FAST_KIND: 0 k_points_pair where the layout allows it, else k_points_col (default),
                         // 1 k_points_tma, 2 k_points_fast, 3 k_points_col always
  bool tma_hoist = false; // $GV_TMA_HOIST=1: k_points_tma keeps FastHot in registers instead of the constant bank
  bool use_tma = true;   // $GV_NO_TMA=1: k_points_fast (per-tile CTAs, LDG) instead of k_points_tma
  int fast_agg = -1;     // $GV_FAST_AGG: 0 one RED per beam, 1 match-any groups, 2 adjacent runs;
                         // -1: by plane size (L2-resident plane: 0, larger: 2)
  bool l2_persist = true;  // $GV_L2_PERSIST=0: no persisting-L2 window on the end-cell plane

  bool has_base = false;
  float Tb[16];
  BinDev bin{};

  unsigned long long *d_stats = nullptr;  // beams, logical, physical, lines, deferred points
  unsigned long long launches = 0;
  unsigned long long beams_bound = 0;  // host-side upper bound of beams since last finalize

  Scratch s[S_COUNT];
  // batch tables of the previous gv_process_batch call (frame layout rarely changes between
  // calls of a replay: identical offsets skip the upload, the tile-table kernel and any sync)
  std::vector<unsigned long long> c_foff;
  std::vector<int> c_boff;
  struct Chunk { int f0, f1; unsigned tile0, ntiles; };
  std::vector<Chunk> c_chunks;  // frame groups launched together (one for device-resident points)
  std::vector<unsigned long long> c_tstart, c_tend;  // host images of the tile table (upload source)
  std::vector<int4> c_tbox;
  std::vector<uint4> c_frames;   // per-frame records for k_points_col
  unsigned c_max_pts = 0;        // largest frame (points)
  bool c_pair_ok = false;        // every frame offset and size even (k_points_pair)
  int c_tile_pts = 0, c_max_boxes = 0;
  bool c_on_device = false, c_valid = false;

#ifdef GV_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  int rank = 0, world = 1;
  // peer-memory (cudaIpc) views of every rank's planes; [rank] is the local pointer
  bool p2p = false;
  Peers<unsigned long long> peer_ends{}, peer_ends_buf[2] = {};
  unsigned *d_flags = nullptr;  // [kMaxPeers] arrival epochs of k_peer_barrier, [kMaxPeers] = timeout flag
  Peers<unsigned> peer_flags{};
  unsigned barrier_epoch = 0;
  Peers<int32_t> peer_hit{}, peer_miss{};
  Peers<float> peer_lo{}, peer_occ{};
  int *d_barrier = nullptr;
  // optional stage timing of gv_grid_finalize_multi ($GV_TIMING=1): events + accumulated ms
  bool timing = false;
  cudaEvent_t tev[12] = {};
  int tev_n = 0;
  double t_acc[12] = {};
  long t_cnt = 0;

  int fail(int code, const char *fmt, ...)
  {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define GV_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return ctx->fail(GV_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                       __FILE__, __LINE__);                                                   \
  } while (0)

#define GV_LAUNCH_CHECK()                                                                     \
  do {                                                                                        \
    ctx->launches++;                                                                          \
    GV_CUDA(cudaPeekAtLastError());                                                           \
  } while (0)

#define GV_TRY(expr)                                                                          \
  do {                                                                                        \
    int rc_ = (expr);                                                                         \
    if (rc_ != GV_OK) return rc_;                                                             \
  } while (0)

#define GV_REQUIRE(cond, code, ...)                                                           \
  do {                                                                                        \
    if (!(cond)) return ctx->fail(code, __VA_ARGS__);                                         \
  } while (0)

namespace {

void stage_mark(gv_ctx *ctx, int i)
{
  if (!ctx->timing) return;
  if (!ctx->tev[i]) cudaEventCreate(&ctx->tev[i]);
  cudaEventRecord(ctx->tev[i], ctx->stream);
  if (i + 1 > ctx->tev_n) ctx->tev_n = i + 1;
}

void stage_collect(gv_ctx *ctx)
{
  if (!ctx->timing || ctx->tev_n < 2) return;
  // never block the host here (that would serialise the merge with the next batch's binning):
  // a chain whose last event has not completed yet is simply not sampled
  if (cudaEventQuery(ctx->tev[ctx->tev_n - 1]) != cudaSuccess) return;
  for (int i = 0; i + 1 < ctx->tev_n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->tev[i], ctx->tev[i + 1]) == cudaSuccess) ctx->t_acc[i] += ms;
  }
  ctx->t_cnt++;
}

int reserve(gv_ctx *ctx, int slot, size_t bytes, void **out)
{
  Scratch &s = ctx->s[slot];
  if (bytes == 0) bytes = 16;
  if (s.cap < bytes) {
    GV_REQUIRE(!ctx->capturing, GV_ERR_STATE,
               "a scratch buffer would have to grow during graph capture: run the call sequence once before gv_graph_begin");
    // growing a scratch slot may free memory a previous async kernel still reads
    GV_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->merge_stream && ctx->merge_stream != ctx->stream) GV_CUDA(cudaStreamSynchronize(ctx->merge_stream));
    if (s.p) GV_CUDA(cudaFree(s.p));
    s.p = nullptr;
    s.cap = 0;
    size_t want = bytes + bytes / 8;
    GV_CUDA(cudaMalloc(&s.p, want));
    s.cap = want;
  }
  *out = s.p;
  return GV_OK;
}

template <typename T>
int reserve_t(gv_ctx *ctx, int slot, size_t count, T **out)
{
  void *p = nullptr;
  GV_TRY(reserve(ctx, slot, count * sizeof(T), &p));
  *out = static_cast<T *>(p);
  return GV_OK;
}

cudaEvent_t get_event(gv_ctx *ctx, size_t i)
{
  while (ctx->events.size() <= i) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ctx->events.push_back(e);
  }
  return ctx->events[i];
}

// The caller's stream waits for every merge still in flight on merge_stream (grid state, count
// planes and both end-cell planes are then safe to touch from the caller's stream).
int join_merge(gv_ctx *ctx)
{
  for (int i = 0; i < 2; ++i)
    if (ctx->merge_busy[i]) {
      GV_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_merge_done[i], 0));
      ctx->merge_busy[i] = false;
    }
  return GV_OK;
}

// Before binning into the current end-cell plane: its previous merge (two batches ago) is done.
int bin_begin(gv_ctx *ctx)
{
  const int i = ctx->ends_cur;
  if (ctx->merge_busy[i]) {
    GV_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_merge_done[i], 0));
    ctx->merge_busy[i] = false;
  }
  return GV_OK;
}

// k_peer_barrier gave up waiting for a rank (it raises the word after the arrival slots)
int check_barrier_timeout(gv_ctx *ctx)
{
  if (!ctx->p2p) return GV_OK;
  unsigned flag = 0;
  GV_CUDA(cudaMemcpy(&flag, ctx->d_flags + kMaxPeers, sizeof(flag), cudaMemcpyDeviceToHost));
  GV_REQUIRE(flag == 0u, GV_ERR_STATE, "a peer-memory barrier timed out: some rank did not reach gv_grid_finalize_multi");
  return GV_OK;
}

// The merge of the batch binned so far runs on merge_stream (ordered after the binning, which is
// on the caller's stream), so that the caller's next batch can be binned into the other plane
// meanwhile.  merge_begin redirects the context stream; merge_end restores it and flips planes.
struct MergeScope {
  cudaStream_t saved;
  bool active;
};

int merge_begin(gv_ctx *ctx, MergeScope *sc)
{
  sc->saved = ctx->stream;
  sc->active = (ctx->overlap < 0 ? ctx->world > 1 : ctx->overlap != 0) && ctx->merge_stream != nullptr;
  if (!sc->active) return join_merge(ctx);
  GV_CUDA(cudaEventRecord(ctx->ev_bin_done, ctx->stream));
  GV_CUDA(cudaStreamWaitEvent(ctx->merge_stream, ctx->ev_bin_done, 0));
  ctx->stream = ctx->merge_stream;
  GV_CUDA(cudaEventRecord(ctx->ev_merge_t0, ctx->stream));
  return GV_OK;
}

int merge_end(gv_ctx *ctx, MergeScope *sc, int rc)
{
  if (sc->active) {
    if (rc == GV_OK) {
      cudaEventRecord(ctx->ev_merge_t1, ctx->stream);
      ctx->merge_timed = true;
    }
    const int i = ctx->ends_cur;
    cudaEventRecord(ctx->ev_merge_done[i], ctx->stream);
    ctx->merge_busy[i] = true;
    ctx->stream = sc->saved;
    ctx->ends_cur ^= 1;
    ctx->d_ends = ctx->d_ends_buf[ctx->ends_cur];
    ctx->peer_ends = ctx->peer_ends_buf[ctx->ends_cur];
  }
  ctx->merges++;
  return rc;
}

inline unsigned blocks_for(unsigned long long n, unsigned per_block)
{
  return (unsigned)((n + per_block - 1) / per_block);
}

bool float_exact(double v) { return (double)(float)v == v; }

bool is_canonical_K(const double *K)
{
  return K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0 &&
         float_exact(K[0]) && float_exact(K[2]) && float_exact(K[4]) && float_exact(K[5]);
}

void copy_T12(float dst[12], const float *T16)
{
  for (int i = 0; i < 12; ++i) dst[i] = T16[i];
}

// exclusive scan of n unsigned values in place (recursive CTA scan)
int scan_u32(gv_ctx *ctx, unsigned *d, unsigned long long n, int level)
{
  if (n == 0) return GV_OK;
  const unsigned nb = blocks_for(n, 1024);
  unsigned *sums = nullptr;
  if (nb > 1) GV_TRY(reserve_t(ctx, S_SCAN1 + level, (size_t)nb, &sums));
  k_scan_block<<<nb, 1024, 0, ctx->stream>>>(d, n, sums);
  GV_LAUNCH_CHECK();
  if (nb > 1) {
    GV_REQUIRE(level < 2, GV_ERR_INVALID, "scan_u32: input too large");
    GV_TRY(scan_u32(ctx, sums, nb, level + 1));
    k_scan_add<<<nb, 1024, 0, ctx->stream>>>(d, n, sums);
    GV_LAUNCH_CHECK();
  }
  return GV_OK;
}

// raw_out != NULL: the rounding is left to the caller's k_box_masks launch (the device copy of the
// raw records is returned), one launch less
int upload_boxes(gv_ctx *ctx, const gv_box *boxes, int nboxes, bool on_device, float4 **out,
                 const BoxRaw **raw_out = nullptr)
{
  if (raw_out) *raw_out = nullptr;
  static_assert(sizeof(gv_box) == 40 && sizeof(BoxRaw) == 40, "BoundingBox layout");
  float4 *d_f4 = nullptr;
  GV_TRY(reserve_t(ctx, S_BOX_F4, (size_t)(nboxes > 0 ? nboxes : 1), &d_f4));
  if (nboxes > 0) {
    const BoxRaw *d_raw = reinterpret_cast<const BoxRaw *>(boxes);
    if (!on_device) {
      BoxRaw *tmp = nullptr;
      GV_TRY(reserve_t(ctx, S_BOX_RAW, (size_t)nboxes, &tmp));
      GV_CUDA(cudaMemcpyAsync(tmp, boxes, (size_t)nboxes * sizeof(BoxRaw), cudaMemcpyHostToDevice,
                              ctx->stream));
      d_raw = tmp;
    }
    if (raw_out) {
      *raw_out = d_raw;
    } else {
      k_round_boxes<<<blocks_for(nboxes, 256), 256, 0, ctx->stream>>>(d_raw, nboxes, d_f4);
      GV_LAUNCH_CHECK();
    }
  }
  *out = d_f4;
  return GV_OK;
}

// image-tile prefilter geometry: square tiles of 2^shift pixels, at most 16 per axis
void mask_geometry(int W, int H, int *shift, int *tx, int *ty)
{
  int sh = 4;
  const int m = W > H ? W : H;
  while (((m + (1 << sh) - 1) >> sh) > 16) ++sh;
  *shift = sh;
  *tx = (W + (1 << sh) - 1) >> sh;
  *ty = (H + (1 << sh) - 1) >> sh;
}

int tile_points_for(unsigned long long n, int num_sms, bool big = false, int cap = 4 * kTilePts)
{
  // bigger CTA tiles amortise the box/mask staging; small clouds keep every SM busy
  const unsigned long long per = n / ((unsigned long long)num_sms * 8ull);
  if (big && per >= 64ull * kTilePts) return 16 * kTilePts;  // k_points_fast on large batches
  if (per >= 4ull * kTilePts && cap >= 4 * kTilePts) return 4 * kTilePts;
  if (per >= 2ull * kTilePts) return 2 * kTilePts;
  return kTilePts;
}

int set_bin_params(gv_ctx *ctx, const gv_accum_params *prm, BinDev *out)
{
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_REQUIRE(ctx->has_base, GV_ERR_STATE, "gv_set_base_transform not called");
  GV_REQUIRE(prm != nullptr, GV_ERR_INVALID, "accum params are NULL");
  GV_REQUIRE(prm->occ_mode == GV_OCC_ALL || prm->occ_mode == GV_OCC_LABELLED, GV_ERR_INVALID,
             "bad occ_mode %d", prm->occ_mode);
  *out = ctx->bin;
  out->occ_mode = prm->occ_mode;
  out->use_z_gate = prm->use_z_gate ? 1 : 0;
  out->z_min = prm->z_min;
  out->z_max = prm->z_max;
  out->cap = prm->r_max > 0.0 ? 1 : 0;
  // float geometry of free-space-only beams: the same single-rounded float expressions as the
  // specification (oracle gvo_beam_geom_init); host floats are IEEE binary32, no contraction
  out->rmaxf = (float)prm->r_max;
  out->rmax2f = out->rmaxf * out->rmaxf;
  const double reach = out->cap ? std::ceil((double)out->rmaxf / ctx->g.res) + 4.0 : INFINITY;
  if (reach > ctx->sweep_reach) ctx->sweep_reach = reach;
  return GV_OK;
}

int note_beams(gv_ctx *ctx, unsigned long long n)
{
  const unsigned long long limit = 2147483647ull / (unsigned long long)ctx->world;
  GV_REQUIRE(ctx->beams_bound + n <= limit, GV_ERR_OVERFLOW,
             "more than %llu beams since the last finalize (int32 count planes)", limit);
  ctx->beams_bound += n;
  return GV_OK;
}

// K1/K2 launch for a single cloud (nframes == 0) in device memory
int launch_points(gv_ctx *ctx, bool fuse, bool bin, PointArgs &a, unsigned ntiles, size_t smem)
{
  if (ntiles == 0) return GV_OK;
  const bool exact_uv = a.pix != nullptr || a.uv != nullptr || a.cell_out != nullptr ||
                        a.flags_out != nullptr;
  // COMMON: flags the specialised hot loop assumes (see k_points); checked here, once
  const bool common = fuse && bin && a.ncam == 1 && a.cam[0].has_T && a.cam[0].t_small &&
                      a.cam[0].canon && a.bin.t_small && a.bin.fast_index_ok && a.mask_words == 1;
  if (fuse && bin && common) k_points<true, true, false, false, true><<<ntiles, kThreads, smem, ctx->stream>>>(a);
  else if (fuse && bin) k_points<true, true, false, false, false><<<ntiles, kThreads, smem, ctx->stream>>>(a);
  else if (fuse && a.ncam > 1 && exact_uv) k_points<true, false, true, true, false><<<ntiles, kThreads, smem, ctx->stream>>>(a);
  else if (fuse && a.ncam > 1) k_points<true, false, true, false, false><<<ntiles, kThreads, smem, ctx->stream>>>(a);
  else if (fuse && exact_uv) k_points<true, false, false, true, false><<<ntiles, kThreads, smem, ctx->stream>>>(a);
  else if (fuse) k_points<true, false, false, false, false><<<ntiles, kThreads, smem, ctx->stream>>>(a);
  else if (exact_uv) k_points<false, true, false, true, false><<<ntiles, kThreads, smem, ctx->stream>>>(a);
  else k_points<false, true, false, false, false><<<ntiles, kThreads, smem, ctx->stream>>>(a);
  GV_LAUNCH_CHECK();
  return GV_OK;
}

void fill_point_args_cloud(gv_ctx *ctx, PointArgs &a, const float *x, const float *y,
                           const float *z, size_t n, int is_dense)
{
  memset(&a, 0, sizeof(a));
  a.x = x;
  a.y = y;
  a.z = z;
  a.n = n;
  a.is_dense = is_dense;
  a.vec_ok = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)z) & 15u) == 0;
  a.tile_pts = tile_points_for(n, ctx->num_sms);
  a.smem_boxes = 1;
  a.mask_words = 1;
}

int fuse_dev_impl(gv_ctx *ctx, const float *d_x, const float *d_y, const float *d_z, size_t n,
                  int is_dense, const gv_box *boxes, int nboxes, bool boxes_on_device,
                  const int32_t *box_cam_offsets, int16_t *d_labels, int32_t *d_pix, float *d_uv)
{
  GV_REQUIRE(ctx->ncam > 0, GV_ERR_STATE, "gv_set_cameras not called");
  GV_REQUIRE(nboxes >= 0 && nboxes <= 32767, GV_ERR_INVALID, "nboxes %d out of range", nboxes);
  GV_REQUIRE(nboxes == 0 || boxes != nullptr, GV_ERR_INVALID, "boxes is NULL");
  GV_REQUIRE(ctx->ncam == 1 || box_cam_offsets != nullptr, GV_ERR_INVALID,
             "box_cam_offsets required for %d cameras", ctx->ncam);
  PointArgs a;
  fill_point_args_cloud(ctx, a, d_x, d_y, d_z, n, is_dense);
  a.ncam = ctx->ncam;
  for (int c = 0; c < ctx->ncam; ++c) {
    a.cam[c] = ctx->cam[c];
    a.cam[c].box_begin = box_cam_offsets ? box_cam_offsets[c] : 0;
    a.cam[c].box_end = box_cam_offsets ? box_cam_offsets[c + 1] : nboxes;
    GV_REQUIRE(a.cam[c].box_begin >= 0 && a.cam[c].box_begin <= a.cam[c].box_end &&
                 a.cam[c].box_end <= nboxes,
               GV_ERR_INVALID, "box_cam_offsets[%d..%d] invalid", c, c + 1);
  }
  float4 *d_f4 = nullptr;
  GV_TRY(upload_boxes(ctx, boxes, nboxes, boxes_on_device, &d_f4));
  a.boxes = d_f4;
  a.labels = d_labels;
  a.pix = d_pix;
  a.uv = d_uv;
  // per-camera tile masks
  int max_set = 1, max_tiles = 1;
  std::vector<int> set_off((size_t)ctx->ncam + 1);
  for (int c = 0; c < ctx->ncam; ++c) {
    int ty;
    mask_geometry(a.cam[c].W, a.cam[c].H, &a.mask_shift[c], &a.mask_tx[c], &ty);
    if (a.mask_tx[c] * ty > max_tiles) max_tiles = a.mask_tx[c] * ty;
    const int nb = a.cam[c].box_end - a.cam[c].box_begin;
    if (nb > max_set) max_set = nb;
    set_off[c] = a.cam[c].box_begin;
    GV_REQUIRE(c == 0 || a.cam[c].box_begin == a.cam[c - 1].box_end, GV_ERR_INVALID,
               "box_cam_offsets must be contiguous");
  }
  set_off[ctx->ncam] = a.cam[ctx->ncam - 1].box_end;
  a.mask_words = (max_set + 63) / 64;
  a.mask_stride = max_tiles * a.mask_words;
  a.smem_boxes = nboxes > 0 ? nboxes : 1;
  unsigned long long *d_masks = nullptr;
  int *d_set_off = nullptr;
  GV_TRY(reserve_t(ctx, S_MASKS, (size_t)ctx->ncam * a.mask_stride, &d_masks));
  GV_TRY(reserve_t(ctx, S_SET_OFF, (size_t)ctx->ncam + 1, &d_set_off));
  GV_CUDA(cudaMemcpyAsync(d_set_off, set_off.data(), set_off.size() * sizeof(int),
                          cudaMemcpyHostToDevice, ctx->stream));
  for (int c = 0; c < ctx->ncam; ++c) {
    int sh, tx, ty;
    mask_geometry(a.cam[c].W, a.cam[c].H, &sh, &tx, &ty);
    k_box_masks<<<1, kThreads, 0, ctx->stream>>>(d_f4, d_set_off + c, sh, tx, ty, a.mask_words,
                                                 a.mask_stride, 0, d_masks + (size_t)c * a.mask_stride, nullptr, nullptr);
    GV_LAUNCH_CHECK();
  }
  GV_CUDA(cudaStreamSynchronize(ctx->stream));  // set_off is a host temporary
  a.masks = d_masks;
  const size_t smem = (size_t)a.smem_boxes * sizeof(float4) +
                      (size_t)ctx->ncam * a.mask_stride * sizeof(unsigned long long);
  GV_REQUIRE(smem <= ctx->max_smem_optin, GV_ERR_INVALID,
             "%d boxes need %zu bytes of shared memory for the box/tile-mask stage; this device allows %zu "
             "(about 14000 boxes per call)", nboxes, smem, ctx->max_smem_optin);
  if (smem > 48 * 1024) {
    GV_CUDA(cudaFuncSetAttribute(k_points<true, false, false, false, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GV_CUDA(cudaFuncSetAttribute(k_points<true, false, false, true, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GV_CUDA(cudaFuncSetAttribute(k_points<true, false, true, false, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GV_CUDA(cudaFuncSetAttribute(k_points<true, false, true, true, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  return launch_points(ctx, true, false, a, blocks_for(n, a.tile_pts), smem);
}

int accumulate_dev_impl(gv_ctx *ctx, const float *d_x, const float *d_y, const float *d_z, size_t n,
                        const int16_t *d_labels, const gv_accum_params *prm, int32_t *d_cell,
                        uint8_t *d_flags)
{
  PointArgs a;
  fill_point_args_cloud(ctx, a, d_x, d_y, d_z, n, 0);
  GV_TRY(set_bin_params(ctx, prm, &a.bin));
  GV_REQUIRE(prm->occ_mode != GV_OCC_LABELLED || d_labels != nullptr, GV_ERR_INVALID,
             "GV_OCC_LABELLED needs labels");
  GV_TRY(note_beams(ctx, n));
  GV_TRY(bin_begin(ctx));
  a.labels_in = d_labels;
  a.ends = ctx->d_ends;
  a.cell_out = d_cell;
  a.flags_out = d_flags;
  ctx->ends_dirty = true;
  return launch_points(ctx, false, true, a, blocks_for(n, a.tile_pts), 32);
}

int raycast_flush_impl(gv_ctx *ctx, unsigned rank, unsigned world, bool p2p_gather = false)
{
  if (!ctx->ends_dirty) return GV_OK;
  ctx->ends_dirty = false;
  if (!ctx->bin.origin_ok || ctx->n_sweep_items == 0) return GV_OK;  // nothing was binned
  // the sweep's counters are left zero by its last kernel (k_miss_fold); memset only at the first
  // sweep or after one that was cut short by an error
  if (ctx->list_count_dirty) GV_CUDA(cudaMemsetAsync(ctx->d_list_count, 0, 4 * sizeof(unsigned), ctx->stream));
  ctx->list_count_dirty = true;
  const unsigned nb = (unsigned)ctx->num_sms * 8u;
  // batch list: every non-empty cell once, each span padded to a multiple of 32
  const size_t max_batches = ctx->ncells / 32 + (size_t)ctx->n_sweep_items + 1;
  int *d_bentry, *d_bmi;
  unsigned *d_bw;
  GV_TRY(reserve_t(ctx, S_BATCH_ENTRY, max_batches, &d_bentry));
  GV_TRY(reserve_t(ctx, S_BATCH_MI, max_batches * 32, &d_bmi));
  GV_TRY(reserve_t(ctx, S_BATCH_W, max_batches * 32, &d_bw));
  stage_mark(ctx, 2);
  // Single GPU with a range cap: the table is sorted by decreasing distance, so the entries no
  // beam of this sweep can reach are a prefix of it (multi-GPU: other ranks' caps are not known here)
  unsigned item0 = 0u;
  int reach = 2147483647 / 2;
  if (world == 1 && ctx->sweep_reach < 1.0e9) {
    reach = (int)ctx->sweep_reach;
    size_t lo = 0, hi = ctx->h_sweep_D.size();  // first entry with D <= reach (D is non-increasing)
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (ctx->h_sweep_D[mid] > reach) lo = mid + 1;
      else hi = mid;
    }
    if (lo < ctx->h_sweep_prefix.size()) item0 = ctx->h_sweep_prefix[lo];
  }
  // one CTA per span this rank owns
  unsigned nbc = (unsigned)(((unsigned long long)(ctx->n_sweep_items - item0) + world - 1) / world);
  if (nbc < 1u) nbc = 1u;
#define GV_COMPACT(PP, CC, CLEAR, PEERS)                                                                          \
  k_sweep_compact<PP, CC><<<nbc, kThreads, 0, ctx->stream>>>(                                                       \
    ctx->d_ends, ctx->d_hit, ctx->d_miss, ctx->d_sweep, ctx->d_sweep_prefix, ctx->d_sweep_item, ctx->n_sweep,       \
    ctx->n_sweep_items, item0, reach, ctx->bin.sx, ctx->bin.sy, ctx->g.nx, rank, world, CLEAR, ctx->d_list_count, d_bentry, d_bmi, \
    d_bw, ctx->d_stats, PEERS)
  // p2p_gather: sums (and clears) the cells it owns in every rank's plane over NVLink
  const int clear = p2p_gather ? 1 : (world == 1 ? 1 : 0);
  if (ctx->span_chunks == 128) {
    if (p2p_gather) GV_COMPACT(true, 128, clear, ctx->peer_ends_buf[ctx->ends_cur]);
    else GV_COMPACT(false, 128, clear, ctx->peer_ends);
  } else if (ctx->span_chunks == 64) {
    if (p2p_gather) GV_COMPACT(true, 64, clear, ctx->peer_ends_buf[ctx->ends_cur]);
    else GV_COMPACT(false, 64, clear, ctx->peer_ends);
  } else {
    if (p2p_gather) GV_COMPACT(true, 32, clear, ctx->peer_ends_buf[ctx->ends_cur]);
    else GV_COMPACT(false, 32, clear, ctx->peer_ends);
  }
#undef GV_COMPACT
  GV_LAUNCH_CHECK();
  stage_mark(ctx, 3);
  // small sweeps (a scan or two): cut the lines into pieces so that no warp walks a whole long line alone
  unsigned nseg = 1u;
  const int seg = ctx->walk_seg;
  if (world == 1 && seg > 0 && ctx->beams_bound <= (unsigned long long)ctx->walk_seg_max_beams) {
    int maxD = ctx->h_sweep_D.empty() ? 0 : ctx->h_sweep_D[0];
    if (reach < maxD) maxD = reach;
    nseg = (unsigned)((maxD + seg - 1) / seg);
    if (nseg < 1u) nseg = 1u;
  }
  k_sweep_walk<<<nb, kThreads, 0, ctx->stream>>>(ctx->d_miss, ctx->d_missT, ctx->d_sweep, ctx->d_list_count, d_bentry,
                                                 d_bmi, d_bw, ctx->bin.sx, ctx->bin.sy, ctx->g.nx, ctx->g.ny, nseg, seg,
                                                 ctx->d_stats);
  GV_LAUNCH_CHECK();
  {
    // fold the x-major lines' transposed plane back: they stay within the range cap of the origin
    int x0 = 0, y0 = 0, x1 = ctx->g.nx - 1, y1 = ctx->g.ny - 1;
    // (multi-GPU: this rank also walks lines other ranks binned, under their own caps: whole map)
    if (world == 1) {
      if (ctx->sweep_reach < 1.0e9) {
        const int R = (int)ctx->sweep_reach;
        x0 = std::max(x0, ctx->bin.sx - R); x1 = std::min(x1, ctx->bin.sx + R);
        y0 = std::max(y0, ctx->bin.sy - R); y1 = std::min(y1, ctx->bin.sy + R);
      }
    }
    if (x1 >= x0 && y1 >= y0) {
      const dim3 grid((unsigned)((x1 - x0) / 32 + 1), (unsigned)((y1 - y0) / 32 + 1));
      k_miss_fold<<<grid, 256, 0, ctx->stream>>>(ctx->d_miss, ctx->d_missT, ctx->g.nx, ctx->g.ny, x0, y0, x1, y1,
                                                 ctx->d_list_count);
      GV_LAUNCH_CHECK();
      ctx->list_count_dirty = false;
    }
    ctx->sweep_reach = 0.0;
  }
  stage_mark(ctx, 4);
  // multi-GPU, NCCL path: the plane holds every rank's (all-reduced) entries but this rank
  // walked only its own items: drop the rest.  (P2P: the owners cleared every rank's cells.)
  if (world > 1 && !p2p_gather)
    GV_CUDA(cudaMemsetAsync(ctx->d_ends, 0, ctx->ncells * sizeof(unsigned long long), ctx->stream));
  ctx->counts_dirty = true;
  return GV_OK;
}

// Sweep table for start cell (sx,sy): one entry per (direction, distance) with a non-empty
// clipped segment, longest distance first; item_prefix counts 32-cell work items.
int build_sweep_table(gv_ctx *ctx)
{
  ctx->n_sweep = 0;
  ctx->n_sweep_items = 0;
  if (!ctx->bin.origin_ok) return GV_OK;
  const int nx = ctx->g.nx, ny = ctx->g.ny, sx = ctx->bin.sx, sy = ctx->bin.sy;
  const int maxD = (nx > ny ? nx : ny);
  // span size: see k_sweep_compact
  ctx->span_chunks = maxD > 2048 ? 128 : 32;
  if (ctx->span_chunks_env == 32 || ctx->span_chunks_env == 64 || ctx->span_chunks_env == 128)
    ctx->span_chunks = ctx->span_chunks_env;
  std::vector<SweepEntry> ent;
  ent.reserve(4 * (size_t)maxD + 1);
  // Longest distance first, the four directions interleaved.  (Direction-major order, meant to keep
  // the batches in flight inside one quadrant of the plane, measured slower on every config: C3
  // raycast 0.44 -> 0.49 ms, C5 1.44 -> 1.51 ms.)
  for (int D = maxD; D >= 1; --D) {
    for (int dir = 0; dir < 4; ++dir) {
      SweepEntry e;
      e.dir = dir;
      e.D = D;
      if (dir < 2) {  // x-major: column sx +- D, |dy| <= D
        const int x = dir == 0 ? sx + D : sx - D;
        if (x < 0 || x >= nx) continue;
        e.m0 = sy - D < 0 ? 0 : sy - D;
        e.m1 = sy + D > ny - 1 ? ny - 1 : sy + D;
      } else {  // y-major: row sy +- D, |dx| < D
        const int y = dir == 2 ? sy + D : sy - D;
        if (y < 0 || y >= ny) continue;
        e.m0 = sx - (D - 1) < 0 ? 0 : sx - (D - 1);
        e.m1 = sx + (D - 1) > nx - 1 ? nx - 1 : sx + (D - 1);
      }
      if (e.m1 < e.m0) continue;
      ent.push_back(e);
    }
  }
  ent.push_back(SweepEntry{4, 0, 0, 0});  // the start cell itself (zero-length beams)
  std::vector<unsigned> prefix(ent.size() + 1);
  unsigned long long items = 0;
  for (size_t i = 0; i < ent.size(); ++i) {
    prefix[i] = (unsigned)items;
    items += (unsigned long long)(ent[i].m1 - ent[i].m0) / (unsigned long long)(32 * ctx->span_chunks) + 1ull;
  }
  GV_REQUIRE(items < 4294967295ull, GV_ERR_INVALID, "sweep table too large");
  prefix[ent.size()] = (unsigned)items;
  void *p = nullptr;
  GV_TRY(reserve(ctx, S_SWEEP, ent.size() * sizeof(SweepEntry), &p));
  ctx->d_sweep = static_cast<SweepEntry *>(p);
  GV_TRY(reserve(ctx, S_SWEEP_PREFIX, prefix.size() * sizeof(unsigned), &p));
  ctx->d_sweep_prefix = static_cast<unsigned *>(p);
  GV_CUDA(cudaMemcpyAsync(ctx->d_sweep, ent.data(), ent.size() * sizeof(SweepEntry),
                          cudaMemcpyHostToDevice, ctx->stream));
  GV_CUDA(cudaMemcpyAsync(ctx->d_sweep_prefix, prefix.data(), prefix.size() * sizeof(unsigned),
                          cudaMemcpyHostToDevice, ctx->stream));
  // item -> entry, so that a span's CTA finds its entry with one load
  std::vector<int> item_entry((size_t)items + 1);
  for (size_t i = 0; i < ent.size(); ++i)
    for (unsigned k = prefix[i]; k < prefix[i + 1]; ++k) item_entry[k] = (int)i;
  GV_TRY(reserve(ctx, S_SWEEP_ITEM, item_entry.size() * sizeof(int), &p));
  ctx->d_sweep_item = static_cast<int *>(p);
  GV_CUDA(cudaMemcpyAsync(ctx->d_sweep_item, item_entry.data(), item_entry.size() * sizeof(int),
                          cudaMemcpyHostToDevice, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));  // host vectors die here
  ctx->n_sweep = (int)ent.size();
  ctx->n_sweep_items = (unsigned)items;
  ctx->h_sweep_D.resize(ent.size());
  for (size_t i = 0; i < ent.size(); ++i) ctx->h_sweep_D[i] = ent[i].D;
  ctx->h_sweep_prefix = prefix;
  return GV_OK;
}

// footprints -> device rect list.  mode 0 corners (n x 8), 1 poses (n x 4), 2 points (n x 2 + labels)
int footprint_rects(gv_ctx *ctx, const double *in, const int32_t *labels, int n, int mode,
                    int4 **d_rects_out)
{
  *d_rects_out = nullptr;
  if (n <= 0) return GV_OK;
  GV_REQUIRE(in != nullptr, GV_ERR_INVALID, "footprint array is NULL");
  const int per = mode == 0 ? 8 : (mode == 1 ? 4 : 2);
  double *d_in = nullptr;
  int32_t *d_lab = nullptr;
  int4 *d_rect = nullptr;
  GV_TRY(reserve_t(ctx, S_FOOT_IN, (size_t)n * per, &d_in));
  GV_TRY(reserve_t(ctx, S_RECT, (size_t)n, &d_rect));
  GV_CUDA(cudaMemcpyAsync(d_in, in, (size_t)n * per * sizeof(double), cudaMemcpyHostToDevice,
                          ctx->stream));
  if (mode == 2) {
    GV_REQUIRE(labels != nullptr, GV_ERR_INVALID, "labels is NULL");
    GV_TRY(reserve_t(ctx, S_FOOT_LAB, (size_t)n, &d_lab));
    GV_CUDA(cudaMemcpyAsync(d_lab, labels, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice,
                            ctx->stream));
  }
  k_footprint_rects<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(d_in, d_lab, n, mode, ctx->g, d_rect);
  GV_LAUNCH_CHECK();
  *d_rects_out = d_rect;
  return GV_OK;
}

int finalize_slab(gv_ctx *ctx, int32_t k_decay, const int4 *d_rects, int nfoot, size_t cell0,
                  size_t ncell, bool counts, bool p2p = false)
{
  if (ncell == 0) return GV_OK;
  FinalizeArgs fa;
  fa.world = ctx->world;
  fa.peer_lo = ctx->peer_lo;
  fa.peer_occ = ctx->peer_occ;
  fa.peer_hit = ctx->peer_hit;
  fa.peer_miss = ctx->peer_miss;
  fa.log_odds = ctx->d_lo;
  fa.occupancy = ctx->d_occ;
  fa.hit = ctx->d_hit;
  fa.miss = ctx->d_miss;
  fa.cell0 = cell0;
  fa.ncell = ncell;
  fa.nx = ctx->g.nx;
  fa.decay = (float)k_decay * -0.2f;  // log_odds_decay_, ref: occupancy_grid.hpp:29
  fa.rects = d_rects;
  fa.nfoot = d_rects ? nfoot : 0;
  const unsigned nb = blocks_for(ncell, kThreads * 4);
  if (p2p) k_finalize<true, true><<<nb, kThreads, 0, ctx->stream>>>(fa);
  else if (counts) k_finalize<true, false><<<nb, kThreads, 0, ctx->stream>>>(fa);
  else k_finalize<false, false><<<nb, kThreads, 0, ctx->stream>>>(fa);
  GV_LAUNCH_CHECK();
  return GV_OK;
}

int finalize_impl(gv_ctx *ctx, int32_t k_decay, const double *in, const int32_t *labels, int n,
                  int mode)
{
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_REQUIRE(n >= 0, GV_ERR_INVALID, "negative footprint count");
  GV_TRY(raycast_flush_impl(ctx, 0, 1));
  int4 *d_rects = nullptr;
  GV_TRY(footprint_rects(ctx, in, labels, n, mode, &d_rects));
  GV_TRY(finalize_slab(ctx, k_decay, d_rects, n, 0, ctx->ncells, ctx->counts_dirty));
  ctx->counts_dirty = false;
  ctx->beams_bound = 0;
  return GV_OK;
}

void close_peers(gv_ctx *ctx)
{
  if (!ctx->p2p) return;
  for (int r = 0; r < ctx->world && r < kMaxPeers; ++r) {
    if (r == ctx->rank) continue;
    if (ctx->peer_ends_buf[0].p[r]) cudaIpcCloseMemHandle(ctx->peer_ends_buf[0].p[r]);
    if (ctx->peer_ends_buf[1].p[r]) cudaIpcCloseMemHandle(ctx->peer_ends_buf[1].p[r]);
    if (ctx->peer_flags.p[r]) cudaIpcCloseMemHandle(ctx->peer_flags.p[r]);
    if (ctx->peer_hit.p[r]) cudaIpcCloseMemHandle(ctx->peer_hit.p[r]);
    if (ctx->peer_miss.p[r]) cudaIpcCloseMemHandle(ctx->peer_miss.p[r]);
    if (ctx->peer_lo.p[r]) cudaIpcCloseMemHandle(ctx->peer_lo.p[r]);
    if (ctx->peer_occ.p[r]) cudaIpcCloseMemHandle(ctx->peer_occ.p[r]);
  }
  ctx->peer_ends = {};
  ctx->peer_ends_buf[0] = {};
  ctx->peer_ends_buf[1] = {};
  ctx->peer_flags = {};
  ctx->peer_hit = {};
  ctx->peer_miss = {};
  ctx->peer_lo = {};
  ctx->peer_occ = {};
  ctx->p2p = false;
  cudaGetLastError();
}

void free_grid(gv_ctx *ctx)
{
  close_peers(ctx);
  cudaFree(ctx->d_lo);
  cudaFree(ctx->d_occ);
  cudaFree(ctx->d_hit);
  cudaFree(ctx->d_miss);
  cudaFree(ctx->d_missT);
  ctx->d_missT = nullptr;
  cudaFree(ctx->d_ends_buf[0]);
  cudaFree(ctx->d_ends_buf[1]);
  ctx->d_ends_buf[0] = ctx->d_ends_buf[1] = nullptr;
  ctx->d_lo = ctx->d_occ = nullptr;
  ctx->d_hit = ctx->d_miss = nullptr;
  ctx->d_ends = nullptr;
  ctx->has_grid = false;
}

int refresh_origin(gv_ctx *ctx)
{
  ctx->multi_checked = false;
  // start cell + continuous index coordinates of the sensor origin, computed by the same
  // device getIndex every beam uses
  if (!(ctx->has_grid && ctx->has_base)) return GV_OK;
  OriginOut *d_o = nullptr;
  GV_TRY(reserve_t(ctx, S_SMALL, 1, &d_o));
  const double ox = (double)ctx->Tb[3], oy = (double)ctx->Tb[7];
  k_origin_setup<<<1, 1, 0, ctx->stream>>>(ox, oy, ctx->g, d_o);
  GV_LAUNCH_CHECK();
  OriginOut o;
  GV_CUDA(cudaMemcpyAsync(&o, d_o, sizeof(o), cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  BinDev &b = ctx->bin;
  const GridGeom &g = ctx->g;
  copy_T12(b.T, ctx->Tb);
  b.g = g;
  b.sx = o.sx;
  b.sy = o.sy;
  b.origin_ok = o.ok;
  b.oxf = ctx->Tb[3];
  b.oyf = ctx->Tb[7];
  b.c0xf = (float)(0.5 * g.len_x + g.pos_x);
  b.c0yf = (float)(0.5 * g.len_y + g.pos_y);
  b.inv_resf = (float)(1.0 / g.res);
  const float dxo = b.c0xf - b.oxf, dyo = b.c0yf - b.oyf;
  b.oaxf = dxo * b.inv_resf;
  b.oayf = dyo * b.inv_resf;
  // certified fixed-point index (grid_get_index_cert): needs 8*2^-53*(|p|+half+|pos|)/res < 2^-21
  // for every in-map p, i.e. (half + |pos|)/res < ~2^28; otherwise always the exact path
  b.c0xd = g.half_x + g.pos_x;
  b.c0yd = g.half_y + g.pos_y;
  b.mres = 65536.0 / g.res;
  b.klim_x = g.nx << 16;
  b.klim_y = g.ny << 16;
  const double span = (g.half_x + std::fabs(g.pos_x) > g.half_y + std::fabs(g.pos_y)
                         ? g.half_x + std::fabs(g.pos_x) : g.half_y + std::fabs(g.pos_y)) / g.res;
  b.fast_index_ok = (span < 134217728.0 && g.nx <= 16384 && g.ny <= 16384) ? 1 : 0;
  b.t_small = 1;
  for (int i = 0; i < 12; ++i)
    if (!(std::fabs(b.T[i]) < 1.0e6f)) b.t_small = 0;
  return build_sweep_table(ctx);
}

int grid_init_impl(gv_ctx *ctx, double length_x, double length_y, double res, double pos_x,
                   double pos_y)
{
  GV_REQUIRE(res > 0.0 && length_x > 0.0 && length_y > 0.0, GV_ERR_INVALID,
             "grid geometry must be positive");
  // grid_map::GridMap::setGeometry: size = (int)round(length/resolution); length = size*res
  const double sx = std::round(length_x / res), sy = std::round(length_y / res);
  GV_REQUIRE(sx >= 1.0 && sy >= 1.0 && sx * sy <= 2147483647.0, GV_ERR_INVALID,
             "grid of %.0f x %.0f cells unsupported", sx, sy);
  GV_TRY(join_merge(ctx));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  free_grid(ctx);
  GridGeom &g = ctx->g;
  g.nx = (int)sx;
  g.ny = (int)sy;
  g.res = res;
  g.len_x = (double)g.nx * res;
  g.len_y = (double)g.ny * res;
  g.pos_x = pos_x;
  g.pos_y = pos_y;
  g.half_x = 0.5 * g.len_x;
  g.half_y = 0.5 * g.len_y;
  ctx->ncells = (size_t)g.nx * (size_t)g.ny;
  const size_t padded = ctx->ncells + kPlanePad;
  GV_CUDA(cudaMalloc(&ctx->d_lo, padded * sizeof(float)));
  GV_CUDA(cudaMalloc(&ctx->d_occ, padded * sizeof(float)));
  GV_CUDA(cudaMalloc(&ctx->d_hit, padded * sizeof(int32_t)));
  GV_CUDA(cudaMalloc(&ctx->d_miss, padded * sizeof(int32_t)));
  GV_CUDA(cudaMalloc(&ctx->d_missT, padded * sizeof(int32_t)));
  GV_CUDA(cudaMemsetAsync(ctx->d_missT, 0, padded * sizeof(int32_t), ctx->stream));
  GV_CUDA(cudaMalloc(&ctx->d_ends_buf[0], padded * sizeof(unsigned long long)));
  GV_CUDA(cudaMalloc(&ctx->d_ends_buf[1], padded * sizeof(unsigned long long)));
  ctx->ends_cur = 0;
  ctx->d_ends = ctx->d_ends_buf[0];
  ctx->has_grid = true;
  GV_CUDA(cudaMemsetAsync(ctx->d_lo, 0, padded * sizeof(float), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_occ, 0, padded * sizeof(float), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_hit, 0, padded * sizeof(int32_t), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_miss, 0, padded * sizeof(int32_t), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_ends_buf[0], 0, padded * sizeof(unsigned long long), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_ends_buf[1], 0, padded * sizeof(unsigned long long), ctx->stream));
  // ref: src/occupancy_grid.cpp:12-13 log_odds = log_odds_prior_ (0.0f), occupancy = 0.5f
  k_fill_f32<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(ctx->d_occ, ctx->ncells, 0.5f);
  GV_LAUNCH_CHECK();
  ctx->counts_dirty = ctx->ends_dirty = false;
  ctx->beams_bound = 0;
  GV_CUDA(cudaMemsetAsync(ctx->d_stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
  return refresh_origin(ctx);
}

}  // namespace

// =====================================================================================
// lifecycle
// =====================================================================================
extern "C" {

int gv_version(void) { return GV_VERSION; }

const char *gv_status_string(int status)
{
  switch (status) {
  case GV_OK: return "ok";
  case GV_ERR_INVALID: return "invalid argument";
  case GV_ERR_CUDA: return "CUDA error";
  case GV_ERR_STATE: return "invalid call order";
  case GV_ERR_NCCL: return "NCCL error";
  case GV_ERR_OVERFLOW: return "count overflow";
  case GV_ERR_NO_DEVICE: return "no usable sm_100 device";
  default: return "unknown status";
  }
}

int gv_create(gv_ctx **out, int device)
{
  if (!out) return GV_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return GV_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GV_ERR_NO_DEVICE;
  // the library carries sm_100a SASS only: refuse anything else instead of failing at launch
  if (prop.major != 10) return GV_ERR_NO_DEVICE;
  gv_ctx *ctx = new (std::nothrow) gv_ctx();
  if (!ctx) return GV_ERR_INVALID;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->max_smem_optin = prop.sharedMemPerBlockOptin;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_bin_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_merge_done[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_merge_done[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&ctx->ev_merge_t0) != cudaSuccess || cudaEventCreate(&ctx->ev_merge_t1) != cudaSuccess ||
      cudaMalloc(&ctx->d_flags, (kMaxPeers + 1) * sizeof(unsigned)) != cudaSuccess ||
      cudaMemset(ctx->d_flags, 0, (kMaxPeers + 1) * sizeof(unsigned)) != cudaSuccess ||
      cudaMalloc(&ctx->d_stats, 8 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMalloc(&ctx->d_list_count, 4 * sizeof(unsigned)) != cudaSuccess ||
      cudaMemset(ctx->d_stats, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) {
    cudaGetLastError();
    gv_destroy(ctx);
    return GV_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  {
    // the merge of batch i runs beside the binning of batch i+1: give it the higher priority so
    // its (few, latency-bound) CTAs are placed as soon as resources free up
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&ctx->merge_stream, cudaStreamNonBlocking, hi) != cudaSuccess) {
      cudaGetLastError();
      gv_destroy(ctx);
      return GV_ERR_CUDA;
    }
  }
  if (const char *u = std::getenv("GV_OVERLAP")) ctx->overlap = std::atoi(u) != 0 ? 1 : 0;
  ctx->timing = std::getenv("GV_TIMING") != nullptr;
  ctx->use_fast = std::getenv("GV_NO_FAST") == nullptr;
  if (const char *u = std::getenv("GV_FAST_U")) ctx->fast_unroll = std::atoi(u);
  if (const char *u = std::getenv("GV_FAST_AGG")) ctx->fast_agg = std::atoi(u);
  ctx->use_tma = std::getenv("GV_NO_TMA") == nullptr;
  if (const char *u = std::getenv("GV_FAST_KIND")) ctx->fast_kind = std::atoi(u);
  if (const char *u = std::getenv("GV_COL_HOIST")) ctx->col_hoist = std::atoi(u) != 0;
  if (const char *u = std::getenv("GV_PAIR_MINB")) ctx->pair_minb = std::atoi(u);
  if (ctx->pair_minb < 3 || ctx->pair_minb > 6) ctx->pair_minb = GV_PAIR_MINB;
  if (const char *u = std::getenv("GV_SPAN_CHUNKS")) ctx->span_chunks_env = std::atoi(u);
  if (const char *u = std::getenv("GV_WALK_SEG")) ctx->walk_seg = std::atoi(u);
  if (ctx->walk_seg < 0 || ctx->walk_seg > 16384) ctx->walk_seg = 64;
  if (const char *u = std::getenv("GV_PAIR_WAVES")) ctx->pair_waves = std::atoi(u);
  if (ctx->pair_waves < 1 || ctx->pair_waves > 64) ctx->pair_waves = 32;
  if (ctx->fast_kind != 1) ctx->use_tma = false;
  if (const char *u = std::getenv("GV_TMA_HOIST")) ctx->tma_hoist = std::atoi(u) != 0;
  if (const char *u = std::getenv("GV_L2_PERSIST")) ctx->l2_persist = std::atoi(u) != 0;
  if (ctx->l2_persist && prop.persistingL2CacheMaxSize > 0) {
    // room for the end-cell plane of a 2048 x 2048 map (32 MiB) plus slack; larger planes get a
    // proportional hit ratio (set_ends_window)
    size_t want = (size_t)48 << 20;
    if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
      ctx->persist_bytes = want;
      ctx->max_window = (size_t)prop.accessPolicyMaxWindowSize;
    }
    cudaGetLastError();
  }
  *out = ctx;
  return GV_OK;
}

void gv_destroy(gv_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->timing && ctx->t_cnt > 0 && ctx->rank == 0) {
    std::fprintf(stderr, "[grid_vision_b200] finalize_multi stages (ms, mean of %ld, %s):", ctx->t_cnt,
                 ctx->p2p ? "p2p" : "nccl");
    for (int i = 0; i + 1 < ctx->tev_n; ++i) std::fprintf(stderr, " %.3f", ctx->t_acc[i] / ctx->t_cnt);
    std::fprintf(stderr, "\n");
  }
  for (auto &e : ctx->tev)
    if (e) cudaEventDestroy(e);
  for (auto &g : ctx->graphs)
    if (g.exec) {
      cudaGraphExecDestroy(g.exec);
      cudaGraphDestroy(g.graph);
    }
#ifdef GV_WITH_NCCL
  if (ctx->comm) ncclCommDestroy(ctx->comm);
#endif
  free_grid(ctx);
  for (auto &s : ctx->s) cudaFree(s.p);
  cudaFree(ctx->d_stats);
  cudaFree(ctx->d_list_count);
  cudaFree(ctx->d_barrier);
  for (auto e : ctx->events) cudaEventDestroy(e);
  if (ctx->merge_stream) cudaStreamDestroy(ctx->merge_stream);
  if (ctx->ev_bin_done) cudaEventDestroy(ctx->ev_bin_done);
  for (auto e : ctx->ev_merge_done)
    if (e) cudaEventDestroy(e);
  if (ctx->ev_merge_t0) cudaEventDestroy(ctx->ev_merge_t0);
  if (ctx->ev_merge_t1) cudaEventDestroy(ctx->ev_merge_t1);
  cudaFree(ctx->d_flags);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  cudaGetLastError();
  delete ctx;
}

const char *gv_last_error(const gv_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gv_synchronize(gv_ctx *ctx)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_TRY(join_merge(ctx));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return check_barrier_timeout(ctx);
}

int gv_join(gv_ctx *ctx)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  return join_merge(ctx);
}

void *gv_stream(gv_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int gv_set_stream(gv_ctx *ctx, void *stream)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_TRY(join_merge(ctx));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  // the handle is used as given: NULL is the legacy default stream (torch's default current
  // stream), not "reset"; gv_stream() of a fresh context returns its private stream
  ctx->stream = (cudaStream_t)stream;
  return GV_OK;
}

static void pair_axis_thresholds(double W, double e6, double e0, float *half_out, float *e_out, float *ain, float *aout);

// Host-only probe (no device, no context): the thresholds k_points_pair's certified image test uses
// for an image axis of `size` pixels with principal point c, exactly as fill_fast_args /
// fill_pair_args compute them.  out = {half, e, ain, aout, e6, e0}.  The CPU tests check the error
// analysis against these very numbers (tests/test_pair_kernel_bounds.py).
int gv_debug_pair_thresholds(double size, float c, float *out)
{
  if (!out || !(size > 0.0)) return GV_ERR_INVALID;
  const float u22 = 2.384185791015625e-07f;
  const float e6 = 6.0f * u22;
  const float e0 = u22 * (1.5f * std::fabs(c) + 1.0f) * 1.0001f;
  pair_axis_thresholds((double)(float)size, (double)e6, (double)e0, &out[0], &out[1], &out[2], &out[3]);
  out[4] = e6;
  out[5] = e0;
  return GV_OK;
}

// ---- CUDA graphs: a captured sequence of device-pointer calls replayed with one launch ----
int gv_graph_begin(gv_ctx *ctx)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(!ctx->capturing, GV_ERR_STATE, "gv_graph_begin: a capture is already open");
  GV_REQUIRE(ctx->world == 1, GV_ERR_STATE, "gv_graph_begin: single-GPU contexts only");
  GV_REQUIRE(ctx->stream != nullptr && ctx->stream != cudaStreamLegacy && ctx->stream != cudaStreamPerThread,
             GV_ERR_STATE, "gv_graph_begin: the context runs on the default stream, which cannot be captured "
                           "(gv_set_stream with a created stream, or keep the context's own)");
  GV_TRY(join_merge(ctx));
  GV_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = true;
  return GV_OK;
}

int gv_graph_end(gv_ctx *ctx, int32_t *graph_id)
{
  if (!ctx || !graph_id) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->capturing, GV_ERR_STATE, "gv_graph_end without gv_graph_begin");
  ctx->capturing = false;
  gv_ctx::GraphSlot g;
  const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g.graph);
  if (e != cudaSuccess || g.graph == nullptr) {
    cudaGetLastError();
    return ctx->fail(GV_ERR_CUDA, "graph capture failed (%s): a captured call synchronised, allocated or copied "
                                   "from pageable memory; run the sequence once before capturing",
                     cudaGetErrorString(e));
  }
  GV_CUDA(cudaGraphGetNodes(g.graph, nullptr, &g.nodes));
  if (cudaGraphInstantiate(&g.exec, g.graph, 0) != cudaSuccess) {
    cudaGraphDestroy(g.graph);
    cudaGetLastError();
    return ctx->fail(GV_ERR_CUDA, "cudaGraphInstantiate failed");
  }
  ctx->graphs.push_back(g);
  *graph_id = (int32_t)ctx->graphs.size() - 1;
  return GV_OK;
}

int gv_graph_launch(gv_ctx *ctx, int32_t graph_id)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_REQUIRE(graph_id >= 0 && (size_t)graph_id < ctx->graphs.size() && ctx->graphs[graph_id].exec, GV_ERR_INVALID,
             "no such graph %d", graph_id);
  GV_REQUIRE(!ctx->capturing, GV_ERR_STATE, "gv_graph_launch during a capture");
  GV_CUDA(cudaGraphLaunch(ctx->graphs[graph_id].exec, ctx->stream));
  ctx->launches += ctx->graphs[graph_id].nodes;
  return GV_OK;
}

int gv_graph_destroy(gv_ctx *ctx, int32_t graph_id)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_REQUIRE(graph_id >= 0 && (size_t)graph_id < ctx->graphs.size(), GV_ERR_INVALID, "no such graph %d", graph_id);
  gv_ctx::GraphSlot &g = ctx->graphs[graph_id];
  if (g.exec) {
    GV_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaGraphExecDestroy(g.exec);
    cudaGraphDestroy(g.graph);
    g.exec = nullptr;
    g.graph = nullptr;
  }
  return GV_OK;
}

int gv_get_stats(gv_ctx *ctx, gv_stats *out)
{
  if (!ctx || !out) return GV_ERR_INVALID;
  GV_TRY(join_merge(ctx));
  unsigned long long h[5];
  GV_CUDA(cudaMemcpyAsync(h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  out->merges = ctx->merges;
  out->merge_ms_last = 0.0;
  if (ctx->merge_timed) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_merge_t0, ctx->ev_merge_t1) == cudaSuccess) out->merge_ms_last = ms;
    cudaGetLastError();
  }
  out->beams = h[0];
  out->cells_logical = h[1];
  out->cells_physical = h[2];
  out->distinct_ends = h[3];
  out->deferred_points = h[4];
  out->kernel_launches = ctx->launches;
  return GV_OK;
}

// =====================================================================================
// fusion
// =====================================================================================
int gv_set_cameras(gv_ctx *ctx, int ncam, const double *K, const float *T_cam_lidar,
                   const int32_t *wh)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_REQUIRE(ncam >= 1 && ncam <= kMaxCam, GV_ERR_INVALID, "ncam %d not in [1,%d]", ncam, kMaxCam);
  GV_REQUIRE(K && wh, GV_ERR_INVALID, "K / wh is NULL");
  for (int c = 0; c < ncam; ++c) {
    GV_REQUIRE(wh[2 * c] > 0 && wh[2 * c + 1] > 0, GV_ERR_INVALID, "camera %d image size", c);
    CamDev &d = ctx->cam[c];
    memset(&d, 0, sizeof(d));
    for (int i = 0; i < 9; ++i) d.K[i] = K[9 * c + i];
    d.has_T = T_cam_lidar != nullptr;
    if (d.has_T) copy_T12(d.T, T_cam_lidar + 16 * c);
    d.W = wh[2 * c];
    d.H = wh[2 * c + 1];
    d.Wf = (float)d.W;
    d.Hf = (float)d.H;
    d.canon = is_canonical_K(d.K) ? 1 : 0;
    d.t_small = 1;
    for (int i = 0; i < 12; ++i)
      if (!(std::fabs(d.T[i]) < 1.0e6f)) d.t_small = 0;
    for (int i = 0; i < 9; ++i)
      if (!(std::fabs(d.K[i]) < 1.0e6)) d.canon = 0;  // fast float path needs sane intrinsics
    // certified float projection constants: E(q) = e6*|q| + e0 = 2^-22 (6|q| + |c| + 1)
    d.fxf = (float)d.K[0];
    d.fyf = (float)d.K[4];
    d.cxf = (float)d.K[2];
    d.cyf = (float)d.K[5];
    d.e6 = 6.0f * 2.384185791015625e-07f;
    d.e0u = 2.384185791015625e-07f * (std::fabs(d.cxf) + 1.0f) * 1.0001f;
    d.e0v = 2.384185791015625e-07f * (std::fabs(d.cyf) + 1.0f) * 1.0001f;
  }
  ctx->ncam = ncam;
  return GV_OK;
}

int gv_fuse_dev(gv_ctx *ctx, const float *d_x, const float *d_y, const float *d_z, size_t n,
                int is_dense, const gv_box *d_boxes, int nboxes, const int32_t *box_cam_offsets,
                int16_t *d_labels_out, int32_t *d_pix_out, float *d_uv_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(n == 0 || (d_x && d_y && d_z), GV_ERR_INVALID, "point planes are NULL");
  return fuse_dev_impl(ctx, d_x, d_y, d_z, n, is_dense, d_boxes, nboxes, true, box_cam_offsets,
                       d_labels_out, d_pix_out, d_uv_out);
}

static int fuse_host_impl(gv_ctx *ctx, const float *d_x, const float *d_y, const float *d_z,
                          size_t n, int is_dense, const gv_box *boxes, int nboxes,
                          const int32_t *box_cam_offsets, int16_t *labels_out, int32_t *pix_out,
                          float *uv_out)
{
  const size_t planes = (size_t)ctx->ncam * n;
  int16_t *d_lab = nullptr;
  int32_t *d_pix = nullptr;
  float *d_uv = nullptr;
  if (labels_out) GV_TRY(reserve_t(ctx, S_LAB, planes, &d_lab));
  if (pix_out) GV_TRY(reserve_t(ctx, S_PIX, planes, &d_pix));
  if (uv_out) GV_TRY(reserve_t(ctx, S_UV, 2 * planes, &d_uv));
  GV_TRY(fuse_dev_impl(ctx, d_x, d_y, d_z, n, is_dense, boxes, nboxes, false, box_cam_offsets,
                       d_lab, d_pix, d_uv));
  if (labels_out)
    GV_CUDA(cudaMemcpyAsync(labels_out, d_lab, planes * sizeof(int16_t), cudaMemcpyDeviceToHost,
                            ctx->stream));
  if (pix_out)
    GV_CUDA(cudaMemcpyAsync(pix_out, d_pix, planes * sizeof(int32_t), cudaMemcpyDeviceToHost,
                            ctx->stream));
  if (uv_out)
    GV_CUDA(cudaMemcpyAsync(uv_out, d_uv, 2 * planes * sizeof(float), cudaMemcpyDeviceToHost,
                            ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

static int upload_cloud(gv_ctx *ctx, const float *x, const float *y, const float *z, size_t n,
                        float **d_x, float **d_y, float **d_z)
{
  GV_REQUIRE(n == 0 || (x && y && z), GV_ERR_INVALID, "point planes are NULL");
  GV_TRY(reserve_t(ctx, S_X, n, d_x));
  GV_TRY(reserve_t(ctx, S_Y, n, d_y));
  GV_TRY(reserve_t(ctx, S_Z, n, d_z));
  if (n) {
    GV_CUDA(cudaMemcpyAsync(*d_x, x, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    GV_CUDA(cudaMemcpyAsync(*d_y, y, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    GV_CUDA(cudaMemcpyAsync(*d_z, z, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  }
  return GV_OK;
}

int gv_fuse(gv_ctx *ctx, const float *x, const float *y, const float *z, size_t n, int is_dense,
            const gv_box *boxes, int nboxes, const int32_t *box_cam_offsets, int16_t *labels_out,
            int32_t *pix_out, float *uv_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->ncam > 0, GV_ERR_STATE, "gv_set_cameras not called");
  float *d_x, *d_y, *d_z;
  GV_TRY(upload_cloud(ctx, x, y, z, n, &d_x, &d_y, &d_z));
  return fuse_host_impl(ctx, d_x, d_y, d_z, n, is_dense, boxes, nboxes, box_cam_offsets,
                        labels_out, pix_out, uv_out);
}

int gv_fuse_aos32(gv_ctx *ctx, const gv_point_xyzi *pts, size_t n, int is_dense,
                  const gv_box *boxes, int nboxes, const int32_t *box_cam_offsets,
                  int16_t *labels_out, int32_t *pix_out, float *uv_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->ncam > 0, GV_ERR_STATE, "gv_set_cameras not called");
  GV_REQUIRE(n == 0 || pts, GV_ERR_INVALID, "pts is NULL");
  static_assert(sizeof(gv_point_xyzi) == 32, "pcl::PointXYZI layout");
  float4 *d_aos = nullptr;
  float *d_x, *d_y, *d_z;
  GV_TRY(reserve_t(ctx, S_AOS, 2 * n, &d_aos));
  GV_TRY(reserve_t(ctx, S_X, n, &d_x));
  GV_TRY(reserve_t(ctx, S_Y, n, &d_y));
  GV_TRY(reserve_t(ctx, S_Z, n, &d_z));
  if (n) {
    GV_CUDA(cudaMemcpyAsync(d_aos, pts, n * sizeof(gv_point_xyzi), cudaMemcpyHostToDevice,
                            ctx->stream));
    k_aos32_to_soa<<<blocks_for(n, kThreads), kThreads, 0, ctx->stream>>>(d_aos, n, d_x, d_y, d_z);
    GV_LAUNCH_CHECK();
  }
  return fuse_host_impl(ctx, d_x, d_y, d_z, n, is_dense, boxes, nboxes, box_cam_offsets,
                        labels_out, pix_out, uv_out);
}

int gv_transform_points(gv_ctx *ctx, int cam, const float *x, const float *y, const float *z,
                        size_t n, int is_dense, float *ox, float *oy, float *oz)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(cam >= 0 && cam < ctx->ncam, GV_ERR_STATE, "camera %d not set", cam);
  GV_REQUIRE(ctx->cam[cam].has_T, GV_ERR_STATE, "camera %d has no extrinsic", cam);
  GV_REQUIRE(n == 0 || (ox && oy && oz), GV_ERR_INVALID, "output planes are NULL");
  float *d_x, *d_y, *d_z, *d_ox, *d_oy, *d_oz;
  GV_TRY(upload_cloud(ctx, x, y, z, n, &d_x, &d_y, &d_z));
  GV_TRY(reserve_t(ctx, S_OX, n, &d_ox));
  GV_TRY(reserve_t(ctx, S_OY, n, &d_oy));
  GV_TRY(reserve_t(ctx, S_OZ, n, &d_oz));
  if (n) {
    k_transform<<<blocks_for(n, kThreads), kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, n, is_dense,
                                                                      ctx->cam[cam], d_ox, d_oy,
                                                                      d_oz);
    GV_LAUNCH_CHECK();
    GV_CUDA(cudaMemcpyAsync(ox, d_ox, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    GV_CUDA(cudaMemcpyAsync(oy, d_oy, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    GV_CUDA(cudaMemcpyAsync(oz, d_oz, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

int gv_project_kdtree(gv_ctx *ctx, int cam, const float *x, const float *y, const float *z,
                      size_t n, float *uvz_out, size_t *m_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(cam >= 0 && cam < ctx->ncam, GV_ERR_STATE, "camera %d not set", cam);
  GV_REQUIRE(m_out != nullptr, GV_ERR_INVALID, "m_out is NULL");
  *m_out = 0;
  if (n == 0) return GV_OK;
  GV_REQUIRE(uvz_out != nullptr, GV_ERR_INVALID, "uvz_out is NULL");
  float *d_x, *d_y, *d_z, *d_uvz;
  unsigned *d_cnt;
  GV_TRY(upload_cloud(ctx, x, y, z, n, &d_x, &d_y, &d_z));
  const unsigned nb = blocks_for(n, kThreads);
  GV_TRY(reserve_t(ctx, S_SCAN0, (size_t)nb + 1, &d_cnt));
  GV_TRY(reserve_t(ctx, S_UVZ, 3 * n, &d_uvz));
  GV_CUDA(cudaMemsetAsync(d_cnt + nb, 0, sizeof(unsigned), ctx->stream));
  k_kdtree_count<<<nb, kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, n, 0, ctx->cam[cam], d_cnt);
  GV_LAUNCH_CHECK();
  GV_TRY(scan_u32(ctx, d_cnt, (unsigned long long)nb + 1, 0));
  k_kdtree_scatter<<<nb, kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, n, 0, ctx->cam[cam], d_cnt,
                                                     d_uvz);
  GV_LAUNCH_CHECK();
  unsigned total = 0;
  GV_CUDA(cudaMemcpyAsync(&total, d_cnt + nb, sizeof(unsigned), cudaMemcpyDeviceToHost,
                          ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (total)
    GV_CUDA(cudaMemcpyAsync(uvz_out, d_uvz, (size_t)total * 3 * sizeof(float),
                            cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  *m_out = total;
  return GV_OK;
}

int gv_partition_by_label(gv_ctx *ctx, const int16_t *labels, size_t n, int nboxes,
                          uint32_t *indices_out, uint64_t *offsets_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(nboxes >= 0 && nboxes <= 1024, GV_ERR_INVALID, "nboxes %d not in [0,1024]", nboxes);
  GV_REQUIRE(offsets_out != nullptr, GV_ERR_INVALID, "offsets_out is NULL");
  GV_REQUIRE(n < 4294967296ull, GV_ERR_INVALID, "n too large for 32-bit indices");
  for (int b = 0; b <= nboxes; ++b) offsets_out[b] = 0;
  if (n == 0 || nboxes == 0) return GV_OK;
  GV_REQUIRE(labels && indices_out, GV_ERR_INVALID, "labels / indices_out is NULL");
  const unsigned nb = blocks_for(n, kThreads);
  const unsigned long long hn = (unsigned long long)nboxes * nb;
  int16_t *d_lab;
  unsigned *d_hist, *d_idx;
  GV_TRY(reserve_t(ctx, S_LABELS_IN, n, &d_lab));
  GV_TRY(reserve_t(ctx, S_SCAN0, (size_t)hn + 1, &d_hist));
  GV_TRY(reserve_t(ctx, S_INDICES, n, &d_idx));
  GV_CUDA(cudaMemcpyAsync(d_lab, labels, n * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  GV_CUDA(cudaMemsetAsync(d_hist + hn, 0, sizeof(unsigned), ctx->stream));
  k_label_hist<<<nb, kThreads, (size_t)nboxes * sizeof(unsigned), ctx->stream>>>(d_lab, n, nboxes,
                                                                                 nb, d_hist);
  GV_LAUNCH_CHECK();
  GV_TRY(scan_u32(ctx, d_hist, hn + 1, 0));
  const size_t smem = (size_t)nboxes * (kThreads / 32) * sizeof(unsigned);
  k_label_scatter<<<nb, kThreads, smem, ctx->stream>>>(d_lab, n, nboxes, nb, d_hist, d_idx);
  GV_LAUNCH_CHECK();
  // per-box offsets are the scanned histogram at block 0 of each label
  std::vector<unsigned> h_off((size_t)nboxes + 1);
  GV_CUDA(cudaMemcpy2DAsync(h_off.data(), sizeof(unsigned), d_hist, (size_t)nb * sizeof(unsigned),
                            sizeof(unsigned), (size_t)nboxes, cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaMemcpyAsync(&h_off[nboxes], d_hist + hn, sizeof(unsigned), cudaMemcpyDeviceToHost,
                          ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b <= nboxes; ++b) offsets_out[b] = h_off[b];
  if (h_off[nboxes])
    GV_CUDA(cudaMemcpyAsync(indices_out, d_idx, (size_t)h_off[nboxes] * sizeof(unsigned),
                            cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

// N1: RANSAC ground-plane removal (ref: src/cloud_detections.cpp:105-138)
int gv_segment_ground(gv_ctx *ctx, const float *x, const float *y, const float *z, size_t n,
                      float threshold, uint32_t seed, int n_hyp, float *ox, float *oy, float *oz,
                      size_t *m_out, float *plane_out, int *found_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(m_out != nullptr, GV_ERR_INVALID, "m_out is NULL");
  GV_REQUIRE(n_hyp >= 1 && n_hyp <= kMaxHyp, GV_ERR_INVALID, "n_hyp %d not in [1,%d]", n_hyp, kMaxHyp);
  GV_REQUIRE(n < 4294967296ull, GV_ERR_INVALID, "n too large");
  *m_out = 0;
  if (found_out) *found_out = 0;
  if (plane_out) plane_out[0] = plane_out[1] = plane_out[2] = plane_out[3] = 0.0f;
  if (n < 3) return GV_OK;  // no planar model: the reference returns an empty cloud
  GV_REQUIRE(ox && oy && oz, GV_ERR_INVALID, "output planes are NULL");
  float *d_x, *d_y, *d_z, *d_ox, *d_oy, *d_oz;
  GV_TRY(upload_cloud(ctx, x, y, z, n, &d_x, &d_y, &d_z));
  float4 *d_planes;
  int *d_valid, *d_scores;
  GroundState *d_st;
  unsigned *d_cnt;
  const unsigned nb = blocks_for(n, kThreads);
  GV_TRY(reserve_t(ctx, S_N1_PLANES, (size_t)n_hyp, &d_planes));
  GV_TRY(reserve_t(ctx, S_N1_VALID, (size_t)n_hyp, &d_valid));
  GV_TRY(reserve_t(ctx, S_N1_SCORES, (size_t)n_hyp, &d_scores));
  GV_TRY(reserve_t(ctx, S_N1_STATE, 1, &d_st));
  GV_TRY(reserve_t(ctx, S_SCAN0, (size_t)nb + 1, &d_cnt));
  GV_TRY(reserve_t(ctx, S_OX, n, &d_ox));
  GV_TRY(reserve_t(ctx, S_OY, n, &d_oy));
  GV_TRY(reserve_t(ctx, S_OZ, n, &d_oz));
  GV_CUDA(cudaMemsetAsync(d_scores, 0, (size_t)n_hyp * sizeof(int), ctx->stream));
  k_n1_hypotheses<<<blocks_for(n_hyp, 128), 128, 0, ctx->stream>>>(d_x, d_y, d_z, (unsigned)n, seed, n_hyp,
                                                                   d_planes, d_valid);
  GV_LAUNCH_CHECK();
  const unsigned sb = (unsigned)ctx->num_sms * 4u;
  k_n1_score<<<sb, kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, n, d_planes, d_valid, n_hyp, threshold, d_scores);
  GV_LAUNCH_CHECK();
  k_n1_best<<<1, 32, 0, ctx->stream>>>(d_planes, d_scores, n_hyp, d_st);
  GV_LAUNCH_CHECK();
  for (int pass = 0; pass < 2; ++pass) {
    k_n1_moments<<<sb, kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, n, threshold, pass, d_st);
    GV_LAUNCH_CHECK();
  }
  k_n1_refine<<<1, 1, 0, ctx->stream>>>(d_st);
  GV_LAUNCH_CHECK();
  GV_CUDA(cudaMemsetAsync(d_cnt + nb, 0, sizeof(unsigned), ctx->stream));
  k_n1_count<<<nb, kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, n, threshold, d_st, d_cnt, nullptr);
  GV_LAUNCH_CHECK();
  GV_TRY(scan_u32(ctx, d_cnt, (unsigned long long)nb + 1, 0));
  k_n1_scatter<<<nb, kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, n, threshold, d_st, d_cnt, d_ox, d_oy, d_oz);
  GV_LAUNCH_CHECK();
  GroundState st;
  unsigned total = 0;
  GV_CUDA(cudaMemcpyAsync(&st, d_st, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaMemcpyAsync(&total, d_cnt + nb, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!st.found) return GV_OK;  // ref: :122-126 empty cloud
  if (found_out) *found_out = 1;
  if (plane_out) {
    plane_out[0] = st.plane.x; plane_out[1] = st.plane.y; plane_out[2] = st.plane.z; plane_out[3] = st.plane.w;
  }
  if (total) {
    GV_CUDA(cudaMemcpyAsync(ox, d_ox, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    GV_CUDA(cudaMemcpyAsync(oy, d_oy, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    GV_CUDA(cudaMemcpyAsync(oz, d_oz, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    GV_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  *m_out = total;
  return GV_OK;
}

// N2: radius-outlier filter + PCA box per label (ref: src/cloud_detections.cpp:140-247)
int gv_bbox_pose(gv_ctx *ctx, const float *x, const float *y, const float *z, size_t n,
                 const int16_t *labels, int nboxes, gv_lshape *out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  static_assert(sizeof(gv_lshape) == sizeof(LShapeDev), "gv_lshape layout");
  GV_REQUIRE(nboxes >= 0 && nboxes <= 1024, GV_ERR_INVALID, "nboxes %d not in [0,1024]", nboxes);
  GV_REQUIRE(n < 4294967296ull, GV_ERR_INVALID, "n too large for 32-bit indices");
  if (nboxes == 0) return GV_OK;
  GV_REQUIRE(out != nullptr, GV_ERR_INVALID, "out is NULL");
  for (int b = 0; b < nboxes; ++b) {
    memset(&out[b], 0, sizeof(gv_lshape));
    out[b].qw = 1.0;
  }
  if (n == 0) return GV_OK;
  GV_REQUIRE(labels != nullptr, GV_ERR_INVALID, "labels is NULL");
  // stable partition of point indices by label, on the device (same kernels as
  // gv_partition_by_label), offsets come back to size the launches
  float *d_x, *d_y, *d_z;
  GV_TRY(upload_cloud(ctx, x, y, z, n, &d_x, &d_y, &d_z));
  const unsigned nb = blocks_for(n, kThreads);
  const unsigned long long hn = (unsigned long long)nboxes * nb;
  int16_t *d_lab;
  unsigned *d_hist, *d_idx;
  GV_TRY(reserve_t(ctx, S_LABELS_IN, n, &d_lab));
  GV_TRY(reserve_t(ctx, S_SCAN0, (size_t)hn + 1, &d_hist));
  GV_TRY(reserve_t(ctx, S_INDICES, n, &d_idx));
  GV_CUDA(cudaMemcpyAsync(d_lab, labels, n * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  GV_CUDA(cudaMemsetAsync(d_hist + hn, 0, sizeof(unsigned), ctx->stream));
  k_label_hist<<<nb, kThreads, (size_t)nboxes * sizeof(unsigned), ctx->stream>>>(d_lab, n, nboxes, nb, d_hist);
  GV_LAUNCH_CHECK();
  GV_TRY(scan_u32(ctx, d_hist, hn + 1, 0));
  k_label_scatter<<<nb, kThreads, (size_t)nboxes * (kThreads / 32) * sizeof(unsigned), ctx->stream>>>(
    d_lab, n, nboxes, nb, d_hist, d_idx);
  GV_LAUNCH_CHECK();
  std::vector<unsigned> h_off((size_t)nboxes + 1);
  GV_CUDA(cudaMemcpy2DAsync(h_off.data(), sizeof(unsigned), d_hist, (size_t)nb * sizeof(unsigned),
                            sizeof(unsigned), (size_t)nboxes, cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaMemcpyAsync(&h_off[nboxes], d_hist + hn, sizeof(unsigned), cudaMemcpyDeviceToHost,
                          ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<unsigned long long> off64(h_off.begin(), h_off.end());
  std::vector<int2> tab;
  for (int b = 0; b < nboxes; ++b)
    for (unsigned q = 0; q < h_off[b + 1] - h_off[b]; q += kThreads) tab.push_back(make_int2(b, (int)q));
  unsigned long long *d_off;
  int2 *d_tab;
  uint8_t *d_keep;
  LShapeDev *d_out;
  GV_TRY(reserve_t(ctx, S_N2_OFF, (size_t)nboxes + 1, &d_off));
  GV_TRY(reserve_t(ctx, S_N2_TAB, tab.size() + 1, &d_tab));
  GV_TRY(reserve_t(ctx, S_N2_KEEP, (size_t)h_off[nboxes] + 1, &d_keep));
  GV_TRY(reserve_t(ctx, S_N2_OUT, (size_t)nboxes, &d_out));
  GV_CUDA(cudaMemcpyAsync(d_off, off64.data(), off64.size() * sizeof(unsigned long long),
                          cudaMemcpyHostToDevice, ctx->stream));
  if (!tab.empty()) {
    GV_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice,
                            ctx->stream));
    // ref: :150-154 setRadiusSearch(0.4), setMinNeighborsInRadius(10); FLANN gets (float)(r*r)
    k_n2_neighbors<<<(unsigned)tab.size(), kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, d_idx, d_tab, d_off,
                                                                      (float)(0.4 * 0.4), 10, d_keep);
    GV_LAUNCH_CHECK();
  }
  k_n2_box<<<nboxes, kThreads, 0, ctx->stream>>>(d_x, d_y, d_z, d_idx, d_off, d_keep, d_out);
  GV_LAUNCH_CHECK();
  GV_CUDA(cudaMemcpyAsync(out, d_out, (size_t)nboxes * sizeof(LShapeDev), cudaMemcpyDeviceToHost,
                          ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

// N4: static-object depth — cloud_detections::computeDepthForBoundingBoxes
// (ref: src/cloud_detections.cpp:43-87, called at src/grid_vision_node.cpp:179)
int gv_box_depths(gv_ctx *ctx, const float *uvz, size_t m, const gv_box *boxes, int nboxes, int k,
                  float *depths_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(nboxes >= 0 && k >= 1 && k <= kMaxKnn, GV_ERR_INVALID, "nboxes %d / k %d out of range (k <= %d)",
             nboxes, k, kMaxKnn);
  GV_REQUIRE(m < 4294967295ull, GV_ERR_INVALID, "too many image points");
  if (nboxes == 0) return GV_OK;
  GV_REQUIRE(boxes && depths_out && (m == 0 || uvz), GV_ERR_INVALID, "boxes / depths_out / uvz is NULL");
  float *d_uvz, *d_depth;
  BoxRaw *d_box;
  GV_TRY(reserve_t(ctx, S_UVZ, 3 * (m ? m : 1), &d_uvz));
  GV_TRY(reserve_t(ctx, S_BOX_RAW, (size_t)nboxes, &d_box));
  GV_TRY(reserve_t(ctx, S_KNN_DEPTH, (size_t)nboxes, &d_depth));
  if (m) GV_CUDA(cudaMemcpyAsync(d_uvz, uvz, 3 * m * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  GV_CUDA(cudaMemcpyAsync(d_box, boxes, (size_t)nboxes * sizeof(BoxRaw), cudaMemcpyHostToDevice, ctx->stream));
  k_box_knn_depth<<<nboxes, kThreads, 0, ctx->stream>>>(d_uvz, (unsigned)m, d_box, k, d_depth);
  GV_LAUNCH_CHECK();
  GV_CUDA(cudaMemcpyAsync(depths_out, d_depth, (size_t)nboxes * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

// cloud_detections::pixelTo3D for a list of box centres (ref: src/cloud_detections.cpp:89-103,
// loop at src/grid_vision_node.cpp:309-335)
int gv_pixels_to_3d(gv_ctx *ctx, const gv_box *boxes, const float *depths, int nboxes, const double *K_inv,
                    double *xyz_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(nboxes >= 0, GV_ERR_INVALID, "negative box count");
  if (nboxes == 0) return GV_OK;
  GV_REQUIRE(boxes && depths && K_inv && xyz_out, GV_ERR_INVALID, "NULL argument");
  BoxRaw *d_box;
  float *d_depth;
  double *d_k;
  GV_TRY(reserve_t(ctx, S_BOX_RAW, (size_t)nboxes, &d_box));
  GV_TRY(reserve_t(ctx, S_KNN_DEPTH, (size_t)nboxes, &d_depth));
  GV_TRY(reserve_t(ctx, S_KNN_XYZ, (size_t)3 * nboxes + 9, &d_k));
  GV_CUDA(cudaMemcpyAsync(d_box, boxes, (size_t)nboxes * sizeof(BoxRaw), cudaMemcpyHostToDevice, ctx->stream));
  GV_CUDA(cudaMemcpyAsync(d_depth, depths, (size_t)nboxes * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  GV_CUDA(cudaMemcpyAsync(d_k, K_inv, 9 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_pixels_to_3d<<<blocks_for(nboxes, 128), 128, 0, ctx->stream>>>(d_box, d_depth, nboxes, d_k, d_k + 9);
  GV_LAUNCH_CHECK();
  GV_CUDA(cudaMemcpyAsync(xyz_out, d_k + 9, (size_t)3 * nboxes * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

// =====================================================================================
// grid state
// =====================================================================================
int gv_grid_init(gv_ctx *ctx, double length_x, double length_y, double resolution, double pos_x,
                 double pos_y)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  return grid_init_impl(ctx, length_x, length_y, resolution, pos_x, pos_y);
}

int gv_grid_init_reference(gv_ctx *ctx, uint8_t grid_x, uint8_t grid_y, double resolution)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  // ref: src/occupancy_grid.cpp:10-11  Length(grid_x, grid_y), Position(grid_x / 3, 0.0):
  // uint8_t / int -> integer division
  return grid_init_impl(ctx, (double)grid_x, (double)grid_y, resolution, (double)(grid_x / 3), 0.0);
}

int gv_grid_get_desc(gv_ctx *ctx, gv_grid_desc *out)
{
  if (!ctx || !out) return GV_ERR_INVALID;
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  out->nx = ctx->g.nx;
  out->ny = ctx->g.ny;
  out->resolution = ctx->g.res;
  out->length_x = ctx->g.len_x;
  out->length_y = ctx->g.len_y;
  out->pos_x = ctx->g.pos_x;
  out->pos_y = ctx->g.pos_y;
  return GV_OK;
}

int gv_grid_reset(gv_ctx *ctx)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_TRY(join_merge(ctx));
  const size_t padded = ctx->ncells + kPlanePad;
  GV_CUDA(cudaMemsetAsync(ctx->d_lo, 0, padded * sizeof(float), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_hit, 0, padded * sizeof(int32_t), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_miss, 0, padded * sizeof(int32_t), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_ends_buf[0], 0, padded * sizeof(unsigned long long), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_ends_buf[1], 0, padded * sizeof(unsigned long long), ctx->stream));
  k_fill_f32<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(ctx->d_occ, ctx->ncells, 0.5f);
  GV_LAUNCH_CHECK();
  GV_CUDA(cudaMemsetAsync(ctx->d_stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
  ctx->counts_dirty = ctx->ends_dirty = false;
  ctx->beams_bound = 0;
  return GV_OK;
}

int gv_grid_upload(gv_ctx *ctx, const float *log_odds, const float *occupancy)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_TRY(join_merge(ctx));
  const size_t bytes = ctx->ncells * sizeof(float);
  if (log_odds)
    GV_CUDA(cudaMemcpyAsync(ctx->d_lo, log_odds, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (occupancy)
    GV_CUDA(cudaMemcpyAsync(ctx->d_occ, occupancy, bytes, cudaMemcpyHostToDevice, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

int gv_grid_download(gv_ctx *ctx, float *log_odds, float *occupancy)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_TRY(join_merge(ctx));
  const size_t bytes = ctx->ncells * sizeof(float);
  if (log_odds)
    GV_CUDA(cudaMemcpyAsync(log_odds, ctx->d_lo, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (occupancy)
    GV_CUDA(cudaMemcpyAsync(occupancy, ctx->d_occ, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

int gv_grid_counts_download(gv_ctx *ctx, int32_t *hit, int32_t *miss)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_TRY(join_merge(ctx));
  GV_TRY(raycast_flush_impl(ctx, 0, 1));
  const size_t bytes = ctx->ncells * sizeof(int32_t);
  if (hit) GV_CUDA(cudaMemcpyAsync(hit, ctx->d_hit, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (miss) GV_CUDA(cudaMemcpyAsync(miss, ctx->d_miss, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

int gv_grid_layers_dev(gv_ctx *ctx, float **d_log_odds, float **d_occupancy, int32_t **d_hit,
                       int32_t **d_miss)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_TRY(join_merge(ctx));
  if (d_log_odds) *d_log_odds = ctx->d_lo;
  if (d_occupancy) *d_occupancy = ctx->d_occ;
  if (d_hit) *d_hit = ctx->d_hit;
  if (d_miss) *d_miss = ctx->d_miss;
  return GV_OK;
}

int gv_grid_get_index(gv_ctx *ctx, const double *xy, int n, int32_t *ixy_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_TRY(join_merge(ctx));
  if (n <= 0) return GV_OK;
  GV_REQUIRE(xy && ixy_out, GV_ERR_INVALID, "xy / ixy_out is NULL");
  double *d_in;
  int32_t *d_out;
  GV_TRY(reserve_t(ctx, S_FOOT_IN, (size_t)2 * n, &d_in));
  GV_TRY(reserve_t(ctx, S_CELL, (size_t)2 * n, &d_out));
  GV_CUDA(cudaMemcpyAsync(d_in, xy, (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice,
                          ctx->stream));
  k_get_index<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(d_in, n, ctx->g, d_out);
  GV_LAUNCH_CHECK();
  GV_CUDA(cudaMemcpyAsync(ixy_out, d_out, (size_t)2 * n * sizeof(int32_t), cudaMemcpyDeviceToHost,
                          ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

// =====================================================================================
// per-frame updates R7 / R8 / R9
// =====================================================================================
int gv_grid_update(gv_ctx *ctx)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_TRY(join_merge(ctx));
  return finalize_impl(ctx, 1, nullptr, nullptr, 0, 0);
}

int gv_grid_update_poses(gv_ctx *ctx, const double *xylw, int n)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_TRY(join_merge(ctx));
  return finalize_impl(ctx, 1, xylw, nullptr, n, 1);
}

int gv_grid_update_points(gv_ctx *ctx, const double *xy, const int32_t *labels, int n)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_TRY(join_merge(ctx));
  return finalize_impl(ctx, 1, xy, labels, n, 2);
}

int gv_grid_update_corners(gv_ctx *ctx, const double *corners, int n)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_TRY(join_merge(ctx));
  return finalize_impl(ctx, 1, corners, nullptr, n, 0);
}

// =====================================================================================
// binning + raycast + finalise
// =====================================================================================
int gv_set_base_transform(gv_ctx *ctx, const float *T_base_lidar)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(T_base_lidar != nullptr, GV_ERR_INVALID, "T_base_lidar is NULL");
  GV_TRY(join_merge(ctx));
  // beams binned under the previous pose belong to the previous start cell: walk them first
  if (ctx->has_grid && ctx->ends_dirty) GV_TRY(raycast_flush_impl(ctx, 0, 1));
  memcpy(ctx->Tb, T_base_lidar, sizeof(ctx->Tb));
  ctx->has_base = true;
  return refresh_origin(ctx);
}

int gv_grid_accumulate_dev(gv_ctx *ctx, const float *d_x, const float *d_y, const float *d_z,
                           size_t n, const int16_t *d_labels, const gv_accum_params *prm,
                           int32_t *d_cell_out, uint8_t *d_flags_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(n == 0 || (d_x && d_y && d_z), GV_ERR_INVALID, "point planes are NULL");
  return accumulate_dev_impl(ctx, d_x, d_y, d_z, n, d_labels, prm, d_cell_out, d_flags_out);
}

int gv_grid_accumulate(gv_ctx *ctx, const float *x, const float *y, const float *z, size_t n,
                       const int16_t *labels, const gv_accum_params *prm, int32_t *cell_out,
                       uint8_t *flags_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  float *d_x, *d_y, *d_z;
  GV_TRY(upload_cloud(ctx, x, y, z, n, &d_x, &d_y, &d_z));
  int16_t *d_lab = nullptr;
  int32_t *d_cell = nullptr;
  uint8_t *d_flags = nullptr;
  if (labels) {
    GV_TRY(reserve_t(ctx, S_LABELS_IN, n, &d_lab));
    GV_CUDA(cudaMemcpyAsync(d_lab, labels, n * sizeof(int16_t), cudaMemcpyHostToDevice,
                            ctx->stream));
  }
  if (cell_out) GV_TRY(reserve_t(ctx, S_CELL, n, &d_cell));
  if (flags_out) GV_TRY(reserve_t(ctx, S_FLAGS, n, &d_flags));
  GV_TRY(accumulate_dev_impl(ctx, d_x, d_y, d_z, n, d_lab, prm, d_cell, d_flags));
  if (cell_out && n)
    GV_CUDA(cudaMemcpyAsync(cell_out, d_cell, n * sizeof(int32_t), cudaMemcpyDeviceToHost,
                            ctx->stream));
  if (flags_out && n)
    GV_CUDA(cudaMemcpyAsync(flags_out, d_flags, n * sizeof(uint8_t), cudaMemcpyDeviceToHost,
                            ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

int gv_grid_raycast_flush(gv_ctx *ctx)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_TRY(join_merge(ctx));
  return raycast_flush_impl(ctx, 0, 1);
}

int gv_grid_finalize(gv_ctx *ctx, int32_t k_decay, const double *corners, int nfoot)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  MergeScope sc;
  GV_TRY(merge_begin(ctx, &sc));
  return merge_end(ctx, &sc, finalize_impl(ctx, k_decay, corners, nullptr, nfoot, 0));
}

// Persisting-L2 access window on the end-cell plane for the kernels launched next on the context
// stream: every beam's RED wants its line in L2, while the point planes stream through the same
// cache (ncu, round 2: half of the RED sectors missed L2 without it).  on = false clears it.
static void set_ends_window(gv_ctx *ctx, bool on)
{
  if (!ctx->persist_bytes || !ctx->max_window || ctx->capturing) return;
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof(v));
  if (on) {
    size_t bytes = ctx->ncells * sizeof(unsigned long long);
    if (bytes > ctx->max_window) bytes = ctx->max_window;
    v.accessPolicyWindow.base_ptr = ctx->d_ends;
    v.accessPolicyWindow.num_bytes = bytes;
    v.accessPolicyWindow.hitRatio = bytes <= ctx->persist_bytes ? 1.0f : (float)ctx->persist_bytes / (float)bytes;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  }
  cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &v);
  cudaGetLastError();  // a hint: failure to set it changes nothing but speed
}

// ---- k_points_fast (gv_points_fast.cuh): eligibility, parameter block, launch ------------
static bool fast_eligible(const gv_ctx *ctx, const BinDev &bin, int max_boxes)
{
  const CamDev &c = ctx->cam[0];
  if (!(ctx->use_fast && c.has_T && c.t_small && c.canon && bin.t_small && bin.fast_index_ok &&
        bin.origin_ok && max_boxes <= kFastBoxes))
    return false;
  // operand ranges of the inlined IEEE division / square-root sequences (div_rn_inrange,
  // sqrt_rn_inrange): the sensor origin is at least 2^-20 cells away from every map edge (so the
  // clip numerators and denominators are normal numbers) and the range cap is a sane length
  const float lim = 9.5367431640625e-07f;  // 2^-20
  const float nxf = (float)bin.g.nx, nyf = (float)bin.g.ny;
  if (!(bin.oaxf >= lim && bin.oayf >= lim && nxf - bin.oaxf >= lim && nyf - bin.oayf >= lim)) return false;
  if (bin.cap && !(bin.rmaxf >= 1.0e-3f && bin.rmaxf <= 1.0e6f)) return false;
  return true;
}

static void fill_fast_args(const PointArgs &a, unsigned *d_defer, FastArgs &f, bool *bounded)
{
  memset(&f, 0, sizeof(f));
  const CamDev &c = a.cam[0];
  const BinDev &b = a.bin;
  FastHot &h = f.hot;
  f.x = a.x; f.y = a.y; f.z = a.z;
  f.labels = a.labels;
  f.ends = a.ends;
  f.defer_bits = d_defer;
  f.boxes = a.boxes;
  f.masks = a.masks;
  f.tile_start = a.tile_start; f.tile_end = a.tile_end; f.tile_boxes = a.tile_boxes;
  f.tile0 = a.tile0;
  f.tile_pts = a.tile_pts;
  f.mask_stride = a.mask_stride;
  f.mask_shift = a.mask_shift[0];
  f.mask_tx = a.mask_tx[0];
  FastWarm &w = f.warm;
  for (int i = 0; i < 8; ++i) w.Tcxy[i] = c.T[i];
  for (int i = 0; i < 4; ++i) h.Tcz[i] = c.T[8 + i];
  for (int i = 0; i < 8; ++i) h.Tb[i] = b.T[i];
  for (int i = 0; i < 4; ++i) f.Tbz[i] = b.T[8 + i];
  w.fx = c.fxf; w.fy = c.fyf; w.cx = c.cxf; w.cy = c.cyf;
  // E(q) = 2^-22 (6|q| + 1.5|c| + 1), see fast_point
  const float u22 = 2.384185791015625e-07f;
  w.e6 = 6.0f * u22;
  w.e0u = u22 * (1.5f * std::fabs(c.cxf) + 1.0f) * 1.0001f;
  w.e0v = u22 * (1.5f * std::fabs(c.cyf) + 1.0f) * 1.0001f;
  w.Wf = c.Wf; w.Hf = c.Hf;
  h.oxf = b.oxf; h.oyf = b.oyf;
  w.rmaxf = b.rmaxf;
  h.rmax2f = b.cap ? b.rmax2f : INFINITY;
  h.lab_min = b.occ_mode == 1 ? 0 : -1;
  f.z_min = b.z_min; f.z_max = b.z_max;
  // index FMA: r = p * (-1/res) + (c0/res + bias + 1.5*2^36); ulp(r) = 2^-16 cells
  const GridGeom &g = b.g;
  const int nmax = g.nx > g.ny ? g.nx : g.ny;
  int bias_cells = 16;
  *bounded = false;
  if (b.cap) {
    const double reach = std::ceil((double)b.rmaxf / g.res) + 4.0;  // cells a beam can extend from the origin
    if (reach + (double)nmax + reach < 32768.0) {
      bias_cells = (int)reach;
      *bounded = true;
    }
  }
  const double magic = 103079215104.0;  // 1.5 * 2^36
  h.nires = -1.0 / g.res;
  h.Cx = b.c0xd / g.res + (double)bias_cells + magic;
  h.Cy = b.c0yd / g.res + (double)bias_cells + magic;
  h.kbm = ((unsigned)bias_cells << 16) + kIdxMargin;
  f.hi0 = 0x42380000u;
  f.klim_x = (unsigned)g.nx << 16;
  f.klim_y = (unsigned)g.ny << 16;
  h.klim_xm = f.klim_x - 2u * kIdxMargin;
  h.klim_ym = f.klim_y - 2u * kIdxMargin;
  h.nx = g.nx;
  f.ny = g.ny;
  // clip geometry: the same single-rounded float expressions as clip_end / oracle gvo_clip_end
  f.c0xf = b.c0xf; f.c0yf = b.c0yf; f.inv_resf = b.inv_resf;
  f.oaxf = b.oaxf; f.oayf = b.oayf;
  f.nxf = (float)g.nx; f.nyf = (float)g.ny;
  f.nxm1f = (float)(g.nx - 1); f.nym1f = (float)(g.ny - 1);
  f.noaxf = 0.0f - b.oaxf; f.noayf = 0.0f - b.oayf;
  f.paxf = f.nxf - b.oaxf; f.payf = f.nyf - b.oayf;
  f.cam = c;
  f.bin = b;
}

// The deferred stage of a certified launch: scan + compact the bitmap, then one thread per entry.
// In a slots variable: a word index with a gw above 2^32 or an overflowing list is handled in place.
static int launch_deferred(gv_ctx *ctx, const FastArgs &f, unsigned long long nwords)
{
  if (nwords == 0) return GV_OK;
  unsigned nb = (unsigned)((nwords + kThreads - 1) / kThreads);
  const unsigned capb = (unsigned)ctx->num_sms * 16u;
  if (nb > capb) nb = capb;
  DeferList dl;
  // per scan CTA: room for 1/4 of its words to be non-zero (the certified paths defer < 1 % of the
  // points); what does not fit is processed in place
  const unsigned long long per_cta = (nwords + nb - 1) / nb;
  // a scan or two (<= 16 k words): the scan processes its words in place (capacity 0), one launch less
  const bool in_place = nwords <= 16384ull;
  dl.capacity = in_place ? 0u : (unsigned)(per_cta / 4 + 64);
  GV_TRY(reserve_t(ctx, S_DEFER_LIST, (size_t)nb * (dl.capacity ? dl.capacity : 1u), &dl.items));
  GV_TRY(reserve_t(ctx, S_DEFER_COUNT, (size_t)nb, &dl.count));
  k_points_deferred<<<nb, kThreads, 0, ctx->stream>>>(f, dl);
  GV_LAUNCH_CHECK();
  if (!in_place) {
    k_points_deferred_list<<<nb, kThreads, 0, ctx->stream>>>(f, dl, ctx->d_stats + 4);
    GV_LAUNCH_CHECK();
  }
  return GV_OK;
}

// tma: k_points_tma (persistent, bulk-copy fed) instead of k_points_fast
static int launch_points_fast(gv_ctx *ctx, FastArgs &f, bool bounded, bool tma, unsigned tile0, unsigned ntiles)
{
  if (ntiles == 0) return GV_OK;
  f.tile0 = tile0;
  f.ntiles = ntiles;
  const bool lab = f.labels != nullptr, zg = f.bin.use_z_gate != 0;
  // an end-cell plane that fits L2 absorbs one RED per beam; a larger one pays a DRAM
  // read-modify-write per RED, so merging the repeats of a warp-row first is worth its instructions
  const int U = ctx->fast_unroll;
  const int G = ctx->fast_agg >= 0 ? ctx->fast_agg : (ctx->ncells * sizeof(unsigned long long) > ((size_t)64 << 20) ? 2 : 0);
  if (tma) {
    const size_t stage = 3 * (size_t)f.tile_pts * 4 + kFastBoxes * 16 + (size_t)f.mask_stride * 8;
    const size_t smem = 2 * stage + (size_t)f.tile_pts * 8;  // two stages + the run slots
    // CTA b walks the contiguous table entries [b*K, (b+1)*K): 3 CTAs per SM, and K <= 32768 so
    // that the 16-bit run counters cannot overflow
    unsigned grid = 3u * (unsigned)ctx->num_sms;
    if (grid > ntiles) grid = ntiles;
    unsigned K = (ntiles + grid - 1) / grid;
    if (K > 32768u) K = 32768u;
    grid = (ntiles + K - 1) / K;
    f.tiles_per_cta = K;
    const bool hoist = ctx->tma_hoist;
#define GV_TMA_LAUNCH1(UU, BB, LL, ZZ, HH)                                                           \
  do {                                                                                               \
    GV_CUDA(cudaFuncSetAttribute(k_points_tma<UU, BB, LL, ZZ, HH>,                                    \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    k_points_tma<UU, BB, LL, ZZ, HH><<<grid, kThreads, smem, ctx->stream>>>(f);                       \
  } while (0)
#define GV_TMA_LAUNCH(UU, BB, LL, ZZ)                \
  do {                                               \
    if (hoist) GV_TMA_LAUNCH1(UU, BB, LL, ZZ, true); \
    else GV_TMA_LAUNCH1(UU, BB, LL, ZZ, false);      \
  } while (0)
#define GV_TMA_BL(UU, ZZ)                                  \
  do {                                                     \
    if (bounded && lab) GV_TMA_LAUNCH(UU, true, true, ZZ); \
    else if (bounded) GV_TMA_LAUNCH(UU, true, false, ZZ);  \
    else if (lab) GV_TMA_LAUNCH(UU, false, true, ZZ);      \
    else GV_TMA_LAUNCH(UU, false, false, ZZ);              \
  } while (0)
    if (zg) GV_TMA_BL(2, true);
    else if (U == 1) GV_TMA_BL(1, false);
    else GV_TMA_BL(2, false);
#undef GV_TMA_BL
#undef GV_TMA_LAUNCH
#undef GV_TMA_LAUNCH1
  } else {
    const size_t smem = (size_t)f.mask_stride * sizeof(unsigned long long);
#define GV_FAST_LAUNCH(UU, BB, LL, ZZ, GG) k_points_fast<UU, BB, LL, ZZ, GG><<<ntiles, kThreads, smem, ctx->stream>>>(f)
#define GV_FAST_BL(UU, ZZ, GG)                                  \
  do {                                                          \
    if (bounded && lab) GV_FAST_LAUNCH(UU, true, true, ZZ, GG); \
    else if (bounded) GV_FAST_LAUNCH(UU, true, false, ZZ, GG);  \
    else if (lab) GV_FAST_LAUNCH(UU, false, true, ZZ, GG);      \
    else GV_FAST_LAUNCH(UU, false, false, ZZ, GG);              \
  } while (0)
    if (zg) GV_FAST_BL(2, true, 1);  // rare configuration: one instantiation per remaining flag
    else if (G == 0 && U == 1) GV_FAST_BL(1, false, 0);
    else if (G == 0 && U == 4) GV_FAST_BL(4, false, 0);
    else if (G == 0) GV_FAST_BL(2, false, 0);
    else if (G == 2 && U == 1) GV_FAST_BL(1, false, 2);
    else if (G == 2 && U == 4) GV_FAST_BL(4, false, 2);
    else if (G == 2) GV_FAST_BL(2, false, 2);
    else if (U == 1) GV_FAST_BL(1, false, 1);
    else if (U == 4) GV_FAST_BL(4, false, 1);
    else GV_FAST_BL(2, false, 1);
#undef GV_FAST_BL
#undef GV_FAST_LAUNCH
  }
  GV_LAUNCH_CHECK();
  // the points whose decisions could not be certified (a bitmap, normally < 0.1 % of the points)
  return launch_deferred(ctx, f, (unsigned long long)ntiles * (unsigned)(f.tile_pts >> 5));
}

// Thresholds of k_points_pair's certified image test and the constants of its tile lookup, from
// E(q) = e6 |q| + e0 (fill_fast_args).  Everything is worked out in double and rounded to the
// safe side: a threshold that is too strict only defers a few more points.
// One image axis of size W with E(q) = e6 |q| + e0: image centre, the constant interval half-width
// of the box tests and the two thresholds of the certified image test (see fill_pair_args).
static void pair_axis_thresholds(double W, double e6, double e0, float *half_out, float *e_out, float *ain, float *aout)
{
  const double slack = 1.0 + 4.76837158203125e-07;  // 1 + 2^-21
  const double half = 0.5 * W;                        // exact in binary32 (W is an image size)
  const double e = (e6 * (W + 1.0) + e0) * 1.001;     // >= E(q) for every |q| <= W + 1
  *half_out = (float)half;
  *e_out = std::nextafterf((float)e, INFINITY);
  // |RN(q - W/2)| < ain  =>  E < q < W - E, hence every value within E(q) of q is in [0, W)
  *ain = std::nextafterf((float)(half - (double)*e_out - W * 4.76837158203125e-07), 0.0f);
  // |RN(q - W/2)| > aout  =>  q - E(q) >= W or q + E(q) < 0
  *aout = std::nextafterf((float)(((W + e0) / (1.0 - e6) - half) * slack), INFINITY);
}

static void fill_pair_args(const FastArgs &f, PairArgs &p)
{
  memset(&p, 0, sizeof(p));
  p.one = 1.0f;
  const FastWarm &w = f.warm;
  const FastHot &h = f.hot;
  for (int i = 0; i < 4; ++i) {
    p.Tcz[i] = h.Tcz[i];
    p.Tcx[i] = w.Tcxy[i];
    p.Tcy[i] = w.Tcxy[4 + i];
    p.Tbx[i] = h.Tb[i];
    p.Tby[i] = h.Tb[4 + i];
  }
  p.fx = w.fx; p.fy = w.fy; p.cx = w.cx; p.cy = w.cy;
  p.oxf = h.oxf; p.oyf = h.oyf; p.noxf = -h.oxf; p.noyf = -h.oyf;
  p.rmax2f = h.rmax2f; p.rmaxf = w.rmaxf;
  p.nires = h.nires; p.Cx = h.Cx; p.Cy = h.Cy;
  p.kbm = h.kbm; p.nx = h.nx; p.klim_xm = h.klim_xm; p.klim_ym = h.klim_ym;
  p.lab_min = h.lab_min;
  p.defer_stride = f.defer_stride;
  pair_axis_thresholds((double)w.Wf, (double)w.e6, (double)w.e0u, &p.half_w, &p.eu, &p.ain_u, &p.aout_u);
  pair_axis_thresholds((double)w.Hf, (double)w.e6, (double)w.e0v, &p.half_h, &p.ev, &p.ain_v, &p.aout_v);
  p.inv_tile = 1.0f / (float)(1 << f.mask_shift);
  p.mask_tx = (unsigned)f.mask_tx;
  p.mask_stride = (unsigned)f.mask_stride;
  p.mask_bias = 0x4340u * (unsigned)(f.mask_tx + 1);
}

// k_points_col / k_points_pair over frames [frame0, frame0 + nframes): grid = (column blocks, frame groups)
static int launch_points_col(gv_ctx *ctx, FastArgs &f, bool bounded, int frame0, int nframes, unsigned max_pts)
{
  if (nframes <= 0 || max_pts == 0) return GV_OK;
  f.frame0 = frame0;
  f.nframes = nframes;
  // k_points_pair: one thread per pair of adjacent points
  const unsigned per_thread = f.pair ? 2u : 1u;
  const unsigned cols = ((max_pts + per_thread - 1) / per_thread + kThreads - 1) / kThreads;
  // enough CTAs for ~16 waves of 5 CTAs per SM, as few frame groups as that allows: the longer a
  // thread stays on its beam index, the more of the beam's repeats it merges before the RED
  const unsigned want = (unsigned)ctx->num_sms * (f.pair ? (unsigned)(ctx->pair_minb * ctx->pair_waves) : 5u * 16u);
  unsigned groups = (want + cols - 1) / cols;
  if (groups > (unsigned)nframes) groups = (unsigned)nframes;
  if (groups < 1u) groups = 1u;
  if (groups > 65535u) groups = 65535u;
  f.frames_per_cta = (nframes + (int)groups - 1) / (int)groups;
  // records are staged per CTA (kColFrames) with 32-bit element offsets relative to the group's
  // first frame: a group may not span 2^32 points
  if (f.frames_per_cta > kColFrames) f.frames_per_cta = kColFrames;
  while (f.frames_per_cta > 1 && (unsigned long long)f.frames_per_cta * max_pts >= 4294967295ull) f.frames_per_cta /= 2;
  groups = (unsigned)((nframes + f.frames_per_cta - 1) / f.frames_per_cta);
  GV_REQUIRE(groups <= 65535u, GV_ERR_INVALID, "too many frames for one launch (%d)", nframes);
  const dim3 grid(cols, groups);
  const bool lab = f.labels != nullptr, zg = f.bin.use_z_gate != 0;
  const bool hoist = ctx->col_hoist;
#define GV_COL_LAUNCH(BB, LL, ZZ)                                                  \
  do {                                                                             \
    if (hoist) k_points_col<BB, LL, ZZ, true><<<grid, kThreads, 0, ctx->stream>>>(f); \
    else k_points_col<BB, LL, ZZ, false><<<grid, kThreads, 0, ctx->stream>>>(f);      \
  } while (0)
#define GV_COL_BL(ZZ)                                  \
  do {                                                 \
    if (bounded && lab) GV_COL_LAUNCH(true, true, ZZ); \
    else if (bounded) GV_COL_LAUNCH(true, false, ZZ);  \
    else if (lab) GV_COL_LAUNCH(false, true, ZZ);      \
    else GV_COL_LAUNCH(false, false, ZZ);              \
  } while (0)
  if (f.pair) {
    PairArgs pa;
    fill_pair_args(f, pa);
#define GV_PAIR_M(BB, LL, ZZ)                                                                         \
  do {                                                                                                \
    switch (ctx->pair_minb) {                                                                         \
    case 3: k_points_pair<BB, LL, ZZ, 3><<<grid, kThreads, 0, ctx->stream>>>(f, pa); break;            \
    case 5: k_points_pair<BB, LL, ZZ, 5><<<grid, kThreads, 0, ctx->stream>>>(f, pa); break;            \
    case 6: k_points_pair<BB, LL, ZZ, 6><<<grid, kThreads, 0, ctx->stream>>>(f, pa); break;            \
    default: k_points_pair<BB, LL, ZZ, 4><<<grid, kThreads, 0, ctx->stream>>>(f, pa); break;           \
    }                                                                                                 \
  } while (0)
#define GV_PAIR_BL(ZZ)                                \
  do {                                                \
    if (bounded && lab) GV_PAIR_M(true, true, ZZ);    \
    else if (bounded) GV_PAIR_M(true, false, ZZ);     \
    else if (lab) GV_PAIR_M(false, true, ZZ);         \
    else GV_PAIR_M(false, false, ZZ);                 \
  } while (0)
    if (zg) GV_PAIR_BL(true);
    else GV_PAIR_BL(false);
#undef GV_PAIR_M
#undef GV_PAIR_BL
  } else if (zg) GV_COL_BL(true);
  else GV_COL_BL(false);
#undef GV_COL_BL
#undef GV_COL_LAUNCH
  GV_LAUNCH_CHECK();
  return launch_deferred(ctx, f, (unsigned long long)nframes * f.defer_stride);
}

// ---- batch: the whole hot path, one kernel pass over the points ------------------------
static int process_batch_impl(gv_ctx *ctx, const float *px, const float *py, const float *pz,
                              bool points_on_device, const uint64_t *frame_offsets, int nframes,
                              const gv_box *boxes, bool boxes_on_device,
                              const int32_t *box_frame_offsets, const gv_accum_params *prm,
                              int16_t *labels_out)
{
  GV_REQUIRE(ctx->ncam >= 1, GV_ERR_STATE, "gv_set_cameras not called");
  GV_REQUIRE(nframes >= 0, GV_ERR_INVALID, "negative frame count");
  if (nframes == 0) return GV_OK;
  GV_REQUIRE(frame_offsets && box_frame_offsets, GV_ERR_INVALID, "offset arrays are NULL");
  const uint64_t p0 = frame_offsets[0];
  const uint64_t n = frame_offsets[nframes] - p0;
  const int nboxes = box_frame_offsets[nframes];
  GV_REQUIRE(box_frame_offsets[0] == 0, GV_ERR_INVALID, "box_frame_offsets[0] must be 0");
  GV_REQUIRE(nboxes == 0 || boxes, GV_ERR_INVALID, "boxes is NULL");
  GV_REQUIRE(n == 0 || (px && py && pz), GV_ERR_INVALID, "point planes are NULL");

  PointArgs a;
  memset(&a, 0, sizeof(a));
  GV_TRY(set_bin_params(ctx, prm, &a.bin));
  GV_TRY(note_beams(ctx, n));
  GV_TRY(bin_begin(ctx));

  // points: device-resident, or staged per chunk from host memory
  const uint64_t base = points_on_device ? 0 : p0;  // host path rebases the staged copy to 0
  // Tile table: one entry per tile of tile_pts point indices of one frame (tiles never straddle
  // frames).  Frames are launched in groups (host path: ~4M-point chunks that pipeline with the
  // PCIe copies; device path: one group) and within a group the table is COLUMN-MAJOR: the same
  // block of point indices of consecutive frames sits in consecutive entries, which is what
  // lets k_points_tma merge a beam with the same beam of the next frame (see gv_points_fast.cuh).
  const int tile_pts = tile_points_for(n, ctx->num_sms, ctx->use_fast && !ctx->use_tma,
                                       ctx->use_fast && ctx->use_tma ? 2 * kTilePts : 4 * kTilePts);
  bool same = ctx->c_valid && ctx->c_tile_pts == tile_pts && ctx->c_on_device == points_on_device &&
              (int)ctx->c_boff.size() == nframes + 1 && (int)ctx->c_foff.size() == nframes + 1 &&
              memcmp(ctx->c_boff.data(), box_frame_offsets, ((size_t)nframes + 1) * sizeof(int)) == 0;
  if (same)
    for (int f = 0; f <= nframes && same; ++f) same = ctx->c_foff[f] == frame_offsets[f] - base;
  if (!same) {
    // a previous call's asynchronous table upload may still be reading the host images
    GV_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->c_valid = false;
    ctx->c_foff.resize((size_t)nframes + 1);
    ctx->c_boff.assign(box_frame_offsets, box_frame_offsets + nframes + 1);
    ctx->c_tile_pts = tile_pts;
    ctx->c_on_device = points_on_device;
    ctx->c_max_boxes = 1;
    unsigned long long ntiles64 = 0;
    for (int f = 0; f < nframes; ++f) {
      GV_REQUIRE(frame_offsets[f + 1] >= frame_offsets[f], GV_ERR_INVALID,
                 "frame_offsets not monotone at %d", f);
      const int nb = box_frame_offsets[f + 1] - box_frame_offsets[f];
      GV_REQUIRE(nb >= 0 && nb <= 32767, GV_ERR_INVALID, "frame %d has %d boxes", f, nb);
      if (nb > ctx->c_max_boxes) ctx->c_max_boxes = nb;
      ntiles64 += (frame_offsets[f + 1] - frame_offsets[f] + tile_pts - 1) / tile_pts;
      GV_REQUIRE(ntiles64 < 2147483647ull, GV_ERR_INVALID, "batch too large");
    }
    for (int f = 0; f <= nframes; ++f) ctx->c_foff[f] = frame_offsets[f] - base;
    const std::vector<unsigned long long> &fo = ctx->c_foff;
    ctx->c_frames.resize((size_t)nframes);
    ctx->c_max_pts = 0;
    ctx->c_pair_ok = true;
    for (int f = 0; f < nframes; ++f) {
      const unsigned long long np = fo[f + 1] - fo[f];
      if ((fo[f] | np) & 1ull) ctx->c_pair_ok = false;
      GV_REQUIRE(np < 4294967295ull, GV_ERR_INVALID, "frame %d too large", f);
      if (np > ctx->c_max_pts) ctx->c_max_pts = (unsigned)np;
      ctx->c_frames[f] = make_uint4((unsigned)(fo[f] & 0xffffffffull), (unsigned)(fo[f] >> 32), (unsigned)np,
                                    (unsigned)box_frame_offsets[f]);
    }
    ctx->c_chunks.clear();
    ctx->c_tstart.clear();
    ctx->c_tend.clear();
    ctx->c_tbox.clear();
    ctx->c_tstart.reserve((size_t)ntiles64);
    ctx->c_tend.reserve((size_t)ntiles64);
    ctx->c_tbox.reserve((size_t)ntiles64);
    const uint64_t chunk_pts = 4ull << 20;
    for (int f0 = 0; f0 < nframes;) {
      int f1 = f0 + 1;
      if (points_on_device) f1 = nframes;
      else while (f1 < nframes && fo[f1 + 1] - fo[f0] <= chunk_pts) ++f1;
      gv_ctx::Chunk ck{f0, f1, (unsigned)ctx->c_tstart.size(), 0u};
      unsigned long long maxcols = 0;
      for (int f = f0; f < f1; ++f) {
        const unsigned long long cols = (fo[f + 1] - fo[f] + tile_pts - 1) / tile_pts;
        if (cols > maxcols) maxcols = cols;
      }
      for (unsigned long long c = 0; c < maxcols; ++c)
        for (int f = f0; f < f1; ++f) {
          const unsigned long long s0 = fo[f] + c * (unsigned long long)tile_pts;
          if (s0 >= fo[f + 1]) continue;
          ctx->c_tstart.push_back(s0);
          ctx->c_tend.push_back(fo[f + 1]);
          ctx->c_tbox.push_back(make_int4(box_frame_offsets[f], box_frame_offsets[f + 1], f, 0));
        }
      ck.ntiles = (unsigned)ctx->c_tstart.size() - ck.tile0;
      ctx->c_chunks.push_back(ck);
      f0 = f1;
    }
  }
  const std::vector<unsigned long long> &foff = ctx->c_foff;
  const int max_boxes = ctx->c_max_boxes;
  const unsigned ntiles = (unsigned)ctx->c_tstart.size();

  unsigned long long *d_tstart, *d_tend;
  int *d_boff;
  int4 *d_tbox;
  GV_TRY(reserve_t(ctx, S_BOX_OFF, (size_t)nframes + 1, &d_boff));
  GV_TRY(reserve_t(ctx, S_TILE_START, (size_t)ntiles + 1, &d_tstart));
  GV_TRY(reserve_t(ctx, S_TILE_END, (size_t)ntiles + 1, &d_tend));
  GV_TRY(reserve_t(ctx, S_TILE_BOX, (size_t)ntiles + 1, &d_tbox));
  uint4 *d_frames;
  GV_TRY(reserve_t(ctx, S_FRAMES, (size_t)nframes + 1, &d_frames));

  const float *d_x = px, *d_y = py, *d_z = pz;
  int16_t *d_lab = labels_out;
  if (!points_on_device) {
    float *tx, *ty, *tz;
    GV_TRY(reserve_t(ctx, S_X, n, &tx));
    GV_TRY(reserve_t(ctx, S_Y, n, &ty));
    GV_TRY(reserve_t(ctx, S_Z, n, &tz));
    d_x = tx;
    d_y = ty;
    d_z = tz;
    if (labels_out) GV_TRY(reserve_t(ctx, S_LAB, n, &d_lab));
  }

  if (!same) {
    // the source vectors are context-owned and stay untouched until the next layout change
    // (which synchronises first), so no stream synchronisation is needed here
    GV_CUDA(cudaMemcpyAsync(d_boff, ctx->c_boff.data(), ((size_t)nframes + 1) * sizeof(int),
                            cudaMemcpyHostToDevice, ctx->stream));
    GV_CUDA(cudaMemcpyAsync(d_frames, ctx->c_frames.data(), (size_t)nframes * sizeof(uint4),
                            cudaMemcpyHostToDevice, ctx->stream));
    if (ntiles) {
      GV_CUDA(cudaMemcpyAsync(d_tstart, ctx->c_tstart.data(), (size_t)ntiles * sizeof(unsigned long long),
                              cudaMemcpyHostToDevice, ctx->stream));
      GV_CUDA(cudaMemcpyAsync(d_tend, ctx->c_tend.data(), (size_t)ntiles * sizeof(unsigned long long),
                              cudaMemcpyHostToDevice, ctx->stream));
      GV_CUDA(cudaMemcpyAsync(d_tbox, ctx->c_tbox.data(), (size_t)ntiles * sizeof(int4),
                              cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->c_valid = true;
  }
  float4 *d_f4 = nullptr;
  const BoxRaw *d_raw = nullptr;  // non-NULL: k_box_masks rounds the boxes itself (sets of <= 256 boxes)
  GV_TRY(upload_boxes(ctx, boxes, nboxes, boxes_on_device, &d_f4, max_boxes <= kThreads ? &d_raw : nullptr));
  // per-frame image-tile box masks
  int mty;
  mask_geometry(ctx->cam[0].W, ctx->cam[0].H, &a.mask_shift[0], &a.mask_tx[0], &mty);
  a.mask_words = (max_boxes + 63) / 64;
  a.mask_stride = (a.mask_tx[0] * mty * a.mask_words + 1) & ~1;  // even: 16-byte sets (bulk copies)
  a.smem_boxes = max_boxes;
  unsigned long long *d_masks = nullptr;
  GV_TRY(reserve_t(ctx, S_MASKS, (size_t)nframes * a.mask_stride, &d_masks));
  const bool fast = fast_eligible(ctx, a.bin, max_boxes);
  k_box_masks<<<nframes, kThreads, 0, ctx->stream>>>(d_f4, d_boff, a.mask_shift[0], a.mask_tx[0],
                                                    mty, a.mask_words, a.mask_stride, fast ? 1 : 0,
                                                    d_masks, d_raw, d_f4);
  GV_LAUNCH_CHECK();
  a.masks = d_masks;
  a.tile_pts = tile_pts;

  a.x = d_x;
  a.y = d_y;
  a.z = d_z;
  a.n = points_on_device ? frame_offsets[nframes] : n;
  a.is_dense = 0;
  a.vec_ok = (((uintptr_t)d_x | (uintptr_t)d_y | (uintptr_t)d_z) & 15u) == 0;
  a.ncam = 1;
  a.cam[0] = ctx->cam[0];
  a.boxes = d_f4;
  a.labels = d_lab;
  a.nframes = nframes;
  a.tile_start = d_tstart;
  a.tile_end = d_tend;
  a.tile_boxes = d_tbox;
  a.ends = ctx->d_ends;
  ctx->ends_dirty = true;
  const size_t smem = (size_t)max_boxes * sizeof(float4) +
                      (size_t)a.mask_stride * sizeof(unsigned long long);
  GV_REQUIRE(smem <= ctx->max_smem_optin, GV_ERR_INVALID,
             "a frame with %d boxes needs %zu bytes of shared memory for the box/tile-mask stage; this "
             "device allows %zu (about 4700 boxes per frame)", max_boxes, smem, ctx->max_smem_optin);
  if (smem > 48 * 1024)
    GV_CUDA(cudaFuncSetAttribute(k_points<true, true, false, false, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  FastArgs fa;
  bool bounded = false, tma = false;
  if (fast) {
    // ballot-word bitmap of deferred points: all-zero between launches (k_points_deferred clears
    // what k_points_fast set), so it is zeroed only when the slot is (re)allocated
    const bool col = ctx->fast_kind == 0 || ctx->fast_kind == 3;
    const unsigned defer_stride = (ctx->c_max_pts + 31u) / 32u;
    const size_t nwords = col ? (size_t)nframes * defer_stride : (size_t)ntiles * (size_t)(tile_pts >> 5);
    unsigned *d_defer = nullptr;
    const void *before = ctx->s[S_DEFER].p;
    const size_t cap_before = ctx->s[S_DEFER].cap;
    GV_TRY(reserve_t(ctx, S_DEFER, nwords, &d_defer));
    if (ctx->s[S_DEFER].p != before || ctx->s[S_DEFER].cap != cap_before)
      GV_CUDA(cudaMemsetAsync(d_defer, 0, ctx->s[S_DEFER].cap, ctx->stream));
    fill_fast_args(a, d_defer, fa, &bounded);
    fa.frames = d_frames;
    fa.defer_stride = defer_stride;
    fa.col_mode = col ? 1 : 0;
    // k_points_pair: 64-bit point loads and 32-bit label stores at even point indices
    fa.pair = col && ctx->fast_kind == 0 && ctx->c_pair_ok && foff[nframes] < 4294967295ull &&
              (unsigned long long)nframes * (unsigned)a.mask_stride < 4000000000ull &&
              (unsigned long long)nframes * defer_stride < 4000000000ull &&
              ((((uintptr_t)d_x | (uintptr_t)d_y | (uintptr_t)d_z) & 7u) == 0) && (((uintptr_t)d_lab & 3u) == 0);
    // bulk copies need 16-byte aligned sources and sizes: plane pointers aligned, every frame
    // boundary (hence every tile start and size) a multiple of 4 points
    tma = ctx->use_tma && a.vec_ok && tile_pts <= 2 * kTilePts;
    for (int f = 0; f <= nframes && tma; ++f) tma = (foff[f] & 3ull) == 0;
  }
  if (points_on_device) {
    a.tile0 = 0;
    if (fast) {
      set_ends_window(ctx, true);
      const int rc = fa.col_mode ? launch_points_col(ctx, fa, bounded, 0, nframes, ctx->c_max_pts)
                                 : launch_points_fast(ctx, fa, bounded, tma, 0, ntiles);
      set_ends_window(ctx, false);
      return rc;
    }
    return launch_points(ctx, true, true, a, ntiles, smem);
  }
  if (fast) set_ends_window(ctx, true);

  // host path: pipeline H2D copies, the fused kernel and the label D2H over three streams,
  // a few frames per chunk (~4M points), so PCIe runs in both directions under the compute
  size_t ev = 0;
  for (const gv_ctx::Chunk &ck : ctx->c_chunks) {
    const uint64_t c0 = foff[ck.f0], c1 = foff[ck.f1];
    const size_t bytes = (size_t)(c1 - c0) * sizeof(float);
    cudaEvent_t e_h2d = get_event(ctx, ev++), e_k = get_event(ctx, ev++);
    GV_REQUIRE(e_h2d && e_k, GV_ERR_CUDA, "cudaEventCreate failed");
    if (bytes) {
      GV_CUDA(cudaMemcpyAsync((float *)d_x + c0, px + p0 + c0, bytes, cudaMemcpyHostToDevice,
                              ctx->h2d_stream));
      GV_CUDA(cudaMemcpyAsync((float *)d_y + c0, py + p0 + c0, bytes, cudaMemcpyHostToDevice,
                              ctx->h2d_stream));
      GV_CUDA(cudaMemcpyAsync((float *)d_z + c0, pz + p0 + c0, bytes, cudaMemcpyHostToDevice,
                              ctx->h2d_stream));
    }
    GV_CUDA(cudaEventRecord(e_h2d, ctx->h2d_stream));
    GV_CUDA(cudaStreamWaitEvent(ctx->stream, e_h2d, 0));
    a.tile0 = ck.tile0;
    if (fast) {
      if (fa.col_mode) GV_TRY(launch_points_col(ctx, fa, bounded, ck.f0, ck.f1 - ck.f0, ctx->c_max_pts));
      else GV_TRY(launch_points_fast(ctx, fa, bounded, tma, ck.tile0, ck.ntiles));
    } else {
      GV_TRY(launch_points(ctx, true, true, a, ck.ntiles, smem));
    }
    if (labels_out && c1 > c0) {
      GV_CUDA(cudaEventRecord(e_k, ctx->stream));
      GV_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, e_k, 0));
      GV_CUDA(cudaMemcpyAsync(labels_out + p0 + c0, d_lab + c0, (size_t)(c1 - c0) * sizeof(int16_t),
                              cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
  }
  if (fast) set_ends_window(ctx, false);
  GV_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

int gv_process_batch(gv_ctx *ctx, const float *x, const float *y, const float *z,
                     const uint64_t *frame_offsets, int nframes, const gv_box *boxes,
                     const int32_t *box_frame_offsets, const gv_accum_params *prm,
                     int16_t *labels_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  return process_batch_impl(ctx, x, y, z, false, frame_offsets, nframes, boxes, false,
                            box_frame_offsets, prm, labels_out);
}

int gv_process_batch_dev(gv_ctx *ctx, const float *d_x, const float *d_y, const float *d_z,
                         const uint64_t *frame_offsets, int nframes, const gv_box *d_boxes,
                         const int32_t *box_frame_offsets, const gv_accum_params *prm,
                         int16_t *d_labels_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  return process_batch_impl(ctx, d_x, d_y, d_z, true, frame_offsets, nframes, d_boxes, true,
                            box_frame_offsets, prm, d_labels_out);
}

// =====================================================================================
// consumers
// =====================================================================================
int gv_grid_to_occupancy(gv_ctx *ctx, int8_t *data_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_TRY(join_merge(ctx));
  GV_REQUIRE(data_out != nullptr, GV_ERR_INVALID, "data_out is NULL");
  int8_t *d_out;
  GV_TRY(reserve_t(ctx, S_FLAGS, ctx->ncells, &d_out));
  k_to_occupancy<<<blocks_for(ctx->ncells, kThreads), kThreads, 0, ctx->stream>>>(ctx->d_occ,
                                                                                 ctx->ncells, d_out);
  GV_LAUNCH_CHECK();
  GV_CUDA(cudaMemcpyAsync(data_out, d_out, ctx->ncells, cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  return GV_OK;
}

// =====================================================================================
// multi-GPU
// =====================================================================================
#ifdef GV_WITH_NCCL
#define GV_NCCL(call)                                                                         \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess)                                                                    \
      return ctx->fail(GV_ERR_NCCL, "%s failed: %s", #call, ncclGetErrorString(r_));          \
  } while (0)
#endif

int gv_nccl_unique_id(void *id128_out)
{
#ifdef GV_WITH_NCCL
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (!id128_out) return GV_ERR_INVALID;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return GV_ERR_NCCL;
  memcpy(id128_out, &id, sizeof(id));
  return GV_OK;
#else
  (void)id128_out;
  return GV_ERR_NCCL;
#endif
}

int gv_nccl_init(gv_ctx *ctx, const void *id128, int rank, int world)
{
  if (!ctx) return GV_ERR_INVALID;
#ifdef GV_WITH_NCCL
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(id128 && world >= 1 && rank >= 0 && rank < world && world <= 64, GV_ERR_INVALID,
             "bad rank/world %d/%d", rank, world);
  GV_REQUIRE(ctx->comm == nullptr, GV_ERR_STATE, "communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  GV_NCCL(ncclCommInitRank(&ctx->comm, world, id, rank));
  ctx->rank = rank;
  ctx->world = world;
  return GV_OK;
#else
  (void)id128; (void)rank; (void)world;
  return ctx->fail(GV_ERR_NCCL, "library built without NCCL");
#endif
}

int gv_nccl_world(gv_ctx *ctx, int *rank_out, int *world_out)
{
  if (!ctx) return GV_ERR_INVALID;
  if (rank_out) *rank_out = ctx->rank;
  if (world_out) *world_out = ctx->world;
  return GV_OK;
}

#ifdef GV_WITH_NCCL
// Every rank must hold the same grid geometry and sensor pose (same start cell, hence the same
// sweep table, span ownership and slab split): a mismatch would merge planes of different maps
// and, over peer memory, read and write out of bounds.  Checked collectively (min == max of a
// small descriptor) the first time after the grid or the pose changed.
static int check_ranks_agree(gv_ctx *ctx)
{
  if (ctx->multi_checked) return GV_OK;
  double h[16] = {(double)ctx->g.nx, (double)ctx->g.ny, ctx->g.res, ctx->g.pos_x, ctx->g.pos_y,
                  (double)ctx->bin.sx, (double)ctx->bin.sy, (double)ctx->bin.origin_ok};
  for (int i = 0; i < 8; ++i) h[8 + i] = -h[i];  // max of the negation = -min: one collective
  double *d = nullptr;
  GV_TRY(reserve_t(ctx, S_SMALL, 16, &d));
  GV_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  GV_NCCL(ncclAllReduce(d, d, 16, ncclDouble, ncclMax, ctx->comm, ctx->stream));
  double r[16];
  GV_CUDA(cudaMemcpyAsync(r, d, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 8; ++i)
    GV_REQUIRE(r[i] == -r[8 + i], GV_ERR_STATE,
               "ranks disagree on the grid geometry / sensor pose (field %d: max %.17g, min %.17g): "
               "every rank must call gv_grid_init* and gv_set_base_transform with the same arguments",
               i, r[i], -r[8 + i]);
  ctx->multi_checked = true;
  return GV_OK;
}
#endif

#ifdef GV_WITH_NCCL
static int finalize_multi_chain(gv_ctx *ctx, int32_t k_decay, const double *corners, int nfoot)
{
  const unsigned world = (unsigned)ctx->world, rank = (unsigned)ctx->rank;
  // equal slabs, 4-cell aligned (k_finalize vector width); planes carry kPlanePad slack cells
  size_t slab = (ctx->ncells + world - 1) / world;
  slab = (slab + 3) & ~(size_t)3;
  GV_REQUIRE(slab * world <= ctx->ncells + kPlanePad, GV_ERR_INVALID, "grid too small for %u ranks", world);
  const size_t c0 = (size_t)rank * slab;
  const size_t c1 = c0 + slab < ctx->ncells ? c0 + slab : ctx->ncells;
  if (ctx->p2p) {
    // ---- fused over peer memory (NVLink P2P).  Barriers are k_peer_barrier launches (flag words
    // in peer memory), no NCCL call on this path.
    auto barrier = [&]() -> int {
      ctx->barrier_epoch++;
      k_peer_barrier<<<1, 32, 0, ctx->stream>>>(ctx->peer_flags, rank, world, ctx->barrier_epoch,
                                                ctx->d_flags + kMaxPeers);
      GV_LAUNCH_CHECK();
      return GV_OK;
    };
    stage_mark(ctx, 0);
    GV_TRY(barrier());  // every rank has finished binning into its current end-cell plane
    stage_mark(ctx, 1);
    ctx->ends_dirty = true;  // a rank with no local beams still owns a share of the lines
    // sweep: sums + clears the cells of its spans in every rank's plane, walks its lines
    GV_TRY(raycast_flush_impl(ctx, rank, world, true));
    stage_mark(ctx, 5);
    GV_TRY(barrier());  // every rank's partial hit/miss planes are complete
    stage_mark(ctx, 6);
    int4 *d_rects = nullptr;
    GV_TRY(footprint_rects(ctx, corners, nullptr, nfoot, 0, &d_rects));
    // slab owner: sums all ranks' counts (reduce-scatter), finalises, writes every rank's grid (all-gather)
    if (c1 > c0) GV_TRY(finalize_slab(ctx, k_decay, d_rects, nfoot, c0, c1 - c0, true, true));
    stage_mark(ctx, 7);
    GV_TRY(barrier());  // every slab is in every rank's grid; all count reads are done
    GV_CUDA(cudaMemsetAsync(ctx->d_hit, 0, ctx->ncells * sizeof(int32_t), ctx->stream));
    GV_CUDA(cudaMemsetAsync(ctx->d_miss, 0, ctx->ncells * sizeof(int32_t), ctx->stream));
    stage_mark(ctx, 8);
    stage_collect(ctx);
    ctx->counts_dirty = false;
    ctx->beams_bound = 0;
    return GV_OK;
  }
  // ---- NCCL collectives
  // 1. every rank learns every rank's binned beams: exact u64 sum of the (total,hit) plane
  stage_mark(ctx, 0);
  GV_NCCL(ncclAllReduce(ctx->d_ends, ctx->d_ends, ctx->ncells, ncclUint64, ncclSum, ctx->comm,
                        ctx->stream));
  // 2. the de-duplicated raycast is split by span, partial planes result
  stage_mark(ctx, 1);
  ctx->ends_dirty = true;  // a rank with no local beams still owns a share of the lines
  GV_TRY(raycast_flush_impl(ctx, rank, world));
  stage_mark(ctx, 5);
  // 3. exact int32 sums of the partial planes, scattered by slab
  GV_NCCL(ncclGroupStart());
  GV_NCCL(ncclReduceScatter(ctx->d_hit, ctx->d_hit + (size_t)rank * slab, slab, ncclInt32, ncclSum,
                            ctx->comm, ctx->stream));
  GV_NCCL(ncclReduceScatter(ctx->d_miss, ctx->d_miss + (size_t)rank * slab, slab, ncclInt32,
                            ncclSum, ctx->comm, ctx->stream));
  GV_NCCL(ncclGroupEnd());
  stage_mark(ctx, 6);
  // 4. finalise the local slab
  int4 *d_rects = nullptr;
  GV_TRY(footprint_rects(ctx, corners, nullptr, nfoot, 0, &d_rects));
  if (c1 > c0) GV_TRY(finalize_slab(ctx, k_decay, d_rects, nfoot, c0, c1 - c0, true));
  const size_t padded = ctx->ncells + kPlanePad;
  GV_CUDA(cudaMemsetAsync(ctx->d_hit, 0, padded * sizeof(int32_t), ctx->stream));
  GV_CUDA(cudaMemsetAsync(ctx->d_miss, 0, padded * sizeof(int32_t), ctx->stream));
  stage_mark(ctx, 7);
  // 5. every rank ends with the full grid
  GV_NCCL(ncclGroupStart());
  GV_NCCL(ncclAllGather(ctx->d_lo + c0, ctx->d_lo, slab, ncclFloat32, ctx->comm, ctx->stream));
  GV_NCCL(ncclAllGather(ctx->d_occ + c0, ctx->d_occ, slab, ncclFloat32, ctx->comm, ctx->stream));
  GV_NCCL(ncclGroupEnd());
  stage_mark(ctx, 8);
  stage_collect(ctx);
  ctx->counts_dirty = false;
  ctx->beams_bound = 0;
  return GV_OK;
}
#endif

int gv_grid_finalize_multi(gv_ctx *ctx, int32_t k_decay, const double *corners, int nfoot)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
#ifdef GV_WITH_NCCL
  if (ctx->world == 1 || ctx->comm == nullptr) return gv_grid_finalize(ctx, k_decay, corners, nfoot);
  GV_TRY(check_ranks_agree(ctx));
  MergeScope sc;
  GV_TRY(merge_begin(ctx, &sc));
  return merge_end(ctx, &sc, finalize_multi_chain(ctx, k_decay, corners, nfoot));
#else
  if (ctx->world == 1) return gv_grid_finalize(ctx, k_decay, corners, nfoot);
  return ctx->fail(GV_ERR_NCCL, "library built without NCCL");
#endif
}

int gv_ipc_export(gv_ctx *ctx, void *blob_out)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_REQUIRE(blob_out != nullptr, GV_ERR_INVALID, "blob_out is NULL");
  static_assert(7 * sizeof(cudaIpcMemHandle_t) == GV_IPC_BLOB_BYTES, "ipc blob size");
  cudaIpcMemHandle_t h[7];
  GV_CUDA(cudaIpcGetMemHandle(&h[0], ctx->d_ends_buf[0]));
  GV_CUDA(cudaIpcGetMemHandle(&h[1], ctx->d_ends_buf[1]));
  GV_CUDA(cudaIpcGetMemHandle(&h[2], ctx->d_hit));
  GV_CUDA(cudaIpcGetMemHandle(&h[3], ctx->d_miss));
  GV_CUDA(cudaIpcGetMemHandle(&h[4], ctx->d_lo));
  GV_CUDA(cudaIpcGetMemHandle(&h[5], ctx->d_occ));
  GV_CUDA(cudaIpcGetMemHandle(&h[6], ctx->d_flags));
  memcpy(blob_out, h, sizeof(h));
  return GV_OK;
}

int gv_ipc_import(gv_ctx *ctx, const void *blobs, int world, int rank)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_REQUIRE(ctx->has_grid, GV_ERR_STATE, "grid not initialised");
  GV_REQUIRE(blobs && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, GV_ERR_INVALID,
             "bad rank/world %d/%d (at most %d peers)", rank, world, kMaxPeers);
  GV_REQUIRE(world == ctx->world && rank == ctx->rank, GV_ERR_STATE,
             "gv_nccl_init must come first and agree on rank/world");
  GV_TRY(join_merge(ctx));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  close_peers(ctx);
  ctx->p2p = true;  // so that close_peers cleans up a partial import
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      ctx->peer_ends_buf[0].p[r] = ctx->d_ends_buf[0];
      ctx->peer_ends_buf[1].p[r] = ctx->d_ends_buf[1];
      ctx->peer_hit.p[r] = ctx->d_hit;
      ctx->peer_miss.p[r] = ctx->d_miss;
      ctx->peer_lo.p[r] = ctx->d_lo;
      ctx->peer_occ.p[r] = ctx->d_occ;
      ctx->peer_flags.p[r] = ctx->d_flags;
      continue;
    }
    cudaIpcMemHandle_t h[7];
    memcpy(h, static_cast<const char *>(blobs) + (size_t)r * GV_IPC_BLOB_BYTES, sizeof(h));
    void *q[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 7; ++k) {
      const cudaError_t e = cudaIpcOpenMemHandle(&q[k], h[k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        for (int m = 0; m < k; ++m) cudaIpcCloseMemHandle(q[m]);
        close_peers(ctx);
        return ctx->fail(GV_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d, plane %d): %s", r, k,
                         cudaGetErrorString(e));
      }
    }
    ctx->peer_ends_buf[0].p[r] = static_cast<unsigned long long *>(q[0]);
    ctx->peer_ends_buf[1].p[r] = static_cast<unsigned long long *>(q[1]);
    ctx->peer_hit.p[r] = static_cast<int32_t *>(q[2]);
    ctx->peer_miss.p[r] = static_cast<int32_t *>(q[3]);
    ctx->peer_lo.p[r] = static_cast<float *>(q[4]);
    ctx->peer_occ.p[r] = static_cast<float *>(q[5]);
    ctx->peer_flags.p[r] = static_cast<unsigned *>(q[6]);
  }
  ctx->peer_ends = ctx->peer_ends_buf[ctx->ends_cur];
  // the barrier epochs of all ranks restart together
  ctx->barrier_epoch = 0;
  GV_CUDA(cudaMemset(ctx->d_flags, 0, (kMaxPeers + 1) * sizeof(unsigned)));
  return GV_OK;
}

int gv_ipc_close(gv_ctx *ctx)
{
  if (!ctx) return GV_ERR_INVALID;
  GV_CUDA(cudaSetDevice(ctx->device));
  GV_TRY(join_merge(ctx));
  GV_CUDA(cudaStreamSynchronize(ctx->stream));
  close_peers(ctx);
  return GV_OK;
}

}  // extern "C"
