// hard_voxel.cuh -- deterministic hard voxelization as a parallel pipeline.
//
// Reference semantics: the sequential scan of
// mmdetection3d/mmdet3d/ops/voxel/src/voxelization_cpu.cpp:45-101
//   * voxel ids in order of each voxel's first point           (:75-88)
//   * a NEW voxel is dropped once max_voxels exist             (:80)
//   * a voxel keeps its first max_points points in point order (:91-97)
// None of the reference's GPU formulation (O(N^2) scan + <<<1,1>>> kernel,
// voxelization_cuda.cu:105-180) is reused.  Everything order-related is derived
// from explicit point indices, so the result does not depend on scheduling.
//
//   P1 insert  (R launches over consecutive index ranges of S points, `hv_pass_kernel<Src, 0>`):
//        voxel cell of every point; {key | min point index} into a hash table (64-bit CAS claims an
//        entry, 64-bit atomicMin lowers the index); the bit "point i is the first of its voxel" is
//        kept current by XOR toggles (claim: toggle i; lowering j -> i: toggle both), so no pass
//        over the table is needed to find the first points.  A frame whose earlier rounds already
//        claimed max_voxels voxels is CLOSED: its CTAs of the later rounds return at once (voxels
//        first seen after that point have rank >= max_voxels and the reference drops them), so the
//        table never holds more than max_voxels + S keys, whatever N is.
//   P2 count   (`hv_count_kernel`) per-chunk popcount of the flags; the frame's last CTA scans the chunk totals
//        (-> voxel_num) and lists the marked cells of the 128 x 128 bird's-eye mask the claims left behind.
//      post    (`hv_post_kernel`, one launch) word prefixes + the list of first points by rank (the r-th set flag
//        is the first point of voxel r: rank = popcount prefix, no pass over the table) + their point -> voxel
//        entries; and P2c cull (depth source): per camera and block of image columns, can a pixel of this block
//        reach any marked bird's-eye cell?  Plane test of the block's viewing wedge against the cells, widened by
//        the proven error bound of the pixel->cell map.  Pseudo point clouds are camera-major, so with max_voxels
//        reached early most cameras cannot contribute any more.
//   P3 lookup  (`hv_pass_kernel<Src, 1>`, ONE launch over all points, frame-major so that a frame's
//        table and slot rows stay in L2): tiles whose column blocks are all culled are skipped without
//        reading their depth; every other point: first-point flag (nothing to add) -> cell -> table -> rank ->
//        sorted insertion of its index into S[r][0..K-1) (atomicMin cascade keeping the smallest indices; a full
//        row rejects a later point with one load).  Its leading CTAs fetch the depths of the first points.
//   P4 emit    one lane per voxel, a warp per 32 consecutive ranks: gathers (or re-unprojects exactly) the first
//        point and the slot points into a warp-private shared-memory tile, writes voxels (coalesced), coors,
//        count and the HardSimpleVFE mean.
//
// A warp of the pass kernels owns a STRIP of `iters` consecutive 128-point tiles: the depth of the next
// tile is in flight while the current one is classified; in-range keys are appended to a per-warp list that
// is worked off only in full 32-lane passes (the remainder is carried to the next tile), and the few points
// that need the exact IEEE path are collected over the strip, so that pass runs once per strip with dense
// lanes instead of once per tile with one or two.
//
// Templated on the point source: a (N,C) point array, or DA3 depth maps
// unprojected on the fly (the point cloud never exists in memory).
#pragma once
#include <stdlib.h>

#include "rd3_common.cuh"

namespace rd3 {

constexpr int kPassThreads = 256;
constexpr int kPassWarps = kPassThreads / 32;
#ifndef RD3_EMIT_THREADS
#define RD3_EMIT_THREADS 256
#endif
constexpr int kEmitThreads = RD3_EMIT_THREADS;
#ifndef RD3_INS_MINB
#define RD3_INS_MINB 5                // resident CTAs per SM the insert pass is compiled for
#endif
#ifndef RD3_LKP_MINB
#define RD3_LKP_MINB 5                // ... and the lookup pass
#endif
#ifndef RD3_EMIT_MINB
#define RD3_EMIT_MINB 6               // resident CTAs per SM the emit kernel is compiled for (shared memory allows 6 at K*C = 30)
#endif
constexpr int kTilePoints = 128;      // points per warp tile (4 per lane)
constexpr int kListCap = 160;         // per-warp item / undecided lists: 31 carried + 128 new
constexpr int kMaxRounds = 64;
constexpr uint32_t kDummyKey = 0xFFFFFFE0u;   // + lane: 32 values no voxel key takes (make_grid: volume <= kDummyKey)
constexpr int kBevDim = 128;          // bird's-eye mask of the kept voxels: kBevDim x kBevDim bits over the x,y grid
                                      // (64 x 64 keeps 13 of 42 column blocks alive in the mixture scene, 128 x 128 eleven)
constexpr int kBevWords = kBevDim * kBevDim / 32;
constexpr int kCullDoubles = 16;      // per camera: inverse of the direct cell map [9], T [3], margin, ok
constexpr int kBevCopies = 8;         // privatised copies of the mask (power of two): spreads the marking atomics
constexpr int kEmitList = 96;         // emit: entries of a warp's list of later points (worked off when it could overflow)
constexpr int kLocalIters = 16;       // insert pass: a warp whose strip has at most this many tiles keeps the first-point
                                      // bits of its own points in shared memory (4 words per tile) and flushes them once

// ---------------------------------------------------------------------------
// packed fp32 pairs (FFMA2 / FADD2 on sm_100a: two IEEE-rounded results per issue slot)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n add.rn.f32x2 rd, ra, rb;\n"
      " mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n sub.rn.f32x2 rd, ra, rb;\n"
      " mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

// ---------------------------------------------------------------------------
// point sources
// ---------------------------------------------------------------------------
// culling masks of the lookup pass (depth source): [B][kMaxCams] live column blocks per camera, or null
struct HvCull {
  const uint32_t *mask;
};

// classification of one lane's 4 consecutive points
struct Quad {
  uint32_t key[4];
  unsigned in, und;     // bit q: point q is surely inside (key valid) / needs the exact path
};

// Both sources hand the pass kernels their work through a WALKER: a warp visits a sequence of 128-element
// tiles, lane L owning 4 consecutive elements of each.  Neither pass depends on the order in which the
// elements of its range are visited (the table keeps minima, the slot rows sort), so the decomposition is free:
//   points : consecutive tiles of the flat (N, C) array
//   depth  : a 128-column block of `iters` consecutive image rows -- the lane's column never changes (no
//            index -> (cam, v, u) division per tile, the row advances by one pointer increment), and a whole
//            CTA of culled columns leaves before its prologue.
struct PointsSource {
  const float *pts;   // (B, N, C)
  int64_t N;
  int C;
  static constexpr bool kIsDepth = false;

  __device__ __forceinline__ bool stage_async(float *, uint64_t *, int) const { return false; }
  int host_num_feats() const { return C; }
  __device__ __forceinline__ int num_feats() const { return C; }
  // CTAs that cover the element range [begin, end) with strips of `iters` tiles per warp
  unsigned host_grid(int64_t begin, int64_t end, int iters) const {
    return (unsigned)ceil_div(end - begin, (int64_t)kPassWarps * iters * kTilePoints);
  }
  int64_t host_round_multiple() const { return 1024; }

  struct Walker {
    int64_t cur, end;     // first element of the warp's current tile, end of its strip
    int64_t cur0;         // first element of the strip
    int lane;
  };
  // element index of position lp (tile number * 128 + offset) of the warp's strip
  __device__ __forceinline__ uint32_t strip_index(const Walker &k, uint32_t lp) const { return (uint32_t)k.cur0 + lp; }
  __device__ __forceinline__ bool cta_live(const HvCull &, int, int64_t, int64_t, int, int) const { return true; }
  __device__ __forceinline__ bool walk_init(Walker &k, int64_t begin, int64_t end, int iters, int bx, int wv, int lane) const {
    const int64_t strip = (int64_t)iters * kTilePoints;
    k.cur = begin + ((int64_t)bx * kPassWarps + wv) * strip;
    k.cur0 = k.cur;
    k.end = k.cur + strip < end ? k.cur + strip : end;
    k.lane = lane;
    return k.cur < end;
  }
  __device__ __forceinline__ bool walk_more(const Walker &k) const { return k.cur < k.end; }
  __device__ __forceinline__ void walk_next(Walker &k) const { k.cur += kTilePoints; }
  __device__ __forceinline__ bool walk_live(const Walker &, const uint32_t *) const { return true; }
  __device__ __forceinline__ uint32_t walk_index(const Walker &k) const { return (uint32_t)k.cur + 4u * k.lane; }
  __device__ __forceinline__ int walk_npx(const Walker &k) const {
    const int64_t left = k.end - (k.cur + 4 * k.lane);
    return left >= 4 ? 4 : (left > 0 ? (int)left : 0);
  }

  struct Pre {};
  __device__ __forceinline__ Pre preload(int, const Walker &, bool) const { return Pre(); }
  struct Cursor { const float *p; int npx; };
  __device__ __forceinline__ Cursor cursor(int b, const Walker &k, const float *, const Pre &) const {
    Cursor c;
    c.npx = walk_npx(k);
    c.p = pts + ((int64_t)b * N + (c.npx ? k.cur + 4 * k.lane : 0)) * C;   // lanes past the end read point 0 (masked)
    return c;
  }
  // the lane's 4 points: keys of the points surely inside (bit q of `in`), undecided ones in `und`
  __device__ __forceinline__ void classify(const Cursor &c, const float *, const VoxelGrid &g, Quad &qd) const {
    qd.in = 0; qd.und = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float *pq = c.p + (q < c.npx ? q : 0) * C;
      const int r = voxel_key_fast(__ldg(pq), __ldg(pq + 1), __ldg(pq + 2), g, qd.key[q]);
      const unsigned live = q < c.npx ? 1u : 0u;
      qd.in |= (r == 1 ? live : 0u) << q;
      qd.und |= (r == 2 ? live : 0u) << q;
    }
  }
  __device__ __forceinline__ bool cell_exact(int b, int64_t i, const float *, const VoxelGrid &g,
                                             int &cx, int &cy, int &cz) const {
    const float *p = pts + ((int64_t)b * N + i) * C;
    return voxel_coor(__ldg(p), __ldg(p + 1), __ldg(p + 2), g, cx, cy, cz);
  }
  // emit: all features of point i into dst[0..C)
  __device__ __forceinline__ void gather(int b, int64_t i, const float *, float *dst) const {
    const float *p = pts + ((int64_t)b * N + i) * C;
    for (int c = 0; c < C; ++c) dst[c] = __ldg(p + c);
  }
  // a voxel's first point, listed by the post kernel: nothing to carry besides the index
  static constexpr bool kCarryFirst = false;
  __device__ __forceinline__ float first_payload(int, uint32_t) const { return 0.0f; }
  __device__ __forceinline__ void gather_first(int b, uint32_t i, float, const float *s_cal, float *dst) const {
    gather(b, i, s_cal, dst);
  }
};

struct DepthSource {
  const float *depth;      // (B, npix)
  const float *conf;       // (B, npix) or null
  const uint8_t *sky;      // (B, npix) or null
  const float *sky_prob;   // (B, npix) or null: raw sky probability, sky iff >= sky_thr (used when sky is null)
  float sky_thr;
  const float *intr;       // (B, ncam, 9)
  const float *c2l;        // (B, ncam, 16)
  const float *cal_table;  // (B, ncam, kCalibFloats) precomputed by calib_kernel, or null
  const double *cull_cal;  // (B, ncam, kCullDoubles) inverse cell map per camera for the culling test (calib_kernel), or null
  const float *cull_planes; // (B, ncam, column blocks, 5 planes, 4) the wedge planes of every (camera, column block), or null
  DepthParams p;
  CellRange rg;            // range filter in cell units (fused path); on = 0 when the grid test implies it
  int vec_ok;              // 16-byte aligned float4 loads of 4 pixels of a row are legal (W % 4 == 0, aligned base)
  int mask_vec_ok;         // ... and the same for the conf / sky / sky_prob maps (16 / 4 / 16-byte loads)
  int cbshift;             // culling: image columns are grouped in blocks of 2^cbshift (>= 128, <= 32 blocks)
  FastDiv div_h;           // global row -> camera
  static constexpr bool kIsDepth = true;

  // copy this frame's calibration into shared memory (no barrier)
  __device__ __forceinline__ void stage(float *s_cal, int b) const {
    if (cal_table) {
      const float *src = cal_table + (int64_t)b * p.ncam * kCalibFloats;
      for (int i = threadIdx.x; i < p.ncam * kCalibFloats; i += blockDim.x) s_cal[i] = __ldg(src + i);
    } else {
      stage_calibration(s_cal, intr + (int64_t)b * p.ncam * 9, c2l + (int64_t)b * p.ncam * 16, p.ncam);
    }
  }
  // the same copy as one TMA bulk transfer issued by thread 0 (true: wait with tma_wait after the
  // CTA's next barrier); falls back to the load / store loop when there is no precomputed table
  __device__ __forceinline__ bool stage_async(float *s_cal, uint64_t *s_bar, int b) const {
    if (!cal_table) {
      stage(s_cal, b);
      return false;
    }
    if (threadIdx.x == 0)
      tma_load_1d(s_cal, cal_table + (int64_t)b * p.ncam * kCalibFloats, (uint32_t)(p.ncam * kCalibFloats * 4), s_bar);
    return true;
  }
  __device__ __forceinline__ void prepare(float *s_cal, int b) const {
    stage(s_cal, b);
    __syncthreads();
  }
  int host_num_feats() const { return 3; }
  __device__ __forceinline__ int num_feats() const { return 3; }
  int host_col_tiles() const { return (p.W + kTilePoints - 1) / kTilePoints; }
  // [begin, end) are multiples of W (whole image rows): a CTA owns 8 * iters rows of one 128-column tile
  unsigned host_grid(int64_t begin, int64_t end, int iters) const {
    const int64_t rows = (end - begin) / p.W;
    return (unsigned)(ceil_div(rows, (int64_t)kPassWarps * iters) * host_col_tiles());
  }
  int64_t host_round_multiple() const { return p.W; }

  __device__ __forceinline__ bool depth_ok(float z, int64_t gidx, int b) const {
    // z > 0 & isfinite(z) [& z <= max_depth] (:338-340); p.zmax = min(max_depth, FLT_MAX)
    bool ok = (z > 0.0f) && (z <= p.zmax);
    if (p.use_masks) {
      if (ok && p.use_conf)
        ok = __ldg(conf + gidx) >= (p.conf_thresh_dev ? __ldg(p.conf_thresh_dev + b) : p.conf_thresh);
      if (ok && p.use_sky) ok = sky ? __ldg(sky + gidx) == 0 : !(__ldg(sky_prob + gidx) >= sky_thr);
    }
    return ok;
  }
  // pixel index -> (cam, v, u)
  __device__ __forceinline__ void pixel_cvu(uint32_t pix, uint32_t &cam, uint32_t &v, uint32_t &u) const {
    cam = fast_div(pix, p.div_hw);
    const uint32_t rem = pix - cam * (uint32_t)p.HW;
    v = fast_div(rem, p.div_w);
    u = rem - v * (uint32_t)p.W;
  }

  struct Walker {
    uint32_t gr, gr_end;   // global image row (cam * H + v) of the warp's current tile, end of its rows
    uint32_t gr0, c0;      // first row of the strip, first column of the tile
    uint32_t cam, v;       // the same row as (camera, row)
    uint32_t u;            // the lane's first column (fixed)
    uint32_t blk;          // culling block of the column tile
    int npx;               // columns u .. u + npx - 1 exist (0 for lanes right of the image)
  };
  // rows [r0, r1) of the CTA bx for the element range [begin, end)
  __device__ __forceinline__ void cta_rows(int64_t begin, int64_t end, int iters, uint32_t bx, uint32_t &r0, uint32_t &r1,
                                           uint32_t &ct) const {
    const uint32_t nct = (uint32_t)((p.W + kTilePoints - 1) / kTilePoints);
    const uint32_t rg = bx / nct;
    ct = bx - rg * nct;
    const uint32_t first = fast_div((uint32_t)begin, p.div_w), last = fast_div((uint32_t)end, p.div_w);   // < 2^30 pixels
    r0 = first + rg * (uint32_t)(kPassWarps * iters);
    r1 = r0 + (uint32_t)(kPassWarps * iters);
    if (r1 > last) r1 = last;
  }
  // lookup pass: can any row of this CTA's column tile reach a kept voxel?  (before the prologue: one or two loads)
  __device__ __forceinline__ bool cta_live(const HvCull &c, int b, int64_t begin, int64_t end, int iters, int bx) const {
    if (!c.mask) return true;
    uint32_t r0, r1, ct;
    cta_rows(begin, end, iters, (uint32_t)bx, r0, r1, ct);
    if (r0 >= r1) return false;
    const uint32_t blk = (ct * kTilePoints) >> cbshift;
    const uint32_t c0 = fast_div(r0, div_h), c1 = fast_div(r1 - 1, div_h);
    uint32_t m = 0;
    for (uint32_t cam = c0; cam <= c1; ++cam) m |= __ldg(c.mask + b * kMaxCams + cam);
    return (m >> blk) & 1u;
  }
  __device__ __forceinline__ bool walk_init(Walker &k, int64_t begin, int64_t end, int iters, int bx, int wv, int lane) const {
    uint32_t r0, r1, ct;
    cta_rows(begin, end, iters, (uint32_t)bx, r0, r1, ct);
    k.gr = r0 + (uint32_t)(wv * iters);
    k.gr0 = k.gr;
    k.c0 = ct * kTilePoints;
    k.gr_end = k.gr + (uint32_t)iters < r1 ? k.gr + (uint32_t)iters : r1;
    k.cam = fast_div(k.gr, div_h);
    k.v = k.gr - k.cam * (uint32_t)p.H;
    k.u = ct * kTilePoints + 4u * lane;
    k.blk = (ct * kTilePoints) >> cbshift;
    const int left = p.W - (int)k.u;
    k.npx = left >= 4 ? 4 : (left > 0 ? left : 0);
    return k.gr < r1;
  }
  __device__ __forceinline__ bool walk_more(const Walker &k) const { return k.gr < k.gr_end; }
  __device__ __forceinline__ void walk_next(Walker &k) const {
    ++k.gr;
    if (++k.v == (uint32_t)p.H) { k.v = 0; ++k.cam; }
  }
  __device__ __forceinline__ bool walk_live(const Walker &k, const uint32_t *s_cull) const {
    return (s_cull[k.cam] >> k.blk) & 1u;
  }
  __device__ __forceinline__ uint32_t walk_index(const Walker &k) const { return k.gr * (uint32_t)p.W + k.u; }
  // pixel index of position lp (row of the strip * 128 + column offset inside the tile)
  __device__ __forceinline__ uint32_t strip_index(const Walker &k, uint32_t lp) const {
    return (k.gr0 + (lp >> 7)) * (uint32_t)p.W + k.c0 + (lp & 127u);
  }
  __device__ __forceinline__ int walk_npx(const Walker &k) const { return k.npx; }

  // The depth load does not depend on the calibration: the load of the NEXT tile is issued before the
  // current one is classified.
  struct Pre { float z[4]; };
  __device__ __forceinline__ Pre preload(int b, const Walker &k, bool on) const {
    Pre r;
    const float *src = depth + (int64_t)b * p.npix + walk_index(k);
    if (on && vec_ok && k.npx == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(src));
      r.z[0] = t.x; r.z[1] = t.y; r.z[2] = t.z; r.z[3] = t.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) r.z[q] = (on && q < k.npx) ? __ldg(src + q) : 0.0f;
    }
    return r;
  }
  // one lane's 4 consecutive pixels of one image row
  struct Cursor {
    float z[4];
    unsigned valid;        // depth/conf/sky mask of the 4 pixels
    uint32_t cam;
    float uf;              // column of the first pixel
    float tx, ty, tz;      // row part of the direct cell map
  };
  __device__ __forceinline__ Cursor cursor(int b, const Walker &k, const float *s_cal, const Pre &pre) const {
    Cursor c;
#pragma unroll
    for (int q = 0; q < 4; ++q) c.z[q] = pre.z[q];
    // z > 0 & isfinite(z) [& z <= max_depth] (:338-340); p.zmax = min(max_depth, FLT_MAX)
    c.valid = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) c.valid |= ((c.z[q] > 0.0f) & (c.z[q] <= p.zmax)) ? (1u << q) : 0u;
    c.valid &= (1u << k.npx) - 1u;
    if (p.use_masks && c.valid) {
      const int64_t gi = (int64_t)b * p.npix + walk_index(k);
      if (vec_ok && k.npx == 4 && mask_vec_ok) {
        // the 4 pixels' confidences with one 16-byte load, their sky bytes with one 4-byte load
        unsigned keep = 0xFu;
        if (p.use_conf) {
          const float thr = p.conf_thresh_dev ? __ldg(p.conf_thresh_dev + b) : p.conf_thresh;
          const float4 cf = __ldg(reinterpret_cast<const float4 *>(conf + gi));
          keep = (cf.x >= thr ? 1u : 0u) | (cf.y >= thr ? 2u : 0u) | (cf.z >= thr ? 4u : 0u) | (cf.w >= thr ? 8u : 0u);
        }
        if (p.use_sky) {
          if (sky) {
            const uint32_t sb = __ldg(reinterpret_cast<const uint32_t *>(sky + gi));
            keep &= ((sb & 0xFFu) ? 0u : 1u) | ((sb & 0xFF00u) ? 0u : 2u) | ((sb & 0xFF0000u) ? 0u : 4u) | ((sb & 0xFF000000u) ? 0u : 8u);
          } else {
            const float4 sp = __ldg(reinterpret_cast<const float4 *>(sky_prob + gi));
            keep &= (sp.x >= sky_thr ? 0u : 1u) | (sp.y >= sky_thr ? 0u : 2u) | (sp.z >= sky_thr ? 0u : 4u) | (sp.w >= sky_thr ? 0u : 8u);
          }
        }
        c.valid &= keep;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (((c.valid >> q) & 1u) && !depth_ok(c.z[q], gi + q, b)) c.valid &= ~(1u << q);
      }
    }
    c.cam = k.cam;
    c.uf = (float)k.u;
    pixel_cell_row((float)k.v, s_cal + c.cam * kCalibFloats, c.tx, c.ty, c.tz);
    return c;
  }
  // The direct pixel->cell map (rd3_common.cuh: pixel_key_fast) decides all but the pixels within its error
  // bound of a cell / range-filter boundary; those (bit q of `und`) are redone by cell_exact.  Computed
  // for every pixel (masked ones yield garbage that is discarded): no divergence.
  // Same arithmetic as pixel_key_fast, two pixels per instruction (fma.rn.f32x2 / add.rn.f32x2: each
  // half is the separately rounded IEEE result) and the three |d| of a pixel reduced by one 3-input max.
  __device__ __forceinline__ void classify(const Cursor &c, const float *s_cal, const VoxelGrid &g, Quad &qd) const {
    const float *cal = s_cal + c.cam * kCalibFloats;
    unsigned in = 0, und = 0;
    if (rg.on) {             // a range-filter plane the grid test does not imply: generic (scalar) form
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = pixel_key_fast(c.z[q], __fadd_rn(c.uf, (float)q), c.tx, c.ty, c.tz, cal, g, rg, qd.key[q]);
        in |= (r == 1 ? 1u : 0u) << q;
        und |= (r == 2 ? 1u : 0u) << q;
      }
    } else {
      const float *k = cal + kCalDirect;
      const float row[3] = {c.tx, c.ty, c.tz};
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {                   // pixels (0,1), then (2,3): half the live registers
        const float2 z = make_float2(c.z[2 * pr], c.z[2 * pr + 1]);
        const float2 u = fadd2(splat2(c.uf), make_float2((float)(2 * pr), (float)(2 * pr + 1)));
        float2 m[3], d[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float2 h = ffma2(z, ffma2(splat2(k[a * 4]), u, splat2(row[a])), splat2(k[a * 4 + 3]));
          m[a] = fadd2(h, splat2(kMagic));
          d[a] = fsub2(h, fsub2(m[a], splat2(kMagic)));
        }
        const float2 thr = ffma2(z, splat2(k[12]), splat2(k[13]));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int q = 2 * pr + e;
          const uint32_t ix = (uint32_t)(__float_as_int(e ? m[0].y : m[0].x) - kMagicBits),
                         iy = (uint32_t)(__float_as_int(e ? m[1].y : m[1].x) - kMagicBits),
                         iz = (uint32_t)(__float_as_int(e ? m[2].y : m[2].x) - kMagicBits);
          const float dmax = fmaxf(fabsf(e ? d[0].y : d[0].x), fmaxf(fabsf(e ? d[1].y : d[1].x), fabsf(e ? d[2].y : d[2].x)));
          const bool decided = dmax < (e ? thr.y : thr.x);
          const bool inside = (ix < (uint32_t)g.grid[0]) & (iy < (uint32_t)g.grid[1]) & (iz < (uint32_t)g.grid[2]);
          qd.key[q] = (iz * (uint32_t)g.grid[1] + iy) * (uint32_t)g.grid[0] + ix;
          in |= ((decided & inside) ? 1u : 0u) << q;
          und |= (decided ? 0u : 1u) << q;
        }
      }
    }
    qd.in = g.fast_ok ? (in & c.valid) : 0u;
    qd.und = g.fast_ok ? (und & c.valid) : c.valid;
  }
  // pixel index -> exact ego-frame point (reference arithmetic); false if the range filter drops it
  __device__ __forceinline__ bool point(int b, int64_t i, const float *s_cal, float &x, float &y,
                                        float &z) const {
    const float d = __ldg(depth + (int64_t)b * p.npix + i);
    uint32_t cam, v, u;
    pixel_cvu((uint32_t)i, cam, v, u);
    return unproject_point(d, (int)u, (int)v, s_cal + cam * kCalibFloats, p, x, y, z);
  }
  __device__ __forceinline__ bool cell_exact(int b, int64_t i, const float *s_cal, const VoxelGrid &g,
                                             int &cx, int &cy, int &cz) const {
    float x, y, z;
    if (!point(b, i, s_cal, x, y, z)) return false;
    return voxel_coor(x, y, z, g, cx, cy, cz);
  }
  __device__ __forceinline__ void gather(int b, int64_t i, const float *s_cal, float *dst) const {
    point(b, i, s_cal, dst[0], dst[1], dst[2]);
  }
  // a voxel's first point: the post kernel stores its depth next to its index (its warps have the latency to spare),
  // so the emit kernel's first points do not start with two dependent DRAM round trips
  static constexpr bool kCarryFirst = true;
  __device__ __forceinline__ float first_payload(int b, uint32_t i) const { return __ldg(depth + (int64_t)b * p.npix + i); }
  __device__ __forceinline__ void gather_first(int, uint32_t i, float d, const float *s_cal, float *dst) const {
    uint32_t cam, v, u;
    pixel_cvu(i, cam, v, u);
    unproject_point(d, (int)u, (int)v, s_cal + cam * kCalibFloats, p, dst[0], dst[1], dst[2]);
  }
};

// ---------------------------------------------------------------------------
// per-call work description (device pointers, all with a leading frame dim)
// ---------------------------------------------------------------------------
struct HvWork {
  unsigned long long *table;  // [B][cap]   {key:32 | min point idx:32}, after P2 {key:32 | rank:32}; empty = ~0
  uint32_t *slots;            // [B][max_voxels][K-1] sorted indices of the points after the first, empty = ~0
  uint32_t *first_of;         // [B][max_voxels] first point of the voxel of rank r (the r-th set bit of flags)
  float *first_z;             // [B][max_voxels] its depth (depth source only)
  uint32_t *flags;            // [B][nwords] bit i: point i is the first of a voxel
  int32_t *wordprefix;        // [B][nwords] exclusive popcount prefix inside the chunk
  int32_t *chunk_base;        // [B][nchunks] totals, then exclusive bases after the chunk scan
  int32_t *round_claims;      // [B][kMaxRounds] voxels claimed per insert round
  int32_t *scan_done;         // [B] CTAs of hv_count_kernel that have finished (the last one scans the chunk totals)
  uint32_t *bev;              // [B][kBevCopies][kBevWords] bird's-eye masks of the kept voxels (OR of the copies), or null
  uint32_t *cull;             // [B][kMaxCams] live column blocks per camera, or null (nothing culled)
  uint16_t *bev_list;         // [B][kBevDim^2] the marked cells of the folded mask (count kernel), bev_count[B] of them
  int32_t *bev_count;
  uint32_t *later;            // [B][lwords] bit r: the voxel of rank r has points after its first one (its slot row is in use)
  int32_t *p2v;               // [B][N] point -> voxel map or null
  const int32_t *vnum;        // [B] voxel_num (valid after the count kernel)
  FastDiv div_gx, div_gy;     // key -> (x, y) cell
  uint32_t bev_kx, bev_ky;    // bird's-eye cell of voxel column i: (i * k) >> 20, k = floor(kBevDim 2^20 / grid)
  FastDiv div_Km1;            // slot-row word -> voxel
  int64_t N;
  int b0;                     // first frame of the group this launch works on
  int64_t cap;
  uint32_t cap_mask;
  int log2cap;
  int direct;                 // grid volume <= cap: slot = key, no probing
  int lwords;
  int nwords;                 // multiple of kChunkWords
  int nchunks;
  int K;                      // max_points
  int max_voxels;
};

struct HvOut {
  float *voxels;          // [B][max_voxels][K][C]
  int32_t *coors;         // [B][max_voxels][3]
  int32_t *num;           // [B][max_voxels]
  float *mean;            // [B][max_voxels][F] or null
  int32_t *voxel_num;     // [B]
  int32_t *point2voxel;   // [B][N] or null (pre-filled with -1)
  int F;
};

// The table is probed linearly from the start of the key's bucket of 4 entries (= one 32-byte
// sector).  Every probe sequence enters a bucket at its entry 0 and entries are never released, so
// a bucket fills in order and "entry 3 is empty" <=> the bucket never overflowed <=> a key that
// hashes here and is not in it is absent: a lookup is one 256-bit load (one L1 request, one
// sector) however it ends, and only moves on when the bucket is full without a match.
__device__ __forceinline__ uint32_t hash_bucket_slot(uint32_t key, int log2cap) {
  return ((key * 2654435769u) >> (34 - log2cap)) << 2;
}

__device__ __forceinline__ void load_bucket(const unsigned long long *p, unsigned long long &e0,
                                            unsigned long long &e1, unsigned long long &e2,
                                            unsigned long long &e3) {
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(e0), "=l"(e1), "=l"(e2), "=l"(e3)
               : "l"(p));
}

// "point idx is / is no longer the first point of its voxel": XOR, so that the owner's toggle and the
// toggle of whoever displaces it commute (an index that was displaced is toggled exactly twice)
__device__ __forceinline__ void first_toggle(uint32_t *flags, uint32_t idx) {
  atomicXor(flags + (idx >> 5), 1u << (idx & 31));
}

// Insert (key, idx).  Slot ownership is permanent (CAS from empty); the payload
// only ever decreases (atomicMin), so a stale read can only cause a redundant
// atomic, never a wrong skip.  Returns 1 iff this call created the entry, 2 iff it lowered the entry's index, else 0:
// in both cases `idx` became the first point of its voxel and the CALLER toggles its flag bit (the insert pass keeps
// the bits of its own points in shared memory); the displaced holder's bit is toggled here.
// (Measured: issuing the CAS first, without the load, is slower -- repeated keys then pay an atomic on a hot
// entry where a load would have told them to leave.)
__device__ __forceinline__ int table_insert(unsigned long long *table, uint32_t *flags, const HvWork &w,
                                            uint32_t key, uint32_t idx) {
  const unsigned long long mine = ((unsigned long long)key << 32) | idx;
  uint32_t slot;
  unsigned long long e;
  if (w.direct) {
    slot = key;
    e = __ldcg(table + slot);
  } else {
    // one 256-bit load finds the first entry of the bucket that holds the key or is empty
    uint32_t s0 = hash_bucket_slot(key, w.log2cap);
    while (true) {
      unsigned long long e0, e1, e2, e3;
      load_bucket(table + s0, e0, e1, e2, e3);
      const uint32_t k0 = (uint32_t)(e0 >> 32), k1 = (uint32_t)(e1 >> 32), k2 = (uint32_t)(e2 >> 32),
                     k3 = (uint32_t)(e3 >> 32);
      int q = 4;
      if (k3 == key || k3 == kEmpty32) { q = 3; e = e3; }
      if (k2 == key || k2 == kEmpty32) { q = 2; e = e2; }
      if (k1 == key || k1 == kEmpty32) { q = 1; e = e1; }
      if (k0 == key || k0 == kEmpty32) { q = 0; e = e0; }
      slot = s0 + q;
      if (q < 4) break;
      s0 = (s0 + 4) & w.cap_mask;
    }
  }
  while (true) {                       // `e` may be stale: the atomics decide
    if (e == kEmpty64) {
      const unsigned long long old = atomicCAS(table + slot, kEmpty64, mine);
      if (old == kEmpty64) return 1;
      e = old;
    }
    if ((uint32_t)(e >> 32) == key) {
      if ((uint32_t)e > idx) {
        const unsigned long long old = atomicMin(table + slot, mine);
        if ((uint32_t)old > idx) {     // this point is the voxel's first one now, the previous holder is not
          first_toggle(flags, (uint32_t)old);
          return 2;
        }
      }
      return 0;
    }
    slot = (slot + 1) & w.cap_mask;    // lost the entry to another key: plain linear probing from here
    e = __ldcg(table + slot);
  }
}

// "the voxel of `key` was claimed": its column is marked in one of the frame's kBevCopies bird's-eye masks
// (privatised: all claims of a frame on one 512-byte mask would serialise in L2).  The mask may hold voxels that
// are dropped later (rank >= max_voxels): it only has to contain the kept ones.
__device__ __forceinline__ uint32_t bev_bit(const HvWork &w, uint32_t key) {
  const uint32_t t = fast_div(key, w.div_gx);
  const uint32_t cx = key - t * w.div_gx.d;
  const uint32_t cy = t - fast_div(t, w.div_gy) * w.div_gy.d;
  return ((cy * w.bev_ky) >> 20) * kBevDim + ((cx * w.bev_kx) >> 20);
}

// rank (first-occurrence order) of the voxel whose first point is `first_idx`
__device__ __forceinline__ int voxel_rank(const HvWork &w, int b, uint32_t first_idx) {
  const uint32_t word = first_idx >> 5;
  const int64_t wi = (int64_t)b * w.nwords + word;
  const uint32_t bits = __ldg(w.flags + wi) & ((1u << (first_idx & 31)) - 1u);
  return __ldg(w.chunk_base + (int64_t)b * w.nchunks + (first_idx >> kChunkShift)) +
         __ldg(w.wordprefix + wi) + __popc(bits);
}

// Lookup (the table is not written any more): the first point of the key's voxel, kEmpty32 when the key is absent.
__device__ __forceinline__ uint32_t table_first(const unsigned long long *table, const HvWork &w, uint32_t key) {
  if (w.direct) return (uint32_t)__ldcg(table + key);            // empty entries read ~0 as well
  uint32_t s0 = hash_bucket_slot(key, w.log2cap);
  while (true) {
    unsigned long long e0, e1, e2, e3;
    load_bucket(table + s0, e0, e1, e2, e3);
    uint32_t r = kEmpty32;
    bool hit = false;
    if ((uint32_t)(e3 >> 32) == key) { r = (uint32_t)e3; hit = true; }
    if ((uint32_t)(e2 >> 32) == key) { r = (uint32_t)e2; hit = true; }
    if ((uint32_t)(e1 >> 32) == key) { r = (uint32_t)e1; hit = true; }
    if ((uint32_t)(e0 >> 32) == key) { r = (uint32_t)e0; hit = true; }
    if (hit || (uint32_t)(e3 >> 32) == kEmpty32) return r;       // no valid key is ~0
    s0 = (s0 + 4) & w.cap_mask;
  }
}

// Keep the K smallest point indices of a voxel, sorted, with atomicMin only.  Every point cascades through
// the row: a value reaches slot k only after losing against the slots before it, so the non-empty prefix is
// strictly increasing at all times and the final content is independent of arrival order (every point is
// offered exactly once).
__device__ __forceinline__ void slot_insert(uint32_t *S, int K, uint32_t idx, uint32_t *later, int r) {
  // The last slot and the first eight are requested together (one memory round trip).  A (possibly stale,
  // hence larger) copy of the last slot that is already smaller than idx: K smaller indices exist.
  const uint32_t last = __ldcg(S + K - 1);
  uint32_t cur = idx;
  int k = 0;
  // Skip the prefix of smaller indices with 8 independent loads at a time instead of one
  // dependent load per slot.  A slot read as smaller than cur can only have decreased since: it stays smaller.
  while (k < K) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (k + i < K) ? __ldcg(S + k + i) : kEmpty32;
    if (last < idx) return;
    int i = 8;
#pragma unroll
    for (int q = 7; q >= 0; --q) i = (v[q] < cur) ? i : q;      // first slot not smaller than cur
    k += i;
    if (i < 8) break;
  }
  for (; k < K; ++k) {
    const uint32_t old = atomicMin(S + k, cur);
    if (old == kEmpty32) {
      // word 0 leaves the empty state exactly once per row: that inserter tells the emit kernel to read the row
      if (k == 0) atomicOr(later + (r >> 5), 1u << (r & 31));
      return;
    }
    if (old > cur) cur = old;            // displaced a larger index: carry it on
  }
}

// P1 / P3 ---------------------------------------------------------------------
// 256 threads; every WARP walks its own sequence of 128-element tiles (Src::Walker) and runs without block
// barriers after the prologue:
//   per tile: one 16-byte load per lane (4 consecutive elements; the next tile's load is already in
//        flight), validity, voxel cell by the conservative fast path; in-range keys are appended to the
//        warp's item list, the few undecided elements to its second list
//   item passes (32 lanes each) while 32 items are waiting: MODE 0 table insert; MODE 1 table lookup -> rank
//        -> slot row
//   undecided elements: exact IEEE arithmetic, once 32 have collected or at the end of the walk
template <class Src, int MODE>
__global__ void __launch_bounds__(kPassThreads, MODE ? RD3_LKP_MINB : RD3_INS_MINB)
    hv_pass_kernel(Src src, VoxelGrid g, HvWork w, int32_t *point2voxel, int64_t begin, int64_t end, int round,
                   int iters, int nfz) {
  __shared__ __align__(16) float s_cal[Src::kIsDepth ? kMaxCams * kCalibFloats : 4];
  __shared__ __align__(8) uint64_t s_bar;        // completion of the calibration's TMA copy
  __shared__ uint2 s_itemb[kPassWarps * kListCap];     // (key, element index) of the in-range elements
  __shared__ uint32_t s_undb[kPassWarps * kListCap];   // indices of the undecided elements
  __shared__ uint2 s_hitb[MODE ? kPassWarps * 64 : 1]; // lookup: (first point of the voxel, element index) of the table hits
  __shared__ uint32_t s_cull[kMaxCams];
  __shared__ int s_prev;
  // insert: per warp, the first-point bits of the strip's own points (4 words per tile) and the bird's-eye marks of its
  // claims; both reach global memory once, at the end of the walk, with one atomic per non-zero word
  __shared__ uint32_t s_lflagb[MODE == 0 ? kPassWarps * kLocalIters * 4 : 1];
  __shared__ uint32_t s_lbev[MODE == 0 ? kBevWords : 1];   // the CTA's bird's-eye marks

  const int b = blockIdx.y + w.b0;
  const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  if (MODE == 1 && Src::kCarryFirst && (int)blockIdx.x < nfz) {
    // The first nfz CTAs of every frame fetch what the emit kernel needs of the voxels' first points besides their
    // index (depth source: the depth value): 1024 ranks per CTA, four independent loads per thread.  These CTAs wait
    // for memory while the lookup CTAs around them issue instructions; in the emit kernel the same two dependent
    // round trips would stall every warp.
    const int vn = __ldg(w.vnum + b);
    uint32_t idx[4];
    float z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = (int)blockIdx.x * 1024 + j * kPassThreads + tid;
      idx[j] = r < vn ? __ldg(w.first_of + (int64_t)b * w.max_voxels + r) : kEmpty32;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = idx[j] != kEmpty32 ? src.first_payload(b, idx[j]) : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = (int)blockIdx.x * 1024 + j * kPassThreads + tid;
      if (r < vn) w.first_z[(int64_t)b * w.max_voxels + r] = z[j];
    }
    return;
  }
  const int bx = (int)blockIdx.x - (MODE == 1 ? nfz : 0);
  if (MODE == 1) {
    // a column tile no kept voxel can be seen from: nothing to do, nothing to stage
    HvCull cl{w.cull};
    if (!src.cta_live(cl, b, begin, end, iters, bx)) return;
  } else {
    // voxels claimed by the previous rounds: once max_voxels exist the frame is closed
    if (wv == 0) {
      int c = 0;
      for (int r = lane; r < round; r += 32) c += __ldcg(w.round_claims + b * kMaxRounds + r);
      for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
      if (lane == 0) s_prev = c;
    }
    __syncthreads();
    if (s_prev >= w.max_voxels) return;
  }
  if (MODE == 1 && Src::kIsDepth && tid < kMaxCams) s_cull[tid] = w.cull ? __ldg(w.cull + b * kMaxCams + tid) : 0xFFFFFFFFu;
  if (MODE == 0) {
    for (int i = tid; i < kPassWarps * kLocalIters * 4; i += kPassThreads) s_lflagb[i] = 0u;
    if (w.bev)
      for (int i = tid; i < kBevWords; i += kPassThreads) s_lbev[i] = 0u;
  }
  const bool cal_async = src.stage_async(s_cal, &s_bar, b);
  __syncthreads();
  if (cal_async) tma_wait(&s_bar);

  typename Src::Walker wk;
  // (a warp whose strip lies beyond the range stays for the block barrier at the end of the insert pass)
  const bool active = src.walk_init(wk, begin, end, iters, bx, wv, lane);
  if (MODE == 1 && !active) return;
  uint2 *s_item = s_itemb + wv * kListCap;
  uint32_t *s_und = s_undb + wv * kListCap;
  uint2 *s_hit = s_hitb + (MODE ? wv * 64 : 0);
  unsigned long long *table = w.table + (int64_t)b * w.cap;
  uint32_t *flags = w.flags + (int64_t)b * w.nwords;
  uint32_t *slots = w.slots + (int64_t)b * w.max_voxels * (w.K - 1);
  uint32_t *later = w.later + (int64_t)b * w.lwords;
  uint32_t *s_lflag = s_lflagb + (MODE == 0 ? wv * kLocalIters * 4 : 0);
  // insert with a short strip: list entries hold the POSITION inside the strip (tile number * 128 + offset), which
  // addresses the warp's local flag words directly; the element index is rebuilt from it when the table needs it
  const bool local = MODE == 0 && iters <= kLocalIters;
  const typename Src::Walker wk0 = wk;
  uint32_t tno = 0;                                               // tiles walked so far
  int cnt = 0, nu = 0, nh = 0, claims = 0;

  bool live = active && (MODE == 0 || !Src::kIsDepth || src.walk_live(wk, s_cull));
  typename Src::Pre pre = src.preload(b, wk, live);
  bool flushing = false;
  // One loop, one copy of each stage (the kernel has to stay inside the instruction cache): hit passes while 32
  // table hits are waiting, item passes while 32 items are waiting, exact passes while 32 undecided elements are
  // waiting, else the next tile; at the end of the walk the lists are flushed with partial passes.  All lists are
  // consumed from their END.
#pragma unroll 1
  while (active) {
    if (MODE == 1 && (nh >= 32 || (flushing && nu == 0 && cnt == 0 && nh > 0))) {
      // ---- hit pass (dense lanes): first point -> rank (#first points before it) -> sorted insertion into the
      //      voxel's slot row.  Voxels of rank >= max_voxels are the ones the reference drops.
      const int n = nh < 32 ? nh : 32;
      if (lane < n) {
        const uint2 h = s_hit[nh - n + lane];
        // a voxel's first point is not kept in the slot rows (the post kernel lists the first points by rank and
        // writes their point -> voxel entries); one only gets this far through the exact path
        if (h.x != h.y) {
          const int r = voxel_rank(w, b, h.x);
          if (r < w.max_voxels) {
            if (w.K > 1) slot_insert(slots + (int64_t)r * (w.K - 1), w.K - 1, h.y, later, r);
            if (point2voxel) point2voxel[(int64_t)b * w.N + h.y] = r;
          }
        }
      }
      nh -= n;
      __syncwarp();
      continue;
    }
    if (cnt >= 32 || (flushing && nu == 0 && cnt > 0)) {
      // ---- item pass: MODE 0 table insert; MODE 1 table lookup, the hits join the hit list ----
      // (measured and dropped: a FIFO with 32 / 64 items of slack whose table sectors are prefetched into L2 when
      // they join the queue -- insert +3 %, lookup +3 %)
      const int n = cnt < 32 ? cnt : 32;
      uint2 it = make_uint2(0u, 0u);
      if (lane < n) it = s_item[cnt - n + lane];
      if (MODE == 0) {
        // (measured: electing one lane per key with __match_any_sync + __reduce_min_sync costs more than the
        // repeated atomics it saves, in the sparse and in the dense scene)
        if (lane < n) {
          const uint32_t idx = local ? src.strip_index(wk0, it.y) : it.y;
          const int r = table_insert(table, flags, w, it.x, idx);
          if (r) {                                               // this point is (for now) the first of its voxel
            if (local) atomicXor(s_lflag + (it.y >> 5), 1u << (it.y & 31));
            else first_toggle(flags, idx);
          }
          if (r == 1) {
            ++claims;
            if (w.bev) {
              const uint32_t bb = bev_bit(w, it.x);
              atomicOr(s_lbev + (bb >> 5), 1u << (bb & 31));
            }
          }
        }
      } else {
        uint32_t first = kEmpty32;
        if (lane < n) first = table_first(table, w, it.x);
        const unsigned hb = __ballot_sync(0xffffffffu, first != kEmpty32);
        if (first != kEmpty32) s_hit[nh + __popc(hb & lt)] = make_uint2(first, it.y);
        nh += __popc(hb);
      }
      cnt -= n;
      __syncwarp();
      continue;
    }
    if (nu >= 32 || (flushing && nu > 0)) {
      // ---- exact pass: IEEE arithmetic of the reference for up to 32 undecided elements; hits join the item list ----
      const int n = nu < 32 ? nu : 32;
      bool in = false;
      uint32_t idx = 0;
      int cx, cy, cz;
      if (lane < n) {
        idx = s_und[nu - n + lane];
        in = src.cell_exact(b, local ? src.strip_index(wk0, idx) : idx, s_cal, g, cx, cy, cz);
      }
      const unsigned b1 = __ballot_sync(0xffffffffu, in);
      if (in) s_item[cnt + __popc(b1 & lt)] = make_uint2(voxel_key(cx, cy, cz, g), idx);
      cnt += __popc(b1);
      nu -= n;
      __syncwarp();
      continue;
    }
    if (!src.walk_more(wk)) {
      if (flushing) break;
      flushing = true;
      continue;
    }
    // ---- next tile ----
    const typename Src::Walker cwk = wk;
    const typename Src::Pre cpre = pre;
    const bool clive = live;
    src.walk_next(wk);
    if (src.walk_more(wk)) {                                // the tile after this one: liveness, depth load
      live = MODE == 0 || !Src::kIsDepth || src.walk_live(wk, s_cull);
      pre = src.preload(b, wk, live);
    }
    if (!clive) continue;                                   // no pixel of this tile can reach a kept voxel
    const uint32_t l0 = src.walk_index(cwk);
    const uint32_t e0 = local ? (tno << 7) + 4u * (uint32_t)lane : l0;    // what the lists hold for the lane's first element
    ++tno;
    typename Src::Cursor cur = src.cursor(b, cwk, s_cal, cpre);
    Quad qd;
    src.classify(cur, s_cal, g, qd);
    unsigned in = qd.in;
    if (MODE == 1) {
      // A voxel's first point is already listed by rank and mapped (post kernel): the lookup has nothing to add for
      // it.  Its flag bit says so without a table probe -- in the region the voxels were claimed from that is nearly
      // every in-range pixel.
      const uint32_t w0 = l0 >> 5, sh = l0 & 31u;
      uint32_t fb = __ldg(flags + w0) >> sh;
      if (sh > 28u && (int)w0 + 1 < w.nwords) fb |= __ldg(flags + w0 + 1) << (32u - sh);
      in &= ~fb;
    }
    // exclusive prefix of the per-lane counts (0..4) from three ballots of the count's bit planes
    const unsigned cin = __popc(in);
    const unsigned p0 = __ballot_sync(0xffffffffu, cin & 1u), p1 = __ballot_sync(0xffffffffu, cin & 2u),
                   p2 = __ballot_sync(0xffffffffu, cin & 4u);
    uint2 *it = s_item + cnt + (__popc(p0 & lt) + 2 * __popc(p1 & lt) + 4 * __popc(p2 & lt));
    // the order inside the list is irrelevant: element q of the lane goes to the lane's slot #(set bits below q)
    if (in & 1u) it[0] = make_uint2(qd.key[0], e0);
    if (in & 2u) it[in & 1u] = make_uint2(qd.key[1], e0 + 1);
    if (in & 4u) it[__popc(in & 3u)] = make_uint2(qd.key[2], e0 + 2);
    if (in & 8u) it[__popc(in & 7u)] = make_uint2(qd.key[3], e0 + 3);
    cnt += __popc(p0) + 2 * __popc(p1) + 4 * __popc(p2);
    if (__any_sync(0xffffffffu, qd.und != 0u)) {
      const unsigned cun = __popc(qd.und);
      const unsigned u0 = __ballot_sync(0xffffffffu, cun & 1u), u1 = __ballot_sync(0xffffffffu, cun & 2u),
                     u2 = __ballot_sync(0xffffffffu, cun & 4u);
      int ua = nu + __popc(u0 & lt) + 2 * __popc(u1 & lt) + 4 * __popc(u2 & lt);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if ((qd.und >> q) & 1u) s_und[ua++] = e0 + q;
      nu += __popc(u0) + 2 * __popc(u1) + 4 * __popc(u2);
    }
    __syncwarp();
  }
  if (MODE == 0) {
    __syncwarp();
    if (local) {
      // local flag word j = 32 consecutive elements of the strip starting at position 32 j
      for (uint32_t j = lane; j < tno * 4u; j += 32) {
        const uint32_t bits = s_lflag[j];
        if (bits) {
          const uint32_t pos = src.strip_index(wk0, j << 5);
          const uint32_t sh = pos & 31u;
          atomicXor(flags + (pos >> 5), bits << sh);
          if (sh && (bits >> (32u - sh))) atomicXor(flags + (pos >> 5) + 1, bits >> (32u - sh));
        }
      }
    }
    for (int d = 16; d > 0; d >>= 1) claims += __shfl_xor_sync(0xffffffffu, claims, d);
    if (lane == 0 && claims) atomicAdd(w.round_claims + b * kMaxRounds + round, claims);
    if (w.bev) {
      __syncthreads();                                           // every warp of the CTA has made its marks
      uint32_t *gb = w.bev + ((int64_t)b * kBevCopies + (bx & (kBevCopies - 1))) * kBevWords;
      for (int j = tid; j < kBevWords; j += kPassThreads) {
        const uint32_t bits = s_lbev[j];
        if (bits) atomicOr(gb + j, bits);
      }
    }
  }
}

// P0 ------------------------------------------------------------------------
// The small scratch arrays of a sub-batch are initialised by ONE launch (a memset each would be a launch each): up
// to kInitRegions word-aligned regions, each filled with its own 32-bit pattern, 16 bytes per store in the body.
constexpr int kInitRegions = 8;
struct HvInit {
  uint32_t *ptr[kInitRegions];
  unsigned long long words[kInitRegions];
  uint32_t val[kInitRegions];
  int n;
};
static __global__ void __launch_bounds__(256) hv_init_kernel(HvInit in) {
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
  for (int r = 0; r < in.n; ++r) {
    uint32_t *p = in.ptr[r];
    const unsigned long long n = in.words[r];
    const uint32_t v = in.val[r];
    unsigned long long head = ((16u - (unsigned)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u) >> 2;   // words up to 16-byte alignment
    if (head > n) head = n;
    const unsigned long long body = (n - head) >> 2;             // uint4 stores
    uint4 *p4 = reinterpret_cast<uint4 *>(p + head);
    const uint4 v4 = make_uint4(v, v, v, v);
    for (unsigned long long i = tid; i < body; i += nthreads) p4[i] = v4;
    const unsigned long long done = head + (body << 2);
    if (tid < head) p[tid] = v;
    if (tid < n - done) p[done + tid] = v;
  }
}

// P2a -----------------------------------------------------------------------
// First-point flags -> chunk totals -> (by the frame's LAST CTA to finish: no second launch, no spinning)
// exclusive chunk bases and voxel_num.  grid (ceil(nchunks / 4), frames), 256 threads: a CTA owns 4 chunks of
// kChunkWords flag words, 64 threads per chunk, one 16-byte load per thread.
static __global__ void __launch_bounds__(256) hv_count_kernel(HvWork w, int32_t *voxel_num) {
  __shared__ int s_part[8];
  __shared__ int s_last, s_carry;
  __shared__ int s_warp[8];
  const int b = blockIdx.y + w.b0;
  const int t = threadIdx.x, lane = t & 31, wv = t >> 5;
  const int c = blockIdx.x * 4 + (t >> 6);
  int cnt = 0;
  if (c < w.nchunks) {
    const uint4 f = __ldcs(reinterpret_cast<const uint4 *>(w.flags + (int64_t)b * w.nwords + (int64_t)c * kChunkWords) + (t & 63));
    cnt = __popc(f.x) + __popc(f.y) + __popc(f.z) + __popc(f.w);
  }
  for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if (lane == 0) s_part[wv] = cnt;
  __syncthreads();
  int32_t *cb = w.chunk_base + (int64_t)b * w.nchunks;
  if ((t & 63) == 0 && c < w.nchunks) cb[c] = s_part[wv] + s_part[wv + 1];
  __threadfence();                                   // the totals are visible before the ticket is taken
  __syncthreads();
  if (t == 0) s_last = atomicAdd(w.scan_done + b, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the last CTA of the frame: exclusive scan of the chunk totals in place
  if (t == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < w.nchunks; base += 256) {
    const int i = base + t;
    const int v = (i < w.nchunks) ? __ldcg(cb + i) : 0;
    const int inc = warp_inclusive_scan(v);
    if (lane == 31) s_warp[wv] = inc;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) before += (k < wv) ? s_warp[k] : 0;
    const int excl = s_carry + before + inc - v;
    if (i < w.nchunks) cb[i] = excl;
    __syncthreads();
    if (t == 255) s_carry = excl + v;
    __syncthreads();
  }
  if (t == 0) voxel_num[b] = s_carry < w.max_voxels ? s_carry : w.max_voxels;
  if (w.bev) {
    // The privatised bird's-eye masks of the frame are folded once, here, and the marked cells are LISTED: the cull
    // CTAs then share the cells evenly among their threads (a thread that owned mask words walked up to 64 cells of a
    // dense word while its neighbours had none).  Thread t folds the words [t * per, t * per + per).
    static_assert(kBevWords % 256 == 0, "mask words per thread");
    constexpr int per = kBevWords / 256;
    const uint32_t *m = w.bev + (int64_t)b * kBevCopies * kBevWords;
    uint32_t v[per];
    int c = 0;
#pragma unroll
    for (int j = 0; j < per; ++j) {
      uint32_t x = 0;
#pragma unroll 8
      for (int k = 0; k < kBevCopies; ++k) x |= __ldcg(m + k * kBevWords + t * per + j);
      v[j] = x;
      c += __popc(x);
    }
    const int inc = warp_inclusive_scan(c);
    __syncthreads();                                   // s_warp is free again
    if (lane == 31) s_warp[wv] = inc;
    __syncthreads();
    int pos = inc - c;
#pragma unroll
    for (int k = 0; k < 8; ++k) pos += (k < wv) ? s_warp[k] : 0;
    uint16_t *list = w.bev_list + (int64_t)b * kBevDim * kBevDim;
    if (t == 255) w.bev_count[b] = pos + c;
    // expansion one word at a time with the WARP (lane L owns bit L: consecutive list positions, coalesced stores);
    // empty words are skipped by vote
    const unsigned ltm = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < per; ++j) {
      unsigned nz = __ballot_sync(0xffffffffu, v[j] != 0u);
      while (nz) {
        const int src = __ffs(nz) - 1;
        nz &= nz - 1;
        const uint32_t W = __shfl_sync(0xffffffffu, v[j], src);
        const int P = __shfl_sync(0xffffffffu, pos, src);
        const int T = __shfl_sync(0xffffffffu, t * per + j, src);
        if ((W >> lane) & 1u) list[P + __popc(W & ltm)] = (uint16_t)(T * 32 + lane);
      }
      pos += __popc(v[j]);
    }
  }
}

// P2f -----------------------------------------------------------------------
// After the chunk bases are known, for 4 chunks per CTA (64 threads per chunk, 4 flag words per thread): the
// exclusive popcount prefix of every word inside its chunk (what voxel_rank adds to the chunk base), and the list
// of first points by rank -- the r-th set bit of the flags is the first point of the voxel of rank r.
template <class Src>
__device__ __forceinline__ void firsts_block(const Src &psrc, const HvWork &w, int b, int blk) {
  __shared__ int s_fw[8];
  const int t = threadIdx.x, lane = t & 31, wv = t >> 5;
  const int c = blk * 4 + (t >> 6);
  const bool on = c < w.nchunks;
  const int64_t w0 = (int64_t)b * w.nwords + (int64_t)(on ? c : 0) * kChunkWords + (t & 63) * 4;
  uint4 f = make_uint4(0u, 0u, 0u, 0u);
  if (on) f = __ldg(reinterpret_cast<const uint4 *>(w.flags + w0));
  const int c0 = __popc(f.x), c1 = __popc(f.y), c2 = __popc(f.z), c3 = __popc(f.w);
  const int mine = c0 + c1 + c2 + c3;
  const int inc = warp_inclusive_scan(mine);
  if (lane == 31) s_fw[wv] = inc;
  __syncthreads();
  const int p0 = inc - mine + ((wv & 1) ? s_fw[wv - 1] : 0);           // two warps per chunk
  if (on) *reinterpret_cast<int4 *>(w.wordprefix + w0) = make_int4(p0, p0 + c0, p0 + c0 + c1, p0 + c0 + c1 + c2);
  // Expansion of the set bits, one flag word at a time with the WARP: lane L owns bit L, so the ranks of a word's
  // first points are consecutive over the lanes and the stores coalesce (a thread walking its own words would
  // scatter 4-byte stores ~50 ranks apart -- measured 44 us for this step alone).  Empty words are skipped by vote.
  const int rbase = on ? __ldg(w.chunk_base + (int64_t)b * w.nchunks + c) + p0 : 0;
  uint32_t *out = w.first_of + (int64_t)b * w.max_voxels;
  const uint32_t wl = (uint32_t)((on ? c : 0) * kChunkWords + (t & 63) * 4);
  const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
  const int rq[4] = {rbase, rbase + c0, rbase + c0 + c1, rbase + c0 + c1 + c2};
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    unsigned nz = __ballot_sync(0xffffffffu, fw[q] != 0u);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      const uint32_t W = __shfl_sync(0xffffffffu, fw[q], src);
      const int R = __shfl_sync(0xffffffffu, rq[q], src);
      const uint32_t I = __shfl_sync(0xffffffffu, wl + q, src);
      if ((W >> lane) & 1u) {
        const int pos = R + __popc(W & lt);
        if (pos < w.max_voxels) {
          const uint32_t idx = (I << 5) + (uint32_t)lane;
          out[pos] = idx;
          if (w.p2v) w.p2v[(int64_t)b * w.N + idx] = pos;          // point -> voxel map of the first points
        }
      }
    }
  }
}

// P2s -----------------------------------------------------------------------
// one CTA per frame: exclusive scan of the chunk totals in place; the frame's
// total (clamped) goes to out_total[b].
static __global__ void __launch_bounds__(1024) scan_chunks_kernel(int32_t *chunk_base, int nchunks,
                                                                  int32_t *out_total, int clamp,
                                                                  int b0 = 0) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int b = blockIdx.x + b0;
  int32_t *cb = chunk_base + (int64_t)b * nchunks;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nchunks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < nchunks) ? cb[i] : 0;
    const int inc = warp_inclusive_scan(v);
    if (lane == 31) s_warp[wv] = inc;
    __syncthreads();
    if (wv == 0) {
      const int t = s_warp[lane];
      const int ti = warp_inclusive_scan(t);
      s_warp[lane] = ti - t;
    }
    __syncthreads();
    const int carry = s_carry;
    const int excl = carry + s_warp[wv] + inc - v;
    if (i < nchunks) cb[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int total = s_carry;
    out_total[b] = total < clamp ? total : clamp;
  }
}

// Plane k (0..4) of the viewing wedge of (camera, column block) for the cull test below, from the camera's inverse
// cell map: rows  w3 | w1 - u0 w3 | u1 w3 - w1 | w2 | (H-1) w3 - w2.  out[0..3] = normal x, y | constant with the
// half-extents of a bird's-eye cell folded in | slack.  Worked out in fp64 once per frame (calib_kernel); the
// per-cell test runs in fp32 (this part's fp64 rate made it the bulk of the post kernel), so the slack carries a
// bound of the fp32 evaluation error -- conversions of the three constants, two FMAs, the cell centre (five
// roundings of 2^-24 of the largest magnitude, taken as 6e-7, plus 1e-3 cells of centre error): it only ever culls less.
// A bird's-eye cell j on an axis covers the voxels i with (i * kx) >> 20 == j, i.e. the cell coordinates
// [j 2^20 / kx, (j + 1) 2^20 / kx + 1).
__device__ __forceinline__ void cull_plane_of(const double *iv, const double *T, double margin, int k, int blk, int cbshift,
                                              int W, int H, const VoxelGrid &g, float *out) {
  const double sx = 1048576.0 / (double)(uint32_t)(((uint64_t)kBevDim << 20) / (uint64_t)g.grid[0]);
  const double sy = 1048576.0 / (double)(uint32_t)(((uint64_t)kBevDim << 20) / (uint64_t)g.grid[1]);
  const double gz = g.grid[2];
  const double u0 = (double)(blk << cbshift);
  double u1 = (double)(((blk + 1) << cbshift) - 1);
  if (u1 > (double)(W - 1)) u1 = (double)(W - 1);
  const double hm1 = (double)(H - 1);
  double n[3];
  for (int a = 0; a < 3; ++a) {
    n[a] = k == 0 ? iv[6 + a]
         : k == 1 ? iv[a] - u0 * iv[6 + a]
         : k == 2 ? u1 * iv[6 + a] - iv[a]
         : k == 3 ? iv[3 + a]
                  : hm1 * iv[6 + a] - iv[3 + a];
  }
  const double hx = 0.5 * (sx + 1.0) + margin, hy = 0.5 * (sy + 1.0) + margin, hz = 0.5 * gz + margin;
  // z part (the box spans the whole height), -n.T and the half-extent terms are the same for every cell
  const double cst = n[2] * (0.5 * gz - T[2]) - n[0] * T[0] - n[1] * T[1] + fabs(n[0]) * hx + fabs(n[1]) * hy + fabs(n[2]) * hz;
  const double slack = 1e-9 * (fabs(n[0]) * ((double)g.grid[0] + fabs(T[0]) + hx) + fabs(n[1]) * ((double)g.grid[1] + fabs(T[1]) + hy) +
                               fabs(n[2]) * (gz + fabs(T[2]) + hz));
  const double mag = fabs(n[0]) * ((double)g.grid[0] + sx + 1.0) + fabs(n[1]) * ((double)g.grid[1] + sy + 1.0) + fabs(cst);
  out[0] = (float)n[0];
  out[1] = (float)n[1];
  out[2] = (float)cst;
  out[3] = (float)((slack + 6e-7 * mag + 1e-3 * (fabs(n[0]) + fabs(n[1]))) * 1.0001);
}

// P2c -----------------------------------------------------------------------
// grid (cameras x column blocks, frames), one thread per word of the bird's-eye mask.  For one camera and one
// block of 2^cbshift image columns: is there a claimed voxel that a pixel of the block can fall into?  A pixel (u, v, z > 0) lands at the continuous cell coordinate
//   P = z * Mc (u, v, 1)^T + T,   Mc = [A B C] of the direct cell map (rd3_common.cuh), T = Th + 0.5,
// and its reference cell is within tol(z) <= tolmax of floor(P) (the proven bound of that map), so with
// w = Mc^-1 (P - T) = (z u, z v, z) the block's pixels fill the wedge
//   w3 >= 0,  w1 - u0 w3 >= 0,  u1 w3 - w1 >= 0,  w2 >= 0,  (H-1) w3 - w2 >= 0.
// A bird's-eye cell (a box over all z, widened by 1 + 2 tolmax cells) that lies entirely on the negative side of
// one of these five planes cannot receive a pixel of the block; the largest value of a plane function over a
// box is its value at the centre plus |n| . half-extents.  Everything is evaluated in fp64; a camera whose
// map is singular / non-finite keeps all its blocks.
__device__ __forceinline__ void cull_block(const DepthSource &src, const VoxelGrid &g, const HvWork &w, int b, int pair) {
  __shared__ float s_plf[5][4];          // per plane: normal x, y | constant | slack (calib_kernel: cull_planes_of)
  __shared__ int s_ok, s_hitflag;
  const int nblk = ((src.p.W - 1) >> src.cbshift) + 1;
  const int cam = pair / nblk, blk = pair - cam * nblk;
  const int ncam = src.p.ncam;
  if (threadIdx.x < 20)
    (&s_plf[0][0])[threadIdx.x] = __ldg(src.cull_planes + (((int64_t)b * ncam + cam) * nblk + blk) * 20 + threadIdx.x);
  if (threadIdx.x == 32) {
    s_ok = src.cull_cal[((int64_t)b * ncam + cam) * kCullDoubles + 13] != 0.0 ? 1 : 0;
    s_hitflag = 0;
  }
  __syncthreads();
  const float sxf = 1048576.0f / (float)w.bev_kx, syf = 1048576.0f / (float)w.bev_ky;
  bool hit = false;
  if (!s_ok) {
    hit = true;
  } else {
    // the frame's marked cells (listed by the count kernel), shared evenly among the threads
    const int ncell = __ldg(w.bev_count + b);
    const uint16_t *list = w.bev_list + (int64_t)b * kBevDim * kBevDim;
    for (int i = threadIdx.x; i < ncell && !hit; i += 256) {
      const int cell = __ldg(list + i);
      const float cxc = fmaf((float)(cell % kBevDim) + 0.5f, sxf, 0.5f), cyc = fmaf((float)(cell / kBevDim) + 0.5f, syf, 0.5f);
      bool out = false;
#pragma unroll
      for (int k = 0; k < 5; ++k) out = out || (fmaf(s_plf[k][0], cxc, fmaf(s_plf[k][1], cyc, s_plf[k][2])) < -s_plf[k][3]);
      hit = !out;
    }
  }
  if (hit) s_hitflag = 1;
  __syncthreads();
  if (threadIdx.x == 0 && s_hitflag) atomicOr(w.cull + b * kMaxCams + cam, 1u << blk);
}

// P2f + P2c in ONE launch (both only need the chunk bases): grid (ncull + ceil(nchunks / 4), frames), 256 threads;
// the first ncull CTAs of a frame (depth source with culling: cameras x column blocks, else 0) decide a cull bit each.
// (compiled for 8 CTAs per SM: the cull part's fp64 plane set would otherwise cap the whole launch at 3)
template <class Src>
__global__ void __launch_bounds__(256, 8) hv_post_kernel(Src src, VoxelGrid g, HvWork w, int ncull) {
  const int b = blockIdx.y + w.b0;
  if ((int)blockIdx.x < ncull) {
    if constexpr (Src::kIsDepth) cull_block(src, g, w, b, (int)blockIdx.x);
    return;
  }
  firsts_block(src, w, b, (int)blockIdx.x - ncull);
}

// P4 ------------------------------------------------------------------------
// A warp owns `vpw` consecutive voxels (ranks), vpw = 32 unless the rows are very long (then a power of two below);
// lane groups of 32 / vpw lanes share a voxel.  No block barrier after the calibration copy.
// dynamic smem per warp: tile[vpw*K*C] floats (the voxels' rows exactly as they lie in the output) + the list of the
// later points (kEmitList entries: idx u32, position u16 -- vpw * K <= 10 K by the shared-memory limit).
//   1. the tile is zeroed (16-byte stores); the voxel's first point is requested; the voxels whose `later` bit is set
//      walk their slot rows (sorted, non-empty prefix; the lanes of a group take consecutive words) and ballot-compact
//      (point, position) into the warp's list -- a voxel without later points never touches its row
//   2. first points (one per voxel) and listed points (dense over the lanes) are gathered / re-unprojected with the
//      reference's exact arithmetic into the tile
//   3. the tile is copied out with 16-byte stores (the warp's rows are one contiguous block of the output); one lane
//      per voxel writes coors (from the first point), the count and the HardSimpleVFE mean (slot order, one __fdiv_rn)
template <class Src>
__global__ void __launch_bounds__(kEmitThreads, RD3_EMIT_MINB) hv_emit_kernel(Src src, VoxelGrid g, HvWork w, HvOut o, int wbytes,
                                                                              int lg_vpw) {
  extern __shared__ __align__(16) unsigned char s_dynb[];
  __shared__ __align__(16) float s_cal[Src::kIsDepth ? kMaxCams * kCalibFloats : 4];
  __shared__ __align__(8) uint64_t s_bar;
  const int b = blockIdx.y + w.b0;
  const int vn = o.voxel_num[b];
  const int vpw = 1 << lg_vpw;
  const int nwarps = (int)blockDim.x >> 5;
  const int r0c = blockIdx.x * nwarps * vpw;
  if (r0c >= vn) return;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const bool cal_async = src.stage_async(s_cal, &s_bar, b);     // one TMA bulk copy, no load / store loop
  const int r0 = r0c + wv * vpw;                                 // a multiple of vpw: the warp's voxels share one `later` word
  const int nv = vn - r0 < vpw ? vn - r0 : vpw;
  const int lg_lpv = 5 - lg_vpw;                                 // lanes per voxel
  const int vl = lane >> lg_lpv, sub = lane & ((1 << lg_lpv) - 1);
  const bool live = vl < nv;
  const int v = r0 + vl;
  const int C = src.num_feats();
  const int K = w.K, Km1 = K - 1, KC = K * C;
  float *tile = reinterpret_cast<float *>(s_dynb + (size_t)wv * wbytes);
  uint32_t *l_idx = reinterpret_cast<uint32_t *>(tile + ((vpw * KC + 3) & ~3));
  uint16_t *l_pos = reinterpret_cast<uint16_t *>(l_idx + kEmitList);
  uint32_t first = kEmpty32, lw = 0;
  float firstz = 0.0f;
  if (nv > 0) {
    if (live) {
      first = __ldg(w.first_of + (int64_t)b * w.max_voxels + v);
      if (Src::kCarryFirst) firstz = __ldg(w.first_z + (int64_t)b * w.max_voxels + v);
    }
    lw = __ldg(w.later + (int64_t)b * w.lwords + (r0 >> 5)) >> (r0 & 31);
    float4 *t4 = reinterpret_cast<float4 *>(tile);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n4 = (vpw * KC + 3) >> 2;
    for (int e = lane; e < n4; e += 32) t4[e] = z4;
  }
  __syncthreads();                       // calibration copy issued (and the tile zeroed, as far as this warp goes)
  if (cal_async) tma_wait(&s_bar);
  if (nv <= 0) return;

  // slot rows of the voxels that have later points: the group's lanes take the words k0 + sub, three steps at a time
  // (a row is one or two sectors: its words cost one round trip, not one each).  The list holds kEmitList entries:
  // it is worked off whenever another round of words might not fit.
  int cnt = 0, n = 0;
  {
    bool open = live && ((lw >> vl) & 1u);
    const uint32_t *row = w.slots + ((int64_t)b * w.max_voxels + v) * Km1;
    const int lpv = 1 << lg_lpv;
    for (int k0 = 0; k0 < Km1; k0 += 3 * lpv) {
      if (!__any_sync(0xffffffffu, open)) break;
      uint32_t wd[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int k = k0 + j * lpv + sub;
        wd[j] = (open && k < Km1) ? __ldg(row + k) : kEmpty32;
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int k = k0 + j * lpv + sub;
        const bool has = wd[j] != kEmpty32;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (n + __popc(bal) > kEmitList) {                       // warp-uniform
          __syncwarp();
          for (int t = lane; t < n; t += 32) src.gather(b, l_idx[t], s_cal, tile + (size_t)l_pos[t] * C);
          __syncwarp();
          n = 0;
        }
        if (has) {
          const int q = n + __popc(bal & lt);
          l_idx[q] = wd[j];
          l_pos[q] = (uint16_t)(vl * K + k + 1);
          ++cnt;
        }
        n += __popc(bal);
      }
      // the row is a non-empty prefix: the group goes on only while its last word was in use
      open = __shfl_sync(0xffffffffu, wd[2] != kEmpty32, (vl << lg_lpv) + lpv - 1);
    }
    for (int d = 1; d < lpv; d <<= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    cnt += 1;                                                    // the first point
  }
  __syncwarp();
  if (live && sub == 0) src.gather_first(b, first, firstz, s_cal, tile + vl * KC);
  for (int t = lane; t < n; t += 32) src.gather(b, l_idx[t], s_cal, tile + (size_t)l_pos[t] * C);
  __syncwarp();

  // voxels: the warp's nv rows are nv*K*C contiguous floats (skipped when the caller only wants coors / num / mean)
  if (o.voxels) {
    float *vout = o.voxels + ((int64_t)b * w.max_voxels + r0) * KC;
    const int nfl = nv * KC;
    if ((reinterpret_cast<uintptr_t>(vout) & 15) == 0) {
      const float4 *t4 = reinterpret_cast<const float4 *>(tile);
      float4 *v4 = reinterpret_cast<float4 *>(vout);
      const int n4 = nfl >> 2;
      for (int e = lane; e < n4; e += 32) v4[e] = t4[e];
      for (int e = (nfl & ~3) + lane; e < nfl; e += 32) vout[e] = tile[e];
    } else {
      for (int e = lane; e < nfl; e += 32) vout[e] = tile[e];
    }
  }

  // one lane per voxel: coors, count and HardSimpleVFE mean
  // (voxel_encoder.py:45-46: sum over ALL K slots in slot order, then one division; the
  // slots beyond the count are zeros, so the running sum stops changing at the count --
  // except that (-0.0) + 0.0 = +0.0, which one extra "+ 0.0f" reproduces)
  if (live && sub == 0) {
    const float *p0 = tile + vl * KC;
    const int64_t vr = (int64_t)b * w.max_voxels + v;
    // coors from the voxel's first point (the tile holds it already): the fast path decides all but points within
    // rounding of a cell boundary
    int cx = 0, cy = 0, cz = 0;
    if (voxel_coor_fast(p0[0], p0[1], p0[2], 0.0f, g, cx, cy, cz) == 2) voxel_coor(p0[0], p0[1], p0[2], g, cx, cy, cz);
    o.coors[vr * 3 + 0] = cz;
    o.coors[vr * 3 + 1] = cy;
    o.coors[vr * 3 + 2] = cx;
    o.num[vr] = cnt;
    if (o.mean) {
      const int F = o.F;
      const float nf = (float)cnt;
      float *mo = o.mean + vr * F;
      if (F == 3) {
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        const float *p = p0;
        for (int k = 0; k < cnt; ++k, p += C) {
          sx = __fadd_rn(sx, p[0]);
          sy = __fadd_rn(sy, p[1]);
          sz = __fadd_rn(sz, p[2]);
        }
        if (cnt < K) { sx = __fadd_rn(sx, 0.0f); sy = __fadd_rn(sy, 0.0f); sz = __fadd_rn(sz, 0.0f); }
        mo[0] = __fdiv_rn(sx, nf);
        mo[1] = __fdiv_rn(sy, nf);
        mo[2] = __fdiv_rn(sz, nf);
      } else {
        for (int f = 0; f < F; ++f) {
          float sacc = 0.0f;
          const float *p = p0 + f;
          for (int k = 0; k < cnt; ++k, p += C) sacc = __fadd_rn(sacc, *p);
          if (cnt < K) sacc = __fadd_rn(sacc, 0.0f);
          mo[f] = __fdiv_rn(sacc, nf);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// host side: workspace carving + launch sequence
// ---------------------------------------------------------------------------
struct HvTuning {           // read once from the environment (experiments); defaults are the measured best
  int rounds;               // target number of insert rounds (RD3_ROUNDS)
  int load_pct;             // worst-case load factor of the table in percent (RD3_TABLE_LOAD_PCT)
  int ins_iters, lkp_iters; // tiles per warp strip (RD3_INS_ITERS, RD3_LKP_ITERS)
  int cull;                 // RD3_CULL=0 disables the camera / column-block culling
  int sm_count;
};
const HvTuning &hv_tuning();

struct HvPlan {
  int64_t N;
  int B;
  int K;
  int max_voxels;
  int64_t S;          // points per insert round (multiple of 1024)
  int rounds;
  int64_t cap;
  int log2cap;
  int nwords, nchunks;
  // [table | slots] are set to 0xFF with one memset, [flags | bev | round_claims] to 0 with another
  int lwords;
  size_t off_table, off_slots, off_first, off_firstz, off_flags, off_bev, off_claims, off_done, off_cull, off_prefix, off_chunk,
      off_later, off_bevlist, off_bevcount, total;
};

// `round_multiple`: a round has to be a whole number of the source's work units (image rows for depth maps)
inline HvPlan hv_plan(int64_t N, int B, int K, int max_voxels, int64_t round_multiple = 1024) {
  const HvTuning &t = hv_tuning();
  HvPlan p;
  p.N = N; p.B = B; p.K = K; p.max_voxels = max_voxels;
  const int64_t n1 = N > 0 ? N : 1;
  // round length: at most `rounds` rounds, but a round should keep the whole GPU busy
  // (>= ~1.2 M points over all frames), so small batches use fewer, longer rounds
  const int64_t rm = round_multiple > 0 ? round_multiple : 1;
  int64_t S = ceil_div(n1, t.rounds);
  const int64_t fill = ceil_div((int64_t)t.sm_count * 8 * 1024, B > 0 ? B : 1);
  if (S < fill) S = fill;
  if (S < 65536) S = 65536;
  S = ceil_div(S, rm) * rm;
  p.S = S;
  p.rounds = (int)ceil_div(n1, S);
  // the table never holds more than max_voxels + S keys (see header), nor more than N
  int64_t keys = (int64_t)max_voxels + S;
  if (keys > n1) keys = n1;
  int64_t want = keys * 100 / t.load_pct;
  int lg = 10;
  while (((int64_t)1 << lg) < want) ++lg;
  p.log2cap = lg;
  p.cap = (int64_t)1 << lg;
  p.nchunks = (int)ceil_div(n1, kChunkPoints);
  p.nwords = p.nchunks * kChunkWords;
  size_t off = 0;
  p.off_table = off; off += align_up((size_t)B * p.cap * 8);
  p.off_slots = off; off += align_up((size_t)B * max_voxels * (K > 1 ? K - 1 : 1) * 4);
  p.off_first = off; off += align_up((size_t)B * max_voxels * 4);
  p.off_firstz = off; off += align_up((size_t)B * max_voxels * 4);
  p.off_flags = off; off += align_up((size_t)B * p.nwords * 4);
  p.off_bev = off; off += align_up((size_t)B * kBevCopies * kBevWords * 4);
  p.off_claims = off; off += align_up((size_t)B * kMaxRounds * 4);
  p.off_done = off; off += align_up((size_t)B * 4);
  p.off_cull = off; off += align_up((size_t)B * kMaxCams * 4);
  p.off_prefix = off; off += align_up((size_t)B * p.nwords * 4);
  p.off_chunk = off; off += align_up((size_t)B * p.nchunks * 4);
  p.lwords = (int)ceil_div(max_voxels, 32);
  p.off_later = off; off += align_up((size_t)B * p.lwords * 4);
  p.off_bevlist = off; off += align_up((size_t)B * kBevDim * kBevDim * 2);
  p.off_bevcount = off; off += align_up((size_t)B * 4);
  p.total = off;
  return p;
}

// culling needs the precomputed calibration table and more than one column block or camera to pay off
template <class Src> struct CullLaunch {
  static bool wanted(const Src &) { return false; }
  static int blocks(const Src &) { return 0; }
};
template <> struct CullLaunch<DepthSource> {
  static bool wanted(const DepthSource &s) { return hv_tuning().cull && s.cal_table != nullptr && s.cull_cal != nullptr && s.cull_planes != nullptr; }
  static int blocks(const DepthSource &s) { return s.p.ncam * (((s.p.W - 1) >> s.cbshift) + 1); }
};

template <class Src>
int hv_run(const Src &src, const VoxelGrid &g, uint64_t volume, const HvPlan &p, void *ws,
           HvOut out, cudaStream_t stream) {
  const HvTuning &tune = hv_tuning();
  char *base = (char *)ws;
  HvWork w;
  w.table = (unsigned long long *)(base + p.off_table);
  w.slots = (uint32_t *)(base + p.off_slots);
  w.first_of = (uint32_t *)(base + p.off_first);
  w.first_z = (float *)(base + p.off_firstz);
  w.flags = (uint32_t *)(base + p.off_flags);
  w.round_claims = (int32_t *)(base + p.off_claims);
  w.scan_done = (int32_t *)(base + p.off_done);
  w.wordprefix = (int32_t *)(base + p.off_prefix);
  w.chunk_base = (int32_t *)(base + p.off_chunk);
  w.later = (uint32_t *)(base + p.off_later);
  w.lwords = p.lwords;
  w.p2v = out.point2voxel;
  w.vnum = out.voxel_num;
  const bool cull = CullLaunch<Src>::wanted(src) && g.fast_ok;
  w.bev = cull ? (uint32_t *)(base + p.off_bev) : nullptr;
  w.cull = cull ? (uint32_t *)(base + p.off_cull) : nullptr;
  w.bev_list = (uint16_t *)(base + p.off_bevlist);
  w.bev_count = (int32_t *)(base + p.off_bevcount);
  w.N = p.N; w.cap = p.cap; w.cap_mask = (uint32_t)(p.cap - 1); w.log2cap = p.log2cap;
  w.direct = (volume <= (uint64_t)p.cap) ? 1 : 0;
  w.nwords = p.nwords; w.nchunks = p.nchunks; w.K = p.K; w.max_voxels = p.max_voxels;
  w.div_gx = make_fastdiv((uint32_t)g.grid[0]);
  w.div_gy = make_fastdiv((uint32_t)g.grid[1]);
  w.bev_kx = (uint32_t)(((uint64_t)kBevDim << 20) / (uint64_t)g.grid[0]);
  w.bev_ky = (uint32_t)(((uint64_t)kBevDim << 20) / (uint64_t)g.grid[1]);
  w.div_Km1 = make_fastdiv((uint32_t)(p.K > 1 ? p.K - 1 : 1));
  if (p.S % src.host_round_multiple() != 0 && p.rounds > 1) return RD3_ERR_INVALID_ARGUMENT;   // plan made for another source

  const int C = src.host_num_feats();
  // emit: per-warp shared memory = the rows of the warp's vpw voxels + the list of their later points.  vpw = 32 unless
  // that exceeds ~12 KB per warp (long rows: max_points 100+), then the largest power of two that fits; as many warps per
  // CTA (<= kEmitThreads / 32) as fit ~96 KB
  auto emit_wbytes = [&](int vpw) { return align_up((size_t)vpw * p.K * C * 4, 16) + (size_t)kEmitList * 6; };
  int lg_vpw = 5;
  while (lg_vpw > 0 && emit_wbytes(1 << lg_vpw) > 12 * 1024) --lg_vpw;
  const size_t wbytes = align_up(emit_wbytes(1 << lg_vpw), 16);
  if (wbytes > 200 * 1024) return RD3_ERR_UNSUPPORTED;
  int emit_warps = (int)((96 * 1024) / wbytes);
  if (emit_warps < 1) emit_warps = 1;
  if (emit_warps > kEmitThreads / 32) emit_warps = kEmitThreads / 32;
  const size_t smem = wbytes * emit_warps;
  if (smem > 48 * 1024)
    RD3_CUDA_TRY(cudaFuncSetAttribute(hv_emit_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  if (out.point2voxel && p.N > 0)
    RD3_CUDA_TRY(cudaMemsetAsync(out.point2voxel, 0xFF, (size_t)p.B * p.N * 4, stream));

  // Frames are independent: the batch is split into up to kMaxLanes sub-batches that run the
  // whole kernel sequence on their own streams (forked from / joined to the caller's stream),
  // so that the latency-bound small kernels of one sub-batch overlap the passes of another.
  // With the stage profiler on, one lane is used.
  LaneLock lane_lock;              // the lanes' events are shared by all callers on this device
  StreamLanes *lanes = nullptr;
  int nl = 1;
  if (!prof_enabled() && p.B >= 2) {
    nl = stream_lane_count();
    if (nl > p.B) nl = p.B;
    if (nl > 1) {
      lanes = get_stream_lanes();
      if (!lanes) nl = 1;
    }
  }
  int status = RD3_OK;
  int forked = 0;                  // side lanes that wait on the fork event and must be joined, also on an error
#define RD3_LANE_TRY(expr)                                                            \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) { rd3::set_last_cuda_error(_e); status = RD3_ERR_CUDA; }   \
  } while (0)
  if (nl > 1) RD3_LANE_TRY(cudaEventRecord(lanes->fork, stream));
  if (status != RD3_OK) return status;
  for (int l = 0; l < nl && status == RD3_OK; ++l) {
    const int b0 = (int)((int64_t)p.B * l / nl), b1 = (int)((int64_t)p.B * (l + 1) / nl);
    const int nb = b1 - b0;
    cudaStream_t st = stream;
    if (l > 0) {
      st = lanes->s[l - 1];
      RD3_LANE_TRY(cudaStreamWaitEvent(st, lanes->fork, 0));
      if (status != RD3_OK) break;
      forked = l;
    }
    w.b0 = b0;
    prof_mark(st, 0);
    {
      HvInit in;
      in.n = 0;
      auto add = [&](void *ptr, size_t bytes, uint32_t val) {
        in.ptr[in.n] = (uint32_t *)ptr; in.words[in.n] = bytes / 4; in.val[in.n] = val; ++in.n;
      };
      // the big regions with the driver's memset (measured faster than a fill kernel), the small ones in one launch
      RD3_LANE_TRY(cudaMemsetAsync(w.table + (size_t)b0 * p.cap, 0xFF, (size_t)nb * p.cap * 8, st));
      if (p.K > 1)
        RD3_LANE_TRY(cudaMemsetAsync(w.slots + (size_t)b0 * p.max_voxels * (p.K - 1), 0xFF,
                                     (size_t)nb * p.max_voxels * (p.K - 1) * 4, st));
      RD3_LANE_TRY(cudaMemsetAsync(w.flags + (size_t)b0 * p.nwords, 0, (size_t)nb * p.nwords * 4, st));
      add(w.later + (size_t)b0 * p.lwords, (size_t)nb * p.lwords * 4, 0u);
      add(w.round_claims + (size_t)b0 * kMaxRounds, (size_t)nb * kMaxRounds * 4, 0u);
      add(w.scan_done + b0, (size_t)nb * 4, 0u);
      if (cull) {
        add(w.bev + (size_t)b0 * kBevCopies * kBevWords, (size_t)nb * kBevCopies * kBevWords * 4, 0u);
        add(w.cull + (size_t)b0 * kMaxCams, (size_t)nb * kMaxCams * 4, 0u);
      }
      hv_init_kernel<<<32, 256, 0, st>>>(in);
    }
    prof_mark(st, 1);
    for (int r = 0; r < p.rounds && p.N > 0; ++r) {
      const int64_t begin = (int64_t)r * p.S;
      const int64_t end = begin + p.S < p.N ? begin + p.S : p.N;
      // the rounds after the second mostly find their frames closed: fatter CTAs, fewer of them to retire
      const int it = r < 2 ? tune.ins_iters : 4 * tune.ins_iters;
      hv_pass_kernel<Src, 0><<<dim3(src.host_grid(begin, end, it), nb), kPassThreads, 0, st>>>(src, g, w, nullptr, begin, end, r, it, 0);
    }
    prof_mark(st, 2);
    hv_count_kernel<<<dim3((unsigned)ceil_div(p.nchunks, 4), nb), 256, 0, st>>>(w, out.voxel_num);
    prof_mark(st, 3);
    {
      const int ncull = cull ? CullLaunch<Src>::blocks(src) : 0;
      hv_post_kernel<Src><<<dim3((unsigned)(ncull + ceil_div(p.nchunks, 4)), nb), 256, 0, st>>>(src, g, w, ncull);
    }
    prof_mark(st, 4);
    if (p.N > 0) {
      const int nfz = Src::kCarryFirst ? (int)ceil_div(p.max_voxels, 1024) : 0;
      hv_pass_kernel<Src, 1><<<dim3(nfz + src.host_grid(0, p.N, tune.lkp_iters), nb), kPassThreads, 0, st>>>(
          src, g, w, out.point2voxel, 0, p.N, 0, tune.lkp_iters, nfz);
    }
    prof_mark(st, 5);
    hv_emit_kernel<Src><<<dim3((unsigned)ceil_div(p.max_voxels, (emit_warps << lg_vpw)), nb), 32 * emit_warps, smem, st>>>(
        src, g, w, out, (int)wbytes, lg_vpw);
    prof_mark(st, 6);
    prof_mark(st, 7);
  }
  // join every side lane that was forked, also after an error: an un-joined lane would leave a stream capture
  // of the caller open and the caller's stream unordered against work already enqueued
  for (int l = 1; l <= forked; ++l) {
    RD3_LANE_TRY(cudaEventRecord(lanes->join[l - 1], lanes->s[l - 1]));
    RD3_LANE_TRY(cudaStreamWaitEvent(stream, lanes->join[l - 1], 0));
  }
#undef RD3_LANE_TRY
  if (status != RD3_OK) return status;
  return check_launch();
}

}  // namespace rd3
