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
//   K1 insert (R rounds over consecutive index ranges of S points):
//        a warp owns a tile of 128 points (4 per lane, one 16-byte load): validity, voxel cell by a
//        conservative fast path (magic-number rounding, proven error bound) with an exact IEEE
//        redo for the few points near a cell boundary, in-range keys compacted into a per-warp
//        list, then table operations with dense lanes.
//        Table entry {key:32 | min point index:32} in buckets of 4 (one 32-byte sector, one
//        256-bit load per probe); 64-bit CAS claims an entry, 64-bit atomicMin lowers the index.
//        A round only INSERTS while fewer than max_voxels voxels were claimed by the previous
//        rounds; later rounds only look keys up.  Voxels first seen after that point have
//        rank >= max_voxels and are dropped by the reference anyway, so the
//        table never holds more than max_voxels + S keys, whatever N is.
//        Points whose voxel is in the table are appended to their tile's candidate region.
//   K2a first : every table entry marks bit[min index] in a per-frame bitmask
//   K2b scan  : per-chunk exclusive popcount prefix; K2s chunk totals -> voxel_num
//   K3 slots  : a warp merges the candidate regions of 4 tiles into one dense list; per candidate:
//        rank r = #first-points before its voxel's first point; if r < max_voxels insert its
//        index into S[r][0..K) with a cascade of atomicMin that keeps the K smallest indices, sorted.
//   K4 emit   : a CTA owns V voxels; gathers (or re-unprojects) the slot points
//        into a shared-memory tile, then writes voxels (coalesced), coors, count
//        and the HardSimpleVFE mean from the tile.
//
// Templated on the point source: a (N,C) point array, or DA3 depth maps
// unprojected on the fly (the point cloud never exists in memory).
#pragma once
#include <stdlib.h>

#include "rd3_common.cuh"

namespace rd3 {

#ifndef RD3_INS_THREADS
#define RD3_INS_THREADS 256
#endif
constexpr int kInsThreads = RD3_INS_THREADS;
#ifndef RD3_SLOT_TILES
#define RD3_SLOT_TILES 4              // tiles whose candidate regions one warp of the slots kernel merges
#endif
constexpr int kSlotTiles = RD3_SLOT_TILES;   // power of two, <= 32
#ifndef RD3_EMIT_THREADS
#define RD3_EMIT_THREADS 256
#endif
constexpr int kEmitThreads = RD3_EMIT_THREADS;
#ifndef RD3_INS_MINB
#define RD3_INS_MINB 8                // resident insert CTAs per SM the register budget is sized for
#endif
constexpr int kTilePoints = 128;      // points per warp tile
constexpr int kInsSpan = 4 * kInsThreads;   // points per insert CTA (4 per thread)
constexpr int kTileShift = 7;
constexpr int kMaxRounds = 64;
#ifndef RD3_RANK_PACKED
#define RD3_RANK_PACKED 0             // EXPERIMENT (not yet run on a GPU): first-point flag word and its popcount prefix
#endif                                // interleaved as {flags, prefix} pairs: one random sector per rank lookup, not two
#ifndef RD3_PREFETCH
#define RD3_PREFETCH 0                // EXPERIMENT (not yet run on a GPU): the table sector of every in-range key is
#endif                                // prefetched into L2 right after the cell decision, ~150 instructions before its probe
#ifndef RD3_EMIT_TMA
#define RD3_EMIT_TMA 0                // EXPERIMENT (not yet run on a GPU): emit stages the calibration with the TMA bulk
#endif                                // copy of the insert kernel (its load / store loop holds 7 % of emit's stall samples)
#ifndef RD3_LATE_CLAIMS
#define RD3_LATE_CLAIMS 0             // EXPERIMENT (profiles/r1_analysis.md, not yet run on a GPU): the claims sum is
#endif                                // loaded after the prologue barrier and waited for (mbarrier) only at stage C
#ifndef RD3_KNOCK
#define RD3_KNOCK 0                   // DIAGNOSTIC builds only (wrong results): lookup-only rounds skip 1: the exact
#endif                                // redo, 2: + the probes, 3: everything after the prologue (tools/knockout.sh)

// ---------------------------------------------------------------------------
// point sources
// ---------------------------------------------------------------------------
// stage-AB result of one lane (4 consecutive points)
struct Quad {
  uint32_t key[4];
  unsigned in, und;
};

struct PointsSource {
  const float *pts;   // (B, N, C)
  int64_t N;
  int C;
  static constexpr bool kIsDepth = false;

  __device__ __forceinline__ void stage(float *, int) const {}
  __device__ __forceinline__ bool stage_async(float *, uint64_t *, int) const { return false; }
  __device__ __forceinline__ void prepare(float *, int) const {}
  int host_num_feats() const { return C; }
  __device__ __forceinline__ int num_feats() const { return C; }

  // stage AB: a lane owns 4 consecutive points
  struct Pre {};
  __device__ __forceinline__ Pre preload(int, int64_t, int64_t) const { return Pre(); }
  struct Cursor { const float *p; int npx; };
  __device__ __forceinline__ Cursor cursor(int b, int64_t i0, int64_t end, const float *, const Pre &) const {
    Cursor c;
    c.p = pts + ((int64_t)b * N + i0) * C;
    const int64_t left = end - i0;
    c.npx = left >= 4 ? 4 : (left > 0 ? (int)left : 0);
    if (c.npx == 0) c.p = pts;                      // lanes past the end read point 0 (masked)
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
  DepthParams p;
  CellRange rg;            // range filter in cell units (fused path)
  int vec_ok;              // 16-byte aligned float4 loads are legal
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
  // stage AB: a lane owns 4 consecutive pixels
  struct Cursor {
    float z[4];
    unsigned valid;        // depth/conf/sky mask of the 4 pixels
    uint32_t cam;
    float uf;              // column of the first pixel
    bool wraps;            // the 4 pixels cross a row boundary (only when W % 4 != 0)
    float tx, ty, tz;      // row part of the direct cell map (valid when !wraps)
  };
  // The depth load does not depend on the calibration: it is issued before the CTA's prologue
  // barrier so that its DRAM latency overlaps the claims / calibration loads.
  struct Pre { float z[4]; };
  __device__ __forceinline__ Pre preload(int b, int64_t i0, int64_t end) const {
    Pre r;
    const int64_t gi = (int64_t)b * p.npix + i0;
    const int64_t left = end - i0;
    const int npx = left >= 4 ? 4 : (left > 0 ? (int)left : 0);
    if (vec_ok && npx == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(depth + gi));
      r.z[0] = t.x; r.z[1] = t.y; r.z[2] = t.z; r.z[3] = t.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) r.z[q] = (q < npx) ? __ldg(depth + gi + q) : 0.0f;
    }
    return r;
  }
  __device__ __forceinline__ Cursor cursor(int b, int64_t i0, int64_t end, const float *s_cal, const Pre &pre) const {
    Cursor c;
    const int64_t gi = (int64_t)b * p.npix + i0;
    const int64_t left = end - i0;
    const int npx = left >= 4 ? 4 : (left > 0 ? (int)left : 0);
#pragma unroll
    for (int q = 0; q < 4; ++q) c.z[q] = pre.z[q];
    // z > 0 & isfinite(z) [& z <= max_depth] (:338-340); p.zmax = min(max_depth, FLT_MAX)
    c.valid = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) c.valid |= ((c.z[q] > 0.0f) & (c.z[q] <= p.zmax)) ? (1u << q) : 0u;
    c.valid &= (1u << npx) - 1u;
    if (p.use_masks) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (((c.valid >> q) & 1u) && !depth_ok(c.z[q], gi + q, b)) c.valid &= ~(1u << q);
    }
    uint32_t v, u;
    pixel_cvu(npx ? (uint32_t)i0 : 0u, c.cam, v, u);   // lanes past the end compute on pixel 0 (masked)
    c.uf = (float)u;
    c.wraps = npx && u + 3 >= (uint32_t)p.W;
    pixel_cell_row((float)v, s_cal + c.cam * kCalibFloats, c.tx, c.ty, c.tz);
    return c;
  }
  // The direct pixel->cell map (pixel_key_fast) decides all but the pixels within its error bound
  // of a cell / range-filter boundary; those (bit q of `und`) are redone by cell_exact.  Computed
  // for every pixel (masked ones yield garbage that is discarded): no divergence.  Lanes whose 4
  // pixels cross a row boundary (only when W % 4 != 0) leave all of them to cell_exact.
  __device__ __forceinline__ void classify(const Cursor &c, const float *s_cal, const VoxelGrid &g, Quad &qd) const {
    const float *cal = s_cal + c.cam * kCalibFloats;
    unsigned in = 0, und = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = pixel_key_fast(c.z[q], __fadd_rn(c.uf, (float)q), c.tx, c.ty, c.tz, cal, g, rg, qd.key[q]);
      in |= (r == 1 ? 1u : 0u) << q;
      und |= (r == 2 ? 1u : 0u) << q;
    }
    const bool fast = g.fast_ok && !c.wraps;
    qd.in = fast ? (in & c.valid) : 0u;
    qd.und = fast ? (und & c.valid) : c.valid;
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
};

// ---------------------------------------------------------------------------
// per-call work description (device pointers, all with a leading frame dim)
// ---------------------------------------------------------------------------
struct HvWork {
  unsigned long long *table;  // [B][cap]   {key:32 | min point idx:32}, empty = ~0
  uint32_t *slots;            // [B][max_voxels*K] sorted point indices, empty = ~0
  uint32_t *flags;            // [B][nwords] bit i: point i is the first of a voxel
  int32_t *wordprefix;        // [B][nwords] exclusive popcount prefix inside the chunk
  int32_t *chunk_base;        // [B][nchunks] totals, then exclusive bases after K2s
  int32_t *round_claims;      // [B][kMaxRounds] voxels claimed per insert round
  uint8_t *cand_cnt;          // [B][ntiles] candidates of each 128-point tile (0..128)
  uint2 *cand;                // [B][N] (point idx, table slot); tile t owns [t*128, t*128+128)
  int ntiles;
  int64_t N;
  int b0;                     // first frame of the group this launch works on
  int64_t cap;
  uint32_t cap_mask;
  int log2cap;
  int direct;                 // grid volume <= cap: slot = key, no probing
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
// hashes here and is not in it is absent: a lookup is one 256-bit load (one L1 request, one DRAM
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

// Insert (key, idx).  Slot ownership is permanent (CAS from empty); the payload
// only ever decreases (atomicMin), so a stale read can only cause a redundant
// atomic, never a wrong skip.  *claimed = 1 iff this call created the entry.
__device__ __forceinline__ uint32_t table_insert(unsigned long long *table, const HvWork &w,
                                                 uint32_t key, uint32_t idx, int *claimed) {
  const unsigned long long mine = ((unsigned long long)key << 32) | idx;
  *claimed = 0;
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
      if (old == kEmpty64) { *claimed = 1; return slot; }
      e = old;
    }
    if ((uint32_t)(e >> 32) == key) {
      if ((uint32_t)e > idx) atomicMin(table + slot, mine);
      return slot;
    }
    slot = (slot + 1) & w.cap_mask;    // lost the entry to another key: plain linear probing from here
    e = __ldcg(table + slot);
  }
}

// Lookup only (the table is not written while lookups run); kEmpty32 when the key is absent.
__device__ __forceinline__ uint32_t table_find(const unsigned long long *table, const HvWork &w,
                                               uint32_t key) {
  if (w.direct) return __ldcg(table + key) == kEmpty64 ? kEmpty32 : key;
  uint32_t s0 = hash_bucket_slot(key, w.log2cap);
  while (true) {
    unsigned long long e0, e1, e2, e3;
    load_bucket(table + s0, e0, e1, e2, e3);
    int q = 4;
    q = (uint32_t)(e3 >> 32) == key ? 3 : q;
    q = (uint32_t)(e2 >> 32) == key ? 2 : q;
    q = (uint32_t)(e1 >> 32) == key ? 1 : q;
    q = (uint32_t)(e0 >> 32) == key ? 0 : q;
    if (q < 4) return s0 + q;
    if ((uint32_t)(e3 >> 32) == kEmpty32) return kEmpty32;    // no valid key is ~0
    s0 = (s0 + 4) & w.cap_mask;
  }
}

// K1 ------------------------------------------------------------------------
// grid (ceil((end-begin)/1024), frames), 256 threads; every WARP owns a tile of 128
// consecutive points and runs its stages without block barriers:
//   AB each lane walks its 4 consecutive points (one 16-byte load): validity, voxel cell
//      by the conservative fast path; in-range keys and the few undecided points are
//      ballot-compacted; the undecided ones are redone with exact IEEE arithmetic
//   C  dense lanes: table insert (or lookup once max_voxels voxels exist);
//      hits go to the tile's own region of the candidate list (no global counter)
template <class Src>
__global__ void __launch_bounds__(kInsThreads, RD3_INS_MINB)
    hv_insert_kernel(Src src, VoxelGrid g, HvWork w, int64_t begin, int64_t end, int round) {
  __shared__ __align__(16) float s_cal[Src::kIsDepth ? kMaxCams * kCalibFloats : 4];
  __shared__ __align__(8) uint64_t s_bar;        // completion of the calibration's TMA copy
  __shared__ uint2 s_itemb[kInsSpan];            // (key, local point id) of the in-range points
  __shared__ uint8_t s_undb[kInsSpan];           // local ids of the undecided points
  __shared__ int s_prev, s_claims, s_done;

  const int b = blockIdx.y + w.b0;
  const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t block_base = begin + (int64_t)blockIdx.x * kInsSpan;
  if (block_base >= end) return;
#if RD3_LATE_CLAIMS
#if RD3_KNOCK
#error "the RD3_KNOCK builds need the claims before the prologue barrier"
#endif
  // The source-level profile puts 14 % of all warp time on the prologue barrier: seven warps wait there for
  // warp 0's load of the round claims, a value that is first needed at stage C.  Here the barrier only covers
  // the mbarrier inits + the calibration copy; warp 0 loads the claims after it and publishes them through a
  // one-shot mbarrier that the other warps look at when they reach stage C (by then it has long completed).
  __shared__ __align__(8) uint64_t s_bar2;
  if (tid == 0) { s_claims = 0; s_done = 0; mbar_init(&s_bar2, 1); }
  const typename Src::Pre pre = src.preload(b, block_base + wv * kTilePoints + 4 * lane, end);
  const bool cal_async = src.stage_async(s_cal, &s_bar, b);
  __syncthreads();
  if (wv == 0) {
    int c = 0;
    for (int r = lane; r < round; r += 32) c += __ldcg(w.round_claims + b * kMaxRounds + r);
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) {
      s_prev = c;
      mbar_arrive(&s_bar2);               // release: s_prev is visible to whoever sees the phase complete
    }
  }
  if (cal_async) tma_wait(&s_bar);
#else
  if (tid == 0) { s_claims = 0; s_done = 0; }
  const typename Src::Pre pre = src.preload(b, block_base + wv * kTilePoints + 4 * lane, end);
  if (wv == 0) {
    // voxels claimed by the previous rounds: insert vs lookup-only for the whole round
    int c = 0;
    for (int r = lane; r < round; r += 32) c += __ldcg(w.round_claims + b * kMaxRounds + r);
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) s_prev = c;
  }
  const bool cal_async = src.stage_async(s_cal, &s_bar, b);
  __syncthreads();
  if (cal_async) tma_wait(&s_bar);
  const bool lookup_only = s_prev >= w.max_voxels;
#endif

  if (block_base + wv * kTilePoints >= end) return;
#if RD3_KNOCK >= 3
  if (lookup_only) {                                        // timing only: prologue + depth load
    if (lane == 0 && sizeof(pre) == 16 && reinterpret_cast<const float *>(&pre)[0] == 123.456f) w.cand_cnt[0] = 1;
    return;
  }
#endif
  uint2 *s_item = s_itemb + wv * kTilePoints;
  uint8_t *s_und = s_undb + wv * kTilePoints;
  unsigned long long *table = w.table + (int64_t)b * w.cap;
  int claims = 0;
  const int64_t base = block_base + wv * kTilePoints;     // this warp's tile

  // ---- stage AB ---------------------------------------------------------------------
  int n2, nu;
  {
    typename Src::Cursor cur = src.cursor(b, base + 4 * lane, end, s_cal, pre);
    Quad qd;
    src.classify(cur, s_cal, g, qd);
#if RD3_PREFETCH
    // 23 % of the kernel's stall samples are the long-scoreboard wait for the probed bucket.  The key is known
    // here, the probe only happens after the compaction: a prefetch (no register, no dependency) starts the
    // DRAM access now, so that the probe finds its sector in L2.
    if (!w.direct) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if ((qd.in >> q) & 1u)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(table + hash_bucket_slot(qd.key[q], w.log2cap)));
    }
#endif
    // exclusive prefix of the per-lane counts (0..4) from three ballots of the count's bit planes
    const unsigned cin = __popc(qd.in);
    const unsigned p0 = __ballot_sync(0xffffffffu, cin & 1u), p1 = __ballot_sync(0xffffffffu, cin & 2u),
                   p2 = __ballot_sync(0xffffffffu, cin & 4u);
    n2 = __popc(p0) + 2 * __popc(p1) + 4 * __popc(p2);
    uint2 *it = s_item + (__popc(p0 & lt) + 2 * __popc(p1 & lt) + 4 * __popc(p2 & lt));
    const unsigned in = qd.in;
    // the order inside the list is irrelevant: point q of the lane goes to the lane's slot #(set bits below q)
    if (in & 1u) it[0] = make_uint2(qd.key[0], (uint32_t)(4 * lane));
    if (in & 2u) it[in & 1u] = make_uint2(qd.key[1], (uint32_t)(4 * lane + 1));
    if (in & 4u) it[__popc(in & 3u)] = make_uint2(qd.key[2], (uint32_t)(4 * lane + 2));
    if (in & 8u) it[__popc(in & 7u)] = make_uint2(qd.key[3], (uint32_t)(4 * lane + 3));
    nu = 0;
    if (__any_sync(0xffffffffu, qd.und != 0u)) {
      const unsigned cun = __popc(qd.und);
      const unsigned u0 = __ballot_sync(0xffffffffu, cun & 1u), u1 = __ballot_sync(0xffffffffu, cun & 2u),
                     u2 = __ballot_sync(0xffffffffu, cun & 4u);
      nu = __popc(u0) + 2 * __popc(u1) + 4 * __popc(u2);
      int ua = __popc(u0 & lt) + 2 * __popc(u1 & lt) + 4 * __popc(u2 & lt);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if ((qd.und >> q) & 1u) s_und[ua++] = (uint8_t)(4 * lane + q);
    }
  }
  __syncwarp();
#if RD3_KNOCK >= 1
  if (lookup_only) nu = 0;                                  // timing only: no exact redo of undecided points
#endif
#if RD3_KNOCK >= 2
  if (lookup_only) {                                        // timing only: no table probes
    if (lane == 0) w.cand_cnt[(int64_t)b * w.ntiles + (base >> kTileShift)] = (uint8_t)(n2 == 77777);
    return;
  }
#endif
#pragma unroll 1
  for (int j0 = 0; j0 < nu; j0 += 32) {       // within the error bound of a boundary: exact arithmetic
    const int j = j0 + lane;
    bool in = false;
    int lid = 0, cx, cy, cz;
    if (j < nu) {
      lid = s_und[j];
      in = src.cell_exact(b, base + lid, s_cal, g, cx, cy, cz);
    }
    const unsigned b1 = __ballot_sync(0xffffffffu, in);
    if (in) s_item[n2 + __popc(b1 & lt)] = make_uint2(voxel_key(cx, cy, cz, g), (uint32_t)lid);
    n2 += __popc(b1);
  }
  __syncwarp();

  // ---- stage C ----------------------------------------------------------------------
#if RD3_LATE_CLAIMS
  tma_wait(&s_bar2);
  const bool lookup_only = s_prev >= w.max_voxels;
#endif
  uint2 *cand = w.cand + (int64_t)b * w.N + base;
  int nc = 0;
#pragma unroll 1
  for (int j0 = 0; j0 < n2; j0 += 32) {
    const int j = j0 + lane;
    uint32_t slot = kEmpty32, idx = 0;
    if (j < n2) {
      const uint2 it = s_item[j];
      idx = (uint32_t)base + it.y;
      if (lookup_only) {
        slot = table_find(table, w, it.x);
      } else {
        int c;
        slot = table_insert(table, w, it.x, idx, &c);
        claims += c;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, slot != kEmpty32);
    if (slot != kEmpty32) cand[nc + __popc(bal & lt)] = make_uint2(idx, slot);
    nc += __popc(bal);
  }
  if (lane == 0) w.cand_cnt[(int64_t)b * w.ntiles + (base >> kTileShift)] = (uint8_t)nc;
  if (!lookup_only) {
    // one global atomic per CTA: the last warp to finish adds the CTA's claims to the round's counter
    for (int d = 16; d > 0; d >>= 1) claims += __shfl_xor_sync(0xffffffffu, claims, d);
    if (lane == 0) {
      const int64_t left = end - block_base;
      const int nwarps = left >= kInsSpan ? kInsThreads / 32 : (int)((left + kTilePoints - 1) >> kTileShift);
      if (claims) atomicAdd(&s_claims, claims);
      __threadfence_block();
      if (atomicAdd(&s_done, 1) == nwarps - 1) {
        const int c = atomicAdd(&s_claims, 0);
        if (c) atomicAdd(w.round_claims + b * kMaxRounds + round, c);
      }
    }
  }
}

// K2a -----------------------------------------------------------------------
// every table entry marks its voxel's first point.  grid (gx, frames), grid-stride over the
// table so that the launch has a few thousand fat blocks instead of cap/256 per frame.
static __global__ void __launch_bounds__(256) hv_first_kernel(HvWork w) {
  const int b = blockIdx.y + w.b0;
  const unsigned long long *table = w.table + (int64_t)b * w.cap;
  uint32_t *flags = w.flags + (int64_t)b * w.nwords * (RD3_RANK_PACKED ? 2 : 1);
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < w.cap;
       s += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long e = __ldg(table + s);
    if (e == kEmpty64) continue;
    const uint32_t first = (uint32_t)e;
    atomicOr(flags + (first >> 5) * (RD3_RANK_PACKED ? 2 : 1), 1u << (first & 31));
  }
}

// K2b -----------------------------------------------------------------------
// grid (nchunks, B), kScanThreads threads, one flag word each.
static __global__ void __launch_bounds__(kScanThreads) hv_flagscan_kernel(HvWork w) {
  __shared__ int s_warp[kScanThreads / 32];
  const int b = blockIdx.y + w.b0;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int64_t wi = (int64_t)b * w.nwords + (int64_t)blockIdx.x * kChunkWords + threadIdx.x;
#if RD3_RANK_PACKED
  const int cnt = __popc(w.flags[2 * wi]);
#else
  const int cnt = __popc(w.flags[wi]);
#endif
  const int inc = warp_inclusive_scan(cnt);
  if (lane == 31) s_warp[wv] = inc;
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int k = 0; k < kScanThreads / 32; ++k) {
    const int t = s_warp[k];
    if (k < wv) base += t;
    total += t;
  }
#if RD3_RANK_PACKED
  w.flags[2 * wi + 1] = (uint32_t)(base + inc - cnt);
#else
  w.wordprefix[wi] = base + inc - cnt;
#endif
  if (threadIdx.x == 0) w.chunk_base[(int64_t)b * w.nchunks + blockIdx.x] = total;
}

// Shared tail of the ordered-flag kernels (depth.cu): lane L of warp wv holds the
// 32-bit flag word (chunk*kChunkWords + wv*32 + L).  Stores the word, its exclusive
// popcount prefix inside the chunk, and the chunk total.
__device__ __forceinline__ void chunk_scan_store(uint32_t my_word, int *s_warp, uint32_t *flags,
                                                 int32_t *wordprefix, int32_t *chunk_total) {
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int cnt = __popc(my_word);
  const int inc = warp_inclusive_scan(cnt);
  if (lane == 31) s_warp[wv] = inc;
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int k = 0; k < kScanThreads / 32; ++k) {
    const int t = s_warp[k];
    if (k < wv) base += t;
    total += t;
  }
  const int64_t wi = (int64_t)blockIdx.x * kChunkWords + wv * 32 + lane;
  flags[wi] = my_word;
  wordprefix[wi] = base + inc - cnt;
  if (threadIdx.x == 0) chunk_total[blockIdx.x] = total;
}

// K2s -----------------------------------------------------------------------
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

// rank (first-occurrence order) of the voxel whose first point is `first_idx`
__device__ __forceinline__ int voxel_rank(const HvWork &w, int b, uint32_t first_idx) {
  const uint32_t word = first_idx >> 5;
  const int64_t wi = (int64_t)b * w.nwords + word;
#if RD3_RANK_PACKED
  const uint2 fw = __ldg(reinterpret_cast<const uint2 *>(w.flags) + wi);
  return __ldg(w.chunk_base + (int64_t)b * w.nchunks + (first_idx >> kChunkShift)) + (int)fw.y +
         __popc(fw.x & ((1u << (first_idx & 31)) - 1u));
#else
  const uint32_t bits = __ldg(w.flags + wi) & ((1u << (first_idx & 31)) - 1u);
  return __ldg(w.chunk_base + (int64_t)b * w.nchunks + (first_idx >> kChunkShift)) +
         __ldg(w.wordprefix + wi) + __popc(bits);
#endif
}

// Keep the K smallest point indices of a voxel, sorted, with atomicMin only.
// Slot 0 is reserved for the voxel's first point (known from the table), which is
// stored without an atomic; every other point cascades through slots 1..K-1:
// a value reaches slot k only after losing against the slots before it, so the
// non-empty prefix is strictly increasing at all times and the final content is
// independent of arrival order.  `s_last` is a (possibly stale, hence larger) copy of
// slot K-1: if it is already smaller than idx, K smaller indices exist.
__device__ __forceinline__ void slot_insert_tail(uint32_t *S, int K, uint32_t idx) {
  uint32_t cur = idx;
  int k = 1;
  // Skip the prefix of smaller indices with 8 independent loads at a time instead of one
  // dependent load per slot (a voxel that already holds 7 points would cost 7 round trips).
  // A slot read as smaller than cur can only have decreased since: it stays smaller.
  while (k < K) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (k + i < K) ? __ldcg(S + k + i) : kEmpty32;
    int i = 8;
#pragma unroll
    for (int q = 7; q >= 0; --q) i = (v[q] < cur) ? i : q;      // first slot not smaller than cur
    k += i;
    if (i < 8) break;
  }
  for (; k < K; ++k) {
    const uint32_t old = atomicMin(S + k, cur);
    if (old == kEmpty32 || old == cur) return;
    if (old > cur) cur = old;            // displaced a larger index: carry it on
  }
}

// K3 ------------------------------------------------------------------------
// grid (ceil(ntiles/(4*kSlotTiles)), frames), 128 threads (small CTAs: warps finish at very different times
// and a CTA's slot is only recycled when all have).  A warp works off the candidate regions of kSlotTiles
// consecutive tiles as ONE list: a warp scan of the counts gives each tile's offset, lane j of
// a pass finds its (tile, k) by a log2(kSlotTiles)-step search over those offsets, so the lanes stay dense however
// few candidates a tile has (the lookup-only rounds leave ~5 per tile).
// Neighbouring pixels often share a voxel: lanes holding the same table slot form a
// group (__match_any_sync); only the group leader walks table -> rank -> last slot and
// broadcasts the result, and a whole group leaves after one load when the voxel is full.
static __global__ void __launch_bounds__(128) hv_slots_kernel(HvWork w, int32_t *point2voxel) {
  const int b = blockIdx.y + w.b0;
  const int lane = threadIdx.x & 31;
  const int t0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kSlotTiles;
  if (t0 >= w.ntiles) return;
  const int mine = (lane < kSlotTiles && t0 + lane < w.ntiles) ? w.cand_cnt[(int64_t)b * w.ntiles + t0 + lane] : 0;
  const int inc = warp_inclusive_scan(mine);
  const int excl = inc - mine;
  const int n = __shfl_sync(0xffffffffu, inc, 31);
  const uint2 *cand = w.cand + (int64_t)b * w.N + ((int64_t)t0 << kTileShift);
  const unsigned long long *table = w.table + (int64_t)b * w.cap;
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    const bool on = j < n;
    // tile of list position j: the largest lane t with excl[t] <= j (empty tiles share their
    // offset with the next non-empty one, which the "largest" picks)
    int t = 0;
#pragma unroll
    for (int step = kSlotTiles / 2; step > 0; step >>= 1) {
      const int e = __shfl_sync(0xffffffffu, excl, t + step);
      if (e <= j) t += step;
    }
    const int k = j - __shfl_sync(0xffffffffu, excl, t);
    const uint2 c = on ? __ldg(cand + (t << kTileShift) + k) : make_uint2(0u, kEmpty32 - lane);   // distinct dummies
    const unsigned grp = __match_any_sync(0xffffffffu, c.y);
    const int leader = __ffs(grp) - 1;
    uint32_t first_idx = 0, last = 0;
    int r = w.max_voxels;
    if (on && lane == leader) {
      first_idx = (uint32_t)__ldg(table + c.y);
      r = voxel_rank(w, b, first_idx);
    }
    first_idx = __shfl_sync(0xffffffffu, first_idx, leader);
    r = __shfl_sync(0xffffffffu, r, leader);
    // the last slot is only needed by points that are not their voxel's first point
    const bool tail = on && r < w.max_voxels && w.K > 1 && c.x != first_idx;
    const unsigned tails = __ballot_sync(0xffffffffu, tail);
    if (tails & grp) {
      if (lane == leader) last = __ldcg(w.slots + ((int64_t)b * w.max_voxels + r) * w.K + (w.K - 1));
      last = __shfl_sync(grp, last, leader);
    }
    if (on && r < w.max_voxels) {
      uint32_t *S = w.slots + ((int64_t)b * w.max_voxels + r) * w.K;
      if (c.x == first_idx) S[0] = c.x;                       // the voxel's first point
      else if (w.K > 1 && !(last < c.x)) slot_insert_tail(S, w.K, c.x);
      if (point2voxel) point2voxel[(int64_t)b * w.N + c.x] = r;
    }
  }
}

// K4 ------------------------------------------------------------------------
// A CTA owns V consecutive voxels (~1024 slot items).  dynamic smem: tile[V*K*C] floats
// (padded to 16 B) + idx[V*K] u32 + list[V*K] u16.
//   1. zero the tile (16-byte stores); every warp loads its share of the slot indices and
//      ballot-compacts the non-empty ones into its own list segment (no atomics, no barrier)
//   2. each warp gathers / re-unprojects (exact reference arithmetic) its listed items
//   3. voxels are copied out with 16-byte stores; one thread per voxel writes coors (from the
//      first point), the count and the HardSimpleVFE mean (slot order, one __fdiv_rn)
template <class Src>
__global__ void __launch_bounds__(kEmitThreads) hv_emit_kernel(Src src, VoxelGrid g, HvWork w, HvOut o, int V) {
  extern __shared__ float s_dyn[];
#if RD3_EMIT_TMA
  __shared__ __align__(16) float s_cal[Src::kIsDepth ? kMaxCams * kCalibFloats : 4];
  __shared__ __align__(8) uint64_t s_bar;
#else
  __shared__ float s_cal[Src::kIsDepth ? kMaxCams * kCalibFloats : 1];
#endif
  const int b = blockIdx.y + w.b0;
  const int vn = o.voxel_num[b];
  const int r0 = blockIdx.x * V;
  if (r0 >= vn) return;
  const int C = src.num_feats();
  const int K = w.K;
  const int nvox = min(V, vn - r0);
  const int items = nvox * K;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  constexpr int nw = kEmitThreads / 32;
  float *tile = s_dyn;
  uint32_t *s_idx = reinterpret_cast<uint32_t *>(s_dyn + (((size_t)V * K * C + 3) & ~(size_t)3));
  uint16_t *s_list = reinterpret_cast<uint16_t *>(s_idx + (size_t)V * K);
  const uint32_t *S = w.slots + ((int64_t)b * w.max_voxels + r0) * K;

  // warp wv owns items [wv*per, wv*per+per): its list segment starts at the same offset
  const int per = ((items + nw - 1) / nw + 31) & ~31;
  const int lo = wv * per, hi = min(items, lo + per);
  // The first 128 slot indices of the warp's share are requested up front (4 independent loads per
  // lane): one DRAM round trip instead of four dependent ones (each pass of the compaction loop below
  // waits for its load), overlapped with the calibration staging and the zero fill (emit 0.362 -> 0.320 ms).
  uint32_t pre[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int it = lo + 32 * q + lane;
    pre[q] = it < hi ? __ldg(S + it) : kEmpty32;
  }
#if RD3_EMIT_TMA
  const bool cal_async = src.stage_async(s_cal, &s_bar, b);     // one bulk copy instead of a load / store loop
#else
  src.stage(s_cal, b);
#endif
  {
    float4 *t4 = reinterpret_cast<float4 *>(tile);
    const int n4 = (items * C + 3) >> 2;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = threadIdx.x; e < n4; e += kEmitThreads) t4[e] = z4;
  }
  int nmine = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int it = lo + 32 * q + lane;
    const uint32_t idx = pre[q];
    if (it < hi) s_idx[it] = idx;
    const unsigned bal = __ballot_sync(0xffffffffu, idx != kEmpty32);
    if (idx != kEmpty32) s_list[lo + nmine + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)it;
    nmine += __popc(bal);
  }
  for (int it0 = lo + 128; it0 < hi; it0 += 32) {
    const int it = it0 + lane;
    uint32_t idx = kEmpty32;
    if (it < hi) {
      idx = __ldg(S + it);
      s_idx[it] = idx;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, idx != kEmpty32);
    if (idx != kEmpty32) s_list[lo + nmine + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)it;
    nmine += __popc(bal);
  }
  __syncthreads();                       // tile zeroed, calibration staged
#if RD3_EMIT_TMA
  if (cal_async) tma_wait(&s_bar);
#endif
  for (int j = lane; j < nmine; j += 32) {
    const int it = s_list[lo + j];
    src.gather(b, s_idx[it], s_cal, tile + it * C);
  }
  __syncthreads();

  // voxels: contiguous nvox*K*C floats (skipped when the caller only wants coors / num / mean)
  float *vout = o.voxels + ((int64_t)b * w.max_voxels + r0) * K * C;
  const int nfl = o.voxels ? items * C : 0;
  if ((reinterpret_cast<uintptr_t>(vout) & 15) == 0) {
    const float4 *t4 = reinterpret_cast<const float4 *>(tile);
    float4 *v4 = reinterpret_cast<float4 *>(vout);
    const int n4 = nfl >> 2;
    for (int e = threadIdx.x; e < n4; e += kEmitThreads) v4[e] = t4[e];
    for (int e = (nfl & ~3) + threadIdx.x; e < nfl; e += kEmitThreads) vout[e] = tile[e];
  } else {
    for (int e = threadIdx.x; e < nfl; e += kEmitThreads) vout[e] = tile[e];
  }

  // one thread per voxel: count, coors from the first point, HardSimpleVFE mean
  // (voxel_encoder.py:45-46: sum over ALL K slots in slot order, then one division; the
  // slots beyond the count are zeros, so the running sum stops changing at the count --
  // except that (-0.0) + 0.0 = +0.0, which one extra "+ 0.0f" reproduces)
  const int F = o.F;
  for (int v = threadIdx.x; v < nvox; v += kEmitThreads) {
    const uint32_t *si = s_idx + v * K;
    int cnt = 0;
    while (cnt < K && si[cnt] != kEmpty32) ++cnt;
    const float *p0 = tile + v * K * C;
    int cx = 0, cy = 0, cz = 0;
    if (voxel_coor_fast(p0[0], p0[1], p0[2], 0.0f, g, cx, cy, cz) == 2)
      voxel_coor(p0[0], p0[1], p0[2], g, cx, cy, cz);
    const int64_t vr = (int64_t)b * w.max_voxels + r0 + v;
    o.coors[vr * 3 + 0] = cz;
    o.coors[vr * 3 + 1] = cy;
    o.coors[vr * 3 + 2] = cx;
    o.num[vr] = cnt;
    if (o.mean) {
      const float n = (float)cnt;
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
        mo[0] = __fdiv_rn(sx, n);
        mo[1] = __fdiv_rn(sy, n);
        mo[2] = __fdiv_rn(sz, n);
      } else {
        for (int f = 0; f < F; ++f) {
          float sacc = 0.0f;
          const float *p = p0 + f;
          for (int k = 0; k < cnt; ++k, p += C) sacc = __fadd_rn(sacc, *p);
          if (cnt < K) sacc = __fadd_rn(sacc, 0.0f);
          mo[f] = __fdiv_rn(sacc, n);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// host side: workspace carving + launch sequence
// ---------------------------------------------------------------------------
struct HvPlan {
  int64_t N;
  int B;
  int K;
  int max_voxels;
  int64_t S;          // points per insert round (multiple of 1024)
  int rounds;
  int64_t cap;
  int log2cap;
  int nwords, nchunks, ntiles;
  // [table | slots] are set to 0xFF with one memset, [flags | round_claims] to 0 with another
  size_t off_table, off_slots, off_flags, off_claims, off_ccount, off_prefix, off_chunk, off_cand, total;
};

inline HvPlan hv_plan(int64_t N, int B, int K, int max_voxels) {
  HvPlan p;
  p.N = N; p.B = B; p.K = K; p.max_voxels = max_voxels;
  const int64_t n1 = N > 0 ? N : 1;
  // round length: at most 8 rounds, but a round should keep the whole GPU busy
  // (>= ~1.2 M points over all frames), so small batches use fewer, longer rounds
  int64_t S = ceil_div(ceil_div(n1, 8), 1024) * 1024;
  const int64_t fill = ceil_div(ceil_div((int64_t)148 * 8 * 1024, B > 0 ? B : 1), 1024) * 1024;
  if (S < fill) S = fill;
  if (S < 65536) S = 65536;
  p.S = S;
  p.rounds = (int)ceil_div(n1, S);
  // the table never holds more than max_voxels + S keys (see header), nor more than N
  int64_t keys = (int64_t)max_voxels + S;
  if (keys > n1) keys = n1;
#ifndef RD3_TABLE_LOAD_PCT
#define RD3_TABLE_LOAD_PCT 50         // worst-case load factor of the table in percent
#endif
  int64_t want = keys * 100 / RD3_TABLE_LOAD_PCT;
  int lg = 10;
  while (((int64_t)1 << lg) < want) ++lg;
  p.log2cap = lg;
  p.cap = (int64_t)1 << lg;
  p.nchunks = (int)ceil_div(n1, kChunkPoints);
  p.nwords = p.nchunks * kChunkWords;
  p.ntiles = (int)ceil_div(n1, kTilePoints);
  size_t off = 0;
  p.off_table = off; off += align_up((size_t)B * p.cap * 8);
  p.off_slots = off; off += align_up((size_t)B * max_voxels * K * 4);
  p.off_flags = off; off += align_up((size_t)B * p.nwords * 4 * (RD3_RANK_PACKED ? 2 : 1));
  p.off_claims = off; off += align_up((size_t)B * kMaxRounds * 4);
  p.off_ccount = off; off += align_up((size_t)B * p.ntiles);
  p.off_prefix = off; off += align_up((size_t)B * p.nwords * 4);
  p.off_chunk = off; off += align_up((size_t)B * p.nchunks * 4);
  p.off_cand = off; off += align_up((size_t)B * n1 * 8);
  p.total = off;
  return p;
}

template <class Src>
int hv_run(const Src &src, const VoxelGrid &g, uint64_t volume, const HvPlan &p, void *ws,
           HvOut out, cudaStream_t stream) {
  char *base = (char *)ws;
  HvWork w;
  w.table = (unsigned long long *)(base + p.off_table);
  w.slots = (uint32_t *)(base + p.off_slots);
  w.flags = (uint32_t *)(base + p.off_flags);
  w.round_claims = (int32_t *)(base + p.off_claims);
  w.cand_cnt = (uint8_t *)(base + p.off_ccount);
  w.ntiles = p.ntiles;
  w.wordprefix = (int32_t *)(base + p.off_prefix);
  w.chunk_base = (int32_t *)(base + p.off_chunk);
  w.cand = (uint2 *)(base + p.off_cand);
  w.N = p.N; w.cap = p.cap; w.cap_mask = (uint32_t)(p.cap - 1); w.log2cap = p.log2cap;
  w.direct = (volume <= (uint64_t)p.cap) ? 1 : 0;
  w.nwords = p.nwords; w.nchunks = p.nchunks; w.K = p.K; w.max_voxels = p.max_voxels;

  const int C = src.host_num_feats();
#ifndef RD3_EMIT_ITEMS
#define RD3_EMIT_ITEMS 4             // slot items per thread of an emit CTA: amortises the CTA prologue
#endif
  int V = RD3_EMIT_ITEMS * kEmitThreads / p.K;
  if (V < 1) V = 1;
  if (V > 32 * RD3_EMIT_ITEMS) V = 32 * RD3_EMIT_ITEMS;
  while (V > 1 && (size_t)V * p.K * (C + 1) * 4 > 10 * 1024 * RD3_EMIT_ITEMS) V /= 2;
  const size_t smem = align_up((size_t)V * p.K * C * 4, 16) + (size_t)V * p.K * 4 +
                      align_up((size_t)V * p.K * 2, 16) + (size_t)V * 4;
  if (smem > 200 * 1024) return RD3_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    RD3_CUDA_TRY(cudaFuncSetAttribute(hv_emit_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  if (out.point2voxel && p.N > 0)
    RD3_CUDA_TRY(cudaMemsetAsync(out.point2voxel, 0xFF, (size_t)p.B * p.N * 4, stream));

  // Frames are independent: the batch is split into up to kMaxLanes sub-batches that run the
  // whole kernel sequence on their own streams (forked from / joined to the caller's stream),
  // so that the issue-bound insert kernel of one sub-batch overlaps the latency-bound
  // first/slots kernels and memsets of another.  With the stage profiler on, one lane is used.
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
  if (nl > 1) RD3_CUDA_TRY(cudaEventRecord(lanes->fork, stream));
  // Inside a lane the frames are processed in groups of G (RD3_GROUP, default: the whole lane).
  int G = p.B;
  if (const char *e = getenv("RD3_GROUP")) G = atoi(e);
  if (G < 1) G = 1;
  for (int l = 0; l < nl; ++l) {
    const int lb0 = (int)((int64_t)p.B * l / nl), lb1 = (int)((int64_t)p.B * (l + 1) / nl);
    cudaStream_t st = stream;
    if (l > 0) {
      st = lanes->s[l - 1];
      RD3_CUDA_TRY(cudaStreamWaitEvent(st, lanes->fork, 0));
    }
    for (int b0 = lb0; b0 < lb1; b0 += G) {
    const int nb = (lb1 - b0 < G) ? lb1 - b0 : G;
    w.b0 = b0;
    prof_mark(st, 0);
    RD3_CUDA_TRY(cudaMemsetAsync(w.table + (size_t)b0 * p.cap, 0xFF, (size_t)nb * p.cap * 8, st));
    RD3_CUDA_TRY(cudaMemsetAsync(w.slots + (size_t)b0 * p.max_voxels * p.K, 0xFF,
                                 (size_t)nb * p.max_voxels * p.K * 4, st));
    RD3_CUDA_TRY(cudaMemsetAsync(w.flags + (size_t)b0 * p.nwords * (RD3_RANK_PACKED ? 2 : 1), 0,
                                 (size_t)nb * p.nwords * 4 * (RD3_RANK_PACKED ? 2 : 1), st));
    RD3_CUDA_TRY(cudaMemsetAsync(w.round_claims + (size_t)b0 * kMaxRounds, 0, (size_t)nb * kMaxRounds * 4,
                                 st));
    prof_mark(st, 1);
    for (int r = 0; r < p.rounds && p.N > 0; ++r) {
      const int64_t begin = (int64_t)r * p.S;
      const int64_t end = begin + p.S < p.N ? begin + p.S : p.N;
      dim3 grid((unsigned)ceil_div(end - begin, kInsSpan), nb);
      hv_insert_kernel<Src><<<grid, kInsThreads, 0, st>>>(src, g, w, begin, end, r);
    }
    prof_mark(st, 2);
    {
      int gx = (int)ceil_div(p.cap, 256 * 8);
      const int lim = (148 * 16 + nb - 1) / nb;
      if (gx > lim) gx = lim;
      if (gx < 1) gx = 1;
      hv_first_kernel<<<dim3(gx, nb), 256, 0, st>>>(w);
    }
    hv_flagscan_kernel<<<dim3(p.nchunks, nb), kScanThreads, 0, st>>>(w);
    prof_mark(st, 3);
    scan_chunks_kernel<<<nb, 1024, 0, st>>>(w.chunk_base, w.nchunks, out.voxel_num, w.max_voxels, b0);
    prof_mark(st, 4);
    if (p.N > 0)
      hv_slots_kernel<<<dim3((unsigned)ceil_div(p.ntiles, 4 * kSlotTiles), nb), 128, 0, st>>>(w, out.point2voxel);
    prof_mark(st, 5);
    hv_emit_kernel<Src><<<dim3((unsigned)ceil_div(p.max_voxels, V), nb), kEmitThreads, smem, st>>>(src, g, w, out, V);
    prof_mark(st, 6);
    prof_mark(st, 7);
    }   // groups of this lane
    if (l > 0) {
      RD3_CUDA_TRY(cudaEventRecord(lanes->join[l - 1], st));
      RD3_CUDA_TRY(cudaStreamWaitEvent(stream, lanes->join[l - 1], 0));
    }
  }
  return check_launch();
}

}  // namespace rd3
