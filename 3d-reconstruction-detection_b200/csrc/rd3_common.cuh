// rd3_common.cuh -- shared device helpers for the sm_100a depth->voxel kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rd3_b200.h"

#ifndef __CUDA_ARCH_LIST__
#endif

namespace rd3 {

constexpr uint32_t kEmpty32 = 0xFFFFFFFFu;
constexpr unsigned long long kEmpty64 = 0xFFFFFFFFFFFFFFFFull;

// ordered first-flag scan geometry: one CTA scans kChunkWords flag words
constexpr int kChunkWords = 256;                 // 8192 points per chunk
constexpr int kChunkPoints = kChunkWords * 32;
constexpr int kChunkShift = 13;                  // log2(kChunkPoints)
constexpr int kScanThreads = 256;

void set_last_cuda_error(cudaError_t e);
int check_launch();

// Optional per-stage timing of the hard-voxel pipeline with CUDA events on the
// launching stream (rd3_profile_enable / rd3_profile_read).  No-op when disabled.
constexpr int kProfStages = 7;   // memset, insert, flags, scan, slots, emit, meta
void prof_mark(cudaStream_t stream, int boundary);   // boundary 0..kProfStages
bool prof_enabled();

// Internal side streams for the frame sub-batches of the hard-voxel pipeline (created once
// per device, non-blocking, event fork/join with the caller's stream).
constexpr int kMaxLanes = 4;
struct StreamLanes {
  cudaStream_t s[kMaxLanes - 1];
  cudaEvent_t fork, join[kMaxLanes - 1];
};
StreamLanes *get_stream_lanes();   // nullptr if creation failed
// The lanes (streams + events) of a device are shared by every call on it: host threads that
// enqueue concurrently on the SAME device take its lock for the duration of their enqueue (no device wait inside).
struct LaneLock {
  LaneLock();
  ~LaneLock();
  int dev;
};
int stream_lane_count();           // RD3_STREAMS (1..kMaxLanes), default 2

#define RD3_CUDA_TRY(expr)                       \
  do {                                           \
    cudaError_t _e = (expr);                     \
    if (_e != cudaSuccess) {                     \
      rd3::set_last_cuda_error(_e);              \
      return RD3_ERR_CUDA;                       \
    }                                            \
  } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// Voxel grid (x,y,z order, like voxel_size / coors_range of the reference).
// ---------------------------------------------------------------------------
struct VoxelGrid {
  float lo[3];
  float vs[3];
  float rvs[3];      // RN(1 / vs): only used by the conservative fast path
  float rvs_max;     // max rvs[i]
  float tolc;        // 2^-21 * (max grid + 2) + 1e-30: rounding slack of the fast path, in cells
  float thrc;        // 0.5 - tolc - 2^-20 (<= 0 disables the fast path): |h - rne(h)| must stay below
  int32_t fast_ok;   // max grid + 2 < 2^20 (else the fast path is disabled)
  int32_t grid[3];
};

// Exact unsigned division by a launch-invariant divisor (Granlund-Montgomery,
// branch-free form): q = n / d for every 32-bit n, d >= 1.
struct FastDiv {
  uint32_t d, m, l;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 1;
  while (l < 32 && (1ull << l) < d) ++l;
  f.l = l;
  f.m = d <= 1 ? 0u : (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv &f) {
  if (f.d == 1) return n;
  const uint32_t q = __umulhi(f.m, n);
  return (((n - q) >> 1) + q) >> (f.l - 1);
}

// host: grid = round((max-min)/vs) in fp32  (voxelization_cpu.cpp:121-124)
int make_grid(const float voxel_size[3], const float coors_range[6], VoxelGrid *g,
              uint64_t *volume);

// One point -> (cx,cy,cz).  Bit-exact restatement of voxelization_cpu.cpp:21-38:
// c = floor((p - min) / vs) with IEEE fp32 subtract and TRUE division, tested
// x then y then z, inside iff 0 <= c < grid.  NaN/Inf/huge fail (the reference's
// int conversion yields INT_MIN there).
__device__ __forceinline__ bool voxel_coor(float px, float py, float pz,
                                           const VoxelGrid &g, int &cx, int &cy,
                                           int &cz) {
  float f = floorf(__fdiv_rn(__fsub_rn(px, g.lo[0]), g.vs[0]));
  if (!(f >= 0.0f && f < 2147483648.0f)) return false;
  cx = (int)f;
  if (cx >= g.grid[0]) return false;
  f = floorf(__fdiv_rn(__fsub_rn(py, g.lo[1]), g.vs[1]));
  if (!(f >= 0.0f && f < 2147483648.0f)) return false;
  cy = (int)f;
  if (cy >= g.grid[1]) return false;
  f = floorf(__fdiv_rn(__fsub_rn(pz, g.lo[2]), g.vs[2]));
  if (!(f >= 0.0f && f < 2147483648.0f)) return false;
  cz = (int)f;
  if (cz >= g.grid[2]) return false;
  return true;
}

// Conservative fast path for the same predicate.  f' = RN(RN(p - lo) * RN(1/vs))
// differs from the exact quotient f = RN(RN(p - lo) / vs) by at most 2^-21 |f'|
// (five roundings of 2^-24 each, with slack); `abs_err` bounds an additional
// absolute error of the INPUT coordinate (0 for stored points; the reciprocal-
// based unprojection passes its own bound).  With
//     tol = abs_err * max(1/vs) + 2^-21 * (max grid + 2) + 1e-30
// a point whose three f' are all farther than tol from every integer has
// floor(f') == floor(f) whenever 0 <= floor(f') < grid, and is outside exactly
// when floor(f') is (for |f'| beyond the grid the error stays below |f'| - grid
// because max grid + 2 < 2^20).  NaN, huge values and exact integers fail the
// distance test and are undecided.
// Returns 1 inside, 0 outside, 2 undecided (caller must run voxel_coor()).
__device__ __forceinline__ int voxel_coor_fast(float px, float py, float pz, float abs_err,
                                               const VoxelGrid &g, int &cx, int &cy, int &cz) {
  if (!g.fast_ok) return 2;
  const float fx = __fmul_rn(__fsub_rn(px, g.lo[0]), g.rvs[0]);
  const float fy = __fmul_rn(__fsub_rn(py, g.lo[1]), g.rvs[1]);
  const float fz = __fmul_rn(__fsub_rn(pz, g.lo[2]), g.rvs[2]);
  const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
  const float tx = fminf(fx - flx, (flx + 1.0f) - fx);
  const float ty = fminf(fy - fly, (fly + 1.0f) - fy);
  const float tz = fminf(fz - flz, (flz + 1.0f) - fz);
  // fminf drops NaNs, so test them through the sum (NaN + x = NaN fails the comparison)
  const float t = fminf(tx, fminf(ty, tz));
  const float tol = fmaf(abs_err, g.rvs_max, g.tolc);
  const float sum = (fx + fy) + fz;
  const bool undecided = !(t >= tol) | !(sum == sum);
  cx = (int)flx; cy = (int)fly; cz = (int)flz;
  const bool inside = ((unsigned)cx < (unsigned)g.grid[0]) & ((unsigned)cy < (unsigned)g.grid[1]) &
                      ((unsigned)cz < (unsigned)g.grid[2]);
  return undecided ? 2 : (inside ? 1 : 0);
}

// Round-to-nearest by adding 1.5 * 2^23: for |h| < 2^22, RN(h + kMagic) = kMagic + rne(h) exactly,
// so the integer sits in the low mantissa bits and rne(h) = (h + kMagic) - kMagic costs two FADDs
// (no FRND / F2I, which run on the quarter-rate conversion pipe).
constexpr float kMagic = 12582912.0f;
constexpr int kMagicBits = 0x4B400000;

// One axis of the cell decision on h = f' - 0.5 (f' = the conservative cell coordinate):
// i = rne(h) as an int (garbage >= 2^22 or negative when |h| >= 2^22), d = h - rne(h).
// If |d| < 0.5 - tol then f_ref (within tol of f') lies strictly inside (i, i + 1), so
// floor(f_ref) == i; and when i falls outside [0, grid) so does floor(f_ref).
__device__ __forceinline__ void magic_cell(float h, int &i, float &d) {
  const float m = __fadd_rn(h, kMagic);
  i = __float_as_int(m) - kMagicBits;
  d = __fsub_rn(h, __fsub_rn(m, kMagic));
}

// Same predicate as voxel_coor_fast (abs_err = 0) in the magic-rounding form:
// h = fma(p - lo, RN(1/vs), -0.5).  Returns 1 inside (key valid), 0 outside, 2 undecided.
__device__ __forceinline__ int voxel_key_fast(float px, float py, float pz, const VoxelGrid &g,
                                              uint32_t &key) {
  const float hx = __fmaf_rn(__fsub_rn(px, g.lo[0]), g.rvs[0], -0.5f);
  const float hy = __fmaf_rn(__fsub_rn(py, g.lo[1]), g.rvs[1], -0.5f);
  const float hz = __fmaf_rn(__fsub_rn(pz, g.lo[2]), g.rvs[2], -0.5f);
  int ix, iy, iz;
  float dx, dy, dz;
  magic_cell(hx, ix, dx);
  magic_cell(hy, iy, dy);
  magic_cell(hz, iz, dz);
  const float dmax = fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz)));
  const float sum = (hx + hy) + hz;                      // fmaxf drops NaNs
  const bool decided = (dmax < g.thrc) & (sum == sum);
  const bool inside = ((unsigned)ix < (unsigned)g.grid[0]) & ((unsigned)iy < (unsigned)g.grid[1]) &
                      ((unsigned)iz < (unsigned)g.grid[2]);
  key = ((uint32_t)iz * (uint32_t)g.grid[1] + (uint32_t)iy) * (uint32_t)g.grid[0] + (uint32_t)ix;
  return decided ? (inside ? 1 : 0) : 2;
}

// linear voxel id in (z,y,x) order == lexicographic order of the output coors
__device__ __forceinline__ uint32_t voxel_key(int cx, int cy, int cz, const VoxelGrid &g) {
  return ((uint32_t)cz * (uint32_t)g.grid[1] + (uint32_t)cy) * (uint32_t)g.grid[0] + (uint32_t)cx;
}

__device__ __forceinline__ void key_to_zyx(uint32_t key, const VoxelGrid &g, int &cz, int &cy, int &cx) {
  cx = (int)(key % (uint32_t)g.grid[0]);
  uint32_t r = key / (uint32_t)g.grid[0];
  cy = (int)(r % (uint32_t)g.grid[1]);
  cz = (int)(r / (uint32_t)g.grid[1]);
}

// ---------------------------------------------------------------------------
// Per-camera calibration staged in shared memory (16 floats per camera).
//   [0]=fx [1]=fy [2]=cx [3]=cy  [4..12]=R row-major (M[:3,:3])  [13..15]=t (M[3,:3])
// ---------------------------------------------------------------------------
constexpr int kCalibFloats = 36;   // +[16..19] spare
                                   // +[20..31] direct cell map A,B,C,T-0.5 per axis  [32]=Qn [33]=Pn
constexpr int kCalDirect = 20;
constexpr int kMaxCams = 16;

struct DepthParams {
  int32_t ncam, H, W;
  int32_t HW;            // H*W
  int32_t npix;          // ncam*H*W
  int32_t use_max_depth;
  float max_depth;
  float zmax;            // min(max_depth, FLT_MAX): z <= zmax  <=>  isfinite(z) [&& z <= max_depth]
  int32_t use_masks;     // use_conf || use_sky
  int32_t use_conf;
  float conf_thresh;
  const float *conf_thresh_dev;   // per-frame thresholds on the device (overrides conf_thresh) or null
  int32_t use_sky;
  int32_t use_range;
  float range[6];
  FastDiv div_hw, div_w;   // pixel index -> (cam, v, u)
};

__device__ __forceinline__ void stage_calibration(float *s_cal, const float *intr,
                                                  const float *c2l, int ncam) {
  for (int i = threadIdx.x; i < ncam * kCalibFloats; i += blockDim.x) {
    int cam = i / kCalibFloats, j = i % kCalibFloats;
    const float *K = intr + cam * 9;
    const float *M = c2l + cam * 16;
    float v;
    if (j == 0) v = K[0];
    else if (j == 1) v = K[4];
    else if (j == 2) v = K[2];
    else if (j == 3) v = K[5];
    else if (j < 13) { int r = (j - 4) / 3, c = (j - 4) % 3; v = M[r * 4 + c]; }
    else if (j < 16) v = M[12 + (j - 13)];
    else v = 0.0f;     // [16..19] spare; [20..33] direct cell map: filled by calib_kernel when a voxel grid is known
    s_cal[i] = v;
  }
}

// Pixel -> ego-frame point.  reconstruction_backbone.py:329-334,338-340,370.
// Every fp32 operation is a separately rounded IEEE op; the only fused ops are
// the two FMAs of the 3x3 product, in the order torch-CPU's sgemm evaluates it
// (oracle/rd3_oracle.c: orc_unproject).  `z` already passed the depth/conf/sky masks.
__device__ __forceinline__ bool unproject_point(float z, int u, int v, const float *cal,
                                                const DepthParams &p, float &ox, float &oy,
                                                float &oz) {
  const float x = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, cal[2]), z), cal[0]);
  const float y = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, cal[3]), z), cal[1]);
  ox = __fadd_rn(__fmaf_rn(z, cal[6], __fmaf_rn(y, cal[5], __fmul_rn(x, cal[4]))), cal[13]);
  oy = __fadd_rn(__fmaf_rn(z, cal[9], __fmaf_rn(y, cal[8], __fmul_rn(x, cal[7]))), cal[14]);
  oz = __fadd_rn(__fmaf_rn(z, cal[12], __fmaf_rn(y, cal[11], __fmul_rn(x, cal[10]))), cal[15]);
  if (p.use_range) {                       // respoint_post_processing.py:190-195, inclusive
    if (!(ox >= p.range[0] && ox <= p.range[3] && oy >= p.range[1] && oy <= p.range[4] &&
          oz >= p.range[2] && oz <= p.range[5]))
      return false;
  }
  return true;
}

// Direct pixel -> cell map used ONLY to decide the voxel cell (fused path).
// With per-camera, per-axis constants (fp64-derived, rounded once to fp32)
//   A = R[a][0] / (fx vs_a)      B = R[a][1] / (fy vs_a)
//   C = (R[a][2] - R[a][0] cx/fx - R[a][1] cy/fy) / vs_a      T = (t_a - lo_a) / vs_a
// the cell coordinate is  f'_a = fma(z, fma(A, u, fma(B, v, C)), T)   (3 FMAs per axis).
// Against the reference chain f_ref = RN(RN(o_ref - lo)/vs) (o_ref = its fp32 unprojection):
//   |o_ref - o*| <= 7 eps S*            (3 roundings in x,y; 4 in the 3x3 product)
//   |f_ref - f*| <= 7 eps S* / vs + 2 eps |f|
//   |f'   - f*| <= eps (3 z D + |T| + |f'|),   D = |A|(W-1) + |B|(H-1) + |C|
// with eps = 2^-24 and S*/vs <= z Q + P1 (Q = (|R0| ex + |R1| ey + |R2|)/vs, ex/ey = ray
// extents, P1 = |t|/vs).  Hence |f' - f_ref| <= eps (z (3D + 10Q) + |T| + 10 P1 + 3 |lo|/vs)
// and   tol = 2^-23 (z Qc + Pc)   (2x slack; Qc, Pc = maxima over the three axes, Pc also
// covers the rounding of the range-filter limits and of Th = T - 0.5).  If every f'_a is at
// least tol away from the nearest integer then floor(f'_a) == floor(f_ref_a): the cell and the
// in/out verdict are the reference's.  Everything else is redone exactly.
struct CellRange {          // inclusive range filter in cell units minus 0.5 (h units), per axis
  float lo[3], hi[3];
  int32_t on;
};

// row part of the map, shared by the pixels of one image row: t_a = fma(B_a, v, C_a)
__device__ __forceinline__ void pixel_cell_row(float vf, const float *cal, float &tx, float &ty,
                                               float &tz) {
  const float *k = cal + kCalDirect;
  tx = __fmaf_rn(k[1], vf, k[2]);
  ty = __fmaf_rn(k[5], vf, k[6]);
  tz = __fmaf_rn(k[9], vf, k[10]);
}

// One pixel of the direct map in the magic-rounding form.  k[a*4+3] holds Th_a = RN(T_a - 0.5), so
// h_a = fma(z, fma(A_a, u, row_a), Th_a) = f'_a - 0.5 (the extra rounding of Th is covered by Pc);
// thr = fma(z, Qn, Pn) <= 0.5 - tol with Qn = -2^-23 Qc, Pn = 0.5 - 2^-23 Pc - 2^-20 (both rounded
// down).  The pixel is decided iff max_a |h_a - rne(h_a)| < thr (NaN compares false); then the cell
// is rne(h_a) and the in/out verdict is the reference's (magic_cell).  Returns 1 inside (key valid),
// 0 outside, 2 undecided.  With the range filter on, a pixel inside the grid must also be surely
// inside / outside the filter box (limits in h units), else it is undecided.
__device__ __forceinline__ int pixel_key_fast(float z, float uf, float rtx, float rty, float rtz,
                                              const float *cal, const VoxelGrid &g,
                                              const CellRange &rg, uint32_t &key) {
  const float *k = cal + kCalDirect;
  const float hx = __fmaf_rn(z, __fmaf_rn(k[0], uf, rtx), k[3]);
  const float hy = __fmaf_rn(z, __fmaf_rn(k[4], uf, rty), k[7]);
  const float hz = __fmaf_rn(z, __fmaf_rn(k[8], uf, rtz), k[11]);
  const float thr = __fmaf_rn(z, k[12], k[13]);
  int ix, iy, iz;
  float dx, dy, dz;
  magic_cell(hx, ix, dx);
  magic_cell(hy, iy, dy);
  magic_cell(hz, iz, dz);
  // z is finite here (validity mask) and non-finite constants make thr NaN / -inf: a NaN d_a can
  // only come from an infinite h_a, whose garbage cell index is outside the grid like the point
  const float dmax = fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz)));
  const bool decided = dmax < thr;
  const bool inside = ((unsigned)ix < (unsigned)g.grid[0]) & ((unsigned)iy < (unsigned)g.grid[1]) &
                      ((unsigned)iz < (unsigned)g.grid[2]);
  key = ((uint32_t)iz * (uint32_t)g.grid[1] + (uint32_t)iy) * (uint32_t)g.grid[0] + (uint32_t)ix;
  int r = decided ? (inside ? 1 : 0) : 2;
  if (rg.on && r == 1) {
    const float tol = 0.5f - thr;
    const bool in_sure = hx - rg.lo[0] >= tol && rg.hi[0] - hx >= tol && hy - rg.lo[1] >= tol &&
                         rg.hi[1] - hy >= tol && hz - rg.lo[2] >= tol && rg.hi[2] - hz >= tol;
    const bool out_sure = rg.lo[0] - hx > tol || hx - rg.hi[0] > tol || rg.lo[1] - hy > tol ||
                          hy - rg.hi[1] > tol || rg.lo[2] - hz > tol || hz - rg.hi[2] > tol;
    r = in_sure ? 1 : (out_sure ? 0 : 2);
  }
  return r;
}

// ---------------------------------------------------------------------------
// TMA bulk copy global -> shared (cp.async.bulk, 1-D) completing on an mbarrier: one thread issues
// it, nobody executes a load / store loop.  dst, src 16-byte aligned, bytes a multiple of 16.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *smem_bar) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem_bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(gmem_src), "r"(bytes), "r"(bar)
               : "memory");
}
// a plain mbarrier used as a one-shot flag inside a CTA: init by one thread (then a __syncthreads()), one arrive
// (release) by the producer after its shared-memory stores, tma_wait (acquire, phase 0) by the consumers
__device__ __forceinline__ void mbar_init(uint64_t *smem_bar, uint32_t count) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem_bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *smem_bar) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem_bar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// every consumer thread, after a __syncthreads() that follows tma_load_1d (phase 0 of a fresh barrier)
__device__ __forceinline__ void tma_wait(uint64_t *smem_bar) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem_bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(bar)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// warp / block scan helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_inclusive_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}

}  // namespace rd3
