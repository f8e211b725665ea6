// depth.cu -- depth maps -> ego-frame points (ordered compaction), and the fused
// depth -> hard voxels path (points never written to memory).
//
// Reference: projects/mmdet3d_plugin/models/backbone/reconstruction_backbone.py:285-386
// (a Python double loop over samples and cameras issuing ~15 small torch ops and
// a boolean-mask compaction each).  Here one launch covers all samples/cameras:
// per-camera calibration is staged in shared memory, every pixel is unprojected
// once with exactly the reference's fp32 operation order, and the row-major /
// camera-major output order is reproduced by a single-pass ordered compaction
// (chunk counts chained by decoupled look-back).
#include <math.h>

#include "hard_voxel.cuh"

namespace rd3 {

// calibration table: (B, ncam, kCalibFloats), one block per frame.  With a voxel grid
// (has_grid) it also holds the direct pixel->cell map and its error-bound constants
// (rd3_common.cuh: pixel_key_fast), derived in fp64 and rounded once.
__global__ void calib_kernel(const float *intr, const float *c2l, int ncam, int H, int W,
                             VoxelGrid g, int has_grid, CellRange rg, float *table, double *cull_cal, float zmax,
                             float *cull_planes, int cbshift) {
  __shared__ float s_cal[kMaxCams * kCalibFloats];
  __shared__ double s_cc[kMaxCams][kCullDoubles];
  const int b = blockIdx.x;
  const float *Kb = intr + (int64_t)b * ncam * 9;
  const float *Mb = c2l + (int64_t)b * ncam * 16;
  stage_calibration(s_cal, Kb, Mb, ncam);
  __syncthreads();
  if (has_grid && threadIdx.x < ncam) {
    const int cam = threadIdx.x;
    const float *K = Kb + cam * 9, *M = Mb + cam * 16;
    float *k = s_cal + cam * kCalibFloats + kCalDirect;
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double ex = fmax(fabs(cx), fabs((double)(W - 1) - cx)) / fabs(fx);
    const double ey = fmax(fabs(cy), fabs((double)(H - 1) - cy)) / fabs(fy);
    double Qc = 0.0, Pc = 0.0;
    for (int a = 0; a < 3; ++a) {
      const double rv = 1.0 / (double)g.vs[a];
      const double r0 = M[a * 4 + 0], r1 = M[a * 4 + 1], r2 = M[a * 4 + 2], t = M[12 + a];
      const float A = (float)(r0 / fx * rv), Bc = (float)(r1 / fy * rv);
      const float C = (float)((r2 - r0 * cx / fx - r1 * cy / fy) * rv);
      const float T = (float)((t - (double)g.lo[a]) * rv - 0.5);      // h = f' - 0.5 (pixel_key_fast)
      k[a * 4 + 0] = A; k[a * 4 + 1] = Bc; k[a * 4 + 2] = C; k[a * 4 + 3] = T;
      const double D = fabs((double)A) * (W - 1) + fabs((double)Bc) * (H - 1) + fabs((double)C);
      const double Q = (fabs(r0) * ex + fabs(r1) * ey + fabs(r2)) * rv;
      const double P1 = fabs(t) * rv;
      double P = fabs((double)T) + 2.0 + 10.0 * P1 + 3.0 * fabs((double)g.lo[a]) * rv;
      if (rg.on) P += fabs((double)rg.lo[a]) + fabs((double)rg.hi[a]) + 1.0;
      Qc = fmax(Qc, 3.0 * D + 10.0 * Q);
      Pc = fmax(Pc, P);
    }
    // thr = fma(z, Qn, Pn) <= 0.5 - 2^-23 (z Qc + Pc): constants rounded towards -inf, 2^-20 covers
    // the rounding of the fma itself.  Non-finite inputs give NaN / -inf: nothing is decided.
    const double eps2 = 1.1920928955078125e-7;   // 2^-23
    k[12] = __double2float_rd(-(Qc * 1.000001) * eps2);
    k[13] = __double2float_rd(0.5 - (Pc * 1.000001) * eps2 - 9.5367431640625e-7);
    if (cull_cal) {
      // culling test of the lookup pass (hard_voxel.cuh: cull_block): inverse of the direct cell map Mc = [A B C],
      // T = Th + 0.5 and the margin 1 + 2 tolmax, in fp64.  A singular / non-finite map keeps all its blocks (ok = 0).
      double *cc = cull_cal + ((int64_t)b * ncam + cam) * kCullDoubles;
      double m[9], T[3];
      for (int a = 0; a < 3; ++a) {
        m[a * 3 + 0] = k[a * 4 + 0]; m[a * 3 + 1] = k[a * 4 + 1]; m[a * 3 + 2] = k[a * 4 + 2];
        T[a] = (double)k[a * 4 + 3] + 0.5;
      }
      const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
      const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
      double scale = 0.0;
      for (int i = 0; i < 9; ++i) scale = fmax(scale, fabs(m[i]));
      const double id = 1.0 / det;
      double inv[9];
      inv[0] = c00 * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
      inv[3] = c01 * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
      inv[6] = c02 * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
      // tol(z) = 0.5 - thr(z) >= the proven bound; largest at the largest valid depth
      const double tolmax = 0.5 - ((double)zmax * (double)k[12] + (double)k[13]);
      bool ok = fabs(det) > 1e-12 * scale * scale * scale && isfinite(id) && tolmax >= 0.0 && tolmax < 1e6;
      for (int i = 0; i < 9; ++i) ok = ok && isfinite(inv[i]);
      for (int a = 0; a < 3; ++a) ok = ok && isfinite(T[a]);
      for (int i = 0; i < 9; ++i) cc[i] = inv[i];
      for (int a = 0; a < 3; ++a) cc[9 + a] = T[a];
      cc[12] = 1.0 + 2.0 * tolmax;
      cc[13] = ok ? 1.0 : 0.0;
      cc[14] = cc[15] = 0.0;
      for (int i = 0; i < kCullDoubles; ++i) s_cc[cam][i] = cc[i];
    }
  }
  __syncthreads();
  if (has_grid && cull_cal && cull_planes) {
    // the five wedge planes of every (camera, column block): one plane per thread and step
    const int nblk = ((W - 1) >> cbshift) + 1;
    for (int i = threadIdx.x; i < ncam * nblk * 5; i += blockDim.x) {
      const int k = i % 5, pair = i / 5;
      const int cam = pair / nblk, blk = pair - cam * nblk;
      cull_plane_of(&s_cc[cam][0], &s_cc[cam][9], s_cc[cam][12], k, blk, cbshift, W, H, g,
                    cull_planes + (((int64_t)b * ncam + cam) * nblk + blk) * 20 + k * 4);
    }
  }
  for (int i = threadIdx.x; i < ncam * kCalibFloats; i += blockDim.x)
    table[(int64_t)b * ncam * kCalibFloats + i] = s_cal[i];
}

// Ordered compaction of the valid pixels in ONE pass over the inputs (decoupled look-back; no flag pass, no re-read).
// A CTA takes the next chunk of kUpChunk pixels of its frame from a ticket counter (so a chunk is always started
// after every chunk before it), works out the validity of its pixels (16 per thread: four groups of 4 consecutive
// pixels, one 16-byte load each), publishes its count, adds up the counts of the chunks before it (spinning only on
// chunks that are already running), and writes its points at their ordered positions.  state[b][c] =
// {flag:32 | value:32}: flag 1 = chunk total, 2 = inclusive prefix.
constexpr int kUpChunk = 4096;
constexpr int kUpThreads = 256;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kUpThreads, 4)
    up_single_kernel(DepthSource src, unsigned long long *state, int32_t *ticket, int nchunks,
                     float *__restrict__ out_points, int32_t *__restrict__ out_pix, int32_t *__restrict__ counts) {
  __shared__ float s_cal[kMaxCams * kCalibFloats];
  __shared__ int s_chunk, s_base;
  __shared__ int s_wsum[kUpThreads / 32][4];
  extern __shared__ __align__(16) unsigned char s_dyn[];       // per warp 4 x 128 points (+ 4 x 128 pixel indices)
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
  if (tid == 0) s_chunk = atomicAdd(ticket + b, 1);
  src.stage(s_cal, b);
  __syncthreads();
  const int chunk = s_chunk;
  const int64_t fbase = (int64_t)b * src.p.npix;
  const bool vec = src.vec_ok && (!src.p.use_masks || src.mask_vec_ok);
  const float thr = src.p.use_conf ? (src.p.conf_thresh_dev ? __ldg(src.p.conf_thresh_dev + b) : src.p.conf_thresh) : 0.0f;

  // validity of the thread's 16 pixels: group g covers pixels chunk*kUpChunk + g*1024 + tid*4 .. +3
  float z[4][4];
  unsigned valid[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int64_t i0 = (int64_t)chunk * kUpChunk + g * 1024 + tid * 4;
    unsigned m = 0;
    if (vec && i0 + 4 <= src.p.npix) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(src.depth + fbase + i0));
      z[g][0] = t.x; z[g][1] = t.y; z[g][2] = t.z; z[g][3] = t.w;
#pragma unroll
      for (int q = 0; q < 4; ++q) m |= ((z[g][q] > 0.0f) & (z[g][q] <= src.p.zmax)) ? (1u << q) : 0u;
      if (m && src.p.use_conf) {
        const float4 cf = __ldg(reinterpret_cast<const float4 *>(src.conf + fbase + i0));
        m &= (cf.x >= thr ? 1u : 0u) | (cf.y >= thr ? 2u : 0u) | (cf.z >= thr ? 4u : 0u) | (cf.w >= thr ? 8u : 0u);
      }
      if (m && src.p.use_sky) {
        if (src.sky) {
          const uint32_t sb = __ldg(reinterpret_cast<const uint32_t *>(src.sky + fbase + i0));
          m &= ((sb & 0xFFu) ? 0u : 1u) | ((sb & 0xFF00u) ? 0u : 2u) | ((sb & 0xFF0000u) ? 0u : 4u) | ((sb & 0xFF000000u) ? 0u : 8u);
        } else {
          const float4 sp = __ldg(reinterpret_cast<const float4 *>(src.sky_prob + fbase + i0));
          m &= (sp.x >= src.sky_thr ? 0u : 1u) | (sp.y >= src.sky_thr ? 0u : 2u) | (sp.z >= src.sky_thr ? 0u : 4u) |
               (sp.w >= src.sky_thr ? 0u : 8u);
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        z[g][q] = 0.0f;
        if (i0 + q < src.p.npix) {
          z[g][q] = __ldg(src.depth + fbase + i0 + q);
          if (src.depth_ok(z[g][q], fbase + i0 + q, b)) m |= 1u << q;
        }
      }
    }
    if (m && src.p.use_range) {                    // the inclusive range filter looks at the transformed point
      uint32_t cam0 = 0, v0 = 0, u0 = 0;
      if (vec) src.pixel_cvu((uint32_t)i0, cam0, v0, u0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if ((m >> q) & 1u) {
          uint32_t cam = cam0, v = v0, u = u0 + q;
          if (!vec) src.pixel_cvu((uint32_t)(i0 + q), cam, v, u);
          float x, y, zz;
          if (!unproject_point(z[g][q], (int)u, (int)v, s_cal + cam * kCalibFloats, src.p, x, y, zz)) m &= ~(1u << q);
        }
      }
    }
    valid[g] = m;
  }
  // ordered positions inside the chunk: groups in order, lanes in order inside a group, pixels in order inside a lane
  int excl[4], wtot[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = __popc(valid[g]);
    const int inc = warp_inclusive_scan(c);
    excl[g] = inc - c;
    wtot[g] = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 31) s_wsum[wv][g] = inc;
  }
  __syncthreads();
  int gbase[4], total = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    int before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < kUpThreads / 32; ++k) {
      const int t = s_wsum[k][g];
      if (k < wv) before += t;
      all += t;
    }
    gbase[g] = total + before + excl[g];
    total += all;
  }
  // The chunk's count is published at once; its points are worked out into shared memory (a warp's points of one
  // group are consecutive in the output) while the chunks before it publish theirs; only then does warp 0 look
  // back (32 predecessors per step -- they were started earlier, so the wait is finite), and the points leave with
  // fully coalesced 4-byte stores (lane-strided 12-byte records would touch 8x the sectors).
  unsigned long long *st = state + (int64_t)b * nchunks;
  if (tid == 0 && chunk > 0) st_release_u64(st + chunk, (1ull << 32) | (unsigned)total);
  float *sp = reinterpret_cast<float *>(s_dyn) + wv * (4 * 128 * 3);
  int32_t *sx = reinterpret_cast<int32_t *>(s_dyn) + (kUpThreads / 32) * (4 * 128 * 3) + wv * (4 * 128);
  int wfirst[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int64_t i0 = (int64_t)chunk * kUpChunk + g * 1024 + tid * 4;
    wfirst[g] = __shfl_sync(0xffffffffu, gbase[g], 0);                 // ordered position of the warp's first point
    int lp = gbase[g] - wfirst[g];
    // with 16-byte loads the group's 4 pixels lie in one image row: one index -> (camera, row, column) division
    uint32_t cam0 = 0, v0 = 0, u0 = 0;
    if (vec && valid[g]) src.pixel_cvu((uint32_t)i0, cam0, v0, u0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if ((valid[g] >> q) & 1u) {
        uint32_t cam = cam0, v = v0, u = u0 + q;
        if (!vec) src.pixel_cvu((uint32_t)(i0 + q), cam, v, u);
        float x, y, zz;
        unproject_point(z[g][q], (int)u, (int)v, s_cal + cam * kCalibFloats, src.p, x, y, zz);   // (the range verdict is known)
        sp[g * 384 + lp * 3 + 0] = x;
        sp[g * 384 + lp * 3 + 1] = y;
        sp[g * 384 + lp * 3 + 2] = zz;
        if (out_pix) sx[g * 128 + lp] = (int32_t)(i0 + q);
        ++lp;
      }
    }
  }
  if (wv == 0) {
    int base = 0;
    if (chunk > 0) {
      int c = chunk - 1;                                   // newest chunk of the window; lane l looks at chunk c - l
      while (true) {
        const int ci = c - lane;
        const unsigned long long v = ci >= 0 ? ld_acquire_u64(st + ci) : (2ull << 32);   // before chunk 0: prefix 0
        const unsigned f = (unsigned)(v >> 32);
        const unsigned ready = __ballot_sync(0xffffffffu, f != 0u);
        const unsigned incl = __ballot_sync(0xffffffffu, f == 2u);
        const int k = incl ? __ffs(incl) - 1 : 31;         // the window counts up to its first inclusive prefix
        const unsigned need = k == 31 ? 0xffffffffu : ((2u << k) - 1u);
        if ((ready & need) != need) { __nanosleep(20); continue; }
        int contrib = lane <= k ? (int)(unsigned)v : 0;
        for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
        base += contrib;
        if (incl) break;
        c -= 32;
      }
    }
    if (lane == 0) {
      st_release_u64(st + chunk, (2ull << 32) | (unsigned)(base + total));
      if (chunk == nchunks - 1) counts[b] = base + total;
      s_base = base;
    }
  }
  __syncthreads();
  const int base = s_base;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int cnt = wtot[g];
    const int64_t o = fbase + base + wfirst[g];
    for (int e = lane; e < cnt * 3; e += 32) out_points[o * 3 + e] = sp[g * 384 + e];
    if (out_pix)
      for (int e = lane; e < cnt; e += 32) out_pix[o + e] = sx[g * 128 + e];
  }
}

struct UpPlan {
  int nchunks;
  size_t off_state, off_ticket, total;
};

static UpPlan up_plan(int B, int64_t npix) {
  UpPlan p;
  p.nchunks = (int)ceil_div(npix > 0 ? npix : 1, kUpChunk);
  size_t off = 0;
  p.off_state = off; off += align_up((size_t)B * p.nchunks * 8);     // [state | ticket] are cleared with one memset
  p.off_ticket = off; off += align_up((size_t)B * 4);
  p.total = off;
  return p;
}

// fp32 cell coordinate of p on axis a exactly as voxel_coor() computes it on the device: IEEE subtract, TRUE
// division (volatile: no contraction, no double evaluation on the host)
static float host_cellf(float p, const VoxelGrid &g, int a) {
  volatile float d = p - g.lo[a];
  volatile float q = d / g.vs[a];
  return q;
}

// Does "the point is inside the voxel grid" imply "the point passes the inclusive range filter"
// (respoint_post_processing.py:190-195) on all six planes?  q(p) = RN(RN(p - lo) / vs) is a monotone
// non-decreasing function of p and the device accepts a point iff 0 <= floor(q) < grid, so
//   upper plane: every p > hi has q(p) >= q(nextafter(hi)); if that is >= grid, p is outside the grid
//   lower plane: every p < lo has q(p) <= q(nextbefore(lo)); if that is < 0 (and not -0, whose floor passes
//                the device's ">= 0" test), p is outside the grid
// NaN anywhere: not implied.  When this returns true the fast cell decision ignores the filter; the exact path
// (unproject_point) still applies it literally.
static bool range_filter_implied(const float range[6], const VoxelGrid &g) {
  for (int a = 0; a < 3; ++a) {
    const float lo = range[a], hi = range[3 + a];
    if (!(lo == lo) || !(hi == hi)) return false;
    if (hi < 3.0e38f) {
      const float q = host_cellf(nextafterf(hi, INFINITY), g, a);
      if (!(q == q) || !((double)q >= (double)g.grid[a])) return false;   // a point above hi can still be in the grid
    }
    if (lo > -3.0e38f) {
      const float q = host_cellf(nextafterf(lo, -INFINITY), g, a);
      if (!(q == q) || q >= 0.0f) return false;                           // q >= 0 is true for -0.0f as well
    }
  }
  return true;
}

static int make_depth_source(const float *depth, const float *intrinsics, const float *cam2lidar,
                             const float *conf, const uint8_t *sky, const rd3_depth_params *p,
                             DepthSource *src) {
  if (!p || !depth || !intrinsics || !cam2lidar) return RD3_ERR_INVALID_ARGUMENT;
  if (p->B <= 0 || p->ncam <= 0 || p->H <= 0 || p->W <= 0) return RD3_ERR_INVALID_ARGUMENT;
  if (p->ncam > kMaxCams || p->B > 65535) return RD3_ERR_UNSUPPORTED;
  const int64_t npix = (int64_t)p->ncam * p->H * p->W;
  if (npix >= ((int64_t)1 << 30)) return RD3_ERR_UNSUPPORTED;
  src->depth = depth;
  src->conf = conf;
  src->sky = sky;
  src->sky_prob = sky ? nullptr : p->sky_prob;
  src->sky_thr = p->sky_prob_thresh;
  src->intr = intrinsics;
  src->c2l = cam2lidar;
  src->cal_table = nullptr;
  src->cull_cal = nullptr;
  src->cull_planes = nullptr;
  src->rg.on = 0;
  DepthParams &d = src->p;
  d.ncam = p->ncam; d.H = p->H; d.W = p->W; d.HW = p->H * p->W; d.npix = (int32_t)npix;
  d.use_max_depth = p->use_max_depth; d.max_depth = p->max_depth;
  d.use_conf = conf != nullptr; d.conf_thresh = p->conf_thresh;
  d.conf_thresh_dev = p->conf_thresh_dev;
  d.use_sky = sky != nullptr || src->sky_prob != nullptr;
  d.use_masks = d.use_conf || d.use_sky;
  d.zmax = 3.402823466e+38f;
  if (p->use_max_depth && p->max_depth < d.zmax) d.zmax = p->max_depth;   // NaN max_depth: comparison false
  if (p->use_max_depth && !(p->max_depth == p->max_depth)) d.zmax = -1.0f; // z <= NaN is never true
  d.use_range = p->use_range;
  for (int i = 0; i < 6; ++i) d.range[i] = p->range[i];
  d.div_hw = make_fastdiv((uint32_t)d.HW);
  d.div_w = make_fastdiv((uint32_t)d.W);
  src->vec_ok = ((p->W & 3) == 0) && ((reinterpret_cast<uintptr_t>(depth) & 15) == 0);
  src->mask_vec_ok = ((p->W & 3) == 0) && (!conf || (reinterpret_cast<uintptr_t>(conf) & 15) == 0) &&
                     (!sky || (reinterpret_cast<uintptr_t>(sky) & 3) == 0) &&
                     (!src->sky_prob || (reinterpret_cast<uintptr_t>(src->sky_prob) & 15) == 0);
  src->div_h = make_fastdiv((uint32_t)p->H);
  src->cbshift = 7;                                  // culling: blocks of >= 128 image columns, at most 32 per row
  while (((p->W - 1) >> src->cbshift) >= 32) ++src->cbshift;
  return RD3_OK;
}

}  // namespace rd3

using namespace rd3;

extern "C" {

size_t rd3_unproject_workspace_bytes(const rd3_depth_params *p) {
  if (!p || p->B <= 0) return 0;
  return up_plan(p->B, (int64_t)p->ncam * p->H * p->W).total;
}

int rd3_unproject(const float *depth, const float *intrinsics, const float *cam2lidar,
                  const float *conf, const uint8_t *sky, const rd3_depth_params *p,
                  float *out_points, int32_t *out_pix, int32_t *d_counts, void *workspace,
                  size_t workspace_bytes, rd3_stream_t stream) {
  DepthSource src;
  int st = make_depth_source(depth, intrinsics, cam2lidar, conf, sky, p, &src);
  if (st != RD3_OK) return st;
  if (!out_points || !d_counts || !workspace) return RD3_ERR_INVALID_ARGUMENT;
  const UpPlan plan = up_plan(p->B, src.p.npix);
  if (workspace_bytes < plan.total) return RD3_ERR_WORKSPACE;
  char *base = (char *)workspace;
  unsigned long long *state = (unsigned long long *)(base + plan.off_state);
  int32_t *ticket = (int32_t *)(base + plan.off_ticket);
  cudaStream_t s = (cudaStream_t)stream;
  RD3_CUDA_TRY(cudaMemsetAsync(base, 0, plan.total, s));
  const size_t smem = (size_t)(kUpThreads / 32) * 4 * 128 * (out_pix ? 16 : 12);
  RD3_CUDA_TRY(cudaFuncSetAttribute(up_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  up_single_kernel<<<dim3(plan.nchunks, p->B), kUpThreads, smem, s>>>(src, state, ticket, plan.nchunks, out_points,
                                                                     out_pix, d_counts);
  return check_launch();
}

size_t rd3_depth_to_voxels_workspace_bytes(const rd3_depth_params *p, int max_points,
                                           int max_voxels) {
  if (!p || p->B <= 0 || max_points <= 0 || max_voxels <= 0) return 0;
  return hv_plan((int64_t)p->ncam * p->H * p->W, p->B, max_points, max_voxels, p->W).total +
         align_up((size_t)p->B * p->ncam * kCalibFloats * 4) + align_up((size_t)p->B * p->ncam * kCullDoubles * 8) +
         align_up((size_t)p->B * p->ncam * 32 * 20 * 4);
}

int rd3_depth_to_voxels(const float *depth, const float *intrinsics, const float *cam2lidar,
                        const float *conf, const uint8_t *sky, const rd3_depth_params *p,
                        const float voxel_size[3], const float coors_range[6], int max_points,
                        int max_voxels, float *voxels, int32_t *coors,
                        int32_t *num_points_per_voxel, float *voxel_mean, int32_t *d_voxel_num,
                        void *workspace, size_t workspace_bytes, rd3_stream_t stream) {
  DepthSource src;
  int st = make_depth_source(depth, intrinsics, cam2lidar, conf, sky, p, &src);
  if (st != RD3_OK) return st;
  if (max_points <= 0 || max_voxels <= 0 || !voxel_size || !coors_range || !coors ||
      !num_points_per_voxel || !d_voxel_num || !workspace)
    return RD3_ERR_INVALID_ARGUMENT;
  if (!voxels && !voxel_mean) return RD3_ERR_INVALID_ARGUMENT;   // voxels may be skipped only for the mean
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return RD3_ERR_INVALID_ARGUMENT;   // 256-bit loads, TMA
  VoxelGrid g;
  uint64_t vol;
  st = make_grid(voxel_size, coors_range, &g, &vol);
  if (st != RD3_OK) return st;
  const HvPlan plan = hv_plan(src.p.npix, p->B, max_points, max_voxels, p->W);
  const size_t cal_bytes = align_up((size_t)p->B * p->ncam * kCalibFloats * 4);
  const size_t cull_bytes = align_up((size_t)p->B * p->ncam * kCullDoubles * 8);
  const size_t planes_bytes = align_up((size_t)p->B * p->ncam * 32 * 20 * 4);     // at most 32 column blocks per camera
  if (workspace_bytes < plan.total + cal_bytes + cull_bytes + planes_bytes) return RD3_ERR_WORKSPACE;
  float *cal_table = (float *)((char *)workspace + plan.total);
  double *cull_cal = (double *)((char *)workspace + plan.total + cal_bytes);
  float *cull_planes = (float *)((char *)workspace + plan.total + cal_bytes + cull_bytes);
  // inclusive range filter in cell units minus 0.5 (pixel_key_fast works on h = f' - 0.5).  When the voxel grid's
  // own test implies every plane of the filter (the usual case: the filter box contains the grid), the fast
  // path does not look at it at all; the exact path still applies it literally.
  src.rg.on = p->use_range && !range_filter_implied(p->range, g);
  for (int a = 0; a < 3; ++a) {
    src.rg.lo[a] = (float)(((double)p->range[a] - (double)g.lo[a]) / (double)g.vs[a] - 0.5);
    src.rg.hi[a] = (float)(((double)p->range[3 + a] - (double)g.lo[a]) / (double)g.vs[a] - 0.5);
  }
  calib_kernel<<<p->B, 128, 0, (cudaStream_t)stream>>>(intrinsics, cam2lidar, p->ncam, p->H, p->W, g, 1,
                                                       src.rg, cal_table, cull_cal, src.p.zmax, cull_planes, src.cbshift);
  src.cal_table = cal_table;
  src.cull_cal = cull_cal;
  src.cull_planes = cull_planes;
  HvOut out{voxels, coors, num_points_per_voxel, voxel_mean, d_voxel_num, nullptr,
            voxel_mean ? 3 : 0};
  return hv_run(src, g, vol, plan, workspace, out, (cudaStream_t)stream);
}

}  // extern "C"
