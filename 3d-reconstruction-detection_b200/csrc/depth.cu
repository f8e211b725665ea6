// depth.cu -- depth maps -> ego-frame points (ordered compaction), and the fused
// depth -> hard voxels path (points never written to memory).
//
// Reference: projects/mmdet3d_plugin/models/backbone/reconstruction_backbone.py:285-386
// (a Python double loop over samples and cameras issuing ~15 small torch ops and
// a boolean-mask compaction each).  Here one launch covers all samples/cameras:
// per-camera calibration is staged in shared memory, every pixel is unprojected
// once per pass with exactly the reference's fp32 operation order, and the
// row-major / camera-major output order is reproduced with a ballot + popcount
// scan instead of a stream compaction primitive.
#include <math.h>

#include "hard_voxel.cuh"

namespace rd3 {

// calibration table: (B, ncam, kCalibFloats), one block per frame.  With a voxel grid
// (has_grid) it also holds the direct pixel->cell map and its error-bound constants
// (rd3_common.cuh: pixel_key_fast), derived in fp64 and rounded once.
__global__ void calib_kernel(const float *intr, const float *c2l, int ncam, int H, int W,
                             VoxelGrid g, int has_grid, CellRange rg, float *table, double *cull_cal, float zmax) {
  __shared__ float s_cal[kMaxCams * kCalibFloats];
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
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncam * kCalibFloats; i += blockDim.x)
    table[(int64_t)b * ncam * kCalibFloats + i] = s_cal[i];
}

// U1: validity flags + chunk-local exclusive prefix.  grid (nchunks, B).
__global__ void __launch_bounds__(kScanThreads)
    up_flags_kernel(DepthSource src, uint32_t *flags, int32_t *wordprefix, int32_t *chunk_total,
                    int nwords, int nchunks) {
  __shared__ float s_cal[kMaxCams * kCalibFloats];
  __shared__ int s_warp[kScanThreads / 32];
  const int b = blockIdx.y;
  src.prepare(s_cal, b);
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int word0 = blockIdx.x * kChunkWords + wv * 32;
  uint32_t my_word = 0;
#pragma unroll 4
  for (int it = 0; it < 32; ++it) {
    const int64_t i = ((int64_t)(word0 + it) << 5) + lane;
    bool valid = false;
    if (i < src.p.npix) {
      const int64_t gi = (int64_t)b * src.p.npix + i;
      const float d = __ldg(src.depth + gi);
      if (src.depth_ok(d, gi, b)) {
        float x, y, z;
        valid = !src.p.use_range || src.point(b, i, s_cal, x, y, z);
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, valid);
    if (lane == it) my_word = bal;
  }
  chunk_scan_store(my_word, s_warp, flags + (int64_t)b * nwords, wordprefix + (int64_t)b * nwords,
                   chunk_total + (int64_t)b * nchunks);
}

// U3: recompute valid pixels and write them at their ordered position.
__global__ void __launch_bounds__(256)
    up_write_kernel(DepthSource src, const uint32_t *__restrict__ flags,
                    const int32_t *__restrict__ wordprefix, const int32_t *__restrict__ chunk_base,
                    int nwords, int nchunks, float *__restrict__ out_points,
                    int32_t *__restrict__ out_pix) {
  __shared__ float s_cal[kMaxCams * kCalibFloats];
  const int b = blockIdx.y;
  src.prepare(s_cal, b);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= src.p.npix) return;
  const int64_t wi = (int64_t)b * nwords + (i >> 5);
  const uint32_t word = __ldg(flags + wi);
  const uint32_t bit = 1u << (i & 31);
  if (!(word & bit)) return;
  float x, y, z;
  src.point(b, i, s_cal, x, y, z);
  const int pos = __ldg(chunk_base + (int64_t)b * nchunks + (i >> kChunkShift)) +
                  __ldg(wordprefix + wi) + __popc(word & (bit - 1u));
  const int64_t o = (int64_t)b * src.p.npix + pos;
  out_points[o * 3 + 0] = x;
  out_points[o * 3 + 1] = y;
  out_points[o * 3 + 2] = z;
  if (out_pix) out_pix[o] = (int32_t)i;
}

struct UpPlan {
  int nwords, nchunks;
  size_t off_flags, off_prefix, off_chunk, total;
};

static UpPlan up_plan(int B, int64_t npix) {
  UpPlan p;
  p.nchunks = (int)ceil_div(npix > 0 ? npix : 1, kChunkPoints);
  p.nwords = p.nchunks * kChunkWords;
  size_t off = 0;
  p.off_flags = off; off += align_up((size_t)B * p.nwords * 4);
  p.off_prefix = off; off += align_up((size_t)B * p.nwords * 4);
  p.off_chunk = off; off += align_up((size_t)B * p.nchunks * 4);
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
  uint32_t *flags = (uint32_t *)(base + plan.off_flags);
  int32_t *prefix = (int32_t *)(base + plan.off_prefix);
  int32_t *chunk = (int32_t *)(base + plan.off_chunk);
  cudaStream_t s = (cudaStream_t)stream;
  up_flags_kernel<<<dim3(plan.nchunks, p->B), kScanThreads, 0, s>>>(src, flags, prefix, chunk,
                                                                   plan.nwords, plan.nchunks);
  scan_chunks_kernel<<<p->B, 1024, 0, s>>>(chunk, plan.nchunks, d_counts, 0x7FFFFFFF);
  up_write_kernel<<<dim3((unsigned)ceil_div(src.p.npix, 256), p->B), 256, 0, s>>>(
      src, flags, prefix, chunk, plan.nwords, plan.nchunks, out_points, out_pix);
  return check_launch();
}

size_t rd3_depth_to_voxels_workspace_bytes(const rd3_depth_params *p, int max_points,
                                           int max_voxels) {
  if (!p || p->B <= 0 || max_points <= 0 || max_voxels <= 0) return 0;
  return hv_plan((int64_t)p->ncam * p->H * p->W, p->B, max_points, max_voxels, p->W).total +
         align_up((size_t)p->B * p->ncam * kCalibFloats * 4) + align_up((size_t)p->B * p->ncam * kCullDoubles * 8);
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
  if (workspace_bytes < plan.total + cal_bytes + cull_bytes) return RD3_ERR_WORKSPACE;
  float *cal_table = (float *)((char *)workspace + plan.total);
  double *cull_cal = (double *)((char *)workspace + plan.total + cal_bytes);
  // inclusive range filter in cell units minus 0.5 (pixel_key_fast works on h = f' - 0.5).  When the voxel grid's
  // own test implies every plane of the filter (the usual case: the filter box contains the grid), the fast
  // path does not look at it at all; the exact path still applies it literally.
  src.rg.on = p->use_range && !range_filter_implied(p->range, g);
  for (int a = 0; a < 3; ++a) {
    src.rg.lo[a] = (float)(((double)p->range[a] - (double)g.lo[a]) / (double)g.vs[a] - 0.5);
    src.rg.hi[a] = (float)(((double)p->range[3 + a] - (double)g.lo[a]) / (double)g.vs[a] - 0.5);
  }
  calib_kernel<<<p->B, 128, 0, (cudaStream_t)stream>>>(intrinsics, cam2lidar, p->ncam, p->H, p->W, g, 1,
                                                       src.rg, cal_table, cull_cal, src.p.zmax);
  src.cal_table = cal_table;
  src.cull_cal = cull_cal;
  HvOut out{voxels, coors, num_points_per_voxel, voxel_mean, d_voxel_num, nullptr,
            voxel_mean ? 3 : 0};
  return hv_run(src, g, vol, plan, workspace, out, (cudaStream_t)stream);
}

}  // extern "C"
