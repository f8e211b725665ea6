// pillar.cu -- the gather side of the pillar encoders that follows hard voxelization
// (SURVEY 8(f)3): PillarFeatureNet's feature decorations and PointPillarsScatter.
//
//   PillarFeatureNet.forward   mmdetection3d/mmdet3d/models/voxel_encoders/pillar_encoder.py:104-146
//     (cluster-centre offsets, pillar-centre offsets, optional distance, padding mask; the
//      ~12 elementwise torch ops + cat + mask multiply in front of the PFN layers)
//   PointPillarsScatter        mmdetection3d/mmdet3d/models/middle_encoders/pillar_scatter.py:39-102
//     (per-sample Python loop: zero canvas, boolean mask, transpose, indexed assignment)
#include "rd3_common.cuh"

namespace rd3 {

struct PillarParams {
  int K, C, Cout;
  int with_cluster, with_center, with_distance, legacy;
  int coors_cols;        // 4: (b,z,y,x)   3: (z,y,x)
  float vx, vy, x_offset, y_offset;
};

// One warp per pillar; lanes walk the K point slots.  Every fp32 operation is the separately
// rounded one the torch expression performs; only the cluster mean is a sum whose order torch
// does not define (sequential slot order here, within 1e-6 of any order).
__global__ void __launch_bounds__(256) pillar_decorate_kernel(const float *__restrict__ voxels,
                                                              const int32_t *__restrict__ num,
                                                              const int32_t *__restrict__ coors, int64_t M,
                                                              PillarParams p, float *__restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const float *v = voxels + m * p.K * p.C;
  const int n = __ldg(num + m);
  float mx = 0.0f, my = 0.0f, mz = 0.0f;
  if (p.with_cluster) {
    // features[:, :, :3].sum(dim=1) / num_points  (:109-112): ALL K slots are summed
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int k = 0; k < p.K; ++k) {      // same sequence in every lane (broadcast loads)
      sx = __fadd_rn(sx, __ldg(v + k * p.C));
      sy = __fadd_rn(sy, __ldg(v + k * p.C + 1));
      sz = __fadd_rn(sz, __ldg(v + k * p.C + 2));
    }
    const float nf = (float)n;
    mx = __fdiv_rn(sx, nf); my = __fdiv_rn(sy, nf); mz = __fdiv_rn(sz, nf);
  }
  float cxo = 0.0f, cyo = 0.0f;
  if (p.with_center) {
    // coors[:, 3] * vx + x_offset, coors[:, 2] * vy + y_offset  (:120-125 / :128-133)
    const int32_t *c = coors + m * p.coors_cols;
    const float xi = (float)__ldg(c + p.coors_cols - 1), yi = (float)__ldg(c + p.coors_cols - 2);
    cxo = __fadd_rn(__fmul_rn(xi, p.vx), p.x_offset);
    cyo = __fadd_rn(__fmul_rn(yi, p.vy), p.y_offset);
  }
  for (int k = lane; k < p.K; k += 32) {
    const float *q = v + k * p.C;
    float *o = out + (m * p.K + k) * p.Cout;
    const bool live = k < n;             // get_paddings_indicator (:141-143): features *= mask
    const float x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
    const float fx = __fsub_rn(x, cxo), fy = __fsub_rn(y, cyo);
    // legacy=True: f_center is a VIEW of features[:, :, :2] (:127), the subtraction also
    // overwrites the raw x, y columns (and the distance below sees the shifted values)
    const bool shifted = p.with_center && p.legacy;
    const float x0 = shifted ? fx : x, y0 = shifted ? fy : y;
    int col = 0;
    for (int c = 0; c < p.C; ++c) {
      const float raw = c == 0 ? x0 : (c == 1 ? y0 : (c == 2 ? z : __ldg(q + c)));
      o[col++] = live ? raw : __fmul_rn(raw, 0.0f);
    }
    if (p.with_cluster) {
      const float a = __fsub_rn(x, mx), b = __fsub_rn(y, my), c = __fsub_rn(z, mz);
      o[col++] = live ? a : __fmul_rn(a, 0.0f);
      o[col++] = live ? b : __fmul_rn(b, 0.0f);
      o[col++] = live ? c : __fmul_rn(c, 0.0f);
    }
    if (p.with_center) {
      o[col++] = live ? fx : __fmul_rn(fx, 0.0f);
      o[col++] = live ? fy : __fmul_rn(fy, 0.0f);
    }
    if (p.with_distance) {
      // torch.norm(features[:, :, :3], 2, 2): sqrt(x^2 + y^2 + z^2)
      const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(y0, y0)), __fmul_rn(z, z)));
      o[col++] = live ? d : __fmul_rn(d, 0.0f);
    }
  }
}

// canvas (B, C, ny*nx), zero-filled by the caller-side memset; one warp per pillar, lanes over channels.
__global__ void __launch_bounds__(256) pillars_scatter_kernel(const float *__restrict__ feats,
                                                              const int32_t *__restrict__ coors, int64_t M, int C,
                                                              int coors_cols, int B, int ny, int nx,
                                                              float *__restrict__ canvas) {
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const int32_t *c = coors + m * coors_cols;
  const int b = coors_cols == 4 ? __ldg(c) : 0;
  const int y = __ldg(c + coors_cols - 2), x = __ldg(c + coors_cols - 1);
  if (b < 0 || b >= B || y < 0 || y >= ny || x < 0 || x >= nx) return;   // the reference would raise
  float *dst = canvas + (int64_t)b * C * ny * nx + (int64_t)y * nx + x;
  for (int ch = lane; ch < C; ch += 32) dst[(int64_t)ch * ny * nx] = __ldg(feats + m * C + ch);
}

// ---------------------------------------------------------------------------
// DynamicVFE.map_voxel_center_to_point (voxel_encoder.py:179-219): per point, the feature row of
// the voxel with the same (b,z,y,x).  The reference scatters voxel indices into a dense int64
// canvas of z*y*x*batch cells (5.3 GB per sample on the nuScenes grid); here the voxel rows go
// into an open-addressing table of 2M..4M int32 indices keyed by the 4 coordinates.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_row(int4 c, uint32_t mask) {
  uint32_t h = (uint32_t)c.x * 0x9E3779B1u;
  h = (h ^ (uint32_t)c.y) * 0x85EBCA77u;
  h = (h ^ (uint32_t)c.z) * 0xC2B2AE3Du;
  h = (h ^ (uint32_t)c.w) * 0x27D4EB2Fu;
  return (h ^ (h >> 15)) & mask;
}

__global__ void __launch_bounds__(256) v2p_build_kernel(const int4 *__restrict__ voxel_coors, int64_t M,
                                                        int32_t *table, uint32_t mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int4 c = __ldg(voxel_coors + i);
  uint32_t s = hash_row(c, mask);
  while (true) {
    const int32_t old = atomicCAS(table + s, -1, (int32_t)i);
    if (old == -1) return;
    const int4 o = __ldg(voxel_coors + old);
    if (o.x == c.x && o.y == c.y && o.z == c.z && o.w == c.w) {   // duplicate row: the highest index wins
      atomicMax(table + s, (int32_t)i);
      return;
    }
    s = (s + 1) & mask;
  }
}

__global__ void __launch_bounds__(256) v2p_lookup_kernel(const int4 *__restrict__ pts_coors, int64_t N,
                                                         const int4 *__restrict__ voxel_coors,
                                                         const float *__restrict__ voxel_feats, int C,
                                                         const int32_t *__restrict__ table, uint32_t mask,
                                                         float *__restrict__ out, int32_t *__restrict__ out_index) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int4 c = __ldg(pts_coors + i);
  uint32_t s = hash_row(c, mask);
  int32_t idx = 0;                         // the reference's canvas is zero-initialised: unknown -> voxel 0
  while (true) {
    const int32_t t = __ldg(table + s);
    if (t < 0) break;
    const int4 o = __ldg(voxel_coors + t);
    if (o.x == c.x && o.y == c.y && o.z == c.z && o.w == c.w) { idx = t; break; }
    s = (s + 1) & mask;
  }
  if (out_index) out_index[i] = idx;
  const float *src = voxel_feats + (int64_t)idx * C;
  float *dst = out + i * C;
  for (int f = 0; f < C; ++f) dst[f] = __ldg(src + f);
}

// ---------------------------------------------------------------------------
// GT-side occupancy of the sparse refinement:
//   SoftVoxelOccupancyVFE.forward  projects/mmdet3d_plugin/models/backbone/voxel_occupancy_encoder.py:60-99
//     p_occ = 1 - exp(-lambda * n - gamma * var),  var = mean_xyz( sum_k mask (p - mean)^2 / (n + eps) )
//   dense scatter                  projects/mmdet3d_plugin/models/backbone/sparse_refinement.py:572-587
//     map[b, z, y, x] = p_occ
// One thread per voxel; sums in slot order (torch leaves the order open: 1e-6).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) soft_occupancy_kernel(const float *__restrict__ voxels,
                                                             const int32_t *__restrict__ num, int64_t M, int K,
                                                             int C, float lambda_n, float gamma_var, float eps,
                                                             int hard, float *__restrict__ occ,
                                                             const int32_t *__restrict__ coors, int coors_cols,
                                                             int B, int Z, int Y, int X, float *__restrict__ map) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int n = __ldg(num + m);
  float p;
  if (hard) {
    p = n > 0 ? 1.0f : 0.0f;                               // HardVoxelOccupancyVFE (:35)
  } else {
    const float *v = voxels + m * K * C;
    const int live = n < K ? (n > 0 ? n : 0) : K;          // mask = arange(M) < num_points (:77-79)
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int k = 0; k < live; ++k) {
      sx = __fadd_rn(sx, __ldg(v + k * C));
      sy = __fadd_rn(sy, __ldg(v + k * C + 1));
      sz = __fadd_rn(sz, __ldg(v + k * C + 2));
    }
    const float denom = __fadd_rn((float)n, eps);          // num_points.float() + eps (:83)
    const float mx = __fdiv_rn(sx, denom), my = __fdiv_rn(sy, denom), mz = __fdiv_rn(sz, denom);
    float vx = 0.0f, vy = 0.0f, vz = 0.0f;
    for (int k = 0; k < live; ++k) {
      const float dx = __fsub_rn(__ldg(v + k * C), mx), dy = __fsub_rn(__ldg(v + k * C + 1), my),
                  dz = __fsub_rn(__ldg(v + k * C + 2), mz);
      vx = __fadd_rn(vx, __fmul_rn(dx, dx));
      vy = __fadd_rn(vy, __fmul_rn(dy, dy));
      vz = __fadd_rn(vz, __fmul_rn(dz, dz));
    }
    // (diff.pow(2).sum(dim=1) / denom).mean(dim=1)  (:88)
    const float var = __fdiv_rn(__fadd_rn(__fadd_rn(__fdiv_rn(vx, denom), __fdiv_rn(vy, denom)), __fdiv_rn(vz, denom)),
                                3.0f);
    // 1 - exp(-lambda * n - gamma * var)  (:92-94)
    const float a = __fsub_rn(__fmul_rn(-lambda_n, (float)n), __fmul_rn(gamma_var, var));
    p = __fsub_rn(1.0f, expf(a));
  }
  if (occ) occ[m] = p;
  if (map) {
    const int32_t *c = coors + m * coors_cols;
    const int b = coors_cols == 4 ? __ldg(c) : 0;
    const int z = __ldg(c + coors_cols - 3), y = __ldg(c + coors_cols - 2), x = __ldg(c + coors_cols - 1);
    if (b >= 0 && b < B && z >= 0 && z < Z && y >= 0 && y < Y && x >= 0 && x < X)
      map[(((int64_t)b * Z + z) * Y + y) * X + x] = p;
  }
}

static int v2p_log2cap(int64_t M) {
  int lg = 10;
  while (((int64_t)1 << lg) < 2 * M) ++lg;
  return lg;
}

}  // namespace rd3

using namespace rd3;

extern "C" {

int rd3_pillar_decorate(const float *voxels, const int32_t *num_points, const int32_t *coors, int64_t M,
                        int max_points, int C, int coors_cols, int with_cluster_center, int with_voxel_center,
                        int with_distance, int legacy, float vx, float vy, float x_offset, float y_offset,
                        float *out, rd3_stream_t stream) {
  if (M < 0 || max_points <= 0 || C < 3 || (coors_cols != 3 && coors_cols != 4)) return RD3_ERR_INVALID_ARGUMENT;
  if (M == 0) return RD3_OK;
  if (!voxels || !num_points || !out || (with_voxel_center && !coors)) return RD3_ERR_INVALID_ARGUMENT;
  PillarParams p;
  p.K = max_points; p.C = C;
  p.with_cluster = with_cluster_center != 0; p.with_center = with_voxel_center != 0;
  p.with_distance = with_distance != 0; p.legacy = legacy != 0;
  p.Cout = C + 3 * p.with_cluster + 2 * p.with_center + p.with_distance;
  p.coors_cols = coors_cols;
  p.vx = vx; p.vy = vy; p.x_offset = x_offset; p.y_offset = y_offset;
  pillar_decorate_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(voxels, num_points, coors, M,
                                                                                     p, out);
  return check_launch();
}

int rd3_pillars_scatter(const float *voxel_features, const int32_t *coors, int64_t M, int C, int coors_cols,
                        int batch_size, int ny, int nx, float *canvas, rd3_stream_t stream) {
  if (M < 0 || C <= 0 || batch_size <= 0 || ny <= 0 || nx <= 0 || (coors_cols != 3 && coors_cols != 4))
    return RD3_ERR_INVALID_ARGUMENT;
  if (!canvas) return RD3_ERR_INVALID_ARGUMENT;
  RD3_CUDA_TRY(cudaMemsetAsync(canvas, 0, (size_t)batch_size * C * ny * nx * 4, (cudaStream_t)stream));
  if (M == 0) return RD3_OK;
  if (!voxel_features || !coors) return RD3_ERR_INVALID_ARGUMENT;
  pillars_scatter_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(
      voxel_features, coors, M, C, coors_cols, batch_size, ny, nx, canvas);
  return check_launch();
}

int rd3_voxel_occupancy(const float *voxels, const int32_t *num_points, int64_t M, int max_points, int C,
                        int hard, float lambda_n, float gamma_var, float eps, float *occupancy,
                        const int32_t *coors, int coors_cols, int batch_size, int Z, int Y, int X,
                        float *dense_map, rd3_stream_t stream) {
  if (M < 0 || max_points <= 0 || C < 3) return RD3_ERR_INVALID_ARGUMENT;
  if (!occupancy && !dense_map) return RD3_ERR_INVALID_ARGUMENT;
  if (dense_map && (batch_size <= 0 || Z <= 0 || Y <= 0 || X <= 0 || (coors_cols != 3 && coors_cols != 4)))
    return RD3_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  if (dense_map) RD3_CUDA_TRY(cudaMemsetAsync(dense_map, 0, (size_t)batch_size * Z * Y * X * 4, s));
  if (M == 0) return RD3_OK;
  if (!num_points || (!hard && !voxels) || (dense_map && !coors)) return RD3_ERR_INVALID_ARGUMENT;
  soft_occupancy_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, s>>>(voxels, num_points, M, max_points, C, lambda_n,
                                                                   gamma_var, eps, hard, occupancy, coors,
                                                                   coors_cols, batch_size, Z, Y, X, dense_map);
  return check_launch();
}

size_t rd3_map_voxel_to_point_workspace_bytes(int64_t M) {
  if (M <= 0) return 256;
  return align_up(((size_t)1 << v2p_log2cap(M)) * 4);
}

int rd3_map_voxel_to_point(const int32_t *pts_coors, int64_t N, const int32_t *voxel_coors,
                           const float *voxel_feats, int64_t M, int C, float *out, int32_t *out_index,
                           void *workspace, size_t workspace_bytes, rd3_stream_t stream) {
  if (N < 0 || M < 0 || C <= 0 || M >= ((int64_t)1 << 30)) return RD3_ERR_INVALID_ARGUMENT;
  if (N == 0) return RD3_OK;
  if (M == 0 || !pts_coors || !voxel_coors || !voxel_feats || !out || !workspace) return RD3_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(pts_coors) & 15) || (reinterpret_cast<uintptr_t>(voxel_coors) & 15))
    return RD3_ERR_INVALID_ARGUMENT;
  const int lg = v2p_log2cap(M);
  const size_t cap = (size_t)1 << lg;
  if (workspace_bytes < cap * 4) return RD3_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  int32_t *table = (int32_t *)workspace;
  RD3_CUDA_TRY(cudaMemsetAsync(table, 0xFF, cap * 4, s));
  v2p_build_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, s>>>((const int4 *)voxel_coors, M, table,
                                                              (uint32_t)(cap - 1));
  v2p_lookup_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, s>>>((const int4 *)pts_coors, N,
                                                               (const int4 *)voxel_coors, voxel_feats, C, table,
                                                               (uint32_t)(cap - 1), out, out_index);
  return check_launch();
}

}  // extern "C"
