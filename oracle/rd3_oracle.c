/*
 * rd3_oracle.c -- CPU restatement of the reference's depth->voxel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and
 * there only as the checker.  The product package never links or calls it.
 *
 * Parity status: PINNED.  orc_hard_voxelize / orc_dynamic_voxelize are checked
 * (tests/test_oracle.py) against
 *   - the known-answer vector of the reference's own test
 *     mmdetection3d/tests/test_models/test_voxel_encoder/test_voxel_generator.py:7-22
 *   - the reference's own C++ CPU op compiled unmodified (oracle/_ref, see
 *     oracle/build_ref.py) on random and adversarial inputs, bit for bit.
 * orc_unproject has no reference test ("pinned by source"): it is checked
 * against a torch-CPU transliteration of
 * projects/mmdet3d_plugin/models/backbone/reconstruction_backbone.py:305-386
 * (oracle/torch_restatement.py) and against committed fixtures in tests/golden.
 *
 * Build: gcc -O2 -fPIC -shared -mfma -ffp-contract=off  (see oracle/build.py).
 * -ffp-contract=off matters: every fp32 operation below is meant to round
 * exactly where it is written; the only fused operations are the explicit
 * fmaf() calls in the 3x3 transform.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* grid size: voxelization_cpu.cpp:121-124 (round of an fp32 quotient)       */
/* ------------------------------------------------------------------------ */
static void orc_grid_size_(const float *voxel_size, const float *coors_range,
                           int *grid) {
  for (int i = 0; i < 3; ++i) {
    float q = (coors_range[3 + i] - coors_range[i]) / voxel_size[i];
    grid[i] = (int)roundf(q);
  }
}

void orc_grid_size(const float *voxel_size, const float *coors_range,
                   int *grid) {
  orc_grid_size_(voxel_size, coors_range, grid);
}

/* One point's voxel coordinate, voxelization_cpu.cpp:21-38.
 * Returns 1 and fills czyx (z,y,x order, :30) when the point is inside
 * [min, min + grid*vs) on x, then y, then z; returns 0 otherwise.
 * The reference converts floor() to int; for NaN/Inf/|v|>=2^31 x86 yields
 * INT_MIN, i.e. "c < 0" -> failed.  Comparing in floating point is the same
 * predicate without the undefined conversion. */
static int orc_point_coor_(const float *p, const float *voxel_size,
                           const float *coors_range, const int *grid,
                           int *czyx) {
  for (int j = 0; j < 3; ++j) {
    float f = floorf((p[j] - coors_range[j]) / voxel_size[j]);
    if (!(f >= 0.0f && f < (float)grid[j])) return 0;
    czyx[2 - j] = (int)f;
  }
  return 1;
}

/* dynamic_voxelize_cpu: voxelization_cpu.cpp:7-43,146-171.
 * points (N, C) fp32 row-major, coors (N, 3) int32: (z,y,x) or (-1,-1,-1). */
void orc_dynamic_voxelize(const float *points, int64_t N, int C,
                          const float *voxel_size, const float *coors_range,
                          int32_t *coors) {
  int grid[3];
  orc_grid_size_(voxel_size, coors_range, grid);
  for (int64_t i = 0; i < N; ++i) {
    int c[3];
    if (orc_point_coor_(points + i * C, voxel_size, coors_range, grid, c)) {
      coors[i * 3 + 0] = c[0];
      coors[i * 3 + 1] = c[1];
      coors[i * 3 + 2] = c[2];
    } else {
      coors[i * 3 + 0] = coors[i * 3 + 1] = coors[i * 3 + 2] = -1;
    }
  }
}

/* hard_voxelize_cpu: voxelization_cpu.cpp:45-101,107-144.
 * The caller passes zero-filled voxels (max_voxels, max_points, C),
 * coors (max_voxels, 3) and num (max_voxels) like voxelize.py:57-61 does.
 * Sequential scan: first-occurrence voxel numbering (:75-88), a NEW voxel is
 * dropped once voxel_num >= max_voxels (:80) while existing voxels keep
 * filling, slot = arrival order, kept iff num < max_points (:91-97).
 * point2voxel (optional, N) receives the voxel id of every point that was
 * looked at and mapped (-1 if out of range or its voxel was dropped); it is
 * an extra output used by the parity tests, not part of the reference API. */
int orc_hard_voxelize(const float *points, int64_t N, int C,
                      const float *voxel_size, const float *coors_range,
                      int max_points, int max_voxels, float *voxels,
                      int32_t *coors, int32_t *num, int32_t *point2voxel) {
  int grid[3];
  orc_grid_size_(voxel_size, coors_range, grid);
  size_t cells = (size_t)grid[0] * (size_t)grid[1] * (size_t)grid[2];
  int32_t *coor_to_voxelidx = (int32_t *)malloc(cells * sizeof(int32_t));
  if (!coor_to_voxelidx) return -1;
  memset(coor_to_voxelidx, 0xFF, cells * sizeof(int32_t)); /* all -1 (:129) */
  int voxel_num = 0;
  for (int64_t i = 0; i < N; ++i) {
    int c[3];
    if (point2voxel) point2voxel[i] = -1;
    if (!orc_point_coor_(points + i * C, voxel_size, coors_range, grid, c))
      continue;
    size_t cell = ((size_t)c[0] * grid[1] + c[1]) * grid[0] + c[2];
    int voxelidx = coor_to_voxelidx[cell];
    if (voxelidx == -1) {
      voxelidx = voxel_num;
      if (max_voxels != -1 && voxel_num >= max_voxels) continue;
      voxel_num += 1;
      coor_to_voxelidx[cell] = voxelidx;
      for (int k = 0; k < 3; ++k) coors[voxelidx * 3 + k] = c[k];
    }
    if (point2voxel) point2voxel[i] = voxelidx;
    int n = num[voxelidx];
    if (max_points == -1 || n < max_points) {
      for (int k = 0; k < C; ++k)
        voxels[((size_t)voxelidx * max_points + n) * C + k] = points[i * C + k];
      num[voxelidx] += 1;
    }
  }
  free(coor_to_voxelidx);
  return voxel_num;
}

/* HardSimpleVFE.forward: mmdet3d/models/voxel_encoders/voxel_encoder.py:45-46
 *   features[:, :, :F].sum(dim=1) / num_points.type_as(features).view(-1, 1)
 * Sum over ALL max_points slots (zeros included), sequentially in slot order,
 * in fp32; one fp32 division.  out (M, F). */
void orc_hard_simple_vfe(const float *voxels, const int32_t *num, int64_t M,
                         int max_points, int C, int F, float *out) {
  for (int64_t m = 0; m < M; ++m) {
    float cnt = (float)num[m];
    for (int f = 0; f < F; ++f) {
      float s = 0.0f;
      for (int k = 0; k < max_points; ++k)
        s = s + voxels[((size_t)m * max_points + k) * C + f];
      out[m * F + f] = s / cnt;
    }
  }
}

/* Same, accumulated in fp64 (tolerance anchor for the 1e-6 relative bar). */
void orc_hard_simple_vfe_f64(const float *voxels, const int32_t *num,
                             int64_t M, int max_points, int C, int F,
                             double *out) {
  for (int64_t m = 0; m < M; ++m) {
    for (int f = 0; f < F; ++f) {
      double s = 0.0;
      for (int k = 0; k < max_points; ++k)
        s += (double)voxels[((size_t)m * max_points + k) * C + f];
      out[m * F + f] = s / (double)num[m];
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Depth -> ego-frame points.                                                */
/* reconstruction_backbone.py:305-386, one sample (the b loop is the caller) */
/* ------------------------------------------------------------------------ */
/*
 * depth  (ncam, H, W) fp32
 * intr   (ncam, 3, 3) fp32   fx=K[0][0] fy=K[1][1] cx=K[0][2] cy=K[1][2] (:325-326)
 * cam2lidar (ncam, 4, 4) fp32; rotation M[:3,:3], translation in ROW 3 (:370)
 * max_depth: applied when use_max_depth != 0 (:339-340)
 * conf/conf_thresh: valid &= conf >= thresh when conf != NULL
 *                   (tools/inference_nuscenes.py:399-402)
 * sky: valid &= !sky when sky != NULL (tools/inference_nuscenes.py:407-414)
 * range6: inclusive range filter min<=p<=max on the TRANSFORMED point when
 *         range6 != NULL (respoint_post_processing.py:190-195)
 * out_points (cap, 3), optional out_pix (cap) = flat pixel index
 *         cam*H*W + v*W + u of every emitted point.
 * Returns number of points P.  Order: cameras in index order, pixels
 * row-major (:336,:348,:376-378).
 *
 * Arithmetic (each op separately rounded to fp32, :329-334):
 *   x = ((u - cx) * z) / fx ;  y = ((v - cy) * z) / fy
 * Transform pts @ R.T + t (:370): torch-CPU's sgemm evaluates each output as
 *   fma(z, R[i][2], fma(y, R[i][1], x * R[i][0]))  then a separate  + t[i]
 * (re-verified in tests/test_oracle.py against
 * torch.matmul on this machine).  That formula is the DEFINITION here.
 */
int64_t orc_unproject(const float *depth, int ncam, int H, int W,
                      const float *intr, const float *cam2lidar,
                      int use_max_depth, float max_depth, const float *conf,
                      float conf_thresh, const uint8_t *sky,
                      const float *range6, float *out_points,
                      int32_t *out_pix, int64_t cap) {
  int64_t P = 0;
  for (int cam = 0; cam < ncam; ++cam) {
    const float *K = intr + cam * 9;
    const float *M = cam2lidar + cam * 16;
    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    for (int v = 0; v < H; ++v) {
      for (int u = 0; u < W; ++u) {
        int64_t pix = ((int64_t)cam * H + v) * W + u;
        float z = depth[pix];
        int valid = (z > 0.0f) && isfinite(z);
        if (use_max_depth) valid = valid && (z <= max_depth);
        if (conf) valid = valid && (conf[pix] >= conf_thresh);
        if (sky) valid = valid && !sky[pix];
        if (!valid) continue;
        float x = (((float)u - cx) * z) / fx;
        float y = (((float)v - cy) * z) / fy;
        float p[3];
        for (int i = 0; i < 3; ++i) {
          float acc = x * M[i * 4 + 0];
          acc = fmaf(y, M[i * 4 + 1], acc);
          acc = fmaf(z, M[i * 4 + 2], acc);
          p[i] = acc + M[12 + i];
        }
        if (range6) {
          if (!(p[0] >= range6[0] && p[0] <= range6[3] && p[1] >= range6[1] &&
                p[1] <= range6[4] && p[2] >= range6[2] && p[2] <= range6[5]))
            continue;
        }
        if (P < cap) {
          out_points[P * 3 + 0] = p[0];
          out_points[P * 3 + 1] = p[1];
          out_points[P * 3 + 2] = p[2];
          if (out_pix) out_pix[P] = (int32_t)pix;
        }
        ++P;
      }
    }
  }
  return P;
}

/* ------------------------------------------------------------------------ */
/* DynamicScatter forward.  The reference has no CPU kernel                   */
/* (voxelization.h:118); this restates the GPU host logic                     */
/* scatter_points_cuda.cu:183-239:                                           */
/*   rows with any negative component are dropped (map -1)        (:202)     */
/*   voxels = lexicographically sorted unique rows                (:204-205) */
/*   inverse map, counts                                          (:214-215) */
/*   sum / mean (= sum then /count, :233-234) / max (init -inf, fmaxf :22-30)*/
/* Sums are accumulated in fp64 in point order and rounded once: the         */
/* reference's own fp32 atomic order is not reproducible, so the oracle is   */
/* the correctly rounded value both sides must sit within 1e-6 relative of.  */
/* reduce_type: 0 sum, 1 mean, 2 max.                                        */
/* out_feats (N, C), out_coors (N, 3), map (N), count (N) sized for the      */
/* worst case; returns M.                                                    */
/* ------------------------------------------------------------------------ */
typedef struct {
  int32_t c[3];
  int64_t idx;
} orc_row_t;

static int orc_row_cmp_(const void *a, const void *b) {
  const orc_row_t *ra = (const orc_row_t *)a, *rb = (const orc_row_t *)b;
  for (int k = 0; k < 3; ++k) {
    if (ra->c[k] < rb->c[k]) return -1;
    if (ra->c[k] > rb->c[k]) return 1;
  }
  return (ra->idx > rb->idx) - (ra->idx < rb->idx);
}

int64_t orc_dynamic_scatter(const float *feats, const int32_t *coors,
                            int64_t N, int C, int reduce_type,
                            float *out_feats, int32_t *out_coors, int32_t *map,
                            int32_t *count) {
  orc_row_t *rows = (orc_row_t *)malloc((size_t)(N > 0 ? N : 1) * sizeof(orc_row_t));
  double *acc = (double *)malloc((size_t)(C > 0 ? C : 1) * sizeof(double));
  int64_t nv = 0;
  for (int64_t i = 0; i < N; ++i) {
    const int32_t *c = coors + i * 3;
    if (c[0] < 0 || c[1] < 0 || c[2] < 0) {
      map[i] = -1;
      continue;
    }
    rows[nv].c[0] = c[0];
    rows[nv].c[1] = c[1];
    rows[nv].c[2] = c[2];
    rows[nv].idx = i;
    ++nv;
  }
  qsort(rows, (size_t)nv, sizeof(orc_row_t), orc_row_cmp_);
  int64_t M = 0;
  int64_t s = 0;
  while (s < nv) {
    int64_t e = s + 1;
    while (e < nv && rows[e].c[0] == rows[s].c[0] &&
           rows[e].c[1] == rows[s].c[1] && rows[e].c[2] == rows[s].c[2])
      ++e;
    for (int k = 0; k < 3; ++k) out_coors[M * 3 + k] = rows[s].c[k];
    count[M] = (int32_t)(e - s);
    for (int f = 0; f < C; ++f) {
      if (reduce_type == 2) {
        float m = -INFINITY;
        // reduceMax (scatter_points_cuda.cu:22-30) is a CAS loop around the DEVICE fmaxf: NaN operands are
        // dropped and +0.0 orders above -0.0 (PTX max.f32); written out because C's fmaxf leaves the zeros open
        for (int64_t j = s; j < e; ++j) {
          const float x = feats[rows[j].idx * C + f];
          if (x > m || (x == m && !signbit(x))) m = x;
        }
        out_feats[M * C + f] = m;
      } else {
        double a = 0.0;
        for (int64_t j = s; j < e; ++j) a += (double)feats[rows[j].idx * C + f];
        float sum32 = (float)a;
        out_feats[M * C + f] =
            (reduce_type == 1) ? sum32 / (float)(e - s) : sum32;
      }
    }
    for (int64_t j = s; j < e; ++j) map[rows[j].idx] = (int32_t)M;
    ++M;
    s = e;
  }
  free(rows);
  free(acc);
  return M;
}
