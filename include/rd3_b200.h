/*
 * rd3_b200.h -- C ABI of the B200-native depth->voxel path.
 *
 * One shared library (librd3_b200.so, sm_100a only), plain pointers and sizes,
 * no torch types, no exceptions, no implicit synchronisation: every entry point
 * enqueues its kernels on the caller's stream and returns an int status
 * (RD3_OK == 0).  All pointers are DEVICE pointers unless a parameter says
 * "host".  Data-dependent counts are written to device memory; the caller
 * decides when to read them back.
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the reference repository root):
 *
 *   pybind module `voxel_layer`
 *       mmdetection3d/mmdet3d/ops/voxel/src/voxelization.cpp:6-11
 *       dispatch: mmdetection3d/mmdet3d/ops/voxel/src/voxelization.h:58-140
 *   HardSimpleVFE.forward
 *       mmdetection3d/mmdet3d/models/voxel_encoders/voxel_encoder.py:30-47
 *   ReconstructionBackbone._backproject_depth_to_points
 *       projects/mmdet3d_plugin/models/backbone/reconstruction_backbone.py:285-386
 *
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef RD3_B200_H_
#define RD3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RD3_API __attribute__((visibility("default")))
#else
#define RD3_API
#endif

/* cudaStream_t without dragging cuda_runtime.h into C callers */
typedef void *rd3_stream_t;

enum {
  RD3_OK = 0,
  RD3_ERR_INVALID_ARGUMENT = 1, /* null pointer, negative size, NDim != 3 ...        */
  RD3_ERR_WORKSPACE = 2,        /* workspace smaller than *_workspace_bytes() says   */
  RD3_ERR_UNSUPPORTED = 3,      /* e.g. voxel grid with more than 2^32-2 cells       */
  RD3_ERR_CUDA = 4              /* a CUDA call failed; see rd3_last_cuda_error()     */
};

enum { RD3_REDUCE_SUM = 0, RD3_REDUCE_MEAN = 1, RD3_REDUCE_MAX = 2 };

RD3_API int rd3_version(void);
RD3_API const char *rd3_status_string(int status);
/* cudaGetErrorString of the last CUDA failure seen by this library (host string). */
RD3_API const char *rd3_last_cuda_error(void);

/* Process-wide state of the library (everything else lives in the caller's workspace):
 *   - per device, up to 3 internal non-blocking streams + fork / join events ("lanes"): the hard-voxel pipeline
 *     runs the frames of a batch in sub-batches on them, forked from and joined to the caller's stream.  Host
 *     threads that enqueue on the same device are serialised by a per-device lock for the duration of the
 *     enqueue (no device wait inside).  RD3_STREAMS=1..4 sets the number of sub-batches (default 2).
 *   - the stage profiler below (off by default; while it is on, one lane is used and calls must not overlap).
 *   - the last CUDA error string.
 *
 * Per-stage timing of the hard-voxel pipeline (rd3_hard_voxelize and
 * rd3_depth_to_voxels) with CUDA events recorded on the launching stream between
 * its kernels.  Stages: 0 memset, 1 insert rounds, 2 count (flags -> voxel_num), 3 post (first points, cull
 * bits), 4 lookup, 5 emit, 6 meta.  rd3_profile_enable(1) resets the counters; rd3_profile_read
 * waits for the recorded events and returns the summed milliseconds per stage
 * (host double[7]) and the number of calls covered (at most 1024). */
#define RD3_PROFILE_STAGES 7
RD3_API int rd3_profile_enable(int on);
RD3_API int rd3_profile_read(double *stage_ms, int *calls);

/* grid = round((max - min) / voxel_size) in fp32, x,y,z order.
 * Replaces the inline computation at voxelization_cpu.cpp:121-124 /
 * voxelize.py:113-121.  Host-only helper (voxel_size, coors_range, grid: host). */
RD3_API int rd3_grid_size(const float voxel_size[3], const float coors_range[6], int32_t grid[3]);

/* ---------------------------------------------------------------------------
 * dynamic_voxelize   (voxelization.h:83-94 -> voxelization_cpu.cpp:146-171)
 *   points (N, C>=3) fp32 row-major -> coors (N, 3) int32 = (z,y,x), or
 *   (-1,-1,-1) when the point falls outside [min, min+grid*vs) on any axis
 *   (CPU semantics; the reference GPU kernel's partial -1 prefix is not kept).
 *   voxel_size / coors_range: host arrays.
 * ------------------------------------------------------------------------- */
RD3_API int rd3_dynamic_voxelize(const float *points, int64_t N, int C,
                         const float voxel_size[3], const float coors_range[6],
                         int32_t *coors, rd3_stream_t stream);

/* ---------------------------------------------------------------------------
 * hard_voxelize      (voxelization.h:58-81 -> voxelization_cpu.cpp:107-144)
 *   Deterministic; reproduces the sequential CPU scan bit for bit:
 *   voxels numbered in order of their first point, a new voxel is dropped once
 *   max_voxels exist, a voxel keeps its first max_points points in point order.
 *
 *   voxels (max_voxels, max_points, C), coors (max_voxels, 3) zyx,
 *   num_points_per_voxel (max_voxels): rows [0, voxel_num) are fully written
 *   (unused slots = 0); rows beyond are left untouched (the reference's caller
 *   pre-zeroes them, voxelize.py:57-61).
 *   d_voxel_num: device int32[1], the value the reference returns as `int`.
 *   workspace: 256-byte aligned, rd3_hard_voxelize_workspace_bytes() bytes.
 *   Optional outputs (may be NULL):
 *     voxel_mean (max_voxels, F): HardSimpleVFE fused (sum of slots / count).
 *     point2voxel (N): voxel id of each point, -1 if out of range / dropped.
 * ------------------------------------------------------------------------- */
RD3_API size_t rd3_hard_voxelize_workspace_bytes(int64_t N, int max_points, int max_voxels);
/* number of insert-kernel launches (index-ordered rounds) a call with N points per frame and
 * B frames makes; kernel launches per call = lanes * (rounds + 5) [+ 1 calibration kernel]. */
RD3_API int rd3_hard_voxel_rounds(int64_t N, int B);

RD3_API int rd3_hard_voxelize(const float *points, int64_t N, int C,
                      const float voxel_size[3], const float coors_range[6],
                      int max_points, int max_voxels, float *voxels,
                      int32_t *coors, int32_t *num_points_per_voxel,
                      int32_t *d_voxel_num, float *voxel_mean, int F,
                      int32_t *point2voxel, void *workspace,
                      size_t workspace_bytes, rd3_stream_t stream);

/* ---------------------------------------------------------------------------
 * HardSimpleVFE.forward   (voxel_encoder.py:45-46)
 *   out (M, F) = voxels[:, :, :F].sum(dim=1) / float(num_points)
 * ------------------------------------------------------------------------- */
RD3_API int rd3_hard_simple_vfe(const float *voxels, const int32_t *num_points, int64_t M,
                        int max_points, int C, int F, float *out,
                        rd3_stream_t stream);

/* ---------------------------------------------------------------------------
 * Depth maps -> ego-frame points
 *   (reconstruction_backbone.py:285-386; live masks of
 *    tools/inference_nuscenes.py:399-414; inclusive range filter of
 *    projects/mmdet3d_plugin/datasets/pipelines/respoint_post_processing.py:190-195)
 *
 *   depth (B, ncam, H, W) fp32; intrinsics (B, ncam, 3, 3); cam2lidar
 *   (B, ncam, 4, 4) with the translation in row 3; conf (B, ncam, H, W) fp32 or
 *   NULL; sky (B, ncam, H, W) uint8/bool or NULL.
 *   valid = z > 0 && isfinite(z) [&& z <= max_depth] [&& conf >= conf_thresh]
 *           [&& !sky] [&& range_min <= p <= range_max].
 *   x = ((u-cx)*z)/fx, y = ((v-cy)*z)/fy, p = fma-chain(R, (x,y,z)) + t.
 * ------------------------------------------------------------------------- */
typedef struct rd3_depth_params {
  int32_t B, ncam, H, W;
  int32_t use_max_depth;  /* 0/1 */
  float max_depth;
  float conf_thresh;      /* used when conf != NULL */
  int32_t use_range;      /* 0/1: inclusive filter on the transformed point */
  float range[6];         /* x0 y0 z0 x1 y1 z1 */
  const float *conf_thresh_dev; /* optional DEVICE float[B]: per-sample thresholds (e.g. the
                                   d_thresh32 output of rd3_conf_percentile, no host round trip);
                                   overrides conf_thresh when not NULL */
  const float *sky_prob;  /* optional DEVICE fp32 (B, ncam, H, W): DA3's raw sky head output; a pixel is sky
                             iff sky_prob >= sky_prob_thresh (output_processor.py:152-168 uses 0.5).  Read
                             only when the `sky` argument of the call is NULL, so the boolean mask is never
                             materialised. */
  float sky_prob_thresh;
} rd3_depth_params;

/* Order-preserving compaction (cameras in index order, pixels row-major).
 *   out_points (B, ncam*H*W, 3): first d_counts[b] rows of sample b are valid.
 *   out_pix (B, ncam*H*W) int32 or NULL: flat pixel index of each emitted point.
 *   d_counts: device int32[B]. */
RD3_API size_t rd3_unproject_workspace_bytes(const rd3_depth_params *p);

RD3_API int rd3_unproject(const float *depth, const float *intrinsics,
                  const float *cam2lidar, const float *conf, const uint8_t *sky,
                  const rd3_depth_params *p, float *out_points, int32_t *out_pix,
                  int32_t *d_counts, void *workspace, size_t workspace_bytes,
                  rd3_stream_t stream);

/* Fused: depth -> hard voxels (+ voxel mean) for B samples in one pass; the
 * point cloud is never written to memory.  Output of sample b is bit-identical
 * to rd3_unproject(b) followed by rd3_hard_voxelize + rd3_hard_simple_vfe.
 *   voxels (B, max_voxels, max_points, 3), coors (B, max_voxels, 3),
 *   num_points_per_voxel (B, max_voxels), voxel_mean (B, max_voxels, 3) or
 *   NULL, d_voxel_num device int32[B].
 *   voxels may be NULL when voxel_mean is given: a caller that feeds the sparse
 *   encoder (HardSimpleVFE features + coors, sparse_refinement.py:382-402) never
 *   reads the padded voxel tensor, and 85 % of the output bytes are not written.
 *   workspace: 256-byte aligned (256-bit table loads, TMA bulk copies of the
 *   calibration table), rd3_depth_to_voxels_workspace_bytes() bytes;
 *   RD3_ERR_INVALID_ARGUMENT otherwise. */
RD3_API size_t rd3_depth_to_voxels_workspace_bytes(const rd3_depth_params *p,
                                           int max_points, int max_voxels);

RD3_API int rd3_depth_to_voxels(const float *depth, const float *intrinsics,
                        const float *cam2lidar, const float *conf,
                        const uint8_t *sky, const rd3_depth_params *p,
                        const float voxel_size[3], const float coors_range[6],
                        int max_points, int max_voxels, float *voxels,
                        int32_t *coors, int32_t *num_points_per_voxel,
                        float *voxel_mean, int32_t *d_voxel_num,
                        void *workspace, size_t workspace_bytes,
                        rd3_stream_t stream);

/* ---------------------------------------------------------------------------
 * Batched sparse-encoder inputs.  Replaces the Python tail of
 * SparseRefinement._voxelize_and_encode
 * (projects/mmdet3d_plugin/models/backbone/sparse_refinement.py:393-402:
 * torch.cat of the per-sample slices + F.pad(coor, (1, 0), value=i)) and of
 * MVXTwoStageDetector.voxelize (mmdet3d/models/detectors/mvx_two_stage.py:211-236)
 * with one launch and no host synchronisation:
 *   voxel_feats (B, max_voxels, F), coors (B, max_voxels, 3) zyx, num_points
 *   (B, max_voxels) or NULL, d_voxel_num device int32[B]  ->  rows
 *   [0, voxel_num[b]) of every sample packed in sample order:
 *   out_feats (sum M, F), out_coors (sum M, 4) = (batch_offset + b, z, y, x),
 *   out_num_points (sum M) or NULL; d_offsets device int32[B + 1] = exclusive
 *   prefix of the counts, d_offsets[B] = sum M.  Outputs must hold B * max_voxels
 *   rows (worst case); out_coors 16-byte aligned.  B <= 65535.
 * ------------------------------------------------------------------------- */
RD3_API int rd3_pack_sparse_inputs(const float *voxel_feats, const int32_t *coors,
                           const int32_t *num_points, const int32_t *d_voxel_num,
                           int B, int max_voxels, int F, int batch_offset,
                           float *out_feats, int32_t *out_coors,
                           int32_t *out_num_points, int32_t *d_offsets,
                           rd3_stream_t stream);

/* ---------------------------------------------------------------------------
 * dynamic_point_to_voxel_forward
 *   (voxelization.h:108-121 -> scatter_points_cuda.cu:183-239; with ncols == 4 also the
 *   per-sample loop of scatter_points.py:86-97, as ONE launch sequence)
 *   feats (N, C) fp32, coors (N, ncols) int32, ncols 3: (c0,c1,c2), ncols 4: (batch,c0,c1,c2).
 *   Rows with any negative component are dropped (map -1); with ncols == 4 so are rows whose
 *   batch index exceeds coors[N-1][0] (the reference derives the batch size from the last row).
 *   Voxels come out in lexicographic order of the ncols columns, i.e. for ncols == 4 the
 *   concatenation of the per-sample results in sample order with the batch column in front.
 *   dims (host int32[4], dims[0] = sample capacity, ignored for ncols == 3; dims[1..3] = exclusive
 *   upper bounds of c0,c1,c2, e.g. the voxel grid (gz,gy,gx)); a valid coordinate >= dims (or more
 *   samples than dims[0]) sets *d_status (device int32[1]) to 1 and the outputs are then
 *   undefined -- call rd3_coors_extent and retry.  prod(dims) must stay below 2^32 - 32.
 *   Outputs sized for the worst case: voxel_feats (N, C), voxel_coors (N, ncols),
 *   point2voxel (N), voxel_count (N); d_num_voxels device int32[1].  ncols == 4: coors and
 *   voxel_coors 16-byte aligned.  sum / mean are accumulated in fp64 and rounded once.
 * ------------------------------------------------------------------------- */
RD3_API int rd3_coors_extent(const int32_t *coors, int64_t N, int ncols, int32_t *d_extent4,
                     rd3_stream_t stream);

RD3_API size_t rd3_dynamic_scatter_workspace_bytes(int64_t N, int C, int ncols, const int32_t dims[4]);

RD3_API int rd3_dynamic_scatter_forward(const float *feats, const int32_t *coors, int64_t N,
                                int C, int ncols, const int32_t dims[4], int reduce_type,
                                float *voxel_feats, int32_t *voxel_coors,
                                int32_t *point2voxel, int32_t *voxel_count,
                                int32_t *d_num_voxels, int32_t *d_status,
                                void *workspace, size_t workspace_bytes,
                                rd3_stream_t stream);

/* dynamic_point_to_voxel_backward
 *   (voxelization.h:123-140 -> scatter_points_cuda.cu:241-308)
 *   grad_feats (N, C) is fully written (zeros for dropped points).
 *   max: the gradient goes to the lowest-index point attaining the maximum. */
RD3_API size_t rd3_dynamic_scatter_backward_workspace_bytes(int64_t M, int C);

RD3_API int rd3_dynamic_scatter_backward(float *grad_feats, const float *grad_voxel_feats,
                                 const float *feats, const float *voxel_feats,
                                 const int32_t *point2voxel,
                                 const int32_t *voxel_count, int64_t N, int64_t M,
                                 int C, int reduce_type, void *workspace,
                                 size_t workspace_bytes, rd3_stream_t stream);

/* ---------------------------------------------------------------------------
 * Pillar encoders, gather side (what follows hard voxelization in the pillar configs,
 * configs/_base_/models/centerpoint_02pillar_second_secfpn_nus.py:1-16).
 *
 * rd3_pillar_decorate: the feature decorations of PillarFeatureNet.forward
 *   (mmdetection3d/mmdet3d/models/voxel_encoders/pillar_encoder.py:104-146), i.e. everything
 *   in front of the PFN layers, in one launch:
 *   voxels (M, max_points, C>=3), num_points (M), coors (M, coors_cols) with the
 *   last two columns (y, x) [coors_cols 4: (b,z,y,x), 3: (z,y,x)]  ->
 *   out (M, max_points, C + 3*with_cluster_center + 2*with_voxel_center + with_distance)
 *   = cat(features, xyz - mean_xyz, (x,y) - pillar centre, |xyz|) * padding mask.
 *   x_offset = vx/2 + pcr[0], y_offset = vy/2 + pcr[1] (:87-88), narrowed to fp32 by the caller.
 *   legacy != 0 reproduces the default legacy=True aliasing (:127-133): the raw x, y columns are
 *   replaced by the centre offsets and the distance is taken from them.
 *
 * rd3_pillars_scatter: PointPillarsScatter.forward_batch / forward_single
 *   (mmdetection3d/mmdet3d/models/middle_encoders/pillar_scatter.py:39-102):
 *   voxel_features (M, C), coors (M, coors_cols) -> canvas (batch_size, C, ny, nx), zero where
 *   no pillar.  canvas is fully written (zero fill included).  Rows whose (b, y, x) fall
 *   outside the canvas are skipped (the reference's indexed assignment would raise).
 * ------------------------------------------------------------------------- */
RD3_API int rd3_pillar_decorate(const float *voxels, const int32_t *num_points,
                        const int32_t *coors, int64_t M, int max_points, int C,
                        int coors_cols, int with_cluster_center, int with_voxel_center,
                        int with_distance, int legacy, float vx, float vy,
                        float x_offset, float y_offset, float *out, rd3_stream_t stream);

RD3_API int rd3_pillars_scatter(const float *voxel_features, const int32_t *coors, int64_t M,
                        int C, int coors_cols, int batch_size, int ny, int nx,
                        float *canvas, rd3_stream_t stream);

/* GT-side occupancy of the sparse refinement:
 *   SoftVoxelOccupancyVFE / HardVoxelOccupancyVFE.forward
 *     (projects/mmdet3d_plugin/models/backbone/voxel_occupancy_encoder.py:22-37,60-99):
 *     soft: p = 1 - exp(-lambda_n * n - gamma_var * var), var = mean over xyz of the masked
 *     variance around the masked mean (denominator n + eps);  hard != 0: p = (n > 0).
 *   and the dense scatter of sparse_refinement.py:572-587: dense_map[b, z, y, x] = p.
 *   voxels (M, max_points, C>=3), num_points (M); occupancy (M) or NULL; dense_map
 *   (batch_size, Z, Y, X) or NULL -- fully written (zero fill included) -- with coors
 *   (M, coors_cols) = (b,z,y,x) or (z,y,x). */
RD3_API int rd3_voxel_occupancy(const float *voxels, const int32_t *num_points, int64_t M,
                        int max_points, int C, int hard, float lambda_n, float gamma_var,
                        float eps, float *occupancy, const int32_t *coors, int coors_cols,
                        int batch_size, int Z, int Y, int X, float *dense_map,
                        rd3_stream_t stream);

/* DynamicVFE.map_voxel_center_to_point / HardVFE's and DynamicPillarFeatureNet's copies of it
 *   (mmdetection3d/mmdet3d/models/voxel_encoders/voxel_encoder.py:179-219,
 *    pillar_encoder.py:235-275): out[i] = voxel_feats[j] where voxel_coors[j] == pts_coors[i]
 *   (all four columns b,z,y,x); a point whose voxel is not listed gets row 0, like the reference's
 *   zero-initialised canvas.  pts_coors (N,4), voxel_coors (M,4) int32, 16-byte aligned, any row
 *   order; voxel_feats (M,C); out (N,C); out_index (N) int32 or NULL (the voxel row used).
 *   The reference's dense z*y*x*batch int64 canvas is replaced by a table of 2M..4M int32. */
RD3_API size_t rd3_map_voxel_to_point_workspace_bytes(int64_t M);

RD3_API int rd3_map_voxel_to_point(const int32_t *pts_coors, int64_t N, const int32_t *voxel_coors,
                           const float *voxel_feats, int64_t M, int C, float *out,
                           int32_t *out_index, void *workspace, size_t workspace_bytes,
                           rd3_stream_t stream);

/* ---------------------------------------------------------------------------
 * Confidence threshold of the depth -> points hand-off, per sample, on the device:
 *   conf_thresh = np.percentile(conf[~sky] if (~sky).sum() > 10 else conf.flatten(), p)
 *   (tools/inference_nuscenes.py:351-361; the GLB export does the same,
 *    depth_anything_3/utils/export/glb.py:227-229) -- a partition of 2.7 M values per
 *   sample on one CPU core in the reference.
 *   conf (B, npix) fp32, sky (B, npix) uint8/bool or NULL, percentile in [0, 100] (host).
 *   Exact: a 3-pass radix select finds the two order statistics that numpy's "linear"
 *   method interpolates between; index, weight and interpolation follow numpy's
 *   arithmetic (numpy/lib/_function_base_impl.py: (n-1)*q, _get_indexes, _lerp):
 *   numpy2_fp32_index != 0 -> fp32 index / weight / result, what NumPy >= 2 computes for
 *   fp32 data and a Python-float percentile; 0 -> fp64 index and interpolation (NumPy < 2,
 *   the reference's pin, requirements.txt:8).
 *   d_thresh device double[B]; d_thresh32 device float[B] or NULL (the value rounded to
 *   fp32); d_count device int32[B] or NULL (number of values selected).  NaN when a sample
 *   has no pixel or when a selected value is NaN (like np.percentile).
 * ------------------------------------------------------------------------- */
RD3_API size_t rd3_conf_percentile_workspace_bytes(int B);

RD3_API int rd3_conf_percentile(const float *conf, const uint8_t *sky, int B, int64_t npix,
                        double percentile, int numpy2_fp32_index, double *d_thresh,
                        float *d_thresh32, int32_t *d_count, void *workspace,
                        size_t workspace_bytes, rd3_stream_t stream);

/* The same with DA3's raw sky head output instead of a boolean mask: a pixel is sky iff
 * sky_prob >= sky_prob_thresh (depth_anything_3/utils/io/output_processor.py:152-168: 0.5), so the
 * `(sky >= 0.5)` tensor between the network and the threshold is never written.  sky_prob may be NULL. */
RD3_API int rd3_conf_percentile_skyprob(const float *conf, const float *sky_prob, float sky_prob_thresh,
                        int B, int64_t npix, double percentile, int numpy2_fp32_index,
                        double *d_thresh, float *d_thresh32, int32_t *d_count, void *workspace,
                        size_t workspace_bytes, rd3_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RD3_B200_H_ */
