// scatter.cu -- DynamicScatter forward / backward.
//
// Reference: mmdetection3d/mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-308
// (at::unique_dim -- a multi-pass row sort -- plus one fp32 atomic per feature per
// point).  Here the sorted-unique step needs no sort: valid coordinates are
// bounded by `dims`, so a voxel's lexicographic rank is the number of occupied
// cells before its linear id, i.e. a popcount prefix over an occupancy bitmap
// (83 M cells = 10.4 MB for the 1440x1440x40 grid, L2 resident on B200).
//
//   D1 mark    : bitmap |= cell(point)            (read-before-atomicOr)
//   D2 scan    : per-chunk exclusive popcount prefix; D2s: chunk totals -> M
//   D3 voxels  : every occupied cell writes its coordinates at its rank and
//                initialises its accumulators
//   D4 reduce  : rank per point -> point2voxel, count, sum (fp64 atomics: exact
//                for same-magnitude fp32 inputs, hence order independent) or
//                max (order independent by construction)
//   D5 finish  : sum -> fp32, mean = fp32(sum) / fp32(count)   (:233-234)
#include "hard_voxel.cuh"

namespace rd3 {

struct DsWork {
  uint32_t *bitmap;       // [nwords]
  int32_t *wordprefix;    // [nwords]
  int32_t *chunk_base;    // [nchunks]
  double *acc;            // [N*C] (only the first M*C are used)
  int nwords, nchunks;
  uint32_t d0, d1, d2;
};

__device__ __forceinline__ bool ds_key(const int32_t *__restrict__ coors, int64_t i, const DsWork &w,
                                       uint32_t &key, bool &overflow) {
  const int32_t c0 = __ldg(coors + i * 3), c1 = __ldg(coors + i * 3 + 1), c2 = __ldg(coors + i * 3 + 2);
  overflow = false;
  if (c0 < 0 || c1 < 0 || c2 < 0) return false;          // scatter_points_cuda.cu:202
  if ((uint32_t)c0 >= w.d0 || (uint32_t)c1 >= w.d1 || (uint32_t)c2 >= w.d2) {
    overflow = true;
    return false;
  }
  key = ((uint32_t)c0 * w.d1 + (uint32_t)c1) * w.d2 + (uint32_t)c2;
  return true;
}

static __global__ void __launch_bounds__(256)
    ds_mark_kernel(const int32_t *__restrict__ coors, int64_t N, DsWork w, int32_t *status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint32_t key;
  bool ovf;
  if (ds_key(coors, i, w, key, ovf)) {
    uint32_t *word = w.bitmap + (key >> 5);
    const uint32_t bit = 1u << (key & 31);
    if (!(__ldcg(word) & bit)) atomicOr(word, bit);
  } else if (ovf) {
    *status = 1;
  }
}

static __global__ void __launch_bounds__(kScanThreads) ds_scan_kernel(DsWork w) {
  __shared__ int s_warp[kScanThreads / 32];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int64_t wi = (int64_t)blockIdx.x * kChunkWords + threadIdx.x;
  const int cnt = __popc(w.bitmap[wi]);
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
  w.wordprefix[wi] = base + inc - cnt;
  if (threadIdx.x == 0) w.chunk_base[blockIdx.x] = total;
}

__device__ __forceinline__ int ds_rank(const DsWork &w, uint32_t key) {
  const uint32_t word = key >> 5;
  const uint32_t bits = __ldg(w.bitmap + word) & ((1u << (key & 31)) - 1u);
  return __ldg(w.chunk_base + (word / kChunkWords)) + __ldg(w.wordprefix + word) + __popc(bits);
}

// one thread per bitmap word: emit coordinates of its occupied cells, init accumulators
static __global__ void __launch_bounds__(256)
    ds_voxels_kernel(DsWork w, int C, int reduce_type, int32_t *__restrict__ voxel_coors,
                     int32_t *__restrict__ voxel_count, float *__restrict__ voxel_feats) {
  const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= w.nwords) return;
  uint32_t bits = w.bitmap[wi];
  if (!bits) return;
  int r = __ldg(w.chunk_base + (wi / kChunkWords)) + __ldg(w.wordprefix + wi);
  while (bits) {
    const int bpos = __ffs(bits) - 1;
    bits &= bits - 1;
    const uint32_t key = (uint32_t)(wi << 5) + bpos;
    const uint32_t c2 = key % w.d2;
    const uint32_t t = key / w.d2;
    voxel_coors[(int64_t)r * 3 + 0] = (int32_t)(t / w.d1);
    voxel_coors[(int64_t)r * 3 + 1] = (int32_t)(t % w.d1);
    voxel_coors[(int64_t)r * 3 + 2] = (int32_t)c2;
    voxel_count[r] = 0;
    for (int c = 0; c < C; ++c) {
      if (reduce_type == RD3_REDUCE_MAX) voxel_feats[(int64_t)r * C + c] = __int_as_float(0xFF800000);
      else w.acc[(int64_t)r * C + c] = 0.0;
    }
    ++r;
  }
}

__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
  if (v != v) return;                       // fmaxf(NaN, m) == m  (scatter_points_cuda.cu:22-30)
  // branch on the SIGN BIT, not on v >= 0: -0.0 (0x80000000 = INT_MIN as a signed int) must take the
  // unsigned branch, where it beats the -inf initial value like every other negative number
  if (__float_as_int(v) >= 0) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

static __global__ void __launch_bounds__(256)
    ds_reduce_kernel(const float *__restrict__ feats, const int32_t *__restrict__ coors, int64_t N,
                     int C, int reduce_type, DsWork w, int32_t *__restrict__ point2voxel,
                     int32_t *voxel_count, float *voxel_feats) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint32_t key;
  bool ovf;
  if (!ds_key(coors, i, w, key, ovf)) {
    point2voxel[i] = -1;
    return;
  }
  const int r = ds_rank(w, key);
  point2voxel[i] = r;
  atomicAdd(voxel_count + r, 1);
  const float *f = feats + i * C;
  if (reduce_type == RD3_REDUCE_MAX) {
    for (int c = 0; c < C; ++c) atomic_max_float(voxel_feats + (int64_t)r * C + c, __ldg(f + c));
  } else {
    for (int c = 0; c < C; ++c) atomicAdd(w.acc + (int64_t)r * C + c, (double)__ldg(f + c));
  }
}

static __global__ void __launch_bounds__(256)
    ds_finish_kernel(DsWork w, int C, int reduce_type, const int32_t *__restrict__ num_voxels,
                     const int32_t *__restrict__ voxel_count, float *__restrict__ voxel_feats) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t M = *num_voxels;
  if (t >= M * C) return;
  const float s = (float)w.acc[t];
  voxel_feats[t] = (reduce_type == RD3_REDUCE_MEAN) ? __fdiv_rn(s, (float)__ldg(voxel_count + t / C)) : s;
}

// valid-row coordinate maxima (+1) per column -> extent[3]
static __global__ void __launch_bounds__(256)
    ds_extent_kernel(const int32_t *__restrict__ coors, int64_t N, int32_t *extent) {
  int m0 = 0, m1 = 0, m2 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c0 = __ldg(coors + i * 3), c1 = __ldg(coors + i * 3 + 1), c2 = __ldg(coors + i * 3 + 2);
    if (c0 < 0 || c1 < 0 || c2 < 0) continue;
    m0 = max(m0, c0 + 1); m1 = max(m1, c1 + 1); m2 = max(m2, c2 + 1);
  }
  for (int d = 16; d > 0; d >>= 1) {
    m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, d));
    m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, d));
    m2 = max(m2, __shfl_xor_sync(0xffffffffu, m2, d));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(extent + 0, m0);
    atomicMax(extent + 1, m1);
    atomicMax(extent + 2, m2);
  }
}

// ---- backward (scatter_points_cuda.cu:105-179,241-308) ---------------------
static __global__ void __launch_bounds__(256)
    ds_bwd_add_kernel(float *__restrict__ grad_feats, const float *__restrict__ grad_voxel,
                      const int32_t *__restrict__ map, const int32_t *__restrict__ count, int64_t N,
                      int C, int reduce_type) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * C) return;
  const int64_t i = t / C;
  const int c = (int)(t - i * C);
  const int r = __ldg(map + i);
  float g = 0.0f;
  if (r >= 0) {
    g = __ldg(grad_voxel + (int64_t)r * C + c);
    if (reduce_type == RD3_REDUCE_MEAN) g = __fdiv_rn(g, (float)__ldg(count + r));
  }
  grad_feats[t] = g;
}

static __global__ void __launch_bounds__(256)
    ds_bwd_fill_kernel(int32_t *reduce_from, int64_t n, int32_t v) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) reduce_from[t] = v;
}

static __global__ void __launch_bounds__(256)
    ds_bwd_argmax_kernel(const float *__restrict__ feats, const float *__restrict__ voxel_feats,
                         const int32_t *__restrict__ map, int32_t *reduce_from, int64_t N, int C) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * C) return;
  const int64_t i = t / C;
  const int c = (int)(t - i * C);
  const int r = __ldg(map + i);
  if (r < 0) return;
  if (__ldg(feats + t) == __ldg(voxel_feats + (int64_t)r * C + c))
    atomicMin(reduce_from + (int64_t)r * C + c, (int32_t)i);
}

static __global__ void __launch_bounds__(256)
    ds_bwd_route_kernel(float *__restrict__ grad_feats, const float *__restrict__ grad_voxel,
                        const int32_t *__restrict__ reduce_from, int64_t M, int64_t N, int C) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * C) return;
  const int c = (int)(t % C);
  const int32_t src = __ldg(reduce_from + t);
  if (src >= 0 && src < N) grad_feats[(int64_t)src * C + c] = __ldg(grad_voxel + t);
}

struct DsPlan {
  int nwords, nchunks;
  size_t off_bitmap, off_prefix, off_chunk, off_acc, total;
};

static int ds_plan(int64_t N, int C, const int32_t dims[3], DsPlan *p) {
  if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return RD3_ERR_INVALID_ARGUMENT;
  uint64_t vol = (uint64_t)dims[0] * (uint64_t)dims[1];
  if (vol > 0xFFFFFFFEull) return RD3_ERR_UNSUPPORTED;
  vol *= (uint64_t)dims[2];
  if (vol > 0xFFFFFFFEull) return RD3_ERR_UNSUPPORTED;
  const int64_t words = (int64_t)((vol + 31) / 32);
  p->nchunks = (int)ceil_div(words, kChunkWords);
  p->nwords = p->nchunks * kChunkWords;
  size_t off = 0;
  p->off_bitmap = off; off += align_up((size_t)p->nwords * 4);
  p->off_prefix = off; off += align_up((size_t)p->nwords * 4);
  p->off_chunk = off; off += align_up((size_t)p->nchunks * 4);
  p->off_acc = off; off += align_up((size_t)(N > 0 ? N : 1) * C * 8);
  p->total = off;
  return RD3_OK;
}

}  // namespace rd3

using namespace rd3;

extern "C" {

int rd3_coors_extent(const int32_t *coors, int64_t N, int32_t *d_extent3, rd3_stream_t stream) {
  if (N < 0 || !d_extent3) return RD3_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  RD3_CUDA_TRY(cudaMemsetAsync(d_extent3, 0, 3 * sizeof(int32_t), s));
  if (N == 0) return RD3_OK;
  if (!coors) return RD3_ERR_INVALID_ARGUMENT;
  const unsigned blocks = (unsigned)(ceil_div(N, 256) < 148 * 8 ? ceil_div(N, 256) : 148 * 8);
  ds_extent_kernel<<<blocks, 256, 0, s>>>(coors, N, d_extent3);
  return check_launch();
}

size_t rd3_dynamic_scatter_workspace_bytes(int64_t N, int C, const int32_t dims[3]) {
  DsPlan p;
  if (N < 0 || C <= 0 || !dims || ds_plan(N, C, dims, &p) != RD3_OK) return 0;
  return p.total;
}

int rd3_dynamic_scatter_forward(const float *feats, const int32_t *coors, int64_t N, int C,
                                const int32_t dims[3], int reduce_type, float *voxel_feats,
                                int32_t *voxel_coors, int32_t *point2voxel, int32_t *voxel_count,
                                int32_t *d_num_voxels, int32_t *d_status, void *workspace,
                                size_t workspace_bytes, rd3_stream_t stream) {
  if (N < 0 || C <= 0 || !dims || !d_num_voxels || !d_status || !workspace)
    return RD3_ERR_INVALID_ARGUMENT;
  if (reduce_type != RD3_REDUCE_SUM && reduce_type != RD3_REDUCE_MEAN && reduce_type != RD3_REDUCE_MAX)
    return RD3_ERR_INVALID_ARGUMENT;
  if (N >= ((int64_t)1 << 31)) return RD3_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  RD3_CUDA_TRY(cudaMemsetAsync(d_num_voxels, 0, sizeof(int32_t), s));
  RD3_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), s));
  if (N == 0) return RD3_OK;
  if (!feats || !coors || !voxel_feats || !voxel_coors || !point2voxel || !voxel_count)
    return RD3_ERR_INVALID_ARGUMENT;
  DsPlan plan;
  int st = ds_plan(N, C, dims, &plan);
  if (st != RD3_OK) return st;
  if (workspace_bytes < plan.total) return RD3_ERR_WORKSPACE;
  char *base = (char *)workspace;
  DsWork w;
  w.bitmap = (uint32_t *)(base + plan.off_bitmap);
  w.wordprefix = (int32_t *)(base + plan.off_prefix);
  w.chunk_base = (int32_t *)(base + plan.off_chunk);
  w.acc = (double *)(base + plan.off_acc);
  w.nwords = plan.nwords; w.nchunks = plan.nchunks;
  w.d0 = (uint32_t)dims[0]; w.d1 = (uint32_t)dims[1]; w.d2 = (uint32_t)dims[2];

  RD3_CUDA_TRY(cudaMemsetAsync(w.bitmap, 0, (size_t)plan.nwords * 4, s));
  const unsigned pblocks = (unsigned)ceil_div(N, 256);
  ds_mark_kernel<<<pblocks, 256, 0, s>>>(coors, N, w, d_status);
  ds_scan_kernel<<<plan.nchunks, kScanThreads, 0, s>>>(w);
  scan_chunks_kernel<<<1, 1024, 0, s>>>(w.chunk_base, plan.nchunks, d_num_voxels, 0x7FFFFFFF);
  ds_voxels_kernel<<<(unsigned)ceil_div(plan.nwords, 256), 256, 0, s>>>(w, C, reduce_type, voxel_coors,
                                                                       voxel_count, voxel_feats);
  ds_reduce_kernel<<<pblocks, 256, 0, s>>>(feats, coors, N, C, reduce_type, w, point2voxel,
                                           voxel_count, voxel_feats);
  if (reduce_type != RD3_REDUCE_MAX)
    ds_finish_kernel<<<(unsigned)ceil_div(N * C, 256), 256, 0, s>>>(w, C, reduce_type, d_num_voxels,
                                                                    voxel_count, voxel_feats);
  return check_launch();
}

size_t rd3_dynamic_scatter_backward_workspace_bytes(int64_t M, int C) {
  if (M < 0 || C <= 0) return 0;
  return align_up((size_t)(M > 0 ? M : 1) * C * 4);
}

int rd3_dynamic_scatter_backward(float *grad_feats, const float *grad_voxel_feats, const float *feats,
                                 const float *voxel_feats, const int32_t *point2voxel,
                                 const int32_t *voxel_count, int64_t N, int64_t M, int C,
                                 int reduce_type, void *workspace, size_t workspace_bytes,
                                 rd3_stream_t stream) {
  if (N < 0 || M < 0 || C <= 0) return RD3_ERR_INVALID_ARGUMENT;
  if (reduce_type != RD3_REDUCE_SUM && reduce_type != RD3_REDUCE_MEAN && reduce_type != RD3_REDUCE_MAX)
    return RD3_ERR_INVALID_ARGUMENT;
  if (N == 0) return RD3_OK;
  if (!grad_feats) return RD3_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  if (M == 0) {                                   // scatter_points_cuda.cu:259-262
    RD3_CUDA_TRY(cudaMemsetAsync(grad_feats, 0, (size_t)N * C * 4, s));
    return RD3_OK;
  }
  if (!grad_voxel_feats || !point2voxel || !voxel_count) return RD3_ERR_INVALID_ARGUMENT;
  const unsigned nblocks = (unsigned)ceil_div(N * C, 256);
  if (reduce_type != RD3_REDUCE_MAX) {
    ds_bwd_add_kernel<<<nblocks, 256, 0, s>>>(grad_feats, grad_voxel_feats, point2voxel, voxel_count,
                                              N, C, reduce_type);
    return check_launch();
  }
  if (!feats || !voxel_feats || !workspace) return RD3_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < rd3_dynamic_scatter_backward_workspace_bytes(M, C)) return RD3_ERR_WORKSPACE;
  int32_t *reduce_from = (int32_t *)workspace;
  const unsigned mblocks = (unsigned)ceil_div(M * C, 256);
  RD3_CUDA_TRY(cudaMemsetAsync(grad_feats, 0, (size_t)N * C * 4, s));
  ds_bwd_fill_kernel<<<mblocks, 256, 0, s>>>(reduce_from, M * C, (int32_t)N);
  ds_bwd_argmax_kernel<<<nblocks, 256, 0, s>>>(feats, voxel_feats, point2voxel, reduce_from, N, C);
  ds_bwd_route_kernel<<<mblocks, 256, 0, s>>>(grad_feats, grad_voxel_feats, reduce_from, M, N, C);
  return check_launch();
}

}  // extern "C"
