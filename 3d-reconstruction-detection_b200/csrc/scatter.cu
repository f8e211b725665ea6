// scatter.cu -- DynamicScatter forward / backward.
//
// Reference: mmdetection3d/mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-308
// (at::unique_dim -- a multi-pass row sort -- plus one fp32 atomic per feature per
// point) and the per-sample Python loop of scatter_points.py:86-97.  Here the
// sorted-unique step needs no sort: valid coordinates are bounded by `dims`, so a
// voxel's lexicographic rank is the number of occupied cells before its linear id,
// i.e. a popcount prefix over an occupancy bitmap (83 M cells = 10.4 MB for the
// 1440x1440x40 grid, L2 resident on B200).  A batch column is just the slowest
// dimension of the key, so a whole batch is ONE launch sequence and its output is
// the reference's concatenation in batch order.
//
//   D1 mark    : bitmap |= cell(point); a second bitmap marks the cells that receive
//                more than one point; the point's key is kept (4 B) for D4
//   D2 scan    : per-chunk exclusive popcount prefix; D2s: chunk totals -> M
//   D3 voxels  : every occupied cell writes its coordinates at its rank; cells with
//                several points initialise their accumulators
//   D4 reduce  : rank per point -> point2voxel.  The only point of a cell stores its
//                features as the result (no atomic, no accumulator).  Points of shared
//                cells: lanes of a warp holding the same cell form a segment
//                (__match_any_sync), the segment is reduced with shuffles by its first
//                lane, which issues ONE atomic per feature (fp64 add: exact for
//                same-magnitude fp32 inputs hence order independent; max: ordered-int)
//   D5 finish  : shared cells: sum -> fp32, mean = fp32(sum) / fp32(count)   (:233-234)
#include "hard_voxel.cuh"

namespace rd3 {

struct DsWork {
  uint32_t *bitmap;       // [nwords] occupied cells
  uint32_t *multi;        // [nwords] cells with more than one point
  int32_t *wordprefix;    // [nwords]
  int32_t *chunk_base;    // [nchunks]
  uint32_t *keys;         // [N] cell of every point (kEmpty32: dropped)
  double *acc;            // [N*C] (only the rows of shared cells are used)
  int nwords, nchunks;
  int ncols;              // 3: (z,y,x)   4: (batch,z,y,x)
  uint32_t d[4];          // exclusive bounds, d[0] = 1 when ncols == 3
  const int32_t *last_row;   // ncols == 4: the last row of coors -- the reference takes batch_size = coors[-1, 0] + 1
                             // (scatter_points.py:86) and never looks at rows of a later batch
};

// cell of point i; false: a negative coordinate (dropped, scatter_points_cuda.cu:202) or beyond dims (overflow)
__device__ __forceinline__ bool ds_key(const int32_t *__restrict__ coors, int64_t i, const DsWork &w,
                                       uint32_t &key, bool &overflow) {
  int32_t c[4];
  c[0] = 0;
  if (w.ncols == 4) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(coors) + i);
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
  } else {
    c[1] = __ldg(coors + i * 3); c[2] = __ldg(coors + i * 3 + 1); c[3] = __ldg(coors + i * 3 + 2);
  }
  overflow = false;
  if (w.ncols == 4) {
    const int32_t bs = __ldg(w.last_row) + 1;
    if (bs > (int32_t)w.d[0]) overflow = true;             // more samples than the bitmap was sized for
    if (c[0] >= bs) return false;                          // the reference's loop stops at coors[-1, 0]
  }
  if ((c[0] | c[1] | c[2] | c[3]) < 0) return false;
  if (overflow) return false;
  if ((uint32_t)c[0] >= w.d[0] || (uint32_t)c[1] >= w.d[1] || (uint32_t)c[2] >= w.d[2] || (uint32_t)c[3] >= w.d[3]) {
    overflow = true;
    return false;
  }
  key = (((uint32_t)c[0] * w.d[1] + (uint32_t)c[1]) * w.d[2] + (uint32_t)c[2]) * w.d[3] + (uint32_t)c[3];
  return true;
}

static __global__ void __launch_bounds__(256)
    ds_mark_kernel(const int32_t *__restrict__ coors, int64_t N, DsWork w, int32_t *status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint32_t key = kEmpty32;
  bool ovf;
  if (ds_key(coors, i, w, key, ovf)) {
    uint32_t *word = w.bitmap + (key >> 5);
    const uint32_t bit = 1u << (key & 31);
    // occupied already (plain read, or the atomic's return value): the cell holds more than one point
    bool shared = __ldcg(word) & bit;
    if (!shared) shared = atomicOr(word, bit) & bit;
    if (shared) {
      uint32_t *mw = w.multi + (key >> 5);
      if (!(__ldcg(mw) & bit)) atomicOr(mw, bit);
    }
  } else {
    key = kEmpty32;
    if (ovf) *status = 1;
  }
  w.keys[i] = key;
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

// one thread per bitmap word: emit coordinates of its occupied cells; cells with one point get count 1 (their
// features are stored by that point), shared cells count 0 and cleared accumulators
static __global__ void __launch_bounds__(256)
    ds_voxels_kernel(DsWork w, int C, int reduce_type, int32_t *__restrict__ voxel_coors,
                     int32_t *__restrict__ voxel_count, float *__restrict__ voxel_feats) {
  const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= w.nwords) return;
  uint32_t bits = w.bitmap[wi];
  if (!bits) return;
  const uint32_t shared = w.multi[wi];
  int r = __ldg(w.chunk_base + (wi / kChunkWords)) + __ldg(w.wordprefix + wi);
  while (bits) {
    const int bpos = __ffs(bits) - 1;
    bits &= bits - 1;
    uint32_t key = (uint32_t)(wi << 5) + bpos;
    const uint32_t c3 = key % w.d[3];
    key /= w.d[3];
    const uint32_t c2 = key % w.d[2];
    key /= w.d[2];
    const uint32_t c1 = key % w.d[1], c0 = key / w.d[1];
    if (w.ncols == 4) {
      reinterpret_cast<int4 *>(voxel_coors)[r] = make_int4((int)c0, (int)c1, (int)c2, (int)c3);
    } else {
      voxel_coors[(int64_t)r * 3 + 0] = (int32_t)c1;
      voxel_coors[(int64_t)r * 3 + 1] = (int32_t)c2;
      voxel_coors[(int64_t)r * 3 + 2] = (int32_t)c3;
    }
    const bool sh = (shared >> bpos) & 1u;
    voxel_count[r] = sh ? 0 : 1;
    if (sh) {
      for (int c = 0; c < C; ++c) {
        if (reduce_type == RD3_REDUCE_MAX) voxel_feats[(int64_t)r * C + c] = __int_as_float(0xFF800000);
        else w.acc[(int64_t)r * C + c] = 0.0;
      }
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

// device fmaxf of the reference's reduceMax: NaN operands are dropped, +0.0 orders above -0.0
__device__ __forceinline__ float max_like_reference(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return (__float_as_int(a) >= 0) ? a : b;
  return a > b ? a : b;
}

static __global__ void __launch_bounds__(256)
    ds_reduce_kernel(const float *__restrict__ feats, int64_t N, int C, int reduce_type, DsWork w,
                     int32_t *__restrict__ point2voxel, int32_t *voxel_count, float *voxel_feats) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t key = i < N ? __ldg(w.keys + i) : kEmpty32;
  int r = -1;
  bool shared = false;
  if (key != kEmpty32) {
    r = ds_rank(w, key);
    shared = (__ldg(w.multi + (key >> 5)) >> (key & 31)) & 1u;
  }
  if (i < N) point2voxel[i] = r;
  const float *f = feats + i * C;
  if (r >= 0 && !shared) {                  // the cell's only point IS the result (sum, mean and max alike) ...
    for (int c = 0; c < C; ++c) {
      float v = __ldg(f + c);
      if (reduce_type == RD3_REDUCE_MAX && v != v) v = __int_as_float(0xFF800000);   // ... except fmaxf(NaN, -inf) = -inf
      voxel_feats[(int64_t)r * C + c] = v;
    }
  }
  // segments: lanes of this warp whose points share a cell (other lanes get keys of their own)
  if (!__any_sync(0xffffffffu, shared)) return;
  const unsigned grp = __match_any_sync(0xffffffffu, shared ? key : (0xFFFFFFE0u + lane));
  const int leader = __ffs(grp) - 1;
  const int size = __popc(grp);
  const int maxsize = __reduce_max_sync(0xffffffffu, shared ? size : 1);
  if (shared && lane == leader) atomicAdd(voxel_count + r, size);
  for (int c = 0; c < C; ++c) {
    const float v = shared ? __ldg(f + c) : 0.0f;
    unsigned rest = grp & ~(1u << leader);
    if (reduce_type == RD3_REDUCE_MAX) {
      float m = v;
      for (int j = 1; j < maxsize; ++j) {
        const int src = rest ? __ffs(rest) - 1 : lane;
        const float o = __shfl_sync(0xffffffffu, v, src);
        if (rest) m = max_like_reference(m, o);
        rest &= rest - 1;
      }
      if (shared && lane == leader) atomic_max_float(voxel_feats + (int64_t)r * C + c, m);
    } else {
      double s = (double)v;
      for (int j = 1; j < maxsize; ++j) {
        const int src = rest ? __ffs(rest) - 1 : lane;
        const float o = __shfl_sync(0xffffffffu, v, src);
        if (rest) s += (double)o;
        rest &= rest - 1;
      }
      if (shared && lane == leader) atomicAdd(w.acc + (int64_t)r * C + c, s);
    }
  }
}

static __global__ void __launch_bounds__(256)
    ds_finish_kernel(DsWork w, int C, int reduce_type, const int32_t *__restrict__ num_voxels,
                     const int32_t *__restrict__ voxel_count, float *__restrict__ voxel_feats) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t M = *num_voxels;
  if (t >= M * C) return;
  const int cnt = __ldg(voxel_count + t / C);
  if (cnt <= 1) return;                     // stored by its only point
  const float s = (float)w.acc[t];
  voxel_feats[t] = (reduce_type == RD3_REDUCE_MEAN) ? __fdiv_rn(s, (float)cnt) : s;
}

// valid-row coordinate maxima (+1) per column -> extent[ncols]
static __global__ void __launch_bounds__(256)
    ds_extent_kernel(const int32_t *__restrict__ coors, int64_t N, int ncols, int32_t *extent) {
  int m[4] = {0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    int32_t c[4];
    bool neg = false;
    for (int k = 0; k < ncols; ++k) {
      c[k] = __ldg(coors + i * ncols + k);
      neg |= c[k] < 0;
    }
    if (neg) continue;
    for (int k = 0; k < ncols; ++k) m[k] = max(m[k], c[k] + 1);
  }
  for (int d = 16; d > 0; d >>= 1)
    for (int k = 0; k < 4; ++k) m[k] = max(m[k], __shfl_xor_sync(0xffffffffu, m[k], d));
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < ncols; ++k) atomicMax(extent + k, m[k]);
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
  size_t off_bitmap, off_multi, off_prefix, off_chunk, off_keys, off_acc, total;
};

static int ds_plan(int64_t N, int C, int ncols, const int32_t dims[4], DsPlan *p) {
  if (ncols != 3 && ncols != 4) return RD3_ERR_INVALID_ARGUMENT;
  uint64_t vol = 1;
  for (int k = (ncols == 4 ? 0 : 1); k < 4; ++k) {
    if (dims[k] <= 0) return RD3_ERR_INVALID_ARGUMENT;
    vol *= (uint64_t)dims[k];
    if (vol > (uint64_t)kDummyKey) return RD3_ERR_UNSUPPORTED;     // keys are 32 bit; the top 32 values are lane dummies
  }
  const int64_t words = (int64_t)((vol + 31) / 32);
  p->nchunks = (int)ceil_div(words, kChunkWords);
  p->nwords = p->nchunks * kChunkWords;
  size_t off = 0;
  p->off_bitmap = off; off += align_up((size_t)p->nwords * 4);      // [bitmap | multi] are cleared with one memset
  p->off_multi = off; off += align_up((size_t)p->nwords * 4);
  p->off_prefix = off; off += align_up((size_t)p->nwords * 4);
  p->off_chunk = off; off += align_up((size_t)p->nchunks * 4);
  p->off_keys = off; off += align_up((size_t)(N > 0 ? N : 1) * 4);
  p->off_acc = off; off += align_up((size_t)(N > 0 ? N : 1) * C * 8);
  p->total = off;
  return RD3_OK;
}

}  // namespace rd3

using namespace rd3;

extern "C" {

int rd3_coors_extent(const int32_t *coors, int64_t N, int ncols, int32_t *d_extent4, rd3_stream_t stream) {
  if (N < 0 || !d_extent4 || (ncols != 3 && ncols != 4)) return RD3_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  RD3_CUDA_TRY(cudaMemsetAsync(d_extent4, 0, 4 * sizeof(int32_t), s));
  if (N == 0) return RD3_OK;
  if (!coors) return RD3_ERR_INVALID_ARGUMENT;
  const int64_t want = ceil_div(N, 256), cap = (int64_t)hv_tuning().sm_count * 8;
  ds_extent_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(coors, N, ncols, d_extent4);
  return check_launch();
}

size_t rd3_dynamic_scatter_workspace_bytes(int64_t N, int C, int ncols, const int32_t dims[4]) {
  DsPlan p;
  if (N < 0 || C <= 0 || !dims || ds_plan(N, C, ncols, dims, &p) != RD3_OK) return 0;
  return p.total;
}

int rd3_dynamic_scatter_forward(const float *feats, const int32_t *coors, int64_t N, int C, int ncols,
                                const int32_t dims[4], int reduce_type, float *voxel_feats,
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
  int st = ds_plan(N, C, ncols, dims, &plan);
  if (st != RD3_OK) return st;
  if (workspace_bytes < plan.total) return RD3_ERR_WORKSPACE;
  if (ncols == 4 && ((reinterpret_cast<uintptr_t>(coors) | reinterpret_cast<uintptr_t>(voxel_coors)) & 15))
    return RD3_ERR_INVALID_ARGUMENT;                       // (N, 4) rows are moved as one 16-byte word
  char *base = (char *)workspace;
  DsWork w;
  w.bitmap = (uint32_t *)(base + plan.off_bitmap);
  w.multi = (uint32_t *)(base + plan.off_multi);
  w.wordprefix = (int32_t *)(base + plan.off_prefix);
  w.chunk_base = (int32_t *)(base + plan.off_chunk);
  w.keys = (uint32_t *)(base + plan.off_keys);
  w.acc = (double *)(base + plan.off_acc);
  w.nwords = plan.nwords; w.nchunks = plan.nchunks;
  w.ncols = ncols;
  w.d[0] = ncols == 4 ? (uint32_t)dims[0] : 1u;
  w.d[1] = (uint32_t)dims[1]; w.d[2] = (uint32_t)dims[2]; w.d[3] = (uint32_t)dims[3];
  w.last_row = coors + (N - 1) * 4;

  RD3_CUDA_TRY(cudaMemsetAsync(w.bitmap, 0, plan.off_prefix - plan.off_bitmap, s));
  const unsigned pblocks = (unsigned)ceil_div(N, 256);
  ds_mark_kernel<<<pblocks, 256, 0, s>>>(coors, N, w, d_status);
  ds_scan_kernel<<<plan.nchunks, kScanThreads, 0, s>>>(w);
  scan_chunks_kernel<<<1, 1024, 0, s>>>(w.chunk_base, plan.nchunks, d_num_voxels, 0x7FFFFFFF);
  ds_voxels_kernel<<<(unsigned)ceil_div(plan.nwords, 256), 256, 0, s>>>(w, C, reduce_type, voxel_coors,
                                                                       voxel_count, voxel_feats);
  ds_reduce_kernel<<<pblocks, 256, 0, s>>>(feats, N, C, reduce_type, w, point2voxel, voxel_count, voxel_feats);
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
