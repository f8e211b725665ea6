// hard_voxel.cuh -- deterministic hard voxelization as a parallel pipeline.
//
// Reference semantics: the sequential scan of
// mmdetection3d/mmdet3d/ops/voxel/src/voxelization_cpu.cpp:45-101
//   * voxel ids in order of each voxel's first point           (:75-88)
//   * a NEW voxel is dropped once max_voxels exist             (:80)
//   * a voxel keeps its first max_points points in point order (:91-97)
// None of the reference's GPU formulation (O(N^2) scan + <<<1,1>>> kernel,
// voxelization_cuda.cu:105-180) is reused.  Everything order-related is derived
// from explicit point indices, so the result does not depend on scheduling:
//
//   K1 insert  : voxel key -> hash/direct table entry {key, min point index}
//                (64-bit atomicMin; read-before-atomic skips most atomics)
//   K2 flags   : point i is a voxel's FIRST point iff entry.min == i; one ballot
//                word per 32 points + per-chunk exclusive popcount scan
//   K2s        : scan of chunk totals (one CTA per frame) -> voxel_num
//   K3 slots   : voxel rank r = #first-points before entry.min; if r < max_voxels
//                insert i into the sorted per-voxel slot array S[r][0..K) with a
//                cascade of atomicMin (keeps the K smallest indices, sorted)
//   K4 emit    : voxels[r][k][:] = features(S[r][k]) (zeros for empty slots)
//   K5 meta    : coors[r], num_points[r], fused HardSimpleVFE mean[r]
//
// The pipeline is templated on the point source: a (N,C) point array, or DA3
// depth maps unprojected on the fly (the point cloud never exists in memory).
#pragma once
#include "rd3_common.cuh"

namespace rd3 {

// ---------------------------------------------------------------------------
// point sources
// ---------------------------------------------------------------------------
struct PointsSource {
  const float *pts;   // (B, N, C)
  int64_t N;
  int C;
  static constexpr bool kNeedsSmem = false;

  __device__ __forceinline__ void prepare(float *, int) const {}
  // load xyz of point i of frame b; false if the point does not exist
  __device__ __forceinline__ bool load(int b, int64_t i, const float *, float &x, float &y,
                                       float &z) const {
    const float *p = pts + ((int64_t)b * N + i) * C;
    x = __ldg(p);
    y = __ldg(p + 1);
    z = __ldg(p + 2);
    return true;
  }
  __host__ __device__ __forceinline__ int num_feats() const { return C; }
  int host_num_feats() const { return C; }
  __device__ __forceinline__ void feats3(int b, int64_t i, const float *, float &x, float &y,
                                         float &z) const {
    load(b, i, nullptr, x, y, z);
  }
  __device__ __forceinline__ float feat(int b, int64_t i, int c, const float *) const {
    return __ldg(pts + ((int64_t)b * N + i) * C + c);
  }
};

struct DepthSource {
  const float *depth;      // (B, npix)
  const float *conf;       // (B, npix) or null
  const uint8_t *sky;      // (B, npix) or null
  const float *intr;       // (B, ncam, 9)
  const float *c2l;        // (B, ncam, 16)
  DepthParams p;
  static constexpr bool kNeedsSmem = true;

  __device__ __forceinline__ void prepare(float *s_cal, int b) const {
    stage_calibration(s_cal, intr + (int64_t)b * p.ncam * 9, c2l + (int64_t)b * p.ncam * 16, p.ncam);
    __syncthreads();
  }
  __device__ __forceinline__ bool load(int b, int64_t i, const float *s_cal, float &x, float &y,
                                       float &z) const {
    const int64_t g = (int64_t)b * p.npix + i;
    const float d = __ldg(depth + g);
    const float cf = p.use_conf ? __ldg(conf + g) : 0.0f;
    const bool sk = p.use_sky ? (__ldg(sky + g) != 0) : false;
    const int pix = (int)i;
    const int cam = pix / p.HW;
    const int rem = pix - cam * p.HW;
    const int v = rem / p.W;
    const int u = rem - v * p.W;
    return unproject_pixel(d, cf, sk, u, v, s_cal + cam * kCalibFloats, p, x, y, z);
  }
  __host__ __device__ __forceinline__ int num_feats() const { return 3; }
  int host_num_feats() const { return 3; }
  __device__ __forceinline__ void feats3(int b, int64_t i, const float *s_cal, float &x, float &y,
                                         float &z) const {
    load(b, i, s_cal, x, y, z);
  }
  __device__ __forceinline__ float feat(int b, int64_t i, int c, const float *s_cal) const {
    float x, y, z;
    load(b, i, s_cal, x, y, z);
    return c == 0 ? x : (c == 1 ? y : z);
  }
};

// ---------------------------------------------------------------------------
// per-call work description (device pointers, all with a leading frame dim)
// ---------------------------------------------------------------------------
struct HvWork {
  unsigned long long *table;  // [B][cap]        {key:32 | min point idx:32}, empty = ~0
  int32_t *pslot;             // [B][N]          table slot of each point, -1 if outside
  uint32_t *flags;            // [B][nwords]     bit i: point i is the first of its voxel
  int32_t *wordprefix;        // [B][nwords]     exclusive popcount prefix inside the chunk
  int32_t *chunk_base;        // [B][nchunks]    totals, then exclusive bases after K2s
  uint32_t *slots;            // [B][max_voxels*K] sorted point indices, empty = ~0
  int64_t N;
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
  int32_t *point2voxel;   // [B][N] or null
  int F;
};

__device__ __forceinline__ uint32_t hash_key(uint32_t key, int log2cap) {
  return (key * 2654435769u) >> (32 - log2cap);
}

// Insert (key, idx); returns the table slot of the key.  Slot ownership is
// permanent (CAS from empty); the payload only ever decreases (atomicMin), so
// a stale read can only cause a redundant atomic, never a wrong skip.
__device__ __forceinline__ uint32_t table_insert(unsigned long long *table, const HvWork &w,
                                                 uint32_t key, uint32_t idx) {
  uint32_t slot = w.direct ? key : hash_key(key, w.log2cap);
  const unsigned long long mine = ((unsigned long long)key << 32) | idx;
  while (true) {
    unsigned long long e = __ldcg(table + slot);
    if (e == kEmpty64) {
      const unsigned long long old = atomicCAS(table + slot, kEmpty64, mine);
      if (old == kEmpty64) return slot;
      e = old;
    }
    if ((uint32_t)(e >> 32) == key) {
      if ((uint32_t)e > idx) atomicMin(table + slot, mine);
      return slot;
    }
    slot = (slot + 1) & w.cap_mask;
  }
}

// K1 ------------------------------------------------------------------------
template <class Src>
__global__ void __launch_bounds__(256) hv_insert_kernel(Src src, VoxelGrid g, HvWork w) {
  __shared__ float s_cal[Src::kNeedsSmem ? kMaxCams * kCalibFloats : 1];
  const int b = blockIdx.y;
  src.prepare(s_cal, b);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w.N) return;
  float x, y, z;
  int32_t ps = -1;
  if (src.load(b, i, s_cal, x, y, z)) {
    int cx, cy, cz;
    if (voxel_coor(x, y, z, g, cx, cy, cz)) {
      ps = (int32_t)table_insert(w.table + (int64_t)b * w.cap, w, voxel_key(cx, cy, cz, g), (uint32_t)i);
    }
  }
  w.pslot[(int64_t)b * w.N + i] = ps;
}

// Shared tail of the ordered-flag kernels: lane L of warp wv holds the 32-bit flag
// word (chunk*kChunkWords + wv*32 + L).  Stores the word, its exclusive popcount
// prefix inside the chunk, and the chunk total.
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

// K2 ------------------------------------------------------------------------
// grid (nchunks, B), 256 threads: warp wv owns words [wv*32, wv*32+32) of the chunk.
static __global__ void __launch_bounds__(kScanThreads) hv_flags_kernel(HvWork w) {
  __shared__ int s_warp[kScanThreads / 32];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const unsigned long long *table = w.table + (int64_t)b * w.cap;
  const int32_t *pslot = w.pslot + (int64_t)b * w.N;
  const int word0 = blockIdx.x * kChunkWords + wv * 32;
  uint32_t my_word = 0;
#pragma unroll 4
  for (int it = 0; it < 32; ++it) {
    const int64_t i = ((int64_t)(word0 + it) << 5) + lane;
    bool first = false;
    if (i < w.N) {
      const int32_t ps = __ldg(pslot + i);
      if (ps >= 0) first = ((uint32_t)__ldg(table + ps) == (uint32_t)i);
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, first);
    if (lane == it) my_word = bal;
  }
  chunk_scan_store(my_word, s_warp, w.flags + (int64_t)b * w.nwords, w.wordprefix + (int64_t)b * w.nwords,
                   w.chunk_base + (int64_t)b * w.nchunks);
}

// K2s -----------------------------------------------------------------------
// one CTA per frame: exclusive scan of the chunk totals in place; the frame's
// total (clamped) goes to out_total[b].
static __global__ void __launch_bounds__(1024) scan_chunks_kernel(int32_t *chunk_base, int nchunks,
                                                           int32_t *out_total, int clamp) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int b = blockIdx.x;
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
  const uint32_t bits = __ldg(w.flags + wi) & ((1u << (first_idx & 31)) - 1u);
  return __ldg(w.chunk_base + (int64_t)b * w.nchunks + (first_idx >> kChunkShift)) +
         __ldg(w.wordprefix + wi) + __popc(bits);
}

// Keep the K smallest point indices of a voxel, sorted, with atomicMin only.
// Invariant: a value reaches slot k only after losing against slots < k, so the
// non-empty prefix is strictly increasing at all times; the final content is
// independent of arrival order.
__device__ __forceinline__ void slot_insert(uint32_t *S, int K, uint32_t idx) {
  if (__ldcg(S + K - 1) < idx) return;   // K smaller indices already present
  uint32_t cur = idx;
  for (int k = 0; k < K; ++k) {
    const uint32_t s = __ldcg(S + k);
    if (s < cur) continue;               // stale reads are larger: conservative
    const uint32_t old = atomicMin(S + k, cur);
    if (old == kEmpty32) return;
    if (old > cur) cur = old;            // displaced a larger index: carry it on
  }
}

// K3 ------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) hv_slots_kernel(HvWork w, int32_t *point2voxel) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w.N) return;
  const int32_t ps = __ldg(w.pslot + (int64_t)b * w.N + i);
  int r = -1;
  if (ps >= 0) {
    const uint32_t first_idx = (uint32_t)__ldg(w.table + (int64_t)b * w.cap + ps);
    r = voxel_rank(w, b, first_idx);
    if (r < w.max_voxels) {
      slot_insert(w.slots + ((int64_t)b * w.max_voxels + r) * w.K, w.K, (uint32_t)i);
    } else {
      r = -1;
    }
  }
  if (point2voxel) point2voxel[(int64_t)b * w.N + i] = r;
}

// K4 ------------------------------------------------------------------------
// one thread per output float of voxels[b][r][k][c]; coalesced stores.
template <class Src>
__global__ void __launch_bounds__(256) hv_emit_kernel(Src src, HvWork w, HvOut o) {
  __shared__ float s_cal[Src::kNeedsSmem ? kMaxCams * kCalibFloats : 1];
  const int b = blockIdx.y;
  const int C = src.num_feats();
  const int64_t per_frame = (int64_t)w.max_voxels * w.K * C;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int vn = o.voxel_num[b];
  // whole block beyond the last valid voxel: nothing to do (uniform exit)
  if ((int64_t)blockIdx.x * blockDim.x >= (int64_t)vn * w.K * C) return;
  src.prepare(s_cal, b);
  if (e >= (int64_t)vn * w.K * C) return;
  const int64_t rk = e / C;
  const int c = (int)(e - rk * C);
  const uint32_t idx = __ldg(w.slots + (int64_t)b * w.max_voxels * w.K + rk);
  float v = 0.0f;
  if (idx != kEmpty32) v = src.feat(b, idx, c, s_cal);
  o.voxels[(int64_t)b * per_frame + e] = v;
}

// K5 ------------------------------------------------------------------------
// one thread per voxel: coors (from its first point), count, HardSimpleVFE mean
// (voxel_encoder.py:45-46: sum over ALL slots in slot order, then one division).
template <class Src>
__global__ void __launch_bounds__(256) hv_meta_kernel(Src src, VoxelGrid g, HvWork w, HvOut o) {
  __shared__ float s_cal[Src::kNeedsSmem ? kMaxCams * kCalibFloats : 1];
  const int b = blockIdx.y;
  const int vn = o.voxel_num[b];
  if ((int64_t)blockIdx.x * blockDim.x >= vn) return;
  src.prepare(s_cal, b);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= vn) return;
  const uint32_t *S = w.slots + ((int64_t)b * w.max_voxels + r) * w.K;
  const int64_t vr = (int64_t)b * w.max_voxels + r;
  int cnt = 0;
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  const int F = o.F;
  for (int k = 0; k < w.K; ++k) {
    const uint32_t idx = __ldg(S + k);
    if (idx == kEmpty32) break;
    ++cnt;
    float x, y, z;
    src.feats3(b, idx, s_cal, x, y, z);
    if (k == 0) {
      int cx, cy, cz;
      voxel_coor(x, y, z, g, cx, cy, cz);
      o.coors[vr * 3 + 0] = cz;
      o.coors[vr * 3 + 1] = cy;
      o.coors[vr * 3 + 2] = cx;
    }
    if (o.mean) {
      sx = __fadd_rn(sx, x);
      sy = __fadd_rn(sy, y);
      sz = __fadd_rn(sz, z);
      // features beyond xyz (points source with C > 3)
      for (int f = 3; f < F; ++f) {
        float *m = o.mean + vr * F + f;
        const float a = src.feat(b, idx, f, s_cal);
        *m = (k == 0) ? a : __fadd_rn(*m, a);
      }
    }
  }
  o.num[vr] = cnt;
  if (o.mean) {
    const float n = (float)cnt;
    if (F > 0) o.mean[vr * F + 0] = __fdiv_rn(sx, n);
    if (F > 1) o.mean[vr * F + 1] = __fdiv_rn(sy, n);
    if (F > 2) o.mean[vr * F + 2] = __fdiv_rn(sz, n);
    for (int f = 3; f < F; ++f) o.mean[vr * F + f] = __fdiv_rn(o.mean[vr * F + f], n);
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
  int64_t cap;
  int log2cap;
  int nwords, nchunks;
  size_t off_table, off_slots, off_pslot, off_flags, off_prefix, off_chunk, total;
};

inline HvPlan hv_plan(int64_t N, int B, int K, int max_voxels) {
  HvPlan p;
  p.N = N; p.B = B; p.K = K; p.max_voxels = max_voxels;
  // capacity: power of two >= 1.25 N (worst case: every point its own voxel -> load <= 0.8)
  int64_t want = N + N / 4 + 1;
  int lg = 10;
  while (((int64_t)1 << lg) < want) ++lg;
  p.log2cap = lg;
  p.cap = (int64_t)1 << lg;
  p.nchunks = (int)ceil_div(N > 0 ? N : 1, kChunkPoints);
  p.nwords = p.nchunks * kChunkWords;
  size_t off = 0;
  // table and slots are adjacent: one memset(0xFF) initialises both
  p.off_table = off; off += align_up((size_t)B * p.cap * 8);
  p.off_slots = off; off += align_up((size_t)B * max_voxels * K * 4);
  p.off_pslot = off; off += align_up((size_t)B * (N > 0 ? N : 1) * 4);
  p.off_flags = off; off += align_up((size_t)B * p.nwords * 4);
  p.off_prefix = off; off += align_up((size_t)B * p.nwords * 4);
  p.off_chunk = off; off += align_up((size_t)B * p.nchunks * 4);
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
  w.pslot = (int32_t *)(base + p.off_pslot);
  w.flags = (uint32_t *)(base + p.off_flags);
  w.wordprefix = (int32_t *)(base + p.off_prefix);
  w.chunk_base = (int32_t *)(base + p.off_chunk);
  w.N = p.N; w.cap = p.cap; w.cap_mask = (uint32_t)(p.cap - 1); w.log2cap = p.log2cap;
  w.direct = (volume <= (uint64_t)p.cap) ? 1 : 0;
  w.nwords = p.nwords; w.nchunks = p.nchunks; w.K = p.K; w.max_voxels = p.max_voxels;

  RD3_CUDA_TRY(cudaMemsetAsync(base + p.off_table, 0xFF, p.off_pslot - p.off_table, stream));
  if (p.N > 0) {
    dim3 gp((unsigned)ceil_div(p.N, 256), p.B);
    hv_insert_kernel<Src><<<gp, 256, 0, stream>>>(src, g, w);
  }
  hv_flags_kernel<<<dim3(p.nchunks, p.B), kScanThreads, 0, stream>>>(w);
  scan_chunks_kernel<<<p.B, 1024, 0, stream>>>(w.chunk_base, w.nchunks, out.voxel_num, w.max_voxels);
  if (p.N > 0) {
    dim3 gp((unsigned)ceil_div(p.N, 256), p.B);
    hv_slots_kernel<<<gp, 256, 0, stream>>>(w, out.point2voxel);
  }
  const int C = src.host_num_feats();
  const int64_t elems = (int64_t)p.max_voxels * p.K * C;
  if (elems > 0) {
    hv_emit_kernel<Src><<<dim3((unsigned)ceil_div(elems, 256), p.B), 256, 0, stream>>>(src, w, out);
    hv_meta_kernel<Src><<<dim3((unsigned)ceil_div(p.max_voxels, 256), p.B), 256, 0, stream>>>(src, g, w, out);
  }
  return check_launch();
}

}  // namespace rd3
