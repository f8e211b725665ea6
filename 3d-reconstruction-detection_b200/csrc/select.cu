// select.cu -- the confidence threshold of the depth->points hand-off:
//   conf_thresh = np.percentile(conf[~sky] if (~sky).sum() > 10 else conf.flatten(), p)
//   (tools/inference_nuscenes.py:351-361; depth_anything_3/utils/export/glb.py:227-229)
// per sample, on the device: an exact 2-pass radix select (16 + 16 bits of the order-preserving
// integer image of the fp32 values) of the two order statistics numpy's "linear" method
// interpolates between, then numpy's own index / gamma / lerp arithmetic.  The reference sorts
// (np.partition) 2.7 M values per sample on one CPU core.
#include "rd3_common.cuh"

namespace rd3 {

// Exact selection in TWO passes over the data (16 + 16 bits of the order-preserving integer image of the values):
//   window  a strided sample of each map gives the 16-bit prefix its smallest values have; the 2048 prefixes from
//           there on (16 binades) are the WINDOW that CTAs histogram in shared memory.  The window is only a
//           performance hint: a value outside it is counted with an atomic on the global histogram instead.
//   pass 0  histogram of the top 16 key bits of the selected (non-sky) values: per-CTA window histogram flushed with
//           one atomic per non-empty bin.  Confidences are continuous (conf = 1 + exp(x)), so a warp's 32 values
//           spread over hundreds of window bins -- the 11-bit digits of a classic radix select put them on ~10.
//   pick 0  count -> numpy's two ranks -> their 16-bit prefixes.  n_nonsky <= 10 switches to "all pixels"
//           (inference_nuscenes.py:357-360): pass 0 is then repeated over all pixels (its CTAs return at once otherwise).
//   pass 1  the low 16 bits of the values carrying one of the two prefixes (a few thousand per map): global atomics,
//           aggregated per warp so that a constant map costs one atomic per warp.
//   pick 1  the two order statistics, then numpy's own index / gamma / lerp arithmetic.
constexpr int kSelBins = 65536;         // one 16-bit digit
constexpr int kSelWin = 2048;           // window bins kept in shared memory
constexpr int kSelThreads = 256;
constexpr int kSelPerThread = 64;       // values per thread: 4 iterations of 16
constexpr int kSelChunk = kSelThreads * kSelPerThread;
constexpr int kSelSamples = 4096;

struct SelState {                 // per sample
  uint32_t prefix[2];             // key bits fixed so far, for the two ranks
  int64_t rank[2];                // rank still to locate inside the prefix bucket
  int32_t use_all;                // 1: every pixel, 0: non-sky pixels only
  int32_t n;                      // number of selected values
  double gamma;                   // interpolation weight (index dtype precision)
  int32_t same;                   // both ranks are the same order statistic
  uint32_t win_base;              // first 16-bit prefix of the shared-memory window
  int32_t redo_all;               // pick 0 found <= 10 non-sky pixels: pass 0 has to run again over all pixels
};

__device__ __forceinline__ uint32_t sel_key(float v) {       // monotone: a < b  <=>  key(a) < key(b); NaN last
  const uint32_t u = __float_as_uint(v);
  if (v != v) return 0xFFFFFFFFu;                            // any NaN sorts last (np.partition)
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sel_unkey(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// window base of every map from kSelSamples strided values (any choice is correct; this one covers the data)
static __global__ void __launch_bounds__(256) sel_window_kernel(const float *__restrict__ conf, int64_t npix, SelState *state) {
  __shared__ uint32_t s_min[8];
  const int b = blockIdx.x;
  const int64_t stride = npix / kSelSamples > 0 ? npix / kSelSamples : 1;
  uint32_t m = 0xFFFFu;
  for (int i = threadIdx.x; i < kSelSamples; i += 256) {
    const int64_t j = (int64_t)i * stride;
    if (j < npix) {
      const uint32_t k = sel_key(__ldg(conf + (int64_t)b * npix + j)) >> 16;
      m = k < m ? k : m;
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    const uint32_t o = __shfl_xor_sync(0xffffffffu, m, d);
    m = o < m ? o : m;
  }
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = s_min[i] < m ? s_min[i] : m;
    // two binades of room below the smallest sampled value, the rest of the window above it
    uint32_t base = m > 256u ? m - 256u : 0u;
    if (base > (uint32_t)(kSelBins - kSelWin)) base = kSelBins - kSelWin;
    state[b].win_base = base;
    state[b].redo_all = 0;
    state[b].use_all = 0;
  }
}

// 16 consecutive values of one map: keys and the "selected" bits (non-sky, or every existing value with `all`)
struct SelVals {
  uint32_t key[16];
  uint32_t on;       // bit j: value j exists
  uint32_t ns;       // bit j: value j exists and is not sky
};
template <bool VEC>
__device__ __forceinline__ void sel_load(const float *__restrict__ conf, const uint8_t *__restrict__ sky,
                                         const float *__restrict__ sky_prob, float sky_thr, int64_t base, int64_t i,
                                         int64_t hi, bool want_sky, SelVals &v) {
  v.on = 0; v.ns = 0;
  if (VEC && i + 16 <= hi) {
    const float4 *c4 = reinterpret_cast<const float4 *>(conf + base + i);
    float4 c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q] = __ldg(c4 + q);
    uint32_t skybits = 0;
    if (want_sky && sky) {
      const uint4 s4 = __ldg(reinterpret_cast<const uint4 *>(sky + base + i));
      const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) skybits |= (((sw[q] >> (8 * e)) & 0xFFu) ? 1u : 0u) << (4 * q + e);
    } else if (want_sky && sky_prob) {
      const float4 *p4 = reinterpret_cast<const float4 *>(sky_prob + base + i);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = __ldg(p4 + q);
        skybits |= ((t.x >= sky_thr ? 1u : 0u) | (t.y >= sky_thr ? 2u : 0u) | (t.z >= sky_thr ? 4u : 0u) |
                    (t.w >= sky_thr ? 8u : 0u)) << (4 * q);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v.key[4 * q + 0] = sel_key(c[q].x); v.key[4 * q + 1] = sel_key(c[q].y);
      v.key[4 * q + 2] = sel_key(c[q].z); v.key[4 * q + 3] = sel_key(c[q].w);
    }
    v.on = 0xFFFFu;
    v.ns = 0xFFFFu & ~skybits;
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const bool on = i + j < hi;
      v.key[j] = on ? sel_key(__ldg(conf + base + i + j)) : 0u;
      bool issky = false;
      if (on && want_sky) issky = sky ? __ldg(sky + base + i + j) != 0 : (sky_prob ? __ldg(sky_prob + base + i + j) >= sky_thr : false);
      v.on |= (on ? 1u : 0u) << j;
      v.ns |= ((on && !issky) ? 1u : 0u) << j;
    }
  }
}

// PASS 0 (ALL = false: non-sky values; ALL = true: every value, only for the maps whose pick 0 asked for it).
// hist0: [B][kSelBins].
template <bool VEC, bool ALL>
__global__ void __launch_bounds__(kSelThreads) sel_hist0_kernel(const float *__restrict__ conf,
                                                                const uint8_t *__restrict__ sky,
                                                                const float *__restrict__ sky_prob, float sky_thr,
                                                                int64_t npix, const SelState *__restrict__ state,
                                                                uint32_t *__restrict__ hist0, uint32_t *__restrict__ nanflag) {
  __shared__ uint32_t s_h[kSelWin];
  const int b = blockIdx.y;
  if (ALL && !state[b].redo_all) return;
  const uint32_t wb = state[b].win_base;
  for (int i = threadIdx.x; i < kSelWin; i += kSelThreads) s_h[i] = 0;
  __syncthreads();
  uint32_t *g = hist0 + (int64_t)b * kSelBins;
  const int64_t lo = (int64_t)blockIdx.x * kSelChunk;
  const int64_t hi = lo + kSelChunk < npix ? lo + kSelChunk : npix;
  const int64_t base = (int64_t)b * npix;
  bool nan_sel = false;
  for (int64_t i = lo + (int64_t)threadIdx.x * 16; i < hi; i += (int64_t)kSelThreads * 16) {
    SelVals v;
    sel_load<VEC>(conf, sky, sky_prob, sky_thr, base, i, hi, !ALL, v);
    const uint32_t sel = ALL ? v.on : v.ns;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if ((sel >> j) & 1u) {
        const uint32_t d = v.key[j] >> 16;
        const uint32_t r = d - wb;
        if (r < (uint32_t)kSelWin) atomicAdd(&s_h[r], 1u);
        else atomicAdd(g + d, 1u);
        nan_sel |= v.key[j] == 0xFFFFFFFFu;
      }
    }
  }
  if (nan_sel) nanflag[2 * b + (ALL ? 1 : 0)] = 1u;           // np.percentile of data with a NaN is NaN
  __syncthreads();
  for (int i = threadIdx.x; i < kSelWin; i += kSelThreads) {
    const uint32_t c = s_h[i];
    if (c) atomicAdd(g + wb + i, c);
  }
}

// PASS 1: low 16 bits of the values that carry the prefix of rank 0 / rank 1.  hist1: [B][2][kSelBins].
template <bool VEC>
__global__ void __launch_bounds__(kSelThreads) sel_hist1_kernel(const float *__restrict__ conf,
                                                                const uint8_t *__restrict__ sky,
                                                                const float *__restrict__ sky_prob, float sky_thr,
                                                                int64_t npix, const SelState *__restrict__ state,
                                                                uint32_t *__restrict__ hist1) {
  const int b = blockIdx.y;
  const SelState st = state[b];
  const uint32_t p0 = st.prefix[0] >> 16, p1 = st.prefix[1] >> 16;
  uint32_t *g = hist1 + (int64_t)b * 2 * kSelBins;
  const int64_t lo = (int64_t)blockIdx.x * kSelChunk;
  const int64_t hi = lo + kSelChunk < npix ? lo + kSelChunk : npix;
  const int64_t base = (int64_t)b * npix;
  for (int64_t i = lo + (int64_t)threadIdx.x * 16; i < hi; i += (int64_t)kSelThreads * 16) {
    SelVals v;
    sel_load<VEC>(conf, sky, sky_prob, sky_thr, base, i, hi, !st.use_all, v);
    const uint32_t sel = st.use_all ? v.on : v.ns;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t d = v.key[j] >> 16;
      const bool on = (sel >> j) & 1u;
      const bool m0 = on && d == p0, m1 = on && d == p1;
      if (m0 || m1) {
        // lanes of the warp that are here with the same key are counted by one of them (a constant map would
        // otherwise put every value of the frame on one address)
        const uint32_t low = v.key[j] & 0xFFFFu;
        const unsigned peers = __match_any_sync(__activemask(), v.key[j]);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
          const uint32_t c = (uint32_t)__popc(peers);
          if (m0) atomicAdd(g + low, c);
          if (m1) atomicAdd(g + kSelBins + low, c);
        }
      }
    }
  }
}

// bin that holds rank r of a histogram of kSelBins counters: the CTA's 256 threads own 256 consecutive bins each
__device__ __forceinline__ void sel_find(const uint32_t *h, int64_t r, uint32_t &bin, int64_t &before, int64_t &total) {
  __shared__ long long s_part[kSelThreads];
  __shared__ uint32_t s_bin;
  __shared__ long long s_before;
  const int t = threadIdx.x;
  constexpr int per = kSelBins / kSelThreads;
  const uint4 *h4 = reinterpret_cast<const uint4 *>(h + t * per);
  long long mine = 0;
  for (int i = 0; i < per / 4; ++i) {
    const uint4 c = h4[i];
    mine += (long long)c.x + c.y + c.z + c.w;
  }
  s_part[t] = mine;
  if (t == 0) { s_bin = 0; s_before = 0; }
  __syncthreads();
  long long excl = 0, tot = 0;
  for (int i = 0; i < kSelThreads; ++i) {          // 256 shared-memory reads per thread: a few microseconds per map
    const long long c = s_part[i];
    if (i < t) excl += c;
    tot += c;
  }
  if (r >= excl && r < excl + mine) {
    long long acc = excl;
    for (int i = 0; i < per; ++i) {
      const uint32_t c = h[t * per + i];
      if (r < acc + c) { s_bin = (uint32_t)(t * per + i); s_before = acc; break; }
      acc += c;
    }
  }
  __syncthreads();
  bin = s_bin;
  before = s_before;
  total = tot;
  __syncthreads();
}

// One CTA per sample.  PASS 0 also derives the ranks from the count exactly as numpy does
// (numpy/lib/_function_base_impl.py: 'linear' virtual index (n - 1) * q, _get_indexes, _get_gamma):
// f32_index != 0 -> index arithmetic in fp32 (NumPy >= 2 with fp32 data and a Python-float
// percentile), else in fp64 (NumPy < 2).  AGAIN: the second pick 0, only for the maps that were redone over all pixels.
template <int PASS, bool AGAIN>
__global__ void __launch_bounds__(kSelThreads) sel_pick_kernel(uint32_t *hist0, uint32_t *hist1, SelState *state, double q64,
                                                               float q32, int f32_index, double *out_thr, float *out_thr32,
                                                               int32_t *out_n, const uint32_t *nanflag) {
  const int b = blockIdx.x;
  SelState st = state[b];
  if (PASS == 0) {
    if (AGAIN && !st.redo_all) return;
    uint32_t *h = hist0 + (int64_t)b * kSelBins;
    uint32_t bin;
    int64_t before, n;
    sel_find(h, 0, bin, before, n);                                    // the count
    if (!AGAIN && n <= 10) {                                           // inference_nuscenes.py:357-360: all pixels instead
      for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) h[i] = 0;
      if (threadIdx.x == 0) { state[b].redo_all = 1; state[b].use_all = 1; }
      return;
    }
    st.n = (int32_t)n;
    double vidx;
    if (f32_index) vidx = (double)__fmul_rn((float)(n - 1), q32);    // (n - 1) * quantiles in fp32
    else vidx = __dmul_rn((double)(n - 1), q64);
    double prev = floor(vidx);
    double next = prev + 1.0;
    st.gamma = f32_index ? (double)__fsub_rn((float)vidx, (float)prev) : vidx - prev;   // before the clamps, like numpy
    // _get_indexes: above bounds -> last (the Python int n - 1 is weak: compared in the index dtype)
    const bool above = f32_index ? ((float)vidx >= (float)(n - 1)) : (vidx >= (double)(n - 1));
    if (above) prev = next = (double)(n - 1);
    if (vidx < 0.0) prev = next = 0.0;
    st.rank[0] = (int64_t)prev;
    st.rank[1] = (int64_t)next;
    st.same = st.rank[0] == st.rank[1];
    if (n <= 0) st.rank[0] = st.rank[1] = 0;
    for (int q = 0; q < 2; ++q) {
      int64_t tot;
      sel_find(h, st.rank[q], bin, before, tot);
      st.prefix[q] = bin << 16;
      st.rank[q] -= before;
    }
    if (threadIdx.x == 0) state[b] = st;
  } else {
    for (int q = 0; q < 2; ++q) {
      uint32_t bin;
      int64_t before, tot;
      sel_find(hist1 + ((int64_t)b * 2 + q) * kSelBins, st.rank[q], bin, before, tot);
      st.prefix[q] |= bin;
    }
    if (threadIdx.x == 0) {
      const float a = sel_unkey(st.prefix[0]), bb = sel_unkey(st.prefix[1]);
      double thr;
      if (st.n <= 0 || nanflag[2 * b + (st.use_all ? 1 : 0)]) {
        thr = __longlong_as_double(0x7FF8000000000000ll);            // np.percentile of nothing / of NaNs: nan
      } else if (f32_index) {                                        // _lerp, everything fp32
        const float t = (float)st.gamma, d = __fsub_rn(bb, a);
        float r = __fadd_rn(a, __fmul_rn(d, t));
        if (t >= 0.5f) r = __fsub_rn(bb, __fmul_rn(d, __fsub_rn(1.0f, t)));
        thr = (double)r;
      } else {                                                       // fp32 difference, fp64 interpolation
        const double t = st.gamma, d = (double)__fsub_rn(bb, a);
        thr = __dadd_rn((double)a, __dmul_rn(d, t));
        if (t >= 0.5) thr = __dsub_rn((double)bb, __dmul_rn(d, __dsub_rn(1.0, t)));
      }
      out_thr[b] = thr;
      if (out_thr32) out_thr32[b] = (float)thr;
      if (out_n) out_n[b] = st.n;
    }
  }
}

}  // namespace rd3

using namespace rd3;

template <bool VEC>
static void sel_launch(const float *conf, const uint8_t *sky, const float *sky_prob, float sky_thr, int B, int64_t npix,
                       double q64, float q32, int f32, double *d_thresh, float *d_thresh32, int32_t *d_count,
                       uint32_t *hist0, uint32_t *hist1, uint32_t *nanflag, SelState *state, cudaStream_t s) {
  const dim3 grid((unsigned)ceil_div(npix, kSelChunk), B);
  sel_window_kernel<<<B, 256, 0, s>>>(conf, npix, state);
  sel_hist0_kernel<VEC, false><<<grid, kSelThreads, 0, s>>>(conf, sky, sky_prob, sky_thr, npix, state, hist0, nanflag);
  sel_pick_kernel<0, false><<<B, kSelThreads, 0, s>>>(hist0, hist1, state, q64, q32, f32, d_thresh, d_thresh32, d_count, nanflag);
  sel_hist0_kernel<VEC, true><<<grid, kSelThreads, 0, s>>>(conf, sky, sky_prob, sky_thr, npix, state, hist0, nanflag);
  sel_pick_kernel<0, true><<<B, kSelThreads, 0, s>>>(hist0, hist1, state, q64, q32, f32, d_thresh, d_thresh32, d_count, nanflag);
  sel_hist1_kernel<VEC><<<grid, kSelThreads, 0, s>>>(conf, sky, sky_prob, sky_thr, npix, state, hist1);
  sel_pick_kernel<1, false><<<B, kSelThreads, 0, s>>>(hist0, hist1, state, q64, q32, f32, d_thresh, d_thresh32, d_count, nanflag);
}


extern "C" {

size_t rd3_conf_percentile_workspace_bytes(int B) {
  if (B <= 0) return 0;
  // [hist0 B x 65536 | hist1 B x 2 x 65536 | nan flags B x 2] (zeroed by one memset), then the per-sample state
  return align_up((size_t)B * (3 * kSelBins + 2) * 4) + align_up((size_t)B * sizeof(SelState));
}

static int conf_percentile_impl(const float *conf, const uint8_t *sky, const float *sky_prob, float sky_thr, int B,
                                int64_t npix, double percentile, int numpy2_fp32_index, double *d_thresh,
                                float *d_thresh32, int32_t *d_count, void *workspace, size_t workspace_bytes,
                                rd3_stream_t stream) {
  if (B <= 0 || B > 65535 || npix <= 0 || npix >= ((int64_t)1 << 31) || !conf || !d_thresh || !workspace)
    return RD3_ERR_INVALID_ARGUMENT;
  if (!(percentile >= 0.0 && percentile <= 100.0)) return RD3_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < rd3_conf_percentile_workspace_bytes(B)) return RD3_ERR_WORKSPACE;
  uint32_t *hist0 = (uint32_t *)workspace;
  uint32_t *hist1 = hist0 + (size_t)B * kSelBins;
  uint32_t *nanflag = hist1 + (size_t)B * 2 * kSelBins;
  SelState *state = (SelState *)((char *)workspace + align_up((size_t)B * (3 * kSelBins + 2) * 4));
  cudaStream_t s = (cudaStream_t)stream;
  // np.true_divide(q, a.dtype.type(100)) with a Python-float q: fp32 / fp32 (NumPy >= 2); q / 100 in fp64 before
  const float q32 = (float)percentile / 100.0f;
  const double q64 = percentile / 100.0;
  RD3_CUDA_TRY(cudaMemsetAsync(hist0, 0, (size_t)B * (3 * kSelBins + 2) * 4, s));
  // 16-byte loads of 16 consecutive values: every map starts 16-byte aligned for all three inputs
  const bool vec = (npix % 16 == 0) && ((reinterpret_cast<uintptr_t>(conf) & 15) == 0) &&
                   (!sky || (reinterpret_cast<uintptr_t>(sky) & 15) == 0) &&
                   (!sky_prob || (reinterpret_cast<uintptr_t>(sky_prob) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
  if (vec)
    sel_launch<true>(conf, sky, sky_prob, sky_thr, B, npix, q64, q32, numpy2_fp32_index, d_thresh, d_thresh32, d_count,
                     hist0, hist1, nanflag, state, s);
  else
    sel_launch<false>(conf, sky, sky_prob, sky_thr, B, npix, q64, q32, numpy2_fp32_index, d_thresh, d_thresh32, d_count,
                      hist0, hist1, nanflag, state, s);
  return check_launch();
}

int rd3_conf_percentile(const float *conf, const uint8_t *sky, int B, int64_t npix, double percentile,
                        int numpy2_fp32_index, double *d_thresh, float *d_thresh32, int32_t *d_count,
                        void *workspace, size_t workspace_bytes, rd3_stream_t stream) {
  return conf_percentile_impl(conf, sky, nullptr, 0.0f, B, npix, percentile, numpy2_fp32_index, d_thresh,
                              d_thresh32, d_count, workspace, workspace_bytes, stream);
}

int rd3_conf_percentile_skyprob(const float *conf, const float *sky_prob, float sky_prob_thresh, int B,
                                int64_t npix, double percentile, int numpy2_fp32_index, double *d_thresh,
                                float *d_thresh32, int32_t *d_count, void *workspace, size_t workspace_bytes,
                                rd3_stream_t stream) {
  return conf_percentile_impl(conf, nullptr, sky_prob, sky_prob_thresh, B, npix, percentile, numpy2_fp32_index,
                              d_thresh, d_thresh32, d_count, workspace, workspace_bytes, stream);
}

}  // extern "C"
