// select.cu -- the confidence threshold of the depth->points hand-off:
//   conf_thresh = np.percentile(conf[~sky] if (~sky).sum() > 10 else conf.flatten(), p)
//   (tools/inference_nuscenes.py:351-361; depth_anything_3/utils/export/glb.py:227-229)
// per sample, on the device: an exact 3-pass radix select (11 + 11 + 10 bits of the order-preserving
// integer image of the fp32 values) of the two order statistics numpy's "linear" method
// interpolates between, then numpy's own index / gamma / lerp arithmetic.  The reference sorts
// (np.partition) 2.7 M values per sample on one CPU core.
#include "rd3_common.cuh"

namespace rd3 {

constexpr int kSelBins = 2048;
constexpr int kSelThreads = 256;
constexpr int kSelChunk = 16384;        // values per histogram CTA

struct SelState {                 // per sample
  uint32_t prefix[2];             // key bits fixed so far, for the two ranks
  int64_t rank[2];                // rank still to locate inside the prefix bucket
  int32_t use_all;                // 1: every pixel, 0: non-sky pixels only
  int32_t n;                      // number of selected values
  double gamma;                   // interpolation weight (index dtype precision)
  int32_t same;                   // both ranks are the same order statistic
};

__device__ __forceinline__ uint32_t sel_key(float v) {       // monotone: a < b  <=>  key(a) < key(b); NaN last
  const uint32_t u = __float_as_uint(v);
  if (v != v) return 0xFFFFFFFFu;                            // any NaN sorts last (np.partition)
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sel_unkey(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// PASS = 0: bins of key >> 21 for (non-sky, all);  1: key >> 10 & 2047 for (rank 0, rank 1) inside
// their 11-bit prefixes;  2: key & 1023 inside their 21-bit prefixes.  hist: [B][2][kSelBins].
template <int PASS>
__global__ void __launch_bounds__(kSelThreads) sel_hist_kernel(const float *__restrict__ conf,
                                                               const uint8_t *__restrict__ sky,
                                                               const float *__restrict__ sky_prob, float sky_thr,
                                                               int64_t npix, const SelState *__restrict__ state,
                                                               uint32_t *__restrict__ hist,
                                                               uint32_t *__restrict__ nanflag) {
  __shared__ uint32_t s_h[2][kSelBins];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * kSelBins; i += kSelThreads) (&s_h[0][0])[i] = 0;
  uint32_t p0 = 0, p1 = 0;
  int use_all = 0;
  if (PASS > 0) {
    const SelState st = state[b];
    p0 = st.prefix[0]; p1 = st.prefix[1];
    use_all = st.use_all;
  }
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * kSelChunk;
  const int64_t hi = lo + kSelChunk < npix ? lo + kSelChunk : npix;
  const int lane = threadIdx.x & 31;
  // the loop bound is rounded up to whole warps: every lane takes part in the warp votes of PASS 0
  const int64_t hi32 = lo + ((hi - lo + 31) & ~(int64_t)31);
  for (int64_t i = lo + threadIdx.x; i < hi32; i += kSelThreads) {
    const bool on = i < hi;
    const uint32_t k = on ? sel_key(__ldg(conf + (int64_t)b * npix + i)) : 0u;
    const bool ns = on && (sky ? __ldg(sky + (int64_t)b * npix + i) == 0
                               : (sky_prob ? !(__ldg(sky_prob + (int64_t)b * npix + i) >= sky_thr) : true));
    if (PASS == 0) {
      // Confidences cluster in a handful of the top-11-bit bins (conf = 1 + exp(x): four exponents), so 32 plain
      // shared-memory atomics of a warp would serialise on a few addresses.  Lanes holding the same bin are
      // counted by ONE lane (__match_any_sync); absent lanes get bins of their own.
      const uint32_t bin = on ? (k >> 21) : (kSelBins + lane);
      const unsigned g_all = __match_any_sync(0xffffffffu, bin);
      const unsigned g_ns = __match_any_sync(0xffffffffu, ns ? bin : (kSelBins + lane));
      if (on && lane == __ffs(g_all) - 1) atomicAdd(&s_h[1][bin], (uint32_t)__popc(g_all));
      if (ns && lane == __ffs(g_ns) - 1) atomicAdd(&s_h[0][bin], (uint32_t)__popc(g_ns));
      if (on && k == 0xFFFFFFFFu) {                         // np.percentile of data with a NaN is NaN
        if (ns) nanflag[2 * b] = 1u;
        nanflag[2 * b + 1] = 1u;
      }
    } else if (ns || (on && use_all)) {
      if (PASS == 1) {
        if ((k >> 21) == (p0 >> 21)) atomicAdd(&s_h[0][(k >> 10) & 2047u], 1u);
        if ((k >> 21) == (p1 >> 21)) atomicAdd(&s_h[1][(k >> 10) & 2047u], 1u);
      } else {
        if ((k >> 10) == (p0 >> 10)) atomicAdd(&s_h[0][k & 1023u], 1u);
        if ((k >> 10) == (p1 >> 10)) atomicAdd(&s_h[1][k & 1023u], 1u);
      }
    }
  }
  __syncthreads();
  uint32_t *g = hist + (int64_t)b * 2 * kSelBins;
  for (int i = threadIdx.x; i < 2 * kSelBins; i += kSelThreads) {
    const uint32_t v = (&s_h[0][0])[i];
    if (v) atomicAdd(g + i, v);
  }
}

// bin that holds rank r of a histogram of kSelBins counters: one warp, 64 bins per lane
__device__ __forceinline__ void sel_find(const uint32_t *h, int64_t r, int lane, uint32_t &bin, int64_t &before) {
  int64_t mine = 0;
  for (int i = 0; i < kSelBins / 32; ++i) mine += h[lane * (kSelBins / 32) + i];
  int64_t inc = mine;
  for (int d = 1; d < 32; d <<= 1) {
    const int64_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  const int64_t excl = inc - mine;
  const bool here = r >= excl && r < inc;
  uint32_t mb = 0;
  int64_t mbefore = 0;
  if (here) {
    int64_t acc = excl;
    for (int i = 0; i < kSelBins / 32; ++i) {
      const uint32_t c = h[lane * (kSelBins / 32) + i];
      if (r < acc + c) { mb = lane * (kSelBins / 32) + i; mbefore = acc; break; }
      acc += c;
    }
  }
  const unsigned who = __ballot_sync(0xffffffffu, here);
  const int src = who ? __ffs(who) - 1 : 0;
  bin = __shfl_sync(0xffffffffu, mb, src);
  before = __shfl_sync(0xffffffffu, mbefore, src);
}

// One warp per sample.  PASS 0 also derives the ranks from the count exactly as numpy does
// (numpy/lib/_function_base_impl.py: 'linear' virtual index (n - 1) * q, _get_indexes, _get_gamma):
// f32_index != 0 -> index arithmetic in fp32 (NumPy >= 2 with fp32 data and a Python-float
// percentile), else in fp64 (NumPy < 2).
template <int PASS>
__global__ void __launch_bounds__(32) sel_pick_kernel(uint32_t *hist, SelState *state, double q64, float q32,
                                                      int f32_index, double *out_thr, float *out_thr32,
                                                      int32_t *out_n, const uint32_t *nanflag) {
  const int b = blockIdx.x, lane = threadIdx.x;
  uint32_t *h = hist + (int64_t)b * 2 * kSelBins;
  SelState st;
  if (PASS == 0) {
    int64_t n_ns = 0, n_all = 0;
    for (int i = lane; i < kSelBins; i += 32) { n_ns += h[i]; n_all += h[kSelBins + i]; }
    for (int d = 16; d > 0; d >>= 1) {
      n_ns += __shfl_xor_sync(0xffffffffu, n_ns, d);
      n_all += __shfl_xor_sync(0xffffffffu, n_all, d);
    }
    st.use_all = n_ns > 10 ? 0 : 1;                                  // inference_nuscenes.py:357-360
    const int64_t n = st.use_all ? n_all : n_ns;
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
    st.prefix[0] = st.prefix[1] = 0;
    if (n <= 0) st.rank[0] = st.rank[1] = 0;
  } else {
    st = state[b];
  }
  const int shift = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
  for (int q = 0; q < 2; ++q) {
    const uint32_t *hq = h + (PASS == 0 ? (st.use_all ? kSelBins : 0) : q * kSelBins);
    uint32_t bin;
    int64_t before;
    sel_find(hq, st.rank[q], lane, bin, before);
    st.prefix[q] |= bin << shift;
    st.rank[q] -= before;
  }
  __syncwarp();
  for (int i = lane; i < 2 * kSelBins; i += 32) h[i] = 0;            // ready for the next pass
  if (lane == 0) {
    state[b] = st;
    if (PASS == 2) {
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

extern "C" {

size_t rd3_conf_percentile_workspace_bytes(int B) {
  if (B <= 0) return 0;
  return align_up((size_t)B * (2 * kSelBins + 2) * 4) + align_up((size_t)B * sizeof(SelState));
}

static int conf_percentile_impl(const float *conf, const uint8_t *sky, const float *sky_prob, float sky_thr, int B,
                                int64_t npix, double percentile, int numpy2_fp32_index, double *d_thresh,
                                float *d_thresh32, int32_t *d_count, void *workspace, size_t workspace_bytes,
                                rd3_stream_t stream) {
  if (B <= 0 || B > 65535 || npix <= 0 || npix >= ((int64_t)1 << 31) || !conf || !d_thresh || !workspace)
    return RD3_ERR_INVALID_ARGUMENT;
  if (!(percentile >= 0.0 && percentile <= 100.0)) return RD3_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < rd3_conf_percentile_workspace_bytes(B)) return RD3_ERR_WORKSPACE;
  uint32_t *hist = (uint32_t *)workspace;
  uint32_t *nanflag = hist + (size_t)B * 2 * kSelBins;
  SelState *state = (SelState *)((char *)workspace + align_up((size_t)B * (2 * kSelBins + 2) * 4));
  cudaStream_t s = (cudaStream_t)stream;
  // np.true_divide(q, a.dtype.type(100)) with a Python-float q: fp32 / fp32 (NumPy >= 2); q / 100 in fp64 before
  const float q32 = (float)percentile / 100.0f;
  const double q64 = percentile / 100.0;
  RD3_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)B * (2 * kSelBins + 2) * 4, s));
  const dim3 grid((unsigned)ceil_div(npix, kSelChunk), B);
  sel_hist_kernel<0><<<grid, kSelThreads, 0, s>>>(conf, sky, sky_prob, sky_thr, npix, state, hist, nanflag);
  sel_pick_kernel<0><<<B, 32, 0, s>>>(hist, state, q64, q32, numpy2_fp32_index, d_thresh, d_thresh32, d_count,
                                      nanflag);
  sel_hist_kernel<1><<<grid, kSelThreads, 0, s>>>(conf, sky, sky_prob, sky_thr, npix, state, hist, nanflag);
  sel_pick_kernel<1><<<B, 32, 0, s>>>(hist, state, q64, q32, numpy2_fp32_index, d_thresh, d_thresh32, d_count,
                                      nanflag);
  sel_hist_kernel<2><<<grid, kSelThreads, 0, s>>>(conf, sky, sky_prob, sky_thr, npix, state, hist, nanflag);
  sel_pick_kernel<2><<<B, 32, 0, s>>>(hist, state, q64, q32, numpy2_fp32_index, d_thresh, d_thresh32, d_count,
                                      nanflag);
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
