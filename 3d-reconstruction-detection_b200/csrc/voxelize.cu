// voxelize.cu -- C-ABI entry points for dynamic / hard voxelization of a point
// array and the standalone HardSimpleVFE, plus library-wide helpers.
#include <math.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "hard_voxel.cuh"

namespace rd3 {

static thread_local cudaError_t g_last_error = cudaSuccess;

void set_last_cuda_error(cudaError_t e) { g_last_error = e; }

int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return RD3_ERR_CUDA;
  }
  return RD3_OK;
}

// ---- per-stage profiler ------------------------------------------------------
namespace {
constexpr int kProfMaxCalls = 1024;
struct Profiler {
  bool enabled = false;
  int calls = 0;
  bool in_call = false;
  cudaEvent_t ev[kProfMaxCalls][kProfStages + 1];
  bool created = false;
} g_prof;
}  // namespace

void prof_mark(cudaStream_t stream, int boundary) {
  if (!g_prof.enabled) return;
  if (boundary == 0) g_prof.in_call = g_prof.calls < kProfMaxCalls;
  if (!g_prof.in_call) return;
  cudaEventRecord(g_prof.ev[g_prof.calls][boundary], stream);
  if (boundary == kProfStages) {
    ++g_prof.calls;
    g_prof.in_call = false;
  }
}

bool prof_enabled() { return g_prof.enabled; }

namespace {
struct LaneSet {
  bool tried = false, ok = false;
  StreamLanes lanes;
};
LaneSet g_lanes[64];
std::mutex g_lane_mutex[64];      // one per device: enqueues on different GPUs of one process do not wait for each other
int lane_mutex_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return dev;
}
}  // namespace

LaneLock::LaneLock() : dev(lane_mutex_index()) { g_lane_mutex[dev].lock(); }
LaneLock::~LaneLock() { g_lane_mutex[dev].unlock(); }

StreamLanes *get_stream_lanes() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  LaneSet &L = g_lanes[dev];
  if (!L.tried) {
    L.tried = true;
    bool ok = cudaEventCreateWithFlags(&L.lanes.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < kMaxLanes - 1; ++i) {
      ok = cudaStreamCreateWithFlags(&L.lanes.s[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&L.lanes.join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    L.ok = ok;
  }
  return L.ok ? &L.lanes : nullptr;
}

static int env_int(const char *name, int dflt, int lo, int hi) {
  int v = dflt;
  if (const char *e = getenv(name)) v = atoi(e);
  return v < lo ? lo : (v > hi ? hi : v);
}

int stream_lane_count() {
  static const int n = env_int("RD3_STREAMS", 2, 1, kMaxLanes);    // read once, not in the launch path
  return n;
}

const HvTuning &hv_tuning() {
  static const HvTuning t = [] {
    HvTuning v;
    v.rounds = env_int("RD3_ROUNDS", 8, 1, kMaxRounds);
    v.load_pct = env_int("RD3_TABLE_LOAD_PCT", 50, 10, 90);
    v.ins_iters = env_int("RD3_INS_ITERS", 4, 1, 64);
    v.lkp_iters = env_int("RD3_LKP_ITERS", 8, 1, 64);
    v.cull = env_int("RD3_CULL", 1, 0, 1);
    v.sm_count = 148;                                              // B200; replaced by the device's own count when one is visible
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      v.sm_count = sms;
    (void)cudaGetLastError();
    return v;
  }();
  return t;
}

int make_grid(const float voxel_size[3], const float coors_range[6], VoxelGrid *g,
              uint64_t *volume) {
  for (int i = 0; i < 3; ++i) {
    if (!(voxel_size[i] > 0.0f)) return RD3_ERR_INVALID_ARGUMENT;
    g->lo[i] = coors_range[i];
    g->vs[i] = voxel_size[i];
    // voxelization_cpu.cpp:121-124: round() of the fp32 quotient
    const float q = (coors_range[3 + i] - coors_range[i]) / voxel_size[i];
    const float r = roundf(q);
    if (!(r >= 1.0f && r < 2147483648.0f)) return RD3_ERR_INVALID_ARGUMENT;
    g->grid[i] = (int32_t)r;
    g->rvs[i] = (float)(1.0 / (double)voxel_size[i]);
  }
  g->rvs_max = fmaxf(g->rvs[0], fmaxf(g->rvs[1], g->rvs[2]));
  {
    int gmax = g->grid[0] > g->grid[1] ? g->grid[0] : g->grid[1];
    if (g->grid[2] > gmax) gmax = g->grid[2];
    g->fast_ok = (gmax < (1 << 20) - 2) ? 1 : 0;
    g->tolc = 4.76837158e-7f * (float)(gmax + 2) + 1e-30f;
    g->thrc = g->fast_ok ? 0.5f - g->tolc - 9.5367431640625e-7f : -1.0f;
  }
  uint64_t vol = 1;
  for (int i = 0; i < 3; ++i) {
    vol *= (uint64_t)g->grid[i];
    if (vol > 0xFFFFFFDFull) {
      *volume = 0;
      return RD3_ERR_UNSUPPORTED;   // grid[] is still valid
    }
  }
  *volume = vol;
  return RD3_OK;
}

// ---------------------------------------------------------------------------
// dynamic_voxelize: one thread per point.  For C == 3 a warp's 32 points are
// 384 contiguous bytes in and out; for C == 4 each point is one 128-bit load.
// ---------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(256) dynamic_voxelize_kernel(const float *__restrict__ pts,
                                                               int64_t N, int C, VoxelGrid g,
                                                               int32_t *__restrict__ coors) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float x, y, z;
  if (CT == 4) {
    const float4 p = __ldg(reinterpret_cast<const float4 *>(pts) + i);
    x = p.x; y = p.y; z = p.z;
  } else {
    const float *p = pts + i * C;
    x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
  }
  int cx, cy, cz;
  const bool ok = voxel_coor(x, y, z, g, cx, cy, cz);
  int32_t *o = coors + i * 3;
  o[0] = ok ? cz : -1;
  o[1] = ok ? cy : -1;
  o[2] = ok ? cx : -1;
}

// ---------------------------------------------------------------------------
// HardSimpleVFE: one thread per (voxel, feature); sequential slot order.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hard_simple_vfe_kernel(const float *__restrict__ voxels,
                                                              const int32_t *__restrict__ num,
                                                              int64_t M, int K, int C, int F,
                                                              float *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * F) return;
  const int64_t m = t / F;
  const int f = (int)(t - m * F);
  const float *v = voxels + m * K * C + f;
  float s = 0.0f;
  for (int k = 0; k < K; ++k) s = __fadd_rn(s, __ldg(v + (int64_t)k * C));
  out[t] = __fdiv_rn(s, (float)__ldg(num + m));
}

// ---------------------------------------------------------------------------
// Batched sparse-encoder inputs: rows [0, voxel_num[b]) of every sample, packed in sample order,
// coors widened to (batch, z, y, x).  grid (ceil(max_voxels / 256), B).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_sparse_kernel(const float *__restrict__ feats,
                                                          const int32_t *__restrict__ coors,
                                                          const int32_t *__restrict__ num,
                                                          const int32_t *__restrict__ voxel_num, int B,
                                                          int max_voxels, int F, int batch_offset,
                                                          float *__restrict__ out_feats,
                                                          int32_t *__restrict__ out_coors,
                                                          int32_t *__restrict__ out_num,
                                                          int32_t *__restrict__ offsets) {
  __shared__ int s_warp[8];
  __shared__ int s_off;
  const int b = blockIdx.y;
  // exclusive prefix of the clamped per-sample counts (every CTA of sample b recomputes it)
  int acc = 0;
  for (int i = threadIdx.x; i < b; i += 256) acc += min(max(__ldg(voxel_num + i), 0), max_voxels);
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; ++k) t += s_warp[k];
    s_off = t;
  }
  __syncthreads();
  const int off = s_off;
  const int m = min(max(__ldg(voxel_num + b), 0), max_voxels);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    offsets[b] = off;
    if (b == B - 1) offsets[B] = off + m;
  }
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v >= m) return;
  const int64_t src = (int64_t)b * max_voxels + v, dst = (int64_t)off + v;
  for (int f = 0; f < F; ++f) out_feats[dst * F + f] = __ldg(feats + src * F + f);
  const int4 c = make_int4(batch_offset + b, __ldg(coors + src * 3), __ldg(coors + src * 3 + 1),
                           __ldg(coors + src * 3 + 2));
  reinterpret_cast<int4 *>(out_coors)[dst] = c;
  if (out_num) out_num[dst] = __ldg(num + src);
}

}  // namespace rd3

using namespace rd3;

extern "C" {

int rd3_version(void) { return 100; }

const char *rd3_status_string(int status) {
  switch (status) {
    case RD3_OK: return "ok";
    case RD3_ERR_INVALID_ARGUMENT: return "invalid argument";
    case RD3_ERR_WORKSPACE: return "workspace too small";
    case RD3_ERR_UNSUPPORTED: return "unsupported configuration";
    case RD3_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}

const char *rd3_last_cuda_error(void) { return cudaGetErrorString(g_last_error); }

int rd3_profile_enable(int on) {
  if (on && !g_prof.created) {
    for (int c = 0; c < kProfMaxCalls; ++c)
      for (int b = 0; b <= kProfStages; ++b) RD3_CUDA_TRY(cudaEventCreate(&g_prof.ev[c][b]));
    g_prof.created = true;
  }
  g_prof.enabled = on != 0;
  g_prof.calls = 0;
  g_prof.in_call = false;
  return RD3_OK;
}

int rd3_profile_read(double *stage_ms, int *calls) {
  if (!stage_ms || !calls) return RD3_ERR_INVALID_ARGUMENT;
  for (int s = 0; s < kProfStages; ++s) stage_ms[s] = 0.0;
  *calls = g_prof.calls;
  for (int c = 0; c < g_prof.calls; ++c) {
    RD3_CUDA_TRY(cudaEventSynchronize(g_prof.ev[c][kProfStages]));
    for (int s = 0; s < kProfStages; ++s) {
      float ms = 0.0f;
      RD3_CUDA_TRY(cudaEventElapsedTime(&ms, g_prof.ev[c][s], g_prof.ev[c][s + 1]));
      stage_ms[s] += ms;
    }
  }
  return RD3_OK;
}

int rd3_grid_size(const float voxel_size[3], const float coors_range[6], int32_t grid[3]) {
  if (!voxel_size || !coors_range || !grid) return RD3_ERR_INVALID_ARGUMENT;
  VoxelGrid g;
  uint64_t vol;
  int st = make_grid(voxel_size, coors_range, &g, &vol);
  if (st != RD3_OK && st != RD3_ERR_UNSUPPORTED) return st;
  for (int i = 0; i < 3; ++i) grid[i] = g.grid[i];
  return RD3_OK;
}

int rd3_dynamic_voxelize(const float *points, int64_t N, int C, const float voxel_size[3],
                         const float coors_range[6], int32_t *coors, rd3_stream_t stream) {
  if (N < 0 || C < 3 || !voxel_size || !coors_range) return RD3_ERR_INVALID_ARGUMENT;
  if (N == 0) return RD3_OK;
  if (!points || !coors) return RD3_ERR_INVALID_ARGUMENT;
  VoxelGrid g;
  uint64_t vol;
  int st = make_grid(voxel_size, coors_range, &g, &vol);
  if (st == RD3_ERR_UNSUPPORTED) st = RD3_OK;  // no linear key needed here
  if (st != RD3_OK) return st;
  const unsigned blocks = (unsigned)ceil_div(N, 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (C == 4 && (reinterpret_cast<uintptr_t>(points) & 15) == 0)
    dynamic_voxelize_kernel<4><<<blocks, 256, 0, s>>>(points, N, C, g, coors);
  else
    dynamic_voxelize_kernel<0><<<blocks, 256, 0, s>>>(points, N, C, g, coors);
  return check_launch();
}

size_t rd3_hard_voxelize_workspace_bytes(int64_t N, int max_points, int max_voxels) {
  if (N < 0 || max_points <= 0 || max_voxels <= 0) return 0;
  return hv_plan(N, 1, max_points, max_voxels).total;
}

int rd3_hard_voxel_rounds(int64_t N, int B) {
  if (N < 0 || B <= 0) return 0;
  return hv_plan(N, B, 1, 1).rounds;
}

int rd3_hard_voxelize(const float *points, int64_t N, int C, const float voxel_size[3],
                      const float coors_range[6], int max_points, int max_voxels, float *voxels,
                      int32_t *coors, int32_t *num_points_per_voxel, int32_t *d_voxel_num,
                      float *voxel_mean, int F, int32_t *point2voxel, void *workspace,
                      size_t workspace_bytes, rd3_stream_t stream) {
  if (N < 0 || C < 3 || max_points <= 0 || max_voxels <= 0 || !voxel_size || !coors_range ||
      !voxels || !coors || !num_points_per_voxel || !d_voxel_num || !workspace)
    return RD3_ERR_INVALID_ARGUMENT;
  if (N > 0 && !points) return RD3_ERR_INVALID_ARGUMENT;
  if (voxel_mean && (F < 1 || F > C)) return RD3_ERR_INVALID_ARGUMENT;
  if (N >= ((int64_t)1 << 30)) return RD3_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return RD3_ERR_INVALID_ARGUMENT;   // 256-bit table loads
  VoxelGrid g;
  uint64_t vol;
  int st = make_grid(voxel_size, coors_range, &g, &vol);
  if (st != RD3_OK) return st;
  const HvPlan plan = hv_plan(N, 1, max_points, max_voxels);
  if (workspace_bytes < plan.total) return RD3_ERR_WORKSPACE;
  PointsSource src{points, N, C};
  HvOut out{voxels, coors, num_points_per_voxel, voxel_mean, d_voxel_num, point2voxel,
            voxel_mean ? F : 0};
  return hv_run(src, g, vol, plan, workspace, out, (cudaStream_t)stream);
}

int rd3_pack_sparse_inputs(const float *voxel_feats, const int32_t *coors, const int32_t *num_points,
                           const int32_t *d_voxel_num, int B, int max_voxels, int F, int batch_offset,
                           float *out_feats, int32_t *out_coors, int32_t *out_num_points,
                           int32_t *d_offsets, rd3_stream_t stream) {
  if (B <= 0 || B > 65535 || max_voxels <= 0 || F <= 0) return RD3_ERR_INVALID_ARGUMENT;
  if (!voxel_feats || !coors || !d_voxel_num || !out_feats || !out_coors || !d_offsets)
    return RD3_ERR_INVALID_ARGUMENT;
  if (out_num_points && !num_points) return RD3_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(out_coors) & 15) return RD3_ERR_INVALID_ARGUMENT;
  pack_sparse_kernel<<<dim3((unsigned)ceil_div(max_voxels, 256), B), 256, 0, (cudaStream_t)stream>>>(
      voxel_feats, coors, num_points, d_voxel_num, B, max_voxels, F, batch_offset, out_feats, out_coors,
      out_num_points, d_offsets);
  return check_launch();
}

int rd3_hard_simple_vfe(const float *voxels, const int32_t *num_points, int64_t M, int max_points,
                        int C, int F, float *out, rd3_stream_t stream) {
  if (M < 0 || max_points <= 0 || C <= 0 || F <= 0 || F > C) return RD3_ERR_INVALID_ARGUMENT;
  if (M == 0) return RD3_OK;
  if (!voxels || !num_points || !out) return RD3_ERR_INVALID_ARGUMENT;
  const unsigned blocks = (unsigned)ceil_div(M * F, 256);
  hard_simple_vfe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(voxels, num_points, M,
                                                                    max_points, C, F, out);
  return check_launch();
}

}  // extern "C"
