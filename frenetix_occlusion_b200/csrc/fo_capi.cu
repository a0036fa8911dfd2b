// Library plumbing: error string, launch counter, version, FP32 probe, host-buffer entry point.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "fo_common.cuh"

namespace fo {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// 8 independent FFMA chains per thread; 2 flops per FFMA.
__global__ void __launch_bounds__(256) fp32_probe_kernel(int iters, float seed, float* sink) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 0.001f;
#pragma unroll 32
  for (int i = 0; i < iters; ++i) {      // 256 FFMAs per loop-control triple: the loop overhead stays below 1.5 %
    a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
    a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
  }
  float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123456.789f) sink[0] = r;
}

// per-thread device workspace for the host-buffer entry point
constexpr int kHostChunksMax = 16;
struct HostWs {
  void* dev = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;        // compute (+ result copies)
  cudaStream_t copy = nullptr;          // host -> device copies of the bundle
  cudaEvent_t ready[kHostChunksMax] = {};
};
static thread_local HostWs g_ws;

static int ws_reserve(size_t bytes) {
  if (!g_ws.stream) {
    FO_CUDA_TRY(cudaStreamCreateWithFlags(&g_ws.stream, cudaStreamNonBlocking));
    FO_CUDA_TRY(cudaStreamCreateWithFlags(&g_ws.copy, cudaStreamNonBlocking));
    for (int i = 0; i < kHostChunksMax; ++i) FO_CUDA_TRY(cudaEventCreateWithFlags(&g_ws.ready[i], cudaEventDisableTiming));
  }
  if (bytes <= g_ws.bytes) return FO_OK;
  if (g_ws.dev) FO_CUDA_TRY(cudaFree(g_ws.dev));
  g_ws.dev = nullptr; g_ws.bytes = 0;
  size_t want = bytes + bytes / 4;
  FO_CUDA_TRY(cudaMalloc(&g_ws.dev, want));
  g_ws.bytes = want;
  return FO_OK;
}

}  // namespace fo

extern "C" int fo_version(void) { return FO_ABI_VERSION; }
extern "C" const char* fo_last_error(void) { return fo::g_err; }
extern "C" uint64_t fo_launch_count(void) { return fo::g_launches.load(std::memory_order_relaxed); }

// ---- peer-mapped buffers (CUDA IPC) ---------------------------------------------------------------------------------
static_assert(sizeof(FoPeerHandle) == sizeof(cudaIpcMemHandle_t), "FoPeerHandle carries a cudaIpcMemHandle_t");

extern "C" int fo_peer_alloc(size_t bytes, void** dev_ptr, FoPeerHandle* handle) {
  if (!dev_ptr || !handle || bytes == 0) { fo::set_error("fo_peer_alloc: bad argument"); return FO_ERR_INVALID_ARG; }
  void* p = nullptr;
  FO_CUDA_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    fo::set_error("fo_peer_alloc: %s", cudaGetErrorString(e));
    return FO_ERR_CUDA;
  }
  memcpy(handle->bytes, &h, sizeof(h));
  *dev_ptr = p;
  return FO_OK;
}

extern "C" int fo_peer_open(const FoPeerHandle* handle, void** dev_ptr) {
  if (!dev_ptr || !handle) { fo::set_error("fo_peer_open: bad argument"); return FO_ERR_INVALID_ARG; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle->bytes, sizeof(h));
  FO_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FO_OK;
}

extern "C" int fo_peer_close(void* dev_ptr) {
  if (dev_ptr) FO_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return FO_OK;
}

extern "C" int fo_peer_free(void* dev_ptr) {
  if (dev_ptr) FO_CUDA_TRY(cudaFree(dev_ptr));
  return FO_OK;
}

extern "C" int fo_probe_fp32_peak(int32_t iters, float* ms, double* flops, void* stream) {
  if (iters <= 0 || !ms || !flops) { fo::set_error("fo_probe_fp32_peak: bad argument"); return FO_ERR_INVALID_ARG; }
  int dev = 0, sms = 0;
  FO_CUDA_TRY(cudaGetDevice(&dev));
  FO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* sink = nullptr;
  FO_CUDA_TRY(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  FO_CUDA_TRY(cudaEventCreate(&e0));
  FO_CUDA_TRY(cudaEventCreate(&e1));
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = sms * 8;
  fo::fp32_probe_kernel<<<grid, 256, 0, st>>>(iters / 8 + 1, 1.0f, sink);  // warm-up
  FO_CUDA_TRY(cudaEventRecord(e0, st));
  fo::fp32_probe_kernel<<<grid, 256, 0, st>>>(iters, 1.0f, sink);
  FO_CUDA_TRY(cudaEventRecord(e1, st));
  fo::count_launch(2);
  FO_CUDA_TRY(cudaEventSynchronize(e1));
  FO_CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
  *flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)grid;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
  return FO_OK;
}

extern "C" int fo_metric_bundle_host(const float* ego_host, int32_t n_traj, int32_t n_states, const FoAgentsRaw* ag,
                                     const FoMetricArgs* params, uint8_t* out_valid, float* out_summary,
                                     uint32_t* out_flags, float* out_pair, float* out_step) {
  if (!params || !ag || (n_traj > 0 && (!ego_host || !out_valid))) {
    fo::set_error("fo_metric_bundle_host: NULL argument");
    return FO_ERR_INVALID_ARG;
  }
  if (n_traj < 0 || n_states < 0 || ag->n_agents < 0 || ag->t_stride < 0) {
    fo::set_error("fo_metric_bundle_host: negative size");
    return FO_ERR_INVALID_ARG;
  }
  if (n_traj == 0) return FO_OK;
  const size_t N = n_traj, T = n_states, A = ag->n_agents, Tp = ag->t_stride;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_ego = al(N * T * 5 * 4), b_f = al(A * Tp * 4), b_a = al(A * 4 + 4);
  const size_t b_tab = al(fo::agent_table_bytes((int)A, (int)Tp) + 16);
  const size_t b_valid = al(N), b_sum = al(N * FO_SUMMARY_K * 4), b_flags = al(N * 4);
  const size_t b_pair = out_pair ? al(N * A * FO_PAIR_K * 4 + 4) : 0;
  const size_t b_step = (out_step && T > 1) ? al(N * A * (T - 1) * FO_STEP_K * 4 + 4) : 0;
  const size_t total = b_ego + 6 * b_f + 6 * b_a + b_tab + b_valid + b_sum + b_flags + b_pair + b_step;
  int rc = fo::ws_reserve(total);
  if (rc != FO_OK) return rc;
  cudaStream_t st = fo::g_ws.stream;
  // Error exits below happen with copies / kernels possibly still queued on the workspace streams: wait for them
  // before handing the caller's buffers back (the message of the first failure is kept).
  auto bail = [&](int code) {
    char keep[sizeof(fo::g_err)];
    memcpy(keep, fo::g_err, sizeof(keep));
    cudaStreamSynchronize(fo::g_ws.copy);
    cudaStreamSynchronize(st);
    cudaGetLastError();
    memcpy(fo::g_err, keep, sizeof(keep));
    return code;
  };
#define FO_HOST_TRY(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      fo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);   \
      return bail(FO_ERR_CUDA);                                                                    \
    }                                                                                              \
  } while (0)
  char* p = (char*)fo::g_ws.dev;
  auto take = [&](size_t b) { char* r = p; p += b; return r; };
  float* d_ego = (float*)take(b_ego);
  FoAgentsRaw d = *ag;
  const float* hsrc[6] = {ag->x, ag->y, ag->yaw, ag->v, ag->var_x, ag->var_y};
  const float** hdst[6] = {&d.x, &d.y, &d.yaw, &d.v, &d.var_x, &d.var_y};
  for (int i = 0; i < 6; ++i) {
    float* dp = (float*)take(b_f);
    if (A * Tp) {
      if (!hsrc[i]) { fo::set_error("fo_metric_bundle_host: NULL agent array"); return bail(FO_ERR_INVALID_ARG); }
      FO_HOST_TRY(cudaMemcpyAsync(dp, hsrc[i], A * Tp * 4, cudaMemcpyHostToDevice, st));
    }
    *hdst[i] = dp;
  }
  const void* asrc[6] = {ag->n_states, ag->kind, ag->length, ag->width, ag->buf_length, ag->buf_width};
  const void** adst[6] = {(const void**)&d.n_states, (const void**)&d.kind, (const void**)&d.length,
                          (const void**)&d.width, (const void**)&d.buf_length, (const void**)&d.buf_width};
  for (int i = 0; i < 6; ++i) {
    void* dp = take(b_a);
    if (A) {
      if (!asrc[i]) { fo::set_error("fo_metric_bundle_host: NULL agent array"); return bail(FO_ERR_INVALID_ARG); }
      FO_HOST_TRY(cudaMemcpyAsync(dp, asrc[i], A * 4, cudaMemcpyHostToDevice, st));
    }
    *adst[i] = dp;
  }
  void* d_tab = take(b_tab);
  FoMetricArgs m = *params;
  m.ego = d_ego; m.n_traj = n_traj; m.n_states = n_states; m.n_agents = (int32_t)A; m.t_stride = (int32_t)Tp;
  m.agent_table = d_tab;
  // Results go to the library's workspace -- or, when the caller names all three DEVICE destinations in `params`, there
  // (a rank's slice of a gather buffer: n_peers / peer_delta then apply as in fo_metric_bundle), and from there to the host.
  const bool dev_out = params->valid && params->summary && params->flags && !out_pair && !out_step;
  uint8_t* ws_valid = (uint8_t*)take(b_valid);
  float* ws_sum = (float*)take(b_sum);
  uint32_t* ws_flags = (uint32_t*)take(b_flags);
  if (dev_out) {
    m.valid = params->valid; m.summary = params->summary; m.flags = params->flags;
  } else {
    m.valid = ws_valid; m.summary = ws_sum; m.flags = ws_flags;
    m.n_peers = 0;
  }
  m.pair = out_pair ? (float*)take(b_pair) : nullptr;
  m.step = (out_step && T > 1) ? (float*)take(b_step) : nullptr;
  if (A > 0) {
    rc = fo_agents_pack(&d, &m.vehicle, d_tab, b_tab, st);
    if (rc != FO_OK) return bail(rc);
  }
  // Large bundles are pipelined: the trajectory range is cut into chunks, chunk k+1 crosses PCIe on the copy stream
  // while chunk k is evaluated (trajectories are independent), results follow each chunk back on the compute stream.
  // The first chunk is small (its copy cannot be hidden), every following one four times larger: copying is several
  // times faster than evaluating, and every launch costs a tail of about one trajectory time per warp.
  const size_t traj_bytes = T * 5 * 4;
  static const char* chunk_env = getenv("FO_HOST_CHUNK_MB");           // measurement switch: size of the first chunk
  const size_t first_bytes = (size_t)(chunk_env && atoi(chunk_env) > 0 ? atoi(chunk_env) : 16) << 20;
  size_t bound[fo::kHostChunksMax + 1];
  int chunks = 1;
  bound[0] = 0; bound[1] = N;
  if (N * traj_bytes >= 2 * first_bytes && !m.pair && !m.step && traj_bytes > 0) {
    size_t step = (first_bytes + traj_bytes - 1) / traj_bytes, at = 0;
    chunks = 0;
    while (at < N && chunks < fo::kHostChunksMax - 1) {
      at = (at + step < N) ? at + step : N;
      bound[++chunks] = at;
      step *= 4;
    }
    if (at < N) bound[++chunks] = N;
  }
  for (int c = 0; c < chunks; ++c) {
    const size_t lo = bound[c], hi = bound[c + 1], n = hi - lo;
    if (n == 0) continue;
    FO_HOST_TRY(cudaMemcpyAsync(d_ego + lo * T * 5, ego_host + lo * T * 5, n * traj_bytes, cudaMemcpyHostToDevice,
                                fo::g_ws.copy));
    FO_HOST_TRY(cudaEventRecord(fo::g_ws.ready[c], fo::g_ws.copy));
    FO_HOST_TRY(cudaStreamWaitEvent(st, fo::g_ws.ready[c], 0));
    FoMetricArgs mc = m;
    mc.ego = d_ego + lo * T * 5;
    mc.n_traj = (int32_t)n;
    mc.valid = m.valid + lo;
    mc.summary = m.summary + lo * FO_SUMMARY_K;
    mc.flags = m.flags + lo;
    if (m.pair) mc.pair = m.pair + lo * A * FO_PAIR_K;
    if (m.step) mc.step = m.step + lo * A * (T - 1) * FO_STEP_K;
    rc = fo_metric_bundle(&mc, st);
    if (rc != FO_OK) return bail(rc);
    FO_HOST_TRY(cudaMemcpyAsync(out_valid + lo, mc.valid, n, cudaMemcpyDeviceToHost, st));
    if (out_summary)
      FO_HOST_TRY(cudaMemcpyAsync(out_summary + lo * FO_SUMMARY_K, mc.summary, n * FO_SUMMARY_K * 4, cudaMemcpyDeviceToHost, st));
    if (out_flags) FO_HOST_TRY(cudaMemcpyAsync(out_flags + lo, mc.flags, n * 4, cudaMemcpyDeviceToHost, st));
  }
  if (m.pair) FO_HOST_TRY(cudaMemcpyAsync(out_pair, m.pair, N * A * FO_PAIR_K * 4, cudaMemcpyDeviceToHost, st));
  if (m.step) FO_HOST_TRY(cudaMemcpyAsync(out_step, m.step, N * A * (T - 1) * FO_STEP_K * 4, cudaMemcpyDeviceToHost, st));
  FO_HOST_TRY(cudaStreamSynchronize(st));
  return FO_OK;
#undef FO_HOST_TRY
}
