// Host entry points of the dense metric core: agent-table packing and kernel dispatch.
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "fo_metric_dev.cuh"

namespace fo {

// ---------------------------------------------------------------------------------------------
// fo_agents_pack kernel: one thread per (agent, state).
__device__ __forceinline__ float obstacle_mass(int kind, float size) {  // harm_model.py:158-190
  switch (kind) {
    case FO_KIND_CAR: case FO_KIND_PRIORITY_VEHICLE: case FO_KIND_PARKED_VEHICLE: case FO_KIND_TAXI:
      return -1333.5f + 526.9f * powf(size, 0.8f);
    case FO_KIND_TRUCK: return 25000.0f;
    case FO_KIND_BUS: return 13000.0f;
    case FO_KIND_BICYCLE: return 90.0f;
    case FO_KIND_PEDESTRIAN: return 75.0f;
    case FO_KIND_TRAIN: return 118800.0f;
    case FO_KIND_MOTORCYCLE: return 250.0f;
    default: return 0.0f;
  }
}
__device__ __forceinline__ int protection_model(int kind) {  // harm_model.py:15-32
  switch (kind) {
    case FO_KIND_BICYCLE: case FO_KIND_PEDESTRIAN: case FO_KIND_MOTORCYCLE: case FO_KIND_UNKNOWN: return 0;
    default: return 1;
  }
}

// Slot order: stable sort by harm model.  One CTA; A is a few hundred at most, so the O(A^2) rank count is free.
__global__ void fo_agents_order_kernel(const int32_t* __restrict__ kind, int A, int Ap, int32_t* orig, int32_t* slot) {
  for (int a = threadIdx.x; a < Ap; a += blockDim.x) {
    if (a >= A) { orig[a] = -1; slot[a] = a; continue; }      // padding slots keep their place behind the real ones
#ifdef FO_NO_SORT
    slot[a] = a; orig[a] = a; continue;
#endif
    const int m = protection_model(kind[a]);
    int rank = 0;
    for (int b = 0; b < A; ++b) {
      const int mb = protection_model(kind[b]);
      rank += (mb < m) | ((mb == m) & (b < a));
    }
    slot[a] = rank;
    orig[rank] = a;
  }
}

__global__ void fo_agents_pack_kernel(FoAgentsRaw raw, float m_ego, float4* s0, float4* s1, float2* s2, AgentParams* prm,
                                      float4* t0, float* tv, float* tpsi, float4* aw, float* avw,
                                      const int32_t* __restrict__ orig, const int32_t* __restrict__ slot, int Ap) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int A = raw.n_agents, Tp = raw.t_stride;
  const int nW = agent_windows(Tp);
  if (idx < Ap * nW) {                       // window boxes for the summary kernel's filter (slot order)
    const int sl = idx / nW, wi = idx - sl * nW;
    float xl = 1e30f, xh = -1e30f, yl = 1e30f, yh = -1e30f, vm = 0.0f;
    if (sl < A) {
      const int a = orig[sl];
      const int nS = min(raw.n_states[a], Tp);
      const int lo = wi * kWinSteps;
      for (int i = max(lo - 1, 0); i < min(lo + kWinSteps, nS); ++i) {
        const float x = raw.x[a * Tp + i], y = raw.y[a * Tp + i];
        xl = fminf(xl, x); xh = fmaxf(xh, x); yl = fminf(yl, y); yh = fmaxf(yh, y);
        if (i >= lo) vm = fmaxf(vm, fabsf(raw.v[a * Tp + i]));
      }
    }
    aw[(size_t)wi * Ap + sl] = make_float4(xl, xh, yl, yh);
    avw[(size_t)wi * Ap + sl] = vm;
  }
  if (idx >= A * Tp && idx < Ap * Tp) {      // padding slots of the time-major copies
    const int sl = idx / Tp, i = idx - sl * Tp;
    t0[(size_t)i * Ap + sl] = make_float4(0, 0, 1, 0);
    tv[(size_t)i * Ap + sl] = 0.0f;
    tpsi[(size_t)i * Ap + sl] = 0.0f;
  }
  if (idx < A * Tp) {
    const int a = idx / Tp, i = idx - a * Tp;
    const int sl = slot[a];
    float4 o0 = make_float4(0, 0, 1, 0), o1 = make_float4(0, 0, 0, 0);
    float2 o2 = make_float2(1, 1);
    if (i < raw.n_states[a]) {
      float yaw = raw.yaw[idx];
      float sn, cs;
      sincosf(yaw, &sn, &cs);
      const int ip = i > 0 ? idx - 1 : idx;   // state i-1 (CP pairs ego step i with agent position / covariance i-1)
      float vx = raw.var_x[ip], vy = raw.var_y[ip];
      if (vx == 0.0f && vy == 0.0f) { vx = 0.1f; vy = 0.1f; }  // collision_probability.py:85-87
      o0 = make_float4(raw.x[idx], raw.y[idx], cs, sn);
      o1 = make_float4(yaw, raw.v[idx], raw.x[ip], raw.y[ip]);
      o2 = make_float2(rsqrtf(2.0f * vx), rsqrtf(2.0f * vy));
    }
    const size_t am = (size_t)sl * Tp + i, tm = (size_t)i * Ap + sl;
    s0[am] = o0;
    s1[am] = o1;
    s2[am] = o2;
    t0[tm] = o0;
    tv[tm] = o1.y;
    tpsi[tm] = o1.x;
  }
  if (idx < A) {
    AgentParams p;
    p.n_states = raw.n_states[idx];
    p.model = protection_model(raw.kind[idx]);
    p.hl = 0.5f * raw.length[idx];
    p.hw = 0.5f * raw.width[idx];
    p.hlb = 0.5f * raw.buf_length[idx];
    float mo = obstacle_mass(raw.kind[idx], raw.buf_length[idx] * raw.buf_width[idx]);  // harm_model.py:73,78
    p.ke = mo / (m_ego + mo);
    p.ko = m_ego / (m_ego + mo);
    p.pad = sqrtf(p.hl * p.hl + p.hw * p.hw);   // circumradius of the unbuffered rectangle (distance lower bound)
    prm[slot[idx]] = p;
  }
}

}  // namespace fo

// =============================================================================================
extern "C" size_t fo_agent_table_bytes(int32_t n_agents, int32_t t_stride) {
  if (n_agents < 0 || t_stride < 0) return 0;
  return fo::agent_table_bytes(n_agents, t_stride);
}

extern "C" int fo_agents_pack(const FoAgentsRaw* raw, const FoVehicle* vehicle, void* table_dev, size_t table_bytes,
                              void* stream) {
  if (!raw || !vehicle) { fo::set_error("fo_agents_pack: NULL argument"); return FO_ERR_INVALID_ARG; }
  const int A = raw->n_agents, Tp = raw->t_stride;
  if (A < 0 || Tp < 0) { fo::set_error("fo_agents_pack: negative size"); return FO_ERR_INVALID_ARG; }
  if (A == 0 || Tp == 0) return FO_OK;
  if (Tp > FO_MAX_STATES) { fo::set_error("fo_agents_pack: t_stride %d > FO_MAX_STATES", Tp); return FO_ERR_UNSUPPORTED; }
  if (!table_dev || table_bytes < fo::agent_table_bytes(A, Tp)) {
    fo::set_error("fo_agents_pack: table buffer too small (%zu < %zu)", table_bytes, fo::agent_table_bytes(A, Tp));
    return FO_ERR_INVALID_ARG;
  }
  if (!raw->x || !raw->y || !raw->yaw || !raw->v || !raw->var_x || !raw->var_y || !raw->n_states || !raw->kind ||
      !raw->length || !raw->width || !raw->buf_length || !raw->buf_width) {
    fo::set_error("fo_agents_pack: NULL array in FoAgentsRaw");
    return FO_ERR_INVALID_ARG;
  }
  fo::AgentTableView v = fo::agent_table_view(table_dev, A, Tp);
  const int total = v.Ap * Tp;
  fo::fo_agents_order_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(raw->kind, A, v.Ap, const_cast<int32_t*>(v.orig),
                                                                  const_cast<int32_t*>(v.slot));
  fo::fo_agents_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      *raw, vehicle->mass, const_cast<float4*>(v.s0), const_cast<float4*>(v.s1), const_cast<float2*>(v.s2),
      const_cast<fo::AgentParams*>(v.prm), const_cast<float4*>(v.t0), const_cast<float*>(v.tv),
      const_cast<float*>(v.tpsi), const_cast<float4*>(v.aw), const_cast<float*>(v.avw), v.orig, v.slot, v.Ap);
  fo::count_launch(2);
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}

// SM count per device (a process may drive several GPUs, one per planner thread): cached per ordinal, racing
// first calls store the same value
static int num_sms_of_current_device(int* out) {
  constexpr int kMaxDev = 64;
  static std::atomic<int> cache[kMaxDev];
  int dev = 0;
  FO_CUDA_TRY(cudaGetDevice(&dev));
  int n = (dev >= 0 && dev < kMaxDev) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (n == 0) {
    FO_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < kMaxDev) cache[dev].store(n, std::memory_order_relaxed);
  }
  *out = n;
  return FO_OK;
}
static int metric_bundle_impl(const FoMetricArgs* a, unsigned long long* stats, void* stream);

extern "C" int fo_metric_bundle(const FoMetricArgs* a, void* stream) { return metric_bundle_impl(a, nullptr, stream); }

extern "C" int fo_metric_stats(const FoMetricArgs* a, uint64_t* counters_dev, void* stream) {
  if (!counters_dev) { fo::set_error("fo_metric_stats: counters_dev must not be NULL"); return FO_ERR_INVALID_ARG; }
  if (a && (a->pair || a->step)) { fo::set_error("fo_metric_stats: summary outputs only"); return FO_ERR_INVALID_ARG; }
  FO_CUDA_TRY(cudaMemsetAsync(counters_dev, 0, FO_STATS_K * sizeof(uint64_t), (cudaStream_t)stream));
  return metric_bundle_impl(a, reinterpret_cast<unsigned long long*>(counters_dev), stream);
}

static int metric_bundle_impl(const FoMetricArgs* a, unsigned long long* stats, void* stream) {
  if (!a) { fo::set_error("fo_metric_bundle: NULL args"); return FO_ERR_INVALID_ARG; }
  if (a->n_traj < 0 || a->n_states < 0 || a->n_agents < 0) { fo::set_error("fo_metric_bundle: negative size"); return FO_ERR_INVALID_ARG; }
  if (a->n_traj == 0) return FO_OK;
  if (!a->ego || !a->valid) { fo::set_error("fo_metric_bundle: ego/valid must not be NULL"); return FO_ERR_INVALID_ARG; }
  if (a->n_states < 1 || a->n_states > FO_MAX_STATES) {
    fo::set_error("fo_metric_bundle: n_states %d outside [1, %d]", a->n_states, FO_MAX_STATES);
    return FO_ERR_UNSUPPORTED;
  }
  if (a->n_agents > 0 && (!a->agent_table || a->t_stride < 1 || a->t_stride > FO_MAX_STATES)) {
    fo::set_error("fo_metric_bundle: agent table missing or t_stride out of range");
    return FO_ERR_INVALID_ARG;
  }
  if ((a->metric_mask & FO_M_BE) && !(a->metric_mask & FO_M_TTC)) {
    fo::set_error("fo_metric_bundle: 'be' needs 'ttc' (reference raises KeyError, be.py:39)");
    return FO_ERR_INVALID_ARG;
  }
  int g_num_sms = 0;
  { const int rc = num_sms_of_current_device(&g_num_sms); if (rc != FO_OK) return rc; }
  fo::MetricKArgs k;
  k.ego = a->ego; k.N = a->n_traj; k.T = a->n_states; k.A = a->n_agents; k.Tp = a->t_stride;
  k.tab = fo::agent_table_view(a->agent_table, a->n_agents, a->t_stride);
  k.hEx = 0.5f * a->vehicle.length; k.hEy = 0.5f * a->vehicle.width;
  k.wb = a->vehicle.wb_rear_axle; k.a_max = a->vehicle.a_max;
  k.L3 = a->vehicle.length / 3.0f; k.L6 = a->vehicle.length / 6.0f; k.W2 = a->vehicle.width / 2.0f;
  k.hc = a->harm; k.dt = (float)a->dt; k.dtd = a->dt;
  k.mmask = a->metric_mask; k.tmask = a->threshold_mask;
  k.thr_harm = a->thr_harm; k.thr_risk = a->thr_risk; k.thr_be = a->thr_be; k.thr_cp = a->thr_cp;
  k.thr_ttc = a->thr_ttc; k.thr_dce = a->thr_dce;
  {
    // smallest r with r / 1000.0 >= thr_dce (the reference compares np.round(d, 3) = r / 1000.0 in float64)
    long long c = (long long)ceil(a->thr_dce * 1000.0);
    if (!(a->thr_dce > 0.0)) c = 0;
    if (c > 0xffffff) c = 0xffffff;
    while (c > 0 && (double)(c - 1) / 1000.0 >= a->thr_dce) --c;
    while (c < 0xffffff && (double)c / 1000.0 < a->thr_dce) ++c;
    k.thr_dce_mm = (uint32_t)c;
    // smallest step with np.round(step * dt, 3) >= thr_ttc
    uint32_t q = 0;
    while (q < (uint32_t)FO_MAX_STATES + 1u && rint((double)q * a->dt * 1000.0) / 1000.0 < a->thr_ttc) ++q;
    k.thr_ttc_col = q;
  }
  k.valid = a->valid; k.summary = a->summary; k.flags = a->flags; k.pair = a->pair; k.step = a->step;
  k.stats = stats;
  { const char* e = getenv("FO_EXACT_DCE"); k.exact_dce = e && e[0] == '1'; }
  k.claim = nullptr;
  if (a->summary && (reinterpret_cast<uintptr_t>(a->summary) & 7u)) { fo::set_error("fo_metric_bundle: summary must be 8-byte aligned"); return FO_ERR_INVALID_ARG; }
  if (a->n_peers < 0 || a->n_peers > FO_MAX_PEERS) { fo::set_error("fo_metric_bundle: n_peers %d outside [0, %d]", a->n_peers, FO_MAX_PEERS); return FO_ERR_INVALID_ARG; }
  if (a->n_peers > 0 && (a->pair || a->step || stats)) { fo::set_error("fo_metric_bundle: peer stores exist on the summary path only"); return FO_ERR_UNSUPPORTED; }
  k.n_peers = a->n_peers;
  for (int p = 0; p < FO_MAX_PEERS; ++p) k.peer_delta[p] = p < a->n_peers ? (long long)a->peer_delta[p] : 0;

  cudaStream_t st = (cudaStream_t)stream;
  if (k.pair || k.step) return fo::launch_metric_detail(k, g_num_sms, st);
  return fo::launch_metric_sweep(k, g_num_sms, st);
}
