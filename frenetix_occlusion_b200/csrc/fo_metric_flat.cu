// Dense metric core, sm_100a -- SUMMARY kernel (the throughput / latency path).
//
// Emits only what the planner consumes per trajectory: the validity mask (metrics/metric.py:50-98),
// the summary vector and the flags.  Mapping (B200-first, not a translation of the reference loops):
//
//   * one warp owns one trajectory; its T ego states are staged once in shared memory;
//   * the (agent, time) plane of an agent tile is FLATTENED over the lanes: lane l of chunk c works
//     on flat index f = 32c + l -> (agent f / T, step f % T).  Every lane is busy for any T
//     (31, 51, ...), consecutive lanes read consecutive agent-table entries (coalesced 16-byte
//     loads through L1/L2; the table is at most a few hundred kB and shared by every warp);
//   * the dense part per (n, a, i) is branch-light: oriented-box distance (SAT + 8 corner/box
//     distances), harm logit, 5 m CP gate.  All per-trajectory quantities the reference derives
//     from them are order-independent min/max reductions, so they are accumulated per lane and
//     reduced once per trajectory (warp REDUX);
//   * the expensive, sparse work is COMPACTED: evaluations inside the 5 m gate are pushed on a
//     per-warp shared-memory queue and drained 32 at a time with all lanes active (36 erfc per
//     item); colliding pairs are recorded with shared-memory atomics and handed to the
//     warp-cooperative BE bisection afterwards;
//   * max_t harm = logistic(max_t logit) (the logistic is monotone), so the dense loop needs no
//     exp/div at all.
//
// Reference semantics per SURVEY.md appendix A; file:line citations are on the helpers in
// fo_metric_dev.cuh and on the detail kernel (fo_metric_detail.cu), which computes the same numbers.
#include <stdlib.h>

#include "fo_metric_dev.cuh"

namespace fo {

constexpr int kFlatWarps = 8;
constexpr int kTileAgents = 256;
constexpr int kQueueCap = 64;

__host__ __device__ inline size_t flat_warp_bytes(int T) {
  size_t b = (size_t)kTileAgents * 8 + (size_t)T * (16 + 8 + 4) + (size_t)kTileAgents * 4 + kQueueCap * 4 + kBeBuckets + 16;
  return (b + 15) & ~(size_t)15;
}

struct WarpSmem {
  unsigned long long* pairkey;  // [kTileAgents] (cp bits << 32 | (0xffff - t) << 16 | 1): argmax_t cp, first index on ties
  float4* egoA;                 // [T] (x, y, cos theta, sin theta)
  float2* egoB;                 // [T] (theta, v)
  float* dist;                  // [T] cumulative chord length (BE)
  uint32_t* colfirst;           // [kTileAgents] first step with rounded distance 0 (0xffffffff = none)
  uint32_t* queue;              // [kQueueCap] gated (agent-in-tile << 8 | step) items
  uint8_t* inv;                 // [kBeBuckets + 1] last state index with dist <= b * dmax / kBeBuckets
};

__device__ __forceinline__ WarpSmem warp_smem(unsigned char* base, int T) {
  WarpSmem w;
  w.pairkey = reinterpret_cast<unsigned long long*>(base);
  w.egoA = reinterpret_cast<float4*>(w.pairkey + kTileAgents);
  w.egoB = reinterpret_cast<float2*>(w.egoA + T);
  w.dist = reinterpret_cast<float*>(w.egoB + T);
  w.colfirst = reinterpret_cast<uint32_t*>(w.dist + T);
  w.queue = w.colfirst + kTileAgents;
  w.inv = reinterpret_cast<uint8_t*>(w.queue + kQueueCap);
  return w;
}

__device__ __forceinline__ AgentParams load_params(const AgentParams* p) {
  const int4* q = reinterpret_cast<const int4*>(p);
  int4 a = __ldg(q), b = __ldg(q + 1);
  AgentParams r;
  r.n_states = a.x; r.model = a.y; r.hl = __int_as_float(a.z); r.hw = __int_as_float(a.w);
  r.hlb = __int_as_float(b.x); r.ke = __int_as_float(b.y); r.ko = __int_as_float(b.z); r.pad = __int_as_float(b.w);
  return r;
}

// harm logits (he = 1/(1+exp(-ze)), ho likewise); harm_model.py:81-105, logistic_regression.py:35-48,71-73
__device__ __forceinline__ void harm_logits(const MetricKArgs& k, const AgentParams& P, float dv, float dxr, float dyr,
                                            float th, float psi, float& ze, float& zo) {
  if (P.model == 0) {
    ze = fmaf(k.hc.ia_speed * P.ke, dv, k.hc.ia_const);
    zo = fmaf(k.hc.ped_speed * P.ko, dv, -k.hc.ped_const);
  } else if (P.model == 1) {
    const float PI_F = 3.14159265358979323846f;
    float rel = atan2f(dyr, dxr);
    float ae = rel - th;
    float ao = PI_F + rel - psi;
    ze = fmaf(k.hc.rs_speed * P.ke, dv, k.hc.rs_const) + lr4s_coef(ae, k.hc.rs_side, k.hc.rs_rear);
    zo = fmaf(k.hc.rs_speed * P.ko, dv, k.hc.rs_const) + lr4s_coef(ao, k.hc.rs_side, k.hc.rs_rear);
  } else {
    ze = CUDART_INF_F;
    zo = CUDART_INF_F;
  }
}

// Dense-loop variant: only the running maxima of the logits are needed, so the LR4S impact-angle
// class (atan2 + comparisons) is evaluated only when the upper bound logit (largest area
// coefficient) could still raise one of the maxima.  Exact: skipped evaluations cannot change a max.
__device__ __forceinline__ void harm_logits_max(const MetricKArgs& k, const AgentParams& P, float dv, float dxr,
                                                float dyr, float th, float psi, float& acc_ze, float& acc_zo) {
  if (P.model == 0) {
    acc_ze = fmaxf(acc_ze, fmaf(k.hc.ia_speed * P.ke, dv, k.hc.ia_const));
    acc_zo = fmaxf(acc_zo, fmaf(k.hc.ped_speed * P.ko, dv, -k.hc.ped_const));
  } else if (P.model == 1) {
    const float cmax = fmaxf(0.0f, fmaxf(k.hc.rs_side, k.hc.rs_rear));
    const float be = fmaf(k.hc.rs_speed * P.ke, dv, k.hc.rs_const), bo = fmaf(k.hc.rs_speed * P.ko, dv, k.hc.rs_const);
    if (be + cmax > acc_ze || bo + cmax > acc_zo) {
      const float PI_F = 3.14159265358979323846f;
      float rel = atan2f(dyr, dxr);
      acc_ze = fmaxf(acc_ze, be + lr4s_coef(rel - th, k.hc.rs_side, k.hc.rs_rear));
      acc_zo = fmaxf(acc_zo, bo + lr4s_coef(PI_F + rel - psi, k.hc.rs_side, k.hc.rs_rear));
    }
  } else {
    acc_ze = CUDART_INF_F;
    acc_zo = CUDART_INF_F;
  }
}

__device__ __forceinline__ float sigmoid(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }

// ---------------------------------------------------------------------------------------------
// MASK: compile-time metric mask (0 = read k.mmask at run time).  PRUNE: skip the exact oriented-box
// distance when a circumcircle lower bound proves it cannot lower the lane's running minimum.
template <uint32_t MASK, bool PRUNE, int MINB>
__global__ void __launch_bounds__(kFlatWarps * 32, MINB) fo_metric_flat_kernel(const __grid_constant__ MetricKArgs k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int T = k.T;
  const WarpSmem w = warp_smem(smem_raw + (size_t)wib * flat_warp_bytes(T), T);
  const int warp0 = blockIdx.x * kFlatWarps + wib;
  const int nwarps = gridDim.x * kFlatWarps;
  const uint32_t mm = MASK ? MASK : k.mmask;
  const bool do_cp = mm & FO_M_CP, do_dce = mm & FO_M_DCE, do_hr = mm & FO_M_HR, do_be = mm & FO_M_BE,
             do_ttc = mm & FO_M_TTC;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float rE = sqrtf(k.hEx * k.hEx + k.hEy * k.hEy);   // ego circumradius

  for (int n = warp0; n < k.N; n += nwarps) {
    // ---- stage the ego trajectory ------------------------------------------------------------
    float amin = 0.0f;
    const float* eg = k.ego + (size_t)n * T * 5;
    __syncwarp();
    for (int i = lane; i < T; i += 32) {
      float x = __ldg(eg + i * 5 + 0), y = __ldg(eg + i * 5 + 1), th = __ldg(eg + i * 5 + 2);
      float v = __ldg(eg + i * 5 + 3);
      amin = fminf(amin, __ldg(eg + i * 5 + 4));
      float sn, cs;
      sincosf(th, &sn, &cs);
      w.egoA[i] = make_float4(x, y, cs, sn);
      w.egoB[i] = make_float2(th, v);
    }
    __syncwarp();

    // per-lane accumulators (order-independent reductions, finished once per trajectory)
    uint32_t acc_rmin = 0xffffffu, acc_col = 0xffffffffu;
    float acc_ze = -CUDART_INF_F, acc_zo = -CUDART_INF_F;
    float acc_er = 0.0f, acc_or = 0.0f, acc_cp = 0.0f, acc_hwc = 0.0f;
    float btn_all = 0.0f, rcd_all = 0.0f;   // warp-uniform
    uint32_t flags = 0;
    bool be_ready = false;
    float be_lo0 = 0.0f;
    const BeView bev{w.egoA, w.egoB, w.dist, w.inv};

    for (int a0 = 0; a0 < k.A; a0 += kTileAgents) {
      const int nAt = min(kTileAgents, k.A - a0);
      for (int j = lane; j < nAt; j += 32) { w.pairkey[j] = 0ull; w.colfirst[j] = 0xffffffffu; }
      __syncwarp();
      const int F = nAt * T;
      int al = 0, i = lane;
      while (i >= T) { i -= T; ++al; }
      int qn = 0;

      for (int f0 = 0;; f0 += 32) {
        const bool more = f0 < F;
        if (more) {
          // ================= dense part: one (agent, step) per lane ===========================
          const bool active = f0 + lane < F;
          bool ingate = false;
          if (active) {
            const int a = a0 + al;
            const AgentParams P = load_params(k.tab.prm + a);
            if (i < P.n_states) {
              const size_t idx = (size_t)a * k.Tp + i;
              const float4 s0 = __ldg(&k.tab.s0[idx]);
              const float4 s1 = __ldg(&k.tab.s1[idx]);
              const float4 EA = w.egoA[i];
              const float2 EB = w.egoB[i];
              const float c = fmaf(EA.z, s0.z, EA.w * s0.w);    // cos(yaw - theta)
              const float s = fmaf(s0.w, EA.z, -s0.z * EA.w);   // sin(yaw - theta)
              const float dxr = s0.x - EA.x, dyr = s0.y - EA.y;
              if (do_dce) {                                     // i < min(T, n_states)
                const float dx = dxr - k.wb * EA.z, dy = dyr - k.wb * EA.w;   // centre to centre
                bool need = true;
                if (PRUNE) {
                  // distance >= |centres| - rE - rO; skip when that bound is >= running min + 2 mm
                  const float lim = rE + P.pad + (float)(acc_rmin + 2u) * 0.001f;
                  need = fmaf(dx, dx, dy * dy) < lim * lim;
                }
                if (need) {
                  float rx = fmaf(dx, EA.z, dy * EA.w), ry = fmaf(dy, EA.z, -dx * EA.w);
                  float d = sqrtf(obb_d2(rx, ry, c, s, k.hEx, k.hEy, P.hl, P.hw));
                  uint32_t r = (uint32_t)__float2int_rn(fminf(d, 8000.0f) * 1000.0f);   // np.round(d, 3)
                  acc_rmin = min(acc_rmin, r);
                  if (r == 0u) atomicMin(&w.colfirst[al], (uint32_t)i);
                }
              }
              if (do_hr && i < T - 1) {                          // t < min(T-1, n_states)
                float dv2 = fmaxf(fmaf(EB.y, EB.y, s1.y * s1.y) - 2.0f * EB.y * s1.y * c, 0.0f);
                float dv = dv2 * rsqrtf(fmaxf(dv2, 1e-30f));        // |dv|, ~2 ulp: harm tolerance is 1e-4
                harm_logits_max(k, P, dv, dxr, dyr, EB.x, s1.x, acc_ze, acc_zo);
              }
              if (do_cp && i >= 1) {                             // 5 m gate, collision_probability.py:61-78
                float mx = s1.z - EA.x, my = s1.w - EA.y;
                float hx = P.hlb * s0.z, hy = P.hlb * s0.w;
                float d0 = fmaf(mx, mx, my * my);
                float d1 = fmaf(mx + hx, mx + hx, (my + hy) * (my + hy));
                float d2 = fmaf(mx - hx, mx - hx, (my - hy) * (my - hy));
                ingate = fminf(d0, fminf(d1, d2)) <= 25.0f;
              }
            }
          }
          const unsigned b = __ballot_sync(kFull, ingate);
          if (b) {
            if (ingate) w.queue[qn + __popc(b & lt_mask)] = ((uint32_t)al << 8) | (uint32_t)i;
            qn += __popc(b);
            __syncwarp();
          }
          i += 32;
          if (T >= 32) { if (i >= T) { i -= T; ++al; } }
          else { while (i >= T) { i -= T; ++al; } }
        }
        // ================= sparse part: drain gated evaluations, 32 per round =====================
        while (qn >= 32 || (!more && qn > 0)) {
          const int cnt = min(qn, 32);
          uint32_t item = 0;
          if (lane < cnt) item = w.queue[qn - cnt + lane];
          qn -= cnt;
          __syncwarp();
          if (lane < cnt) {
            const int ial = (int)(item >> 8), ii = (int)(item & 0xffu), t = ii - 1;
            const int a = a0 + ial;
            const AgentParams P = load_params(k.tab.prm + a);
            const size_t idx = (size_t)a * k.Tp + ii;
            const float4 s0i = __ldg(&k.tab.s0[idx]);
            const float4 s1i = __ldg(&k.tab.s1[idx]);
            const float2 s2i = __ldg(&k.tab.s2[idx]);
            const float4 Ei = w.egoA[ii];
            const float cp = cp_gauss_boxes(s1i.z - Ei.x, s1i.w - Ei.y, P.hlb * s0i.z, P.hlb * s0i.w, k.L3 * Ei.z,
                                            k.L3 * Ei.w, s2i, k.L6, k.W2);
            acc_cp = fmaxf(acc_cp, cp);
            if (do_hr) {                       // risk[t] = harm[t] * cp[t], cp[t] = CP of step t+1 (hr.py:78-79)
              const float4 s0t = __ldg(&k.tab.s0[idx - 1]);
              const float4 s1t = __ldg(&k.tab.s1[idx - 1]);
              const float4 Et = w.egoA[t];
              const float2 EtB = w.egoB[t];
              const float ct = fmaf(Et.z, s0t.z, Et.w * s0t.w);
              const float dv = sqrtf(fmaxf(fmaf(EtB.y, EtB.y, s1t.y * s1t.y) - 2.0f * EtB.y * s1t.y * ct, 0.0f));
              float ze, zo;
              harm_logits(k, P, dv, s0t.x - Et.x, s0t.y - Et.y, EtB.x, s1t.x, ze, zo);
              acc_er = fmaxf(acc_er, sigmoid(ze) * cp);
              acc_or = fmaxf(acc_or, sigmoid(zo) * cp);
              if (cp > 0.01f)                  // candidates for obst_harm[argmax cp] (hr.py:81-84)
                atomicMax(&w.pairkey[ial], ((unsigned long long)__float_as_uint(cp) << 32) |
                                               ((unsigned long long)(0xffffu - (unsigned)t) << 16) | 1ull);
            }
          }
          __syncwarp();
        }
        if (!more) break;
      }

      // ================= tile epilogue: harm_with_cp, wttc, BE ======================================
      for (int j0 = 0; j0 < nAt; j0 += 32) {
        const int j = j0 + lane;
        const unsigned long long key = (j < nAt) ? w.pairkey[j] : 0ull;
        const uint32_t cf = (j < nAt) ? w.colfirst[j] : 0xffffffffu;
        if (key != 0ull) {
          const int t = (int)(0xffffu - (unsigned)((key >> 16) & 0xffffu));
          const int a = a0 + j;
          const AgentParams P = load_params(k.tab.prm + a);
          const size_t idx = (size_t)a * k.Tp + t;
          const float4 s0t = __ldg(&k.tab.s0[idx]);
          const float4 s1t = __ldg(&k.tab.s1[idx]);
          const float4 Et = w.egoA[t];
          const float2 EtB = w.egoB[t];
          const float ct = fmaf(Et.z, s0t.z, Et.w * s0t.w);
          const float dv = sqrtf(fmaxf(fmaf(EtB.y, EtB.y, s1t.y * s1t.y) - 2.0f * EtB.y * s1t.y * ct, 0.0f));
          float ze, zo;
          harm_logits(k, P, dv, s0t.x - Et.x, s0t.y - Et.y, EtB.x, s1t.x, ze, zo);
          acc_hwc = fmaxf(acc_hwc, sigmoid(zo));
        }
        if (do_ttc) acc_col = min(acc_col, cf);
        if (do_be && do_ttc) {
          unsigned m = __ballot_sync(kFull, cf != 0xffffffffu && cf > 0u);   // be.py:49-50
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int a = a0 + j0 + src;
            const AgentParams P = load_params(k.tab.prm + a);
            if (!be_ready) {
              be_prepare(bev, T, lane);
              float am = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(fabsf(amin))));
              be_lo0 = rintf(am * 100.0f) / 100.0f;                          // be.py:68
              be_ready = true;
            }
            bool range_err = false;
            unsigned probes = 0;
            const float rcd = be_bisect(k, bev, a, P.n_states, P.hl, P.hw, be_lo0, lane, range_err, probes);
            if (range_err) flags |= FO_F_BE_RANGE;
            rcd_all = fmaxf(rcd_all, rcd);
            btn_all = fmaxf(btn_all, rcd / k.a_max);
          }
        }
      }
      __syncwarp();
    }  // agent tiles

    // ---- per-trajectory reduction and threshold mask (metric.py:50-98) ----------------------------
    const float er = umaxf(acc_er), orr = umaxf(acc_or), cpm = umaxf(acc_cp), hwc_all = umaxf(acc_hwc);
    float ze = acc_ze, zo = acc_zo;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ze = fmaxf(ze, __shfl_xor_sync(kFull, ze, o));
      zo = fmaxf(zo, __shfl_xor_sync(kFull, zo, o));
    }
    const uint32_t rmin = __reduce_min_sync(kFull, acc_rmin);
    const uint32_t col = __reduce_min_sync(kFull, acc_col);
    if (lane == 0) {
      const float eh = sigmoid(ze), oh = sigmoid(zo);   // logistic is monotone: max harm = logistic(max logit)
      const bool has_agents = k.A > 0 && mm != 0;
      const double dce_min = (double)rmin / 1000.0;
      const bool has_col = col != 0xffffffffu;
      const double wttc = has_col ? rint((double)col * k.dtd * 1000.0) / 1000.0 : (double)CUDART_INF;
      bool ok = true;
      if (has_agents) {
        if (do_be && (k.tmask & FO_T_BE) && (double)btn_all > k.thr_be) ok = false;
        if (do_hr && (k.tmask & FO_T_HARM) && (double)hwc_all > k.thr_harm) ok = false;
        if (do_hr && (k.tmask & FO_T_RISK) && (double)orr > k.thr_risk) ok = false;
        if (do_hr && (k.tmask & FO_T_CP) && (double)cpm > k.thr_cp) ok = false;
        if (do_ttc && (k.tmask & FO_T_TTC) && has_col && wttc < k.thr_ttc) ok = false;
        if (do_dce && (k.tmask & FO_T_DCE) && rmin != 0xffffffu && dce_min < k.thr_dce) ok = false;
        if (flags & FO_F_BE_RANGE) ok = false;
      }
      k.valid[n] = ok ? 1 : 0;
      if (k.flags) k.flags[n] = flags;
      if (k.summary) {
        float* sm = k.summary + (size_t)n * FO_SUMMARY_K;
        sm[0] = er; sm[1] = orr; sm[2] = do_hr ? eh : 0.0f; sm[3] = do_hr ? oh : 0.0f; sm[4] = cpm; sm[5] = hwc_all;
        sm[6] = (rmin == 0xffffffu || !do_dce) ? CUDART_INF_F : (float)dce_min;
        sm[7] = has_col ? (float)wttc : CUDART_INF_F;
        sm[8] = (flags & FO_F_BE_RANGE) ? CUDART_NAN_F : btn_all;
        sm[9] = (flags & FO_F_BE_RANGE) ? CUDART_NAN_F : rcd_all;
      }
    }
  }
}

template <uint32_t MASK, bool PRUNE, int MINB>
static int launch_flat_inst(const MetricKArgs& k, int num_sms, cudaStream_t st) {
  const size_t smem = flat_warp_bytes(k.T) * kFlatWarps;
  static size_t configured = 0;
  static int per_sm = 1;
  if (smem != configured) {
    FO_CUDA_TRY(cudaFuncSetAttribute(fo_metric_flat_kernel<MASK, PRUNE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    FO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fo_metric_flat_kernel<MASK, PRUNE, MINB>,
                                                              kFlatWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    configured = smem;
  }
  const int ctas_needed = (k.N + kFlatWarps - 1) / kFlatWarps;
  const int full = num_sms * per_sm;
  const int grid = ctas_needed < full ? ctas_needed : full;   // persistent: warps stride over trajectories
  fo_metric_flat_kernel<MASK, PRUNE, MINB><<<grid, kFlatWarps * 32, smem, st>>>(k);
  count_launch();
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}

int launch_metric_flat(const MetricKArgs& k, int num_sms, cudaStream_t st) {
  constexpr uint32_t kAll = FO_M_CP | FO_M_DCE | FO_M_TTC | FO_M_HR | FO_M_BE | FO_M_TTCE | FO_M_WTTC;
  constexpr uint32_t kDefault = kAll & ~FO_M_BE;   // occlusion.yaml:12-18
  static const bool no_prune = getenv("FO_NO_PRUNE") != nullptr;   // measurement switch
  static const bool minb4 = getenv("FO_FLAT_MINB4") != nullptr;     // occupancy experiment switch (64 regs, spills)
  if (no_prune) return launch_flat_inst<0u, false, 3>(k, num_sms, st);
  if (k.mmask == kAll) return minb4 ? launch_flat_inst<kAll, true, 4>(k, num_sms, st) : launch_flat_inst<kAll, true, 3>(k, num_sms, st);
  if (k.mmask == kDefault) return launch_flat_inst<kDefault, true, 3>(k, num_sms, st);
  return launch_flat_inst<0u, true, 3>(k, num_sms, st);
}

}  // namespace fo
