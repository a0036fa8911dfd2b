// Dense metric core of the Frenetix-Occlusion assessment path, sm_100a -- DETAIL kernel.
//
// Used when the caller wants the per-pair / per-step arrays of the reference result dict
// (drop-in Metric.evaluate_metrics); the throughput / latency path is fo_metric_sweep.cu.
//
// One warp owns one ego trajectory; lane = time index (T = 31 fits one warp, longer horizons run
// NP = ceil(T/32) register passes); the warp loops over all phantom predictions.  Per
// (trajectory n, prediction a, step i) it evaluates what the reference does in Python loops:
//   CP   metrics/utils/collision_probability.py:37-124   (Gaussian mass over three ego boxes, 5 m gate)
//   DCE  metrics/dce.py:52-99 + utils/convert_dynamic_obstacle.py  (oriented-box distance, round 1e-3)
//   TTC / TTCE / WTTC  metrics/ttc.py, ttce.py, wttc.py   (post-processing of DCE)
//   HR   metrics/utils/harm_model.py:58-105, logistic_regression.py, hr.py:76-114
//   BE   metrics/be.py:31-193                              (bisection on constant deceleration)
// followed by the threshold mask of metrics/metric.py:50-98.  Reductions over time are warp
// REDUX/shuffles; nothing but the per-trajectory results (and, on request, the per-pair / per-step
// detail the reference's result dict carries) is written back to HBM.
#include "fo_metric_dev.cuh"

namespace fo {

// ---------------------------------------------------------------------------------------------
// BE: warp-cooperative bisection for one (trajectory, prediction) pair.  be.py:66-193.
// sm_* hold the ego arrays of this warp (arc length `dist`, x, y, theta) for the interpolation.
template <int NP>
__device__ float be_bisect(const MetricKArgs& k, const EgoState (&e)[NP], float lo0, const float* sm_dist,
                           const float* sm_x, const float* sm_y, const float* sm_th, int a, const AgentParams& P,
                           int lane, bool& range_err) {
  const int T = k.T;
  const int nA = min(T, P.n_states);
  const float v0 = __shfl_sync(kFull, e[0].v, 0);
  const float v1 = __shfl_sync(kFull, e[0].v, 1);
  const float dmax = sm_dist[T - 1];
  float lo = lo0, hi = 5.0f, cur = 0.0f;
  for (int it = 0; it < 10; ++it) {
    cur = 0.5f * (lo + hi);
    bool hit = false, over = false;
    float carry = 0.0f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int i = lane + 32 * p;
      // v_new (be.py:109) and its exclusive prefix sum dist_new (be.py:113)
      float vn = (i == 0) ? v0 : fmaxf(fmaf(-cur, (float)(i - 1) * k.dt, v1), 0.0f);
      float inc = (i < T) ? vn * k.dt : 0.0f;
      float sc = inc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(kFull, sc, o);
        if (lane >= o) sc += t;
      }
      float q = carry + sc - inc;
      carry += __shfl_sync(kFull, sc, 31);
      if (i < T) {
        if (q > dmax) over = true;  // interp1d bounds_error (be.py:117-124)
        // numpy.interp: j = last index with dist[j] <= q
        int lo_j = 0, hi_j = T - 1;
        while (lo_j < hi_j) {
          int mid = (lo_j + hi_j + 1) >> 1;
          if (sm_dist[mid] <= q) lo_j = mid; else hi_j = mid - 1;
        }
        const int j = lo_j;
        float xn, yn, tn;
        float dj = sm_dist[j];
        if (j == T - 1 || dj == q) {
          xn = sm_x[j]; yn = sm_y[j]; tn = sm_th[j];
        } else {
          float w = (q - dj);
          float inv = 1.0f / (sm_dist[j + 1] - dj);
          xn = fmaf((sm_x[j + 1] - sm_x[j]) * inv, w, sm_x[j]);
          yn = fmaf((sm_y[j + 1] - sm_y[j]) * inv, w, sm_y[j]);
          tn = fmaf((sm_th[j + 1] - sm_th[j]) * inv, w, sm_th[j]);
        }
        if (i < nA) {
          float sn, cn;
          sincosf(tn, &sn, &cn);
          float4 s0 = __ldg(&k.tab.s0[(size_t)a * k.Tp + i]);
          float dx = (s0.x - xn) - k.wb * cn;
          float dy = (s0.y - yn) - k.wb * sn;
          float rx = fmaf(dx, cn, dy * sn), ry = fmaf(dy, cn, -dx * sn);
          float c = fmaf(cn, s0.z, sn * s0.w), s = fmaf(s0.w, cn, -s0.z * sn);
          hit |= obb_hit(rx, ry, c, s, k.hEx, k.hEy, P.hl, P.hw);
        }
      }
    }
    if (__any_sync(kFull, over)) { range_err = true; return CUDART_NAN_F; }
    bool any_hit = __any_sync(kFull, hit);
    if (nA > 0 && !any_hit) hi = cur; else lo = cur;   // be.py:74-77
    if (hi - lo < 0.1f) break;                         // be.py:79
  }
  return cur;
}

// ---------------------------------------------------------------------------------------------
// One copy each of the two big per-state bodies (36 erfcf; atan2f + two angle classes): the kernel evaluates NP
// register passes per agent and was instruction-cache-bound with them inlined NP times.
// Collision probability of one gated step, collision_probability.py:94-122: same term order as the reference's loops.
static __device__ __noinline__ float detail_cp(float mx, float my, float hx, float hy, float bx, float by, float pix,
                                               float piy, float L6, float W2) {
  float prob = 0.0f;
#pragma unroll 1
  for (int m = 0; m < 3; ++m) {
    const float ux = (m == 0) ? mx : (m == 1 ? mx + hx : mx - hx);
    const float uy = (m == 0) ? my : (m == 1 ? my + hy : my - hy);
#pragma unroll 1
    for (int b = 0; b < 3; ++b) {
      const float cxb = (b == 0) ? 0.0f : (b == 1 ? bx : -bx);
      const float cyb = (b == 0) ? 0.0f : (b == 1 ? by : -by);
      const float px = half_derf((cxb - L6 - ux) * pix, (cxb + L6 - ux) * pix);
      const float py = half_derf((cyb - W2 - uy) * piy, (cyb + W2 - uy) * piy);
      prob = fmaf(px, py, prob);
    }
  }
  return prob * (1.0f / 3.0f);
}

// LR4S impact-angle coefficients of ego and obstacle (logistic_regression.py:35-48, harm_model.py:81-105)
static __device__ __noinline__ float2 detail_lr4s_pair(float dyr, float dxr, float th, float psi, float side, float rear) {
  const float PI_F = 3.14159265358979323846f;
  const float rel = atan2f(dyr, dxr);
  return make_float2(lr4s_coef(rel - th, side, rear), lr4s_coef(PI_F + rel - psi, side, rear));
}

template <int NP>
__global__ void __launch_bounds__(kWarpsPerCta * 32) fo_metric_kernel(const __grid_constant__ MetricKArgs k) {
  __shared__ float sm_be[kWarpsPerCta][4][FO_MAX_STATES];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp0 = blockIdx.x * kWarpsPerCta + wib;
  const int nwarps = gridDim.x * kWarpsPerCta;
  const int T = k.T;
  const bool do_cp = k.mmask & FO_M_CP, do_dce = k.mmask & FO_M_DCE, do_hr = k.mmask & FO_M_HR,
             do_be = k.mmask & FO_M_BE, do_ttc = k.mmask & FO_M_TTC;
  const float PI_F = 3.14159265358979323846f;

  for (int n = warp0; n < k.N; n += nwarps) {
    // ---- ego states into registers (lane = time) --------------------------------------------
    EgoState e[NP];
    float amin = 0.0f;
    const float* eg = k.ego + (size_t)n * T * 5;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int i = lane + 32 * p;
      e[p] = EgoState{0, 0, 0, 0, 1, 0};
      if (i < T) {
        e[p].x = __ldg(eg + i * 5 + 0);
        e[p].y = __ldg(eg + i * 5 + 1);
        e[p].th = __ldg(eg + i * 5 + 2);
        e[p].v = __ldg(eg + i * 5 + 3);
        amin = fminf(amin, __ldg(eg + i * 5 + 4));
        sincosf(e[p].th, &e[p].s, &e[p].c);
      }
    }
    bool be_ready = false;  // arc-length table built lazily
    float be_lo0 = 0.0f;

    // per-lane accumulators over all agents (reduced once per trajectory)
    float acc_er = 0.0f, acc_or = 0.0f, acc_eh = 0.0f, acc_oh = 0.0f, acc_cp = 0.0f;
    uint32_t acc_rmin = 0xffffffu;    // min over (a, i) of round(d*1000); 0xffffff = none
    // warp-uniform accumulators
    float hwc_all = 0.0f, btn_all = 0.0f, rcd_all = 0.0f;
    int wttc_idx = 0x7fffffff;
    uint32_t flags = 0;

    for (int a = 0; a < k.A; ++a) {
      const AgentParams P = k.tab.prm[a];
      const int nA = min(T, P.n_states);       // DCE / CP range (dce.py:87-88, collision_probability.py:73)
      const int nH = min(T - 1, P.n_states);   // harm range (harm_model.py:67)
      float cp[NP], he[NP], ho[NP];
      uint32_t key[NP];
      bool tie[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int i = lane + 32 * p;
        float4 s0 = make_float4(0, 0, 1, 0), s1 = make_float4(0, 0, 0, 0);
        float2 s2 = make_float2(1, 1);
        if (i < P.n_states && i < T) {
          s0 = __ldg(&k.tab.s0[(size_t)a * k.Tp + i]);
          s1 = __ldg(&k.tab.s1[(size_t)a * k.Tp + i]);
          s2 = __ldg(&k.tab.s2[(size_t)a * k.Tp + i]);
        }
        // agent position / covariance of state i-1 (pre-shifted by fo_agents_pack)
        const float ppx = s1.z, ppy = s1.w, pix = s2.x, piy = s2.y;

        const EgoState E = e[p];
        const float c = fmaf(E.c, s0.z, E.s * s0.w);    // cos(yaw - theta)
        const float s = fmaf(s0.w, E.c, -s0.z * E.s);   // sin(yaw - theta)
        const float dxr = s0.x - E.x, dyr = s0.y - E.y;  // agent centre relative to the raw ego point

        // ---- DCE: oriented-box distance, axle-shifted ego centre, unbuffered agent shape ----
        key[p] = 0xffffffffu;
        tie[p] = false;
        if (do_dce && i < nA) {
          float dx = dxr - k.wb * E.c, dy = dyr - k.wb * E.s;
          float rx = fmaf(dx, E.c, dy * E.s), ry = fmaf(dy, E.c, -dx * E.s);
          float d = sqrtf(obb_d2(rx, ry, c, s, k.hEx, k.hEy, P.hl, P.hw));
          uint32_t r = (uint32_t)__float2int_rn(fminf(d, 8000.0f) * 1000.0f);  // np.round(d, 3), dce.py:79
          key[p] = (r << 8) | (uint32_t)i;
          tie[p] = near_rounding_boundary(d);
        }

        // ---- harm at state t = i (same index both sides), harm_model.py:81-105 ----------------
        he[p] = 0.0f; ho[p] = 0.0f;
        if (do_hr && i < nH) {
          float dv = sqrtf(fmaxf(fmaf(E.v, E.v, s1.y * s1.y) - 2.0f * E.v * s1.y * c, 0.0f));
          if (P.model == 1) {
            const float2 cls = detail_lr4s_pair(dyr, dxr, E.th, s1.x, k.hc.rs_side, k.hc.rs_rear);
            he[p] = logistic_neg(-k.hc.rs_const - k.hc.rs_speed * (P.ke * dv) - cls.x);
            ho[p] = logistic_neg(-k.hc.rs_const - k.hc.rs_speed * (P.ko * dv) - cls.y);
          } else if (P.model == 0) {
            he[p] = logistic_neg(-k.hc.ia_const - k.hc.ia_speed * (P.ke * dv));
            ho[p] = logistic_neg(k.hc.ped_const - k.hc.ped_speed * (P.ko * dv));
          } else {
            he[p] = 1.0f; ho[p] = 1.0f;
          }
        }

        // ---- CP of ego step i against agent position i-1 / yaw i / covariance i-1 -------------
        cp[p] = 0.0f;
        if (do_cp && i >= 1 && i < nA) {
          float mx = ppx - E.x, my = ppy - E.y;          // obstacle centre point relative to ego point
          float hx = P.hlb * s0.z, hy = P.hlb * s0.w;    // front/back offset uses yaw[i]
          float d0 = fmaf(mx, mx, my * my);
          float d1 = fmaf(mx + hx, mx + hx, (my + hy) * (my + hy));
          float d2 = fmaf(mx - hx, mx - hx, (my - hy) * (my - hy));
          if (fminf(d0, fminf(d1, d2)) <= 25.0f) {        // strict "> 5.0" gate, collision_probability.py:65-67
            cp[p] = detail_cp(mx, my, hx, hy, k.L3 * E.c, k.L3 * E.s, pix, piy, k.L6, k.W2);
          }
        }
      }

      // ---- rounding ties: candidates for the pair's minimum that sit next to a x.xxx5 boundary are re-rounded from a
      //      float64 evaluation (a float32 value is off by at most one unit there, hence the margin of two) ----------
      if (do_dce) {
        uint32_t kpre = 0xffffffffu;
#pragma unroll
        for (int p = 0; p < NP; ++p) kpre = min(kpre, key[p]);
        kpre = __reduce_min_sync(kFull, kpre);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int i = lane + 32 * p;
          if (tie[p] && (key[p] >> 8) <= (kpre >> 8) + 2u) {
            const float4 s0 = __ldg(&k.tab.s0[(size_t)a * k.Tp + i]);
            const float yaw = __ldg(&k.tab.s1[(size_t)a * k.Tp + i]).x;
            const uint32_t r = obb_round_mm_f64(e[p].x, e[p].y, e[p].th, k.wb, k.hEx, k.hEy, s0.x, s0.y, yaw, P.hl, P.hw);
            key[p] = (r << 8) | (uint32_t)i;
          }
        }
      }

      // ---- align cp with harm: cpn[t] = CP of step t+1 (hr.py:78-79) --------------------------
      float cpn[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        float dn = __shfl_down_sync(kFull, cp[p], 1);
        float nx = (p + 1 < NP) ? __shfl_sync(kFull, cp[(p + 1 < NP) ? p + 1 : p], 0) : 0.0f;
        cpn[p] = (lane == 31) ? nx : dn;
      }

      // ---- per-pair reductions over time -------------------------------------------------------
      float er_l = 0.0f, or_l = 0.0f, eh_l = 0.0f, oh_l = 0.0f, cp_l = 0.0f;
      uint32_t key_l = 0xffffffffu;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        er_l = fmaxf(er_l, he[p] * cpn[p]);
        or_l = fmaxf(or_l, ho[p] * cpn[p]);
        eh_l = fmaxf(eh_l, he[p]);
        oh_l = fmaxf(oh_l, ho[p]);
        cp_l = fmaxf(cp_l, cpn[p]);
        key_l = min(key_l, key[p]);
      }
      acc_er = fmaxf(acc_er, er_l); acc_or = fmaxf(acc_or, or_l);
      acc_eh = fmaxf(acc_eh, eh_l); acc_oh = fmaxf(acc_oh, oh_l);
      acc_cp = fmaxf(acc_cp, cp_l);
      acc_rmin = min(acc_rmin, key_l >> 8);

      // harm_with_cp = obst_harm[argmax cp] if max cp > 0.01 (hr.py:81-84); first index on ties
      float hwc = 0.0f, cpmax = 0.0f;
      int cp_arg = 0;
      if ((do_hr && __any_sync(kFull, cp_l > 0.01f)) || (k.pair && do_cp)) {
        cpmax = umaxf(cp_l);
#pragma unroll
        for (int p = NP - 1; p >= 0; --p) {
          unsigned b = __ballot_sync(kFull, cpn[p] == cpmax);
          if (b) {
            int src = __ffs(b) - 1;
            cp_arg = src + 32 * p;
            float h = __shfl_sync(kFull, ho[p], src);
            hwc = (cpmax > 0.01f) ? h : 0.0f;
          }
        }
        if (!do_hr) hwc = 0.0f;
        hwc_all = fmaxf(hwc_all, hwc);
      }

      // dce / first collision index for this pair (ttc.py:40-46)
      uint32_t kmin = __reduce_min_sync(kFull, key_l);
      const bool collides = do_dce && (kmin >> 8) == 0u && kmin != 0xffffffffu;
      const int t_col = (int)(kmin & 0xffu);
      if (collides && do_ttc) wttc_idx = min(wttc_idx, t_col);

      // ---- BE for colliding pairs with ttc > 0 (be.py:49-56) -----------------------------------
      float rcd = 0.0f, btn = 0.0f;
      if (do_be && do_ttc && collides && t_col > 0) {
        if (!be_ready) {
          // arc length of the ego polyline (be.py:99) and lower bisection bound (be.py:68)
          float carry = 0.0f;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const int i = lane + 32 * p;
            // state i-1: neighbour lane, or lane 31 of the previous pass (all lanes run the shuffles)
            float px = __shfl_up_sync(kFull, e[p].x, 1), py = __shfl_up_sync(kFull, e[p].y, 1);
            float qx = __shfl_sync(kFull, e[p > 0 ? p - 1 : 0].x, 31), qy = __shfl_sync(kFull, e[p > 0 ? p - 1 : 0].y, 31);
            if (lane == 0) { px = qx; py = qy; }
            float seg = (i >= 1 && i < T) ? sqrtf((e[p].x - px) * (e[p].x - px) + (e[p].y - py) * (e[p].y - py)) : 0.0f;
            float sc = seg;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              float t = __shfl_up_sync(kFull, sc, o);
              if (lane >= o) sc += t;
            }
            if (i < T) {
              sm_be[wib][0][i] = carry + sc;
              sm_be[wib][1][i] = e[p].x;
              sm_be[wib][2][i] = e[p].y;
              sm_be[wib][3][i] = e[p].th;
            }
            carry += __shfl_sync(kFull, sc, 31);
          }
          float am = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(fabsf(amin))));  // |min(min a, 0)|
          be_lo0 = rintf(am * 100.0f) / 100.0f;
          __syncwarp();
          be_ready = true;
        }
        bool range_err = false;
        rcd = be_bisect<NP>(k, e, be_lo0, sm_be[wib][0], sm_be[wib][1], sm_be[wib][2], sm_be[wib][3], a, P, lane, range_err);
        if (range_err) flags |= FO_F_BE_RANGE;
        btn = rcd / k.a_max;
        rcd_all = fmaxf(rcd_all, rcd);   // fmaxf ignores NaN
        btn_all = fmaxf(btn_all, btn);
      }

      // ---- optional detail outputs ----------------------------------------------------------------
      if (k.step) {
        float* st = k.step + ((size_t)n * k.A + a) * (size_t)(T - 1) * FO_STEP_K;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int t = lane + 32 * p;
          if (t < T - 1) {
            st[t * FO_STEP_K + 0] = cpn[p];
            st[t * FO_STEP_K + 1] = (do_hr && t < nH) ? he[p] : CUDART_NAN_F;
            st[t * FO_STEP_K + 2] = (do_hr && t < nH) ? ho[p] : CUDART_NAN_F;
          }
        }
      }
      if (k.pair) {
        float er_m = umaxf(er_l), or_m = umaxf(or_l), eh_m = umaxf(eh_l), oh_m = umaxf(oh_l);
        int or_arg = 0;
#pragma unroll
        for (int p = NP - 1; p >= 0; --p) {
          const int t = lane + 32 * p;
          unsigned b = __ballot_sync(kFull, (ho[p] * cpn[p] == or_m) && t < max(nH, 1));
          if (b) or_arg = __ffs(b) - 1 + 32 * p;
        }
        if (lane == 0) {
          float* pr = k.pair + ((size_t)n * k.A + a) * FO_PAIR_K;
          pr[0] = (kmin == 0xffffffffu) ? CUDART_INF_F : (float)((double)(kmin >> 8) / 1000.0);
          pr[1] = (kmin == 0xffffffffu) ? 0.0f : (float)t_col;
          pr[2] = er_m; pr[3] = or_m; pr[4] = (float)or_arg; pr[5] = hwc; pr[6] = eh_m; pr[7] = oh_m;
          pr[8] = cpmax; pr[9] = rcd; pr[10] = btn; pr[11] = (float)cp_arg;
        }
      }
    }  // agents

    // ---- per-trajectory reduction, threshold mask (metric.py:50-98) ------------------------------
    float er = umaxf(acc_er), orr = umaxf(acc_or), eh = umaxf(acc_eh), oh = umaxf(acc_oh), cpm = umaxf(acc_cp);
    uint32_t rmin = __reduce_min_sync(kFull, acc_rmin);
    if (lane == 0) {
      const bool has_agents = k.A > 0 && k.mmask != 0;
      const double dce_min = (double)rmin / 1000.0;
      const bool has_col = wttc_idx != 0x7fffffff;
      const double wttc = has_col ? rint((double)wttc_idx * k.dtd * 1000.0) / 1000.0 : (double)CUDART_INF;
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
        sm[0] = er; sm[1] = orr; sm[2] = eh; sm[3] = oh; sm[4] = cpm; sm[5] = hwc_all;
        sm[6] = (rmin == 0xffffffu || !do_dce) ? CUDART_INF_F : (float)dce_min;
        sm[7] = has_col ? (float)wttc : CUDART_INF_F;
        sm[8] = (flags & FO_F_BE_RANGE) ? CUDART_NAN_F : btn_all;
        sm[9] = (flags & FO_F_BE_RANGE) ? CUDART_NAN_F : rcd_all;
      }
    }
  }
}


int launch_metric_detail(const MetricKArgs& k, int num_sms, cudaStream_t st) {
  const int ctas_needed = (k.N + kWarpsPerCta - 1) / kWarpsPerCta;
  const int grid = ctas_needed < num_sms * 8 ? ctas_needed : num_sms * 8;
  const int np = (k.T + 31) / 32;
  if (np == 1) fo_metric_kernel<1><<<grid, kWarpsPerCta * 32, 0, st>>>(k);
  else if (np == 2) fo_metric_kernel<2><<<grid, kWarpsPerCta * 32, 0, st>>>(k);
  else fo_metric_kernel<4><<<grid, kWarpsPerCta * 32, 0, st>>>(k);
  count_launch();
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}

}  // namespace fo
