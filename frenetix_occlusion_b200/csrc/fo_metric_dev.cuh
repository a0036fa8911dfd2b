// Device-side helpers shared by the metric kernels (detail kernel: fo_metric_detail.cu,
// summary kernel: fo_metric_sweep.cu).
#pragma once
#include <math_constants.h>

#include "fo_common.cuh"

namespace fo {


constexpr int kWarpsPerCta = 8;
constexpr unsigned kFull = 0xffffffffu;

struct MetricKArgs {
  const float* ego;
  int N, T, A, Tp;
  AgentTableView tab;
  float hEx, hEy;      // ego half extents
  float wb, a_max;
  float L3, L6, W2;    // L/3, L/6, W/2 (CP boxes)
  FoHarmCoeffs hc;
  float dt;
  double dtd;
  uint32_t mmask, tmask;
  double thr_harm, thr_risk, thr_be, thr_cp, thr_ttc, thr_dce;
  // the two discrete threshold clauses in the units the kernels hold (host-computed, exact in float64):
  //   dce < thr_dce   <=>  round(d * 1000) < thr_dce_mm        (np.round(d, 3) = r / 1000.0, dce.py:79, metric.py:92-98)
  //   ttc < thr_ttc   <=>  first colliding step < thr_ttc_col  (np.round(step * dt, 3), ttc.py:43, metric.py:85-89)
  uint32_t thr_dce_mm, thr_ttc_col;
  bool exact_dce;      // summary kernel: re-round np.round(d, 3) ties in float64 even when no threshold needs it (FO_EXACT_DCE=1)
  uint8_t* valid;
  float* summary;
  uint32_t* flags;
  float* pair;
  float* step;
  unsigned long long* stats;   // optional [8] work counters (see fo_metric_stats), NULL in production launches
  unsigned int* claim;         // zeroed per launch: next-trajectory counter of the summary kernel (NULL = static striding)
  int n_peers;                 // summary kernel: results are also stored at these byte offsets (peer-mapped gather buffers)
  long long peer_delta[FO_MAX_PEERS];
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ AgentParams load_params(const AgentParams* p) {   // two 16-byte loads
  const int4* q = reinterpret_cast<const int4*>(p);
  int4 a = __ldg(q), b = __ldg(q + 1);
  AgentParams r;
  r.n_states = a.x; r.model = a.y; r.hl = __int_as_float(a.z); r.hw = __int_as_float(a.w);
  r.hlb = __int_as_float(b.x); r.ke = __int_as_float(b.y); r.ko = __int_as_float(b.z); r.pad = __int_as_float(b.w);
  return r;
}

// r / 1000 as the float the outputs carry (r = round(d * 1000) is exact in float32; correctly rounded division)
__device__ __forceinline__ float mm_to_m(uint32_t r) { return __fdiv_rn((float)r, 1000.0f); }
// np.round(step * dt, 3) as a float (ttc.py:43)
__device__ __forceinline__ float step_to_s(uint32_t step, double dt) { return (float)(rint((double)step * dt * 1000.0) * 0.001); }

__device__ __forceinline__ float umaxf(float v) {  // warp max of non-negative floats (bit order == value order)
  return __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(v)));
}

__device__ __forceinline__ float pt_box_d2(float px, float py, float hx, float hy) {
  float dx = fmaxf(fabsf(px) - hx, 0.0f);
  float dy = fmaxf(fabsf(py) - hy, 0.0f);
  return fmaf(dx, dx, dy * dy);
}

// Oriented-box test in the frame of box E (half extents hEx,hEy): box O centred at (rx, ry) in that
// frame, rotated by the angle whose cos/sin are (c, s), half extents (hl, hw).
// Returns squared distance (0 when the closed boxes intersect).  dce.py:75-79 / be.py:181.
__device__ __forceinline__ float obb_d2(float rx, float ry, float c, float s, float hEx, float hEy, float hl, float hw) {
  float ac = fabsf(c), as = fabsf(s);
  // centre of E in O's frame is -(rox, roy)
  float rox = fmaf(rx, c, ry * s);
  float roy = fmaf(ry, c, -rx * s);
  bool sep = (fabsf(rx) > hEx + fmaf(hl, ac, hw * as)) | (fabsf(ry) > hEy + fmaf(hl, as, hw * ac)) |
             (fabsf(rox) > hl + fmaf(hEx, ac, hEy * as)) | (fabsf(roy) > hw + fmaf(hEx, as, hEy * ac));
  if (!sep) return 0.0f;
  // corners of O in E's frame: r +- hl*(c,s) +- hw*(-s,c)
  float ux = hl * c, uy = hl * s, wx = -hw * s, wy = hw * c;
  float d2 = pt_box_d2(rx + ux + wx, ry + uy + wy, hEx, hEy);
  d2 = fminf(d2, pt_box_d2(rx + ux - wx, ry + uy - wy, hEx, hEy));
  d2 = fminf(d2, pt_box_d2(rx - ux + wx, ry - uy + wy, hEx, hEy));
  d2 = fminf(d2, pt_box_d2(rx - ux - wx, ry - uy - wy, hEx, hEy));
  // corners of E in O's frame: -ro +- hEx*(c,-s) +- hEy*(s,c)
  ux = hEx * c; uy = -hEx * s; wx = hEy * s; wy = hEy * c;
  d2 = fminf(d2, pt_box_d2(-rox + ux + wx, -roy + uy + wy, hl, hw));
  d2 = fminf(d2, pt_box_d2(-rox + ux - wx, -roy + uy - wy, hl, hw));
  d2 = fminf(d2, pt_box_d2(-rox - ux + wx, -roy - uy + wy, hl, hw));
  d2 = fminf(d2, pt_box_d2(-rox - ux - wx, -roy - uy - wy, hl, hw));
  return d2;
}

// ---- rounding-tie refinement of np.round(d, 3) (dce.py:79) ---------------------------------------------------------
// The float32 distance carries an error of a few 1e-5 m; when d * 1000 lies within kTieBand of a rounding boundary
// (x.5) the float64 reference may round the other way.  Such candidates are re-evaluated here in double from the
// same float32 inputs (the formulation of oracle/geometry.py: SAT, then the 8 corner-to-box distances), so dce,
// time_dce and the first-collision step agree with the float64 reference except when the exact distance itself sits
// within 1e-12 of a boundary.  Rare path (a few per cent of the exact distance evaluations), kept out of line.
constexpr float kTieBand = 0.08f;

__device__ __forceinline__ bool near_rounding_boundary(float d) {
  const float t = d * 1000.0f;
  return fabsf(t - floorf(t) - 0.5f) < kTieBand;
}

__device__ __forceinline__ double pt_box_d2_f64(double px, double py, double hx, double hy) {
  const double dx = fmax(fabs(px) - hx, 0.0), dy = fmax(fabs(py) - hy, 0.0);
  return dx * dx + dy * dy;
}

// sin / cos in float64 of a float32 angle.  CUDA's sincos(double) drags an 8 kB Payne-Hanek slow path into every
// kernel that calls it (the metric kernels live next to a 32 kB instruction cache); headings are float32 values of
// moderate size, so a two-term Cody-Waite reduction by pi/2 plus the fdlibm kernel polynomials (|r| <= pi/4, error
// below 1e-16) is exact enough for a tie-break that only matters within 1e-12 of a rounding boundary.
static __device__ __noinline__ void sincos_f64_of_f32(float af, double* sn, double* cs) {
  const double a = (double)af;
  const double q = rint(a * 0.63661977236758134308);                 // 2 / pi
  double r = fma(-q, 1.57079632679489655800e+00, a);                 // pi/2 high part
  r = fma(-q, 6.12323399573676603587e-17, r);                        // pi/2 low part
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double s = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
  const int n = (int)(long long)q & 3;
  *sn = (n == 0) ? s : (n == 1) ? c : (n == 2) ? -s : -c;
  *cs = (n == 0) ? c : (n == 1) ? -s : (n == 2) ? -c : s;
}

// ego reference point (ex, ey, eth), rectangle centre wb ahead; agent rectangle (ax, ay, ayaw); returns round(d * 1000)
static __device__ __noinline__ uint32_t obb_round_mm_f64(float ex, float ey, float eth, float wb, float hEx, float hEy,
                                                         float ax, float ay, float ayaw, float hl, float hw) {
  double se, ce, sa, ca;
  sincos_f64_of_f32(eth, &se, &ce);
  sincos_f64_of_f32(ayaw, &sa, &ca);
  const double cx = (double)ex + (double)wb * ce, cy = (double)ey + (double)wb * se;
  const double rx = (double)ax - cx, ry = (double)ay - cy;
  const double HEx = hEx, HEy = hEy, HL = hl, HW = hw;
  const double rax = rx * ce + ry * se, ray = -rx * se + ry * ce;     // agent centre in the ego frame
  const double rbx = rx * ca + ry * sa, rby = -rx * sa + ry * ca;     // the same offset in the agent frame
  const double c = fabs(ce * ca + se * sa), s = fabs(sa * ce - ca * se);
  const bool sep = (fabs(rax) > HEx + HL * c + HW * s) | (fabs(ray) > HEy + HL * s + HW * c) |
                   (fabs(rbx) > HL + HEx * c + HEy * s) | (fabs(rby) > HW + HEx * s + HEy * c);
  if (!sep) return 0u;
  double best = 1.0e300;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double sx = (q & 1) ? 1.0 : -1.0, sy = (q & 2) ? 1.0 : -1.0;
    double wx = rx + sx * HL * ca - sy * HW * sa, wy = ry + sx * HL * sa + sy * HW * ca;     // agent corner vs ego box
    best = fmin(best, pt_box_d2_f64(wx * ce + wy * se, -wx * se + wy * ce, HEx, HEy));
    wx = -rx + sx * HEx * ce - sy * HEy * se; wy = -ry + sx * HEx * se + sy * HEy * ce;      // ego corner vs agent box
    best = fmin(best, pt_box_d2_f64(wx * ca + wy * sa, -wx * sa + wy * ca, HL, HW));
  }
  return (uint32_t)llrint(fmin(sqrt(best), 8000.0) * 1000.0);
}

__device__ __forceinline__ bool obb_hit(float rx, float ry, float c, float s, float hEx, float hEy, float hl, float hw) {
  float ac = fabsf(c), as = fabsf(s);
  float rox = fmaf(rx, c, ry * s);
  float roy = fmaf(ry, c, -rx * s);
  bool sep = (fabsf(rx) > hEx + fmaf(hl, ac, hw * as)) | (fabsf(ry) > hEy + fmaf(hl, as, hw * ac)) |
             (fabsf(rox) > hl + fmaf(hEx, ac, hEy * as)) | (fabsf(roy) > hw + fmaf(hEx, as, hEy * ac));
  return !sep;
}

// 0.5*(erf(b) - erf(a)) for a < b, evaluated on the tails (no cancellation far from the mean).
__device__ __forceinline__ float half_derf(float a, float b) {
  float ea = erfcf(fabsf(a)), eb = erfcf(fabsf(b));
  bool same = (a > 0.0f) == (b > 0.0f);
  float r = same ? fabsf(ea - eb) : (2.0f - ea - eb);
  return 0.5f * r;
}

// Single MUFU forms.  __fdividef / __expf wrap the same MUFU.RCP / MUFU.EX2 in range guards (denormal divisor, result
// below 2^-126: 4 and 3 more instructions per call) that cannot fire for the arguments used here; inside the guards'
// dead range the results are bit-identical.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// erfc(z), z >= 0, fractional error < 1.2e-7 in exact arithmetic (Chebyshev fit of Numerical Recipes' erfcc:
// t = 1/(1 + z/2), erfc = t exp(-z^2 + P9(t))); about 4e-6 relative in float32 because of the exponent's
// magnitude -- two orders below the 1e-4 parity tolerance.  17 instructions instead of erfcf's ~45.
// RAW_EXP: exp as one FMUL + MUFU.EX2 (flushes results below 2^-126 to zero).  The summary kernel's cut-off keeps one end
// of every evaluated interval within 4.7, so a flushed far end (< 1e-38 against >= 3e-11) cannot change a bit of the
// factor; the detail kernel (no cut-off, per-step values compared for their first index) keeps the guarded __expf.
template <bool RAW_EXP>
__device__ __forceinline__ float erfc_pos_fast(float z) {
  const float t = rcp_approx(fmaf(0.5f, z, 1.0f));      // 1 + z/2 >= 1: same bits as __fdividef(1, .)
  float p = 0.17087277f;
  p = fmaf(p, t, -0.82215223f);
  p = fmaf(p, t, 1.48851587f);
  p = fmaf(p, t, -1.13520398f);
  p = fmaf(p, t, 0.27886807f);
  p = fmaf(p, t, -0.18628806f);
  p = fmaf(p, t, 0.09678418f);
  p = fmaf(p, t, 0.37409196f);
  p = fmaf(p, t, 1.00002368f);
  p = fmaf(p, t, -1.26551223f);
  const float a = fmaf(-z, z, p);
  return t * (RAW_EXP ? ex2_approx(a * 1.4426950216293334961f) : __expf(a));
}

// Collision probability of one gated step (collision_probability.py:94-122): Gaussian mass of the 3 obstacle
// points over the 3 axis-aligned ego boxes, divided by 3.  (mx, my) = obstacle position i-1 minus ego position i,
// (hx, hy) = half buffered length along yaw_i, (bx, by) = (L/3)(cos, sin theta_i), s2 = 1/(sqrt2 sigma_{x,y}).
// CUT (summary kernel): a factor whose standardised interval lies beyond 4.7 (erfc < 3e-11) contributes less than the
// 1e-10 that the comparison floor of 2e-7 can resolve and is skipped before any erfc is evaluated.  The detail kernel
// evaluates all nine terms: its per-step values feed first-index decisions of the reference's result dict
// (max_obst_risk_index = first maximum of harm x cp, hr.py:87-98), where a 1e-12 must not become an exact zero.
template <bool CUT = true>
__device__ __forceinline__ float cp_gauss_boxes(float mx, float my, float hx, float hy, float bx, float by, float2 s2,
                                                float L6, float W2) {
  constexpr float kFar = 4.7f;
  float prob = 0.0f;
  float fm = 0.0f;                 // obstacle point: centre, front, back
#pragma unroll 1
  for (int m = 0; m < 3; ++m) {
    const float uy = fmaf(fm, hy, my), ux = fmaf(fm, hx, mx);
    float fb = 0.0f;               // ego box: middle, front, rear
#pragma unroll 1
    for (int bb = 0; bb < 3; ++bb) {
      const float cyb = fb * by, cxb = fb * bx;
      fb = (bb == 0) ? 1.0f : -1.0f;
      const float ya = (cyb - W2 - uy) * s2.y, yb = (cyb + W2 - uy) * s2.y;     // ya < yb
      if (CUT && (ya > kFar || yb < -kFar)) continue;
      const float xa = (cxb - L6 - ux) * s2.x, xb = (cxb + L6 - ux) * s2.x;
      if (CUT && (xa > kFar || xb < -kFar)) continue;
      const float eya = erfc_pos_fast<CUT>(fabsf(ya)), eyb = erfc_pos_fast<CUT>(fabsf(yb));
      const float py = ((ya > 0.0f) == (yb > 0.0f)) ? fabsf(eya - eyb) : (2.0f - eya - eyb);
      const float exa = erfc_pos_fast<CUT>(fabsf(xa)), exb = erfc_pos_fast<CUT>(fabsf(xb));
      const float px = ((xa > 0.0f) == (xb > 0.0f)) ? fabsf(exa - exb) : (2.0f - exa - exb);
      prob = fmaf(0.25f * px, py, prob);
    }
    fm = (m == 0) ? 1.0f : -1.0f;
  }
  return prob * (1.0f / 3.0f);
}

__device__ __forceinline__ float logistic_neg(float z) {  // 1 / (1 + exp(z))
  return __fdividef(1.0f, 1.0f + __expf(z));
}

// LR4S angle coefficient, logistic_regression.py:35-42 (angles are NOT wrapped)
__device__ __forceinline__ float lr4s_coef(float ang, float side, float rear) {
  const float ta = 0.78539816339744830962f, tb = 2.35619449019234492885f;
  float c = rear;
  if ((ang >= ta && ang < tb) || (ang <= -ta && ang > -tb)) c = side;
  if (ang > -ta && ang < ta) c = 0.0f;
  return c;
}

// ---- BE (be.py:66-193) shared by the summary kernels: warp-cooperative bisection on a re-timed ego path -----
// The helpers are out of line (one copy per kernel).  Their view of the staged ego trajectory is a set of BYTE OFFSETS
// into the kernel's dynamic shared memory, and the few kernel arguments they need travel by value: pointers and a
// `const MetricKArgs&` would reach a noinline function as generic addresses (LD.E plus a uniform-register pair per access
// instead of LDS / a register).
#ifndef FO_BE_BUCKETS
#define FO_BE_BUCKETS 64
#endif
constexpr int kBeBuckets = FO_BE_BUCKETS;   // arc-length -> state-index lookup used by the interpolation
struct BeView {
  uint32_t egoA;   // float4 [T] (x, y, cos theta, sin theta)
  uint32_t egoB;   // float2 [T] (theta, v)
  uint32_t dist;   // float [T] cumulative chord length (be.py:99)
  uint32_t inv;    // uint8 [kBeBuckets + 1] last state index with dist <= b * dmax / kBeBuckets
  uint32_t seg;    // float4 [T] (SEG variants only): per segment j -> j + 1 the slopes ((x1-x0) inv, (y1-y0) inv, (th1-th0) inv, 0),
                   // inv = 1 / (dist[j+1] - dist[j]) -- the products be_bisect would form per lane and probe, formed once
};
struct BeConst {
  const float4* s0;   // agent-major states of the table
  int T, Tp;
  float dt, wb, hEx, hEy;
};
__device__ __forceinline__ BeConst be_const(const MetricKArgs& k) {
  return BeConst{k.tab.s0, k.T, k.Tp, k.dt, k.wb, k.hEx, k.hEy};
}

template <bool SEG>
static __device__ __noinline__ void be_prepare(const BeView v, int T, int lane) {
  extern __shared__ __align__(16) unsigned char fo_dyn_smem[];
  const float4* const egoA = reinterpret_cast<const float4*>(fo_dyn_smem + v.egoA);
  const float2* const egoB = reinterpret_cast<const float2*>(fo_dyn_smem + v.egoB);
  float4* const seg = reinterpret_cast<float4*>(fo_dyn_smem + v.seg);
  float* const dist = reinterpret_cast<float*>(fo_dyn_smem + v.dist);
  uint8_t* const inv = fo_dyn_smem + v.inv;
  float carry = 0.0f;
#pragma unroll 1
  for (int i0 = 0; i0 < T; i0 += 32) {
    const int i = i0 + lane;
    float seg = 0.0f;
    if (i >= 1 && i < T) {
      float4 p = egoA[i], q = egoA[i - 1];
      seg = sqrtf((p.x - q.x) * (p.x - q.x) + (p.y - q.y) * (p.y - q.y));
    }
    float sc = seg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(kFull, sc, o);
      if (lane >= o) sc += t;
    }
    if (i < T) dist[i] = carry + sc;
    carry += __shfl_sync(kFull, sc, 31);
  }
  __syncwarp();
  if (SEG) {
#pragma unroll 1
    for (int j = lane; j < T - 1; j += 32) {
      const float4 A0 = egoA[j], A1 = egoA[j + 1];
      const float inv = 1.0f / (dist[j + 1] - dist[j]);     // a zero-length segment is never selected by the lookup
      seg[j] = make_float4((A1.x - A0.x) * inv, (A1.y - A0.y) * inv, (egoB[j + 1].x - egoB[j].x) * inv, 0.0f);
    }
  }
  const float bw = dist[T - 1] / (float)kBeBuckets;
#pragma unroll 1
  for (int b = lane; b <= kBeBuckets; b += 32) {
    const float q = (b == kBeBuckets) ? CUDART_INF_F : (float)b * bw;
    int lo_j = 0, hi_j = T - 1;
    while (lo_j < hi_j) {
      const int mid = (lo_j + hi_j + 1) >> 1;
      if (dist[mid] <= q) lo_j = mid; else hi_j = mid - 1;
    }
    inv[b] = (uint8_t)lo_j;
  }
  __syncwarp();
}

// lanes = time steps.  v_new[0] = v0, v_new[k+1] = max(v1 - d k dt, 0) (be.py:109); dist_new[i] = dt sum_{k<i} v_new[k]
// (be.py:113) in closed form: the clipped arithmetic series has floor(v1 / (d dt)) + 1 positive terms.  dist_new is
// non-decreasing in i, so scipy interp1d's bounds_error (be.py:117-124) fires iff its LAST element overruns.
__device__ __forceinline__ float be_arclen(float dt, float v0, float v1, float step, float mpos, int i) {
  const float m = fminf((float)(i - 1), mpos);
  return (i == 0) ? 0.0f : dt * (v0 + fmaf(m, v1, -0.5f * step * m * (m - 1.0f)));
}

// Returns (required constant deceleration, number of probes as int bits); the deceleration is NaN when the re-timed
// path overruns the planned one (the reference raises, be.py:117-124).
// floor_rcd (summary kernel; negative = off): the caller keeps only the MAXIMUM over a trajectory's pairs.  The bracket
// [lo, hi] only shrinks and the result is its last midpoint, so once hi <= floor_rcd this pair cannot raise the maximum
// and the bisection stops -- provided no later probe could overrun the planned path: dist_new[T-1] falls with the
// deceleration, the smallest deceleration the rest of the bisection can probe is the end of its all-miss branch, and
// that one is checked (with a margin for the float32 closed form; inside the margin the bisection simply runs on).
template <bool SEG>
static __device__ __noinline__ float2 be_bisect(const BeConst k, const BeView v, int a, int n_states, float hl, float hw,
                                                float lo0, int lane, float floor_rcd) {
  extern __shared__ __align__(16) unsigned char fo_dyn_smem[];
  const float4* const egoA = reinterpret_cast<const float4*>(fo_dyn_smem + v.egoA);
  const float2* const egoB = reinterpret_cast<const float2*>(fo_dyn_smem + v.egoB);
  const float* const dist = reinterpret_cast<const float*>(fo_dyn_smem + v.dist);
  const uint8_t* const inv_tab = fo_dyn_smem + v.inv;
  const float4* const seg = reinterpret_cast<const float4*>(fo_dyn_smem + v.seg);
  const int T = k.T;
  const int nA = min(T, n_states);
  const float v0 = egoB[0].y, v1 = egoB[T > 1 ? 1 : 0].y;
  const float dmax = dist[T - 1];
  const float inv_w = dmax > 0.0f ? (float)kBeBuckets / dmax : 0.0f;
  // the agent's states do not depend on the probe: the first two 32-step chunks stay in registers for the whole
  // bisection (covers T <= 64, i.e. every horizon the reference uses), later chunks are re-read per probe
  const float4* sa = k.s0 + (size_t)a * k.Tp;
  const float4 pre0 = (lane < nA) ? __ldg(sa + lane) : make_float4(0, 0, 1, 0);
  const float4 pre1 = (32 + lane < nA) ? __ldg(sa + 32 + lane) : make_float4(0, 0, 1, 0);
  float lo = lo0, hi = 5.0f, cur = 0.0f;
  int probes = 0;
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    cur = 0.5f * (lo + hi);
    ++probes;
    const float step = cur * k.dt;
    // step in [4e-4, 0.5]: v1 * rcp == __fdividef(v1, step) bit for bit
    const float mpos = (step > 0.0f) ? fmaxf(floorf(v1 * rcp_approx(step)) + 1.0f, 0.0f) : 1.0e9f;
    if (be_arclen(k.dt, v0, v1, step, mpos, T - 1) > dmax) return make_float2(CUDART_NAN_F, __int_as_float(probes));
    bool any_hit = false;
#pragma unroll 1
    for (int i0 = 0; i0 < nA && !any_hit; i0 += 32) {
      const int i = i0 + lane;
      bool hit = false;
      if (i < nA) {
        const float q = be_arclen(k.dt, v0, v1, step, mpos, i);
        // numpy.interp: j = last index with dist[j] <= q.  The bucket of q brackets it: j0 = inv[b] (stepped one bucket
        // down when float32 rounding of the bucket index put it too high), j1 = inv[b + 1]; a bisection of [j0, j1] --
        // most buckets hold at most one state -- and a final upward check for a q that rounded across the bucket's end
        int b = min(__float2int_rd(q * inv_w), kBeBuckets - 1);
        int j = inv_tab[b];
        if (dist[j] > q) { b = max(b - 1, 0); j = inv_tab[b]; }
        int jh = inv_tab[b + 1];
        while (j < jh) {
          const int mid = (j + jh + 1) >> 1;
          if (dist[mid] <= q) j = mid; else jh = mid - 1;
        }
        while (j < T - 1 && dist[j + 1] <= q) ++j;
        const float dj = dist[j];
        const float4 A0 = egoA[j];
        float xn = A0.x, yn = A0.y, tn = egoB[j].x;
        if (j != T - 1 && dj != q) {
          const float wq = q - dj;
          if (SEG) {
            const float4 sl = seg[j];
            xn = fmaf(sl.x, wq, A0.x);
            yn = fmaf(sl.y, wq, A0.y);
            tn = fmaf(sl.z, wq, tn);
          } else {
            const float4 A1 = egoA[j + 1];
            const float inv = 1.0f / (dist[j + 1] - dj);
            xn = fmaf((A1.x - A0.x) * inv, wq, A0.x);
            yn = fmaf((A1.y - A0.y) * inv, wq, A0.y);
            tn = fmaf((egoB[j + 1].x - tn) * inv, wq, tn);
          }
        }
        float sn, cn;
        __sincosf(tn, &sn, &cn);
        const float4 s0 = (i0 == 0) ? pre0 : ((i0 == 32) ? pre1 : __ldg(sa + i));
        const float dx = (s0.x - xn) - k.wb * cn;
        const float dy = (s0.y - yn) - k.wb * sn;
        const float rx = fmaf(dx, cn, dy * sn), ry = fmaf(dy, cn, -dx * sn);
        const float c = fmaf(cn, s0.z, sn * s0.w), s = fmaf(s0.w, cn, -s0.z * sn);
        hit = obb_hit(rx, ry, c, s, k.hEx, k.hEy, hl, hw);
      }
      any_hit = __any_sync(kFull, hit);
    }
    if (nA > 0 && !any_hit) hi = cur; else lo = cur;   // be.py:74-77 (an agent that never exists counts as a hit)
    if (hi - lo < 0.1f) break;                         // be.py:79
    if (hi <= floor_rcd) {
      float h = hi, c = cur;
      for (int j = it + 1; j < 10; ++j) {              // the all-miss branch of what is left, same arithmetic
        c = 0.5f * (lo + h);
        h = c;
        if (h - lo < 0.1f) break;
      }
      const float st = c * k.dt;
      const float mp = (st > 0.0f) ? fmaxf(floorf(v1 * rcp_approx(st)) + 1.0f, 0.0f) : 1.0e9f;
      if (be_arclen(k.dt, v0, v1, st, mp, T - 1) < dmax * 0.99999f) break;
    }
  }
  return make_float2(cur, __int_as_float(probes));
}

struct EgoState {
  float x, y, th, v, c, s;
};


// one zeroed-per-launch device counter (trajectory claims of the persistent kernels), see fo_metric_sweep.cu
unsigned int* claim_slot(cudaStream_t st);
int launch_metric_detail(const MetricKArgs& k, int num_sms, cudaStream_t st);
int launch_metric_sweep(const MetricKArgs& k, int num_sms, cudaStream_t st);

}  // namespace fo
