// Device-side helpers shared by the metric kernels (detail kernel: fo_metric_detail.cu,
// summary kernel: fo_metric_flat.cu).
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
  uint8_t* valid;
  float* summary;
  uint32_t* flags;
  float* pair;
  float* step;
};

// ---------------------------------------------------------------------------------------------
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

struct EgoState {
  float x, y, th, v, c, s;
};


int launch_metric_detail(const MetricKArgs& k, int num_sms, cudaStream_t st);
int launch_metric_flat(const MetricKArgs& k, int num_sms, cudaStream_t st);

}  // namespace fo
