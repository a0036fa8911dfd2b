// Point classification shared by fo_points.cu (explicit query points) and fo_spawn.cu (device-generated rasters).
#pragma once
#include "fo_common.cuh"

namespace fo {

constexpr int kPtThreads = 256;
constexpr int kPtTile = 1024;

struct PointClass {
  uint32_t flags;   // FO_PT_*
  int owner;        // first opaque thing the segment ego -> point meets
  uint64_t lan;     // lanelet membership bits
};

// Lanelet membership of a point given in the polygons' (the caller's) frame: even-odd rule, same formulation as the
// host / oracle restatement.  Returns "inside some polygon"; bit p of `lan` = inside polygon p (p < 64).
__device__ __forceinline__ bool lanelet_membership(const float* __restrict__ poly_xy, const int32_t* __restrict__ poly_off,
                                                   const int n_polygons, const float qx, const float qy, uint64_t& lan) {
  bool on_road = false;
  for (int p = 0; p < n_polygons; ++p) {
    const int v0 = poly_off[p], v1 = poly_off[p + 1];
    if (v1 - v0 < 3) continue;
    int cross = 0;
    float x0 = poly_xy[2 * (v1 - 1)], y0 = poly_xy[2 * (v1 - 1) + 1];
    for (int v = v0; v < v1; ++v) {
      const float x1 = poly_xy[2 * v], y1 = poly_xy[2 * v + 1];
      if ((y0 > qy) != (y1 > qy)) {
        const float xin = (x1 - x0) * (qy - y0) / (y1 - y0) + x0;
        cross += (qx < xin) ? 1 : 0;
      }
      x0 = x1; y0 = y1;
    }
    if (cross & 1) {
      on_road = true;
      if (p < 64) lan |= (1ull << p);
    }
  }
  return on_road;
}

// Block-cooperative: EVERY thread of a kPtThreads CTA calls it (the frame's edges are staged in `sg` / `st` tile by
// tile); `live` lanes classify the point (pwx, pwy) given in the caller's frame.
__device__ __forceinline__ PointClass classify_point_block(const FoPointQueryArgs& k, const bool live, const float pwx,
                                                           const float pwy, float4* sg, float2* st) {
  const float ex0 = k.ego[0], ey0 = k.ego[1], heading = k.ego[2];
  float px = 0.0f, py = 0.0f;
  if (live) { px = pwx - ex0; py = pwy - ey0; }

  // ---- shadow test: nearest crossing of the segment origin -> p ---------------------------------------
  float best_num = 2.0f, best_den = 1.0f;   // t = num / den, start above 1
  int owner = FO_HIT_NONE;
  bool in_obst = false, focus_cross = false, in_focus = false, near_focus = false;
  const int n_rect_edges = k.n_obstacles * 4;
  const int n_cand = n_rect_edges + k.n_boundary;
  for (int base = 0; base < n_cand; base += kPtTile) {
    const int cnt = min(kPtTile, n_cand - base);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += kPtThreads) {
      const int q = base + j;
      float ax = 0, ay = 0, bx = 0, by = 0;
      int own = FO_HIT_NONE;
      if (q < n_rect_edges) {
        const int o = q >> 2, e = q & 3;
        const uint8_t fl = k.rect_flags[o];
        if ((fl & FO_RECT_EXISTS) && !(fl & FO_RECT_TRANSPARENT)) {
          const float cx = k.rect[o * 5 + 0] - ex0, cy = k.rect[o * 5 + 1] - ey0;
          const float hl = k.rect[o * 5 + 3], hw = k.rect[o * 5 + 4];
          float sn, cs;
          sincosf(k.rect[o * 5 + 2], &sn, &cs);
          const float sx0 = (e == 0 || e == 1) ? -1.0f : 1.0f, sy0 = (e == 0 || e == 3) ? -1.0f : 1.0f;
          const float sx1 = (e == 0 || e == 3) ? -1.0f : 1.0f, sy1 = (e == 0 || e == 1) ? 1.0f : -1.0f;
          ax = cx + sx0 * hl * cs - sy0 * hw * sn; ay = cy + sx0 * hl * sn + sy0 * hw * cs;
          bx = cx + sx1 * hl * cs - sy1 * hw * sn; by = cy + sx1 * hl * sn + sy1 * hw * cs;
          own = o;
        }
      } else {
        const float4 b = reinterpret_cast<const float4*>(k.boundary)[q - n_rect_edges];
        ax = b.x - ex0; ay = b.y - ey0; bx = b.z - ex0; by = b.w - ey0;
        own = FO_HIT_BOUNDARY;
      }
      const float exx = bx - ax, eyy = by - ay;   // degenerate (skipped) edges have e = 0 -> D = 0 -> never hit
      sg[j] = make_float4(ax, ay, exx, eyy);
      st[j] = make_float2(ax * eyy - ay * exx, __int_as_float(own));
    }
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const float4 g = sg[j];
        const float2 t = st[j];
        const float D = px * g.w - py * g.z;          // cross(p, e)
        const float un = g.x * py - g.y * px;         // cross(a, p)
        const float aD = fabsf(D), at = fabsf(t.x);
        // t = t.x / D in [0, 1], u = un / D in [0, 1]
        const bool hit = (D != 0.0f) & (t.x * D >= 0.0f) & (un * D >= 0.0f) & (fabsf(un) <= aD) & (at <= aD);
        if (hit && __float_as_int(t.y) == k.focus_obstacle) focus_cross = true;
        if (hit && at * best_den < best_num * aD) {
          best_num = at; best_den = aD;
          owner = __float_as_int(t.y);
        }
      }
    }
  }
  // ---- inside an opaque obstacle rectangle (closed) ------------------------------------------------------
  if (live) {
    for (int o = 0; o < k.n_obstacles; ++o) {
      const uint8_t fl = k.rect_flags[o];
      if ((fl & FO_RECT_EXISTS) && !(fl & FO_RECT_TRANSPARENT)) {
        const float dx = px - (k.rect[o * 5 + 0] - ex0), dy = py - (k.rect[o * 5 + 1] - ey0);
        float sn, cs;
        sincosf(k.rect[o * 5 + 2], &sn, &cs);
        const float lx = dx * cs + dy * sn, ly = -dx * sn + dy * cs;
        if (o == k.focus_obstacle) {
          const float ex = fmaxf(fabsf(lx) - k.rect[o * 5 + 3], 0.0f), ey = fmaxf(fabsf(ly) - k.rect[o * 5 + 4], 0.0f);
          near_focus = fmaf(ex, ex, ey * ey) <= k.focus_margin * k.focus_margin;
        }
        if (fabsf(lx) <= k.rect[o * 5 + 3] && fabsf(ly) <= k.rect[o * 5 + 4]) {
          in_obst = true;
          if (o == k.focus_obstacle) in_focus = true;
          if (owner == FO_HIT_NONE) owner = o;
        }
      }
    }
  }
  PointClass out{0u, FO_HIT_NONE, 0ull};
  if (!live) return out;

  // ---- lanelet membership (even-odd rule, same formulation as the host/oracle restatement) ----------------
  uint64_t lan = 0;
  const bool on_road = lanelet_membership(k.poly_xy, k.poly_off, k.n_polygons, px + ex0, py + ey0, lan);

  // ---- sensor sector, occluded sector ------------------------------------------------------------------------
  const float PI_F = 3.14159265358979323846f;
  const float dist2 = px * px + py * py;
  float rel = atan2f(py, px) - heading;                      // wrap to [-pi, pi)
  rel -= 2.0f * PI_F * floorf((rel + PI_F) / (2.0f * PI_F));
  bool in_sensor = dist2 <= k.sensor_radius * k.sensor_radius;
  if (k.sensor_angle_deg < 359.9f) in_sensor = in_sensor && fabsf(rel) <= 0.5f * k.sensor_angle_deg * (PI_F / 180.0f);
  const bool blocked = (best_num <= best_den);
  const bool visible = in_sensor && on_road && !blocked && !in_obst;
  const bool occluded = on_road && !visible && fabsf(rel) <= 0.5f * PI_F && dist2 <= k.occluded_radius * k.occluded_radius;
  uint32_t f = 0;
  if (in_sensor) f |= FO_PT_IN_SENSOR;
  if (on_road) f |= FO_PT_ON_ROAD;
  if (blocked) f |= FO_PT_SHADOWED;
  if (in_obst) f |= FO_PT_IN_OBSTACLE;
  if (visible) f |= FO_PT_VISIBLE;
  if (occluded) f |= FO_PT_OCCLUDED;
  if (k.focus_obstacle >= 0 && focus_cross && !in_focus) f |= FO_PT_FOCUS_SHADOW;
  if (k.focus_obstacle >= 0 && near_focus) f |= FO_PT_FOCUS_NEAR;
  out.flags = f;
  out.owner = owner;
  out.lan = lan;
  return out;
}

}  // namespace fo
