// Stage 1: sensor visibility by ray casting, sm_100a.
//
// One CTA = 8 warps = a fan of 256 consecutive rays of one frame; one lane = one ray.  The frame's
// occluder edges (4 per obstacle rectangle + shared road-border segments) are transformed into the
// ego frame, culled (outside the sensor disc / outside the fan's angular sector) and compacted into
// shared memory cooperatively; then every lane walks the staged edge list (all lanes read the same
// shared-memory word: broadcast, conflict free) with a division-free ray/segment test.  Replaces the
// shapely clipping of sensor_model.py:103-193 (one polygon difference per border vertex / obstacle).
#include <math_constants.h>

#include "fo_common.cuh"

namespace fo {

constexpr int kVisThreads = 256;
constexpr int kVisTile = 1024;   // staged edges per tile: 1024 * (16 + 8) B = 24 KB
constexpr int kVisTrCap = 512;   // listed transparent obstacles per frame (more: fall back to scanning all flags)

struct VisEdge {
  float4 g;   // a.x, a.y, e.x, e.y   (segment a -> a + e, ego frame)
};

// conservative cull: segment entirely outside the disc of radius R, or entirely outside the angular
// sector spanned (counter-clockwise) by unit vectors d0 -> d1 with mid direction dm (width <= 180 deg)
__device__ __forceinline__ bool edge_relevant(float ax, float ay, float bx, float by, float R2, bool use_sector,
                                              float2 d0, float2 d1, float2 dm) {
  // distance^2 from the origin to the segment
  float ex = bx - ax, ey = by - ay;
  float l2 = fmaf(ex, ex, ey * ey);
  float t = l2 > 0.0f ? fminf(fmaxf(-(ax * ex + ay * ey) / l2, 0.0f), 1.0f) : 0.0f;
  float px = fmaf(t, ex, ax), py = fmaf(t, ey, ay);
  if (fmaf(px, px, py * py) > R2) return false;
  if (use_sector) {
    float ca = d0.x * ay - d0.y * ax, cb = d0.x * by - d0.y * bx;   // cross(d0, p): < 0 -> clockwise of the fan
    if (ca < 0.0f && cb < 0.0f) return false;
    ca = ax * d1.y - ay * d1.x; cb = bx * d1.y - by * d1.x;         // cross(p, d1): < 0 -> beyond the fan
    if (ca < 0.0f && cb < 0.0f) return false;
    if (ax * dm.x + ay * dm.y < 0.0f && bx * dm.x + by * dm.y < 0.0f) return false;   // behind the ego
  }
  return true;
}

__device__ __forceinline__ void ray_angle_params(const FoVisibilityArgs& k, float heading, float& a0, float& da) {
  const float PI_F = 3.14159265358979323846f;
  if (k.sensor_angle_deg >= 359.9f) {            // sensor_model.py:119-120
    a0 = heading - PI_F;
    da = 2.0f * PI_F / (float)k.n_rays;
  } else {
    const float fov = k.sensor_angle_deg * (PI_F / 180.0f);
    a0 = heading - 0.5f * fov;
    da = k.n_rays > 1 ? fov / (float)(k.n_rays - 1) : 0.0f;
  }
}

__global__ void __launch_bounds__(kVisThreads) fo_visibility_kernel(const FoVisibilityArgs k) {
  __shared__ float4 sg[kVisTile];
  __shared__ float2 st[kVisTile];   // (cross(a, e), owner as int bits)
  __shared__ int s_count;
  __shared__ int s_ntr;                 // transparent (bicycle) obstacles of this frame
  __shared__ uint16_t s_tr[kVisTrCap];
  const int f = blockIdx.y;
  const int r = blockIdx.x * kVisThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const float ex0 = k.ego[f * 3 + 0], ey0 = k.ego[f * 3 + 1], heading = k.ego[f * 3 + 2];
  float a0, da;
  ray_angle_params(k, heading, a0, da);
  const float R = k.sensor_radius, R2 = R * R;

  // this lane's ray
  float c = 1.0f, s = 0.0f;
  sincosf(a0 + da * (float)r, &s, &c);
  float best = R;
  int owner = FO_HIT_NONE;

  // angular sector of this CTA's fan (for culling); fans wider than 180 deg are not culled by angle
  const int r_lo = blockIdx.x * kVisThreads, r_hi = min(r_lo + kVisThreads, k.n_rays) - 1;
  const float span = da * (float)(r_hi - r_lo);
  const bool use_sector = span < 3.0f;
  float2 d0, d1, dm;
  sincosf(a0 + da * (float)r_lo - 1e-4f, &d0.y, &d0.x);
  sincosf(a0 + da * (float)r_hi + 1e-4f, &d1.y, &d1.x);
  sincosf(a0 + da * 0.5f * (float)(r_lo + r_hi), &dm.y, &dm.x);

  const int n_rect_edges = k.n_obstacles * 4;
  const int n_cand = n_rect_edges + k.n_boundary;
  const float* rect = k.rect + (size_t)f * k.n_obstacles * 5;
  const uint8_t* flags = k.rect_flags + (size_t)f * k.n_obstacles;

  // transparent obstacles (type 'bicycle', sensor_model.py:177) cast no shadow but can be seen: list them once
  if (threadIdx.x == 0) s_ntr = 0;
  __syncthreads();
  if (k.visible) {
    for (int o = threadIdx.x; o < k.n_obstacles; o += kVisThreads) {
      const uint8_t fl = flags[o];
      if ((fl & FO_RECT_EXISTS) && (fl & FO_RECT_TRANSPARENT)) {
        const int p = atomicAdd(&s_ntr, 1);
        if (p < kVisTrCap) s_tr[p] = (uint16_t)o;
      }
    }
  }
  __syncthreads();

  for (int base = 0; base < n_cand; base += kVisTile) {
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // ---- stage: transform, cull, compact ---------------------------------------------------------
    for (int q0 = base; q0 < min(base + kVisTile, n_cand); q0 += kVisThreads) {
      const int q = q0 + threadIdx.x;
      bool keep = false;
      float ax = 0, ay = 0, bx = 0, by = 0;
      int own = FO_HIT_NONE;
      if (q < min(base + kVisTile, n_cand)) {
        if (q < n_rect_edges) {
          const int o = q >> 2, e = q & 3;
          const uint8_t fl = flags[o];
          if ((fl & FO_RECT_EXISTS) && !(fl & FO_RECT_TRANSPARENT)) {
            const float cx = rect[o * 5 + 0] - ex0, cy = rect[o * 5 + 1] - ey0, yaw = rect[o * 5 + 2];
            const float hl = rect[o * 5 + 3], hw = rect[o * 5 + 4];
            float sn, cs;
            sincosf(yaw, &sn, &cs);
            // corner ring (-l,-w), (-l,+w), (+l,+w), (+l,-w)  (commonroad Rectangle vertex order)
            const float sx0 = (e == 0 || e == 1) ? -1.0f : 1.0f, sy0 = (e == 0 || e == 3) ? -1.0f : 1.0f;
            const float sx1 = (e == 0 || e == 3) ? -1.0f : 1.0f, sy1 = (e == 0 || e == 1) ? 1.0f : -1.0f;
            ax = cx + sx0 * hl * cs - sy0 * hw * sn; ay = cy + sx0 * hl * sn + sy0 * hw * cs;
            bx = cx + sx1 * hl * cs - sy1 * hw * sn; by = cy + sx1 * hl * sn + sy1 * hw * cs;
            own = o;
            keep = edge_relevant(ax, ay, bx, by, R2, use_sector, d0, d1, dm);
          }
        } else {
          const float4 b = reinterpret_cast<const float4*>(k.boundary)[q - n_rect_edges];
          ax = b.x - ex0; ay = b.y - ey0; bx = b.z - ex0; by = b.w - ey0;
          own = FO_HIT_BOUNDARY;
          keep = edge_relevant(ax, ay, bx, by, R2, use_sector, d0, d1, dm);
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      int pos = 0;
      if (lane == 0 && m) pos = atomicAdd(&s_count, __popc(m));
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if (keep) {
        const int w = pos + __popc(m & ((1u << lane) - 1u));
        float exx = bx - ax, eyy = by - ay;
        float tn = ax * eyy - ay * exx;                      // cross(a, e)
        if (tn < 0.0f) {                                     // orient the edge so that cross(a, e) >= 0: a hit then
          ax = bx; ay = by; exx = -exx; eyy = -eyy; tn = -tn;  // needs cross(d, e) > 0 and the ray test loses its
        }                                                    // sign products and absolute values
        sg[w] = make_float4(ax, ay, exx, eyy);
        st[w] = make_float2(tn, __int_as_float(own));
      }
    }
    __syncthreads();
    // ---- cast: every lane against every staged edge (shared-memory broadcast) ------------------------
    const int cnt = s_count;
    if (r < k.n_rays) {
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const float4 g = sg[j];
        const float2 t = st[j];
        const float D = c * g.w - s * g.z;           // cross(d, e)
        const float un = g.x * s - g.y * c;          // cross(a, d)
        // t = tn / D >= 0 with tn >= 0, u = un / D in [0, 1], t < best   (division-free, strict improvement only)
        const bool okk = (D > 0.0f) & (un >= 0.0f) & (un <= D) & (t.x < best * D);
        if (okk) {
          best = t.x / D;
          owner = __float_as_int(t.y);
        }
      }
    }
    __syncthreads();
  }

  if (r < k.n_rays) {
    k.range[(size_t)f * k.n_rays + r] = best;
    k.hit[(size_t)f * k.n_rays + r] = owner;
    if (k.visible) {
      if (owner >= 0) k.visible[(size_t)f * k.n_obstacles + owner] = 1;
      // transparent obstacles (bicycles): visible when the ray crosses them before its first opaque hit
      const int ntr = s_ntr;
      const bool listed = ntr <= kVisTrCap && k.n_obstacles <= 65535;
      const int n_scan = listed ? ntr : k.n_obstacles;
      for (int q = 0; q < n_scan; ++q) {
        const int o = listed ? (int)s_tr[q] : q;
        const uint8_t fl = flags[o];
        if ((fl & FO_RECT_EXISTS) && (fl & FO_RECT_TRANSPARENT)) {
          const float cx = rect[o * 5 + 0] - ex0, cy = rect[o * 5 + 1] - ey0;
          float sn, cs;
          sincosf(rect[o * 5 + 2], &sn, &cs);
          const float hl = rect[o * 5 + 3], hw = rect[o * 5 + 4];
          // ray vs box in the box frame (slab test)
          const float ox = -(cx * cs + cy * sn), oy = -(-cx * sn + cy * cs);
          const float dx = c * cs + s * sn, dy = -c * sn + s * cs;
          float t0 = 0.0f, t1 = best;
          bool hitb = true;
          if (fabsf(dx) > 1e-12f) { float ta = (-hl - ox) / dx, tb = (hl - ox) / dx; t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb)); }
          else if (fabsf(ox) > hl) hitb = false;
          if (fabsf(dy) > 1e-12f) { float ta = (-hw - oy) / dy, tb = (hw - oy) / dy; t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb)); }
          else if (fabsf(oy) > hw) hitb = false;
          if (hitb && t0 <= t1) k.visible[(size_t)f * k.n_obstacles + o] = 1;
        }
      }
    }
  }
}

// ---- stage 2: constant-velocity rollout (tiny; double arithmetic so the float32 table is the correctly
// rounded image of the reference's float64 numbers) ------------------------------------------------------
__global__ void fo_rollout_cv_kernel(const FoRolloutCvArgs k) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = idx / k.t_stride, i = idx - a * k.t_stride;
  if (a >= k.n_agents) return;
  float x = 0, y = 0, yaw = 0, v = 0, var = 0;
  if (i < k.n_states) {
    const double phi = k.phi[a], sp = k.v[a];
    const double vx = rint(sp * cos(phi) * 1000.0) / 1000.0;   // round(v cos phi, 3), agent.py:492
    const double vy = rint(sp * sin(phi) * 1000.0) / 1000.0;   // agent.py:493
    const double t = (double)i * k.dt;                           // agent.py:499
    x = (float)(k.x0[a] + t * vx);
    y = (float)(k.y0[a] + t * vy);
    yaw = (float)phi;
    v = (float)sp;
    var = (float)(k.var0 * pow(k.var_factor, (double)i));        // agent.py:272
  }
  const size_t o = (size_t)a * k.t_stride + i;
  k.x[o] = x; k.y[o] = y; k.yaw[o] = yaw; k.vel[o] = v; k.var_x[o] = var; k.var_y[o] = var;
}

}  // namespace fo

extern "C" int fo_visibility_raycast(const FoVisibilityArgs* a, void* stream) {
  if (!a) { fo::set_error("fo_visibility_raycast: NULL args"); return FO_ERR_INVALID_ARG; }
  if (a->n_frames < 0 || a->n_rays < 0 || a->n_obstacles < 0 || a->n_boundary < 0) {
    fo::set_error("fo_visibility_raycast: negative size");
    return FO_ERR_INVALID_ARG;
  }
  if (a->n_frames == 0 || a->n_rays == 0) return FO_OK;
  if (!a->ego || !a->range || !a->hit || (a->n_obstacles > 0 && (!a->rect || !a->rect_flags)) ||
      (a->n_boundary > 0 && !a->boundary)) {
    fo::set_error("fo_visibility_raycast: NULL array");
    return FO_ERR_INVALID_ARG;
  }
  if (!(a->sensor_radius > 0.0f)) { fo::set_error("fo_visibility_raycast: sensor_radius must be positive"); return FO_ERR_INVALID_ARG; }
  if (a->n_frames > 65535) { fo::set_error("fo_visibility_raycast: at most 65535 frames per call"); return FO_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  if (a->visible && a->n_obstacles > 0)
    FO_CUDA_TRY(cudaMemsetAsync(a->visible, 0, (size_t)a->n_frames * a->n_obstacles, st));
  dim3 grid((a->n_rays + fo::kVisThreads - 1) / fo::kVisThreads, a->n_frames);
  fo::fo_visibility_kernel<<<grid, fo::kVisThreads, 0, st>>>(*a);
  fo::count_launch();
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}

extern "C" int fo_rollout_cv(const FoRolloutCvArgs* a, void* stream) {
  if (!a) { fo::set_error("fo_rollout_cv: NULL args"); return FO_ERR_INVALID_ARG; }
  if (a->n_agents < 0 || a->n_states < 0 || a->t_stride < a->n_states) {
    fo::set_error("fo_rollout_cv: bad sizes");
    return FO_ERR_INVALID_ARG;
  }
  if (a->n_agents == 0 || a->t_stride == 0) return FO_OK;
  if (!a->x0 || !a->y0 || !a->v || !a->phi || !a->x || !a->y || !a->yaw || !a->vel || !a->var_x || !a->var_y) {
    fo::set_error("fo_rollout_cv: NULL array");
    return FO_ERR_INVALID_ARG;
  }
  const int total = a->n_agents * a->t_stride;
  fo::fo_rollout_cv_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*a);
  fo::count_launch();
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}
