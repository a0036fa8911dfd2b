// Stage 1 -> spawn locator: point-wise visibility / occlusion / road membership, sm_100a.
//
// One lane = one query point.  The frame's opaque edges (4 per non-bicycle obstacle rectangle + the road
// border segments) are transformed into the ego frame and staged in shared memory tile by tile; every lane
// tests the segment ego -> point against every staged edge (shared-memory broadcast, division free) and
// keeps the nearest crossing.  Lanelet membership is an even-odd crossing count over the polygon rings
// (uniform global loads, served by the L1 broadcast path).  Replaces the shapely `within` / `intersects`
// predicates of spawn_locator.py against SensorModel's visible_area / occluded_area / road_polygon.
#include "fo_points_dev.cuh"

namespace fo {

__global__ void __launch_bounds__(kPtThreads) fo_points_kernel(const FoPointQueryArgs k) {
  __shared__ float4 sg[kPtTile];   // a.x, a.y, e.x, e.y (ego frame)
  __shared__ float2 st[kPtTile];   // cross(a, e), owner
  const int m = blockIdx.x * kPtThreads + threadIdx.x;
  const bool live = m < k.n_points;
  float pwx = 0.0f, pwy = 0.0f;
  if (live) { pwx = k.points[2 * m]; pwy = k.points[2 * m + 1]; }
  const PointClass c = classify_point_block(k, live, pwx, pwy, sg, st);
  if (!live) return;
  k.flags[m] = c.flags;
  if (k.blocker) k.blocker[m] = c.owner;
  if (k.lanelets) k.lanelets[m] = c.lan;
}

}  // namespace fo

extern "C" int fo_visibility_points(const FoPointQueryArgs* a, void* stream) {
  if (!a) { fo::set_error("fo_visibility_points: NULL args"); return FO_ERR_INVALID_ARG; }
  if (a->n_points < 0 || a->n_obstacles < 0 || a->n_boundary < 0 || a->n_polygons < 0) {
    fo::set_error("fo_visibility_points: negative size");
    return FO_ERR_INVALID_ARG;
  }
  if (a->n_points == 0) return FO_OK;
  if (!a->ego || !a->points || !a->flags || (a->n_obstacles > 0 && (!a->rect || !a->rect_flags)) ||
      (a->n_boundary > 0 && !a->boundary) || (a->n_polygons > 0 && (!a->poly_xy || !a->poly_off))) {
    fo::set_error("fo_visibility_points: NULL array");
    return FO_ERR_INVALID_ARG;
  }
  if (!(a->sensor_radius > 0.0f)) { fo::set_error("fo_visibility_points: sensor_radius must be positive"); return FO_ERR_INVALID_ARG; }
  const int grid = (a->n_points + fo::kPtThreads - 1) / fo::kPtThreads;
  fo::fo_points_kernel<<<grid, fo::kPtThreads, 0, (cudaStream_t)stream>>>(*a);
  fo::count_launch();
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}
