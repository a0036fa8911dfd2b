// Spawn locator, behind-dynamic-obstacle finder on the device, sm_100a (C-ABI: fo_spawn_region, fo_spawn_rect).
//
// The reference clips polygons (spawn_locator.py:254-287, 695-726); here the same sets are rasters of cell centres that
// never leave the GPU: generated in float64, classified by the code fo_visibility_points uses (one lane = one cell, the
// frame's edges staged in shared memory), labelled with a lock-free union-find over the 4-neighbourhood, and reduced
// (largest part, area, centroid, probe, one-cell dilation; for the candidate boxes: area, centroid, outline cells) by
// one CTA in a fixed order, so results do not depend on the schedule.  The host reads one 32-byte record per raster.
#include "fo_points_dev.cuh"

namespace fo {

__device__ __forceinline__ void raster_point(const FoRasterSpec& r, double cx, double cy, int i, int j, double& X, double& Y) {
  // float64, evaluated exactly as the host restatement writes it: no fused multiply-adds
  const double gx = __dsub_rn(__dmul_rn((double)i + 0.5, r.cell), r.hx);
  const double gy = __dsub_rn(__dmul_rn((double)j + 0.5, r.cell), r.hy);
  X = __dsub_rn(__dadd_rn(cx, __dmul_rn(gx, r.cs)), __dmul_rn(gy, r.sn));
  Y = __dadd_rn(__dadd_rn(cy, __dmul_rn(gx, r.sn)), __dmul_rn(gy, r.cs));
}

__device__ __forceinline__ void raster_centre(const FoRasterSpec& r, const FoRasterResult* from, double& cx, double& cy) {
  cx = r.cx; cy = r.cy;
  if (from && from->count >= 3) { cx = from->sum_x / (double)from->count; cy = from->sum_y / (double)from->count; }
}

__device__ __forceinline__ bool region_predicate(const FoRegionPredicate& p, const PointClass& c, double X, double Y) {
  bool in = (c.lan & p.lanelet_mask) != 0ull;
  in = in && (c.flags & p.want_flags) != 0u && (c.flags & p.reject_flags) == 0u;
  return in && hypot(X - p.disc_x, Y - p.disc_y) <= p.disc_r;
}

// ---- raster classification: region (labels initialised) and candidate box (restricted to the dilated part) ---------
__global__ void __launch_bounds__(kPtThreads) fo_region_raster_kernel(const FoSpawnRegionArgs k) {
  __shared__ float4 sg[kPtTile];
  __shared__ float2 st[kPtTile];
  const int m = blockIdx.x * kPtThreads + threadIdx.x;
  const int n_cells = k.raster.nx * k.raster.ny;
  const bool live = m < n_cells;
  double X = 0.0, Y = 0.0;
  if (live) raster_point(k.raster, k.raster.cx, k.raster.cy, m / k.raster.ny, m % k.raster.ny, X, Y);
  const PointClass c = classify_point_block(k.frame, live, (float)(X - k.raster.org_x), (float)(Y - k.raster.org_y), sg, st);
  if (!live) return;
  k.label[m] = region_predicate(k.pred, c, X, Y) ? m : -1;
  k.size[m] = 0;
  if (m == 0) *k.best = 0ull;
}

__global__ void __launch_bounds__(kPtThreads) fo_rect_raster_kernel(const FoSpawnRectArgs k) {
  __shared__ float4 sg[kPtTile];
  __shared__ float2 st[kPtTile];
  const int m = blockIdx.x * kPtThreads + threadIdx.x;
  const int n_cells = k.raster.nx * k.raster.ny;
  const bool live = m < n_cells;
  double cx, cy, X = 0.0, Y = 0.0;
  raster_centre(k.raster, k.centre_from, cx, cy);
  if (live) raster_point(k.raster, cx, cy, m / k.raster.ny, m % k.raster.ny, X, Y);
  const PointClass c = classify_point_block(k.frame, live, (float)(X - k.raster.org_x), (float)(Y - k.raster.org_y), sg, st);
  if (!live) return;
  bool in = region_predicate(k.pred, c, X, Y);
  // restrict to the selected connected part of the allowed area (coarse raster, one cell of slack)
  const double fi = floor((X - k.region_ox) / k.region_cell), fj = floor((Y - k.region_oy) / k.region_cell);
  in = in && fi >= 0.0 && fi < (double)k.region_n && fj >= 0.0 && fj < (double)k.region_n &&
       k.region_mask[(int)fi * k.region_n + (int)fj] != 0;
  k.mask[m] = in ? 1 : 0;
}

// ---- connected components: union-find with atomicMin links (root = smallest cell index of the part) -----------------
__device__ __forceinline__ int uf_find(const int32_t* label, int a) {
  int p = label[a];
  while (p != a) { a = p; p = label[a]; }
  return a;
}
__device__ __forceinline__ void uf_union(int32_t* label, int a, int b) {
  for (;;) {
    a = uf_find(label, a);
    b = uf_find(label, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }      // a > b: hang the larger root under the smaller
    const int old = atomicMin(&label[a], b);
    if (old == a) return;
    a = old;
  }
}
__global__ void fo_region_merge_kernel(int32_t* label, int nx, int ny) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nx * ny || label[m] < 0) return;
  const int j = m % ny;
  if (j > 0 && label[m - 1] >= 0) uf_union(label, m, m - 1);
  if (m >= ny && label[m - ny] >= 0) uf_union(label, m, m - ny);
}
__global__ void fo_region_flatten_kernel(int32_t* label, int32_t* size, int n_cells) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_cells || label[m] < 0) return;
  const int root = uf_find(label, m);
  atomicAdd(&size[root], 1);
  label[m] = root;      // path compression; safe while others still walk: any value written is an ancestor in the same set
}
__global__ void fo_region_select_kernel(int32_t* label, const int32_t* size, unsigned long long* best, int n_cells) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_cells || label[m] != m) return;           // roots only
  // largest part; the first one in raster order on ties (scipy.ndimage.label numbers parts in that order)
  atomicMax(best, ((unsigned long long)(unsigned)size[m] << 32) | (unsigned long long)(0xffffffffu - (unsigned)m));
}

constexpr int kRedThreads = 1024;
// fixed-order block reduction of (count, sum x, sum y)
__device__ __forceinline__ void block_sum(int& cnt, double& sx, double& sy, int* s_c, double* s_x, double* s_y) {
  const int t = threadIdx.x;
  s_c[t] = cnt; s_x[t] = sx; s_y[t] = sy;
  __syncthreads();
  for (int o = kRedThreads / 2; o > 0; o >>= 1) {
    if (t < o) { s_c[t] += s_c[t + o]; s_x[t] += s_x[t + o]; s_y[t] += s_y[t + o]; }
    __syncthreads();
  }
  cnt = s_c[0]; sx = s_x[0]; sy = s_y[0];
}

__global__ void __launch_bounds__(kRedThreads) fo_region_final_kernel(const FoSpawnRegionArgs k) {
  __shared__ int s_c[kRedThreads];
  __shared__ double s_x[kRedThreads], s_y[kRedThreads];
  __shared__ int s_roots;
  const int nx = k.raster.nx, ny = k.raster.ny, n_cells = nx * ny;
  const unsigned long long key = *k.best;
  const int best = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : -2;
  if (threadIdx.x == 0) s_roots = 0;
  __syncthreads();
  int cnt = 0, roots = 0;
  double sx = 0.0, sy = 0.0;
  for (int m = threadIdx.x; m < n_cells; m += kRedThreads) {
    const int lab = k.label[m];
    roots += lab == m;
    const bool in = lab >= 0 && uf_find(k.label, m) == best;
    bool dil = in;
    if (!dil) {
      const int j = m % ny;
      auto part = [&](int q) { const int l = k.label[q]; return l >= 0 && uf_find(k.label, q) == best; };
      dil = (j > 0 && part(m - 1)) || (j + 1 < ny && part(m + 1)) || (m >= ny && part(m - ny)) || (m + ny < n_cells && part(m + ny));
    }
    k.mask_dilated[m] = dil ? 1 : 0;
    if (in) {
      double X, Y;
      raster_point(k.raster, k.raster.cx, k.raster.cy, m / ny, m % ny, X, Y);
      ++cnt; sx += X; sy += Y;
    }
  }
  if (roots) atomicAdd(&s_roots, roots);
  block_sum(cnt, sx, sy, s_c, s_x, s_y);
  if (threadIdx.x == 0) {
    FoRasterResult r;
    r.sum_x = sx; r.sum_y = sy; r.count = cnt; r.n_outline = 0; r.n_components = s_roots;
    // probe: the cell the point falls into (lower-left corner of the raster = C - (hx, hy); axis-aligned rasters)
    const double fi = floor((k.probe_x - (k.raster.cx - k.raster.hx)) / k.raster.cell);
    const double fj = floor((k.probe_y - (k.raster.cy - k.raster.hy)) / k.raster.cell);
    r.contains = 0;
    if (fi >= 0.0 && fi < (double)nx && fj >= 0.0 && fj < (double)ny) {
      const int q = (int)fi * ny + (int)fj;
      r.contains = (k.label[q] >= 0 && uf_find(k.label, q) == best) ? 1 : 0;
    }
    *k.result = r;
  }
}

__global__ void __launch_bounds__(kRedThreads) fo_rect_final_kernel(const FoSpawnRectArgs k) {
  __shared__ int s_c[kRedThreads];
  __shared__ double s_x[kRedThreads], s_y[kRedThreads];
  __shared__ int s_out;
  const int nx = k.raster.nx, ny = k.raster.ny, n_cells = nx * ny;
  double cx, cy;
  raster_centre(k.raster, k.centre_from, cx, cy);
  if (threadIdx.x == 0) s_out = 0;
  __syncthreads();
  int cnt = 0;
  double sx = 0.0, sy = 0.0;
  for (int m = threadIdx.x; m < n_cells; m += kRedThreads) {
    if (!k.mask[m]) continue;
    const int i = m / ny, j = m % ny;
    double X, Y;
    raster_point(k.raster, cx, cy, i, j, X, Y);
    ++cnt; sx += X; sy += Y;
    // outline: a 4-neighbour is missing (cells beyond the box count as missing)
    const bool core = j > 0 && k.mask[m - 1] && j + 1 < ny && k.mask[m + 1] && i > 0 && k.mask[m - ny] && i + 1 < nx && k.mask[m + ny];
    if (!core) {
      const int p = atomicAdd(&s_out, 1);
      if (p < k.outline_cap) { k.outline[2 * p] = X; k.outline[2 * p + 1] = Y; }
    }
  }
  block_sum(cnt, sx, sy, s_c, s_x, s_y);
  if (threadIdx.x == 0) {
    FoRasterResult r;
    r.sum_x = sx; r.sum_y = sy; r.count = cnt; r.contains = 0; r.n_outline = s_out; r.n_components = 0;
    *k.result = r;
  }
}

// ---- obstacles seen on the road (sensor_model.py:59-76) -------------------------------------------------------------
__global__ void fo_hits_on_road_kernel(const FoHitsOnRoadArgs k) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= k.n_rays) return;
  const int o = k.hit[r];
  if (o < 0 || o >= k.n_obstacles) return;
  const double ang = __dadd_rn(k.angle0, __dmul_rn(k.dangle, (double)r));
  const double t = (double)(k.range[r] - 1e-3f);                      // float32 subtraction, as the host restatement
  const double X = __dadd_rn(k.ego_x, __dmul_rn(t, cos(ang))), Y = __dadd_rn(k.ego_y, __dmul_rn(t, sin(ang)));
  const float ex0 = k.ego[0], ey0 = k.ego[1];
  const float px = (float)(X - k.org_x) - ex0, py = (float)(Y - k.org_y) - ey0;
  uint64_t lan = 0;
  if (lanelet_membership(k.poly_xy, k.poly_off, k.n_polygons, px + ex0, py + ey0, lan)) k.on_road[o] = 1;
}

static bool frame_ok(const FoPointQueryArgs& f) {
  return f.n_obstacles >= 0 && f.n_boundary >= 0 && f.n_polygons >= 0 && f.ego && (f.n_obstacles == 0 || (f.rect && f.rect_flags)) &&
         (f.n_boundary == 0 || f.boundary) && (f.n_polygons == 0 || (f.poly_xy && f.poly_off)) && f.sensor_radius > 0.0f;
}
static bool raster_ok(const FoRasterSpec& r) {
  return r.nx > 0 && r.ny > 0 && (long long)r.nx * r.ny <= (1LL << 30) && r.cell > 0.0;
}

}  // namespace fo

extern "C" int fo_spawn_region(const FoSpawnRegionArgs* a, void* stream) {
  if (!a) { fo::set_error("fo_spawn_region: NULL args"); return FO_ERR_INVALID_ARG; }
  if (!fo::frame_ok(a->frame) || !fo::raster_ok(a->raster)) { fo::set_error("fo_spawn_region: bad frame or raster"); return FO_ERR_INVALID_ARG; }
  if (!a->label || !a->size || !a->best || !a->mask_dilated || !a->result) { fo::set_error("fo_spawn_region: NULL array"); return FO_ERR_INVALID_ARG; }
  if (a->raster.cs != 1.0 || a->raster.sn != 0.0) { fo::set_error("fo_spawn_region: the region raster is axis-aligned"); return FO_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  const int n = a->raster.nx * a->raster.ny;
  const int g = (n + fo::kPtThreads - 1) / fo::kPtThreads;
  fo::fo_region_raster_kernel<<<g, fo::kPtThreads, 0, st>>>(*a);
  fo::fo_region_merge_kernel<<<g, fo::kPtThreads, 0, st>>>(a->label, a->raster.nx, a->raster.ny);
  fo::fo_region_flatten_kernel<<<g, fo::kPtThreads, 0, st>>>(a->label, a->size, n);
  fo::fo_region_select_kernel<<<g, fo::kPtThreads, 0, st>>>(a->label, a->size, a->best, n);
  fo::fo_region_final_kernel<<<1, fo::kRedThreads, 0, st>>>(*a);
  for (int i = 0; i < 5; ++i) fo::count_launch();
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}

extern "C" int fo_spawn_rect(const FoSpawnRectArgs* a, void* stream) {
  if (!a) { fo::set_error("fo_spawn_rect: NULL args"); return FO_ERR_INVALID_ARG; }
  if (!fo::frame_ok(a->frame) || !fo::raster_ok(a->raster)) { fo::set_error("fo_spawn_rect: bad frame or raster"); return FO_ERR_INVALID_ARG; }
  if (!a->region_mask || a->region_n <= 0 || !(a->region_cell > 0.0) || !a->mask || !a->result || a->outline_cap < 0 ||
      (a->outline_cap > 0 && !a->outline)) {
    fo::set_error("fo_spawn_rect: NULL array or bad region raster");
    return FO_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int n = a->raster.nx * a->raster.ny;
  fo::fo_rect_raster_kernel<<<(n + fo::kPtThreads - 1) / fo::kPtThreads, fo::kPtThreads, 0, st>>>(*a);
  fo::fo_rect_final_kernel<<<1, fo::kRedThreads, 0, st>>>(*a);
  fo::count_launch(); fo::count_launch();
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}

extern "C" int fo_visibility_hits_on_road(const FoHitsOnRoadArgs* a, void* stream) {
  if (!a) { fo::set_error("fo_visibility_hits_on_road: NULL args"); return FO_ERR_INVALID_ARG; }
  if (a->n_rays < 0 || a->n_obstacles < 0 || a->n_polygons < 0) { fo::set_error("fo_visibility_hits_on_road: negative size"); return FO_ERR_INVALID_ARG; }
  if (a->n_obstacles == 0) return FO_OK;
  if (!a->on_road || !a->ego || (a->n_rays > 0 && (!a->range || !a->hit)) || (a->n_polygons > 0 && (!a->poly_xy || !a->poly_off))) {
    fo::set_error("fo_visibility_hits_on_road: NULL array");
    return FO_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  FO_CUDA_TRY(cudaMemsetAsync(a->on_road, 0, (size_t)a->n_obstacles, st));
  if (a->n_rays == 0) return FO_OK;
  fo::fo_hits_on_road_kernel<<<(a->n_rays + 255) / 256, 256, 0, st>>>(*a);
  fo::count_launch();
  FO_CUDA_TRY(cudaGetLastError());
  return FO_OK;
}
