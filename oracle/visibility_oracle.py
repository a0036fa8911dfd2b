"""Oracle for stage 1 (sensor visibility) and stage 2 (constant-velocity rollout).  TEST INFRASTRUCTURE.

PARITY UNPINNED at the polygon level: the reference builds its visible area with shapely/GEOS
boolean operations (``sensor_model.py:103-193``), which cannot be installed here, and ships no test
vectors.  Two restatements are kept instead:

* ``reference_point_visible``  -- the reference's own construction evaluated point-wise, polygon-library
  free: a point is visible iff it lies in the sensor sector, in none of the per-border-edge shadow
  quads ``[v1, v2, v2+100(v2-ego), v1+100(v1-ego)]`` (``sensor_model.py:137-155``,
  ``helper_functions.py:79-96``) and in none of the per-obstacle shadows built from the corner pair
  with the largest subtended angle (``helper_functions.py:139-176``) united with the obstacle itself
  (``sensor_model.py:174-191``); bicycles cast no shadow (``sensor_model.py:177``).
* ``raycast`` -- float64 brute-force restatement of the ray-cast formulation the CUDA kernel uses
  (first hit of every ray of the fan), bit-for-bit the same angle convention and closed-interval
  hit rule.  The tests check kernel == ``raycast`` (ranges to 1e-5, hit ids exactly except grazing
  ties) and ``raycast`` == ``reference_point_visible`` up to the angular resolution of the fan.
"""
from __future__ import annotations

import numpy as np

HIT_NONE, HIT_BOUNDARY = -1, -2
RECT_EXISTS, RECT_TRANSPARENT = 1, 2


def ray_angles(heading: float, fov_deg: float, n_rays: int) -> np.ndarray:
    if fov_deg >= 359.9:                                   # sensor_model.py:119-120 (full disc)
        return heading - np.pi + 2.0 * np.pi * np.arange(n_rays) / n_rays
    fov = np.radians(fov_deg)
    return heading - 0.5 * fov + (fov / max(n_rays - 1, 1)) * np.arange(n_rays)


def rect_corners(rect: np.ndarray) -> np.ndarray:
    """[O,5] (cx, cy, yaw, hl, hw) -> [O,4,2] corner ring in commonroad ``Rectangle.vertices`` order,
    rotated and translated as ``hf.calc_corner_points`` does (helper_functions.py:99-112)."""
    rect = np.asarray(rect, dtype=np.float64).reshape(-1, 5)
    sx = np.array([-1.0, -1.0, 1.0, 1.0])
    sy = np.array([-1.0, 1.0, 1.0, -1.0])
    lx = sx[None] * rect[:, 3:4]
    ly = sy[None] * rect[:, 4:5]
    c, s = np.cos(rect[:, 2:3]), np.sin(rect[:, 2:3])
    return np.stack((rect[:, 0:1] + lx * c - ly * s, rect[:, 1:2] + lx * s + ly * c), -1)


def _edges(ego_xy, rect, flags, boundary):
    rect = np.asarray(rect, dtype=np.float64).reshape(-1, 5)
    flags = np.asarray(flags).reshape(-1)
    segs, owner = [], []
    cor = rect_corners(rect)
    for o in range(len(rect)):
        if (flags[o] & RECT_EXISTS) and not (flags[o] & RECT_TRANSPARENT):
            for e in range(4):
                segs.append(np.concatenate((cor[o, e], cor[o, (e + 1) % 4])))
                owner.append(o)
    if boundary is not None and len(boundary):
        for b in np.asarray(boundary, dtype=np.float64).reshape(-1, 4):
            segs.append(b)
            owner.append(HIT_BOUNDARY)
    segs = np.asarray(segs, dtype=np.float64).reshape(-1, 4)
    segs = segs - np.tile(np.asarray(ego_xy, dtype=np.float64), 2)
    return segs, np.asarray(owner, dtype=np.int64)


def raycast(ego, rect, flags, boundary, sensor_radius, fov_deg, n_rays):
    """First-hit range / hit id per ray and per-obstacle visibility for ONE frame (float64)."""
    ego = np.asarray(ego, dtype=np.float64)
    ang = ray_angles(ego[2], fov_deg, n_rays)
    c, s = np.cos(ang)[:, None], np.sin(ang)[:, None]
    segs, owner = _edges(ego[:2], rect, flags, boundary)
    rng = np.full(n_rays, float(sensor_radius))
    hit = np.full(n_rays, HIT_NONE, dtype=np.int64)
    if len(segs):
        ax, ay = segs[None, :, 0], segs[None, :, 1]
        ex, ey = segs[None, :, 2] - ax, segs[None, :, 3] - ay
        D = c * ey - s * ex
        tn = ax * ey - ay * ex
        un = ax * s - ay * c
        with np.errstate(divide="ignore", invalid="ignore"):
            t = tn / D
            u = un / D
        ok = (D != 0) & (t >= 0) & (u >= 0) & (u <= 1) & (t < sensor_radius)
        t = np.where(ok, t, np.inf)
        j = t.argmin(1)
        tb = t[np.arange(n_rays), j]
        better = tb < rng
        rng = np.where(better, tb, rng)
        hit = np.where(better, owner[j], hit)
    rect = np.asarray(rect, dtype=np.float64).reshape(-1, 5)
    flags = np.asarray(flags).reshape(-1)
    visible = np.zeros(len(rect), dtype=np.uint8)
    for o in np.unique(hit[hit >= 0]):
        visible[o] = 1
    for o in range(len(rect)):                             # transparent obstacles: slab test up to the first hit
        if (flags[o] & RECT_EXISTS) and (flags[o] & RECT_TRANSPARENT):
            cx, cy, yaw, hl, hw = rect[o]
            cx, cy = cx - ego[0], cy - ego[1]
            cs, sn = np.cos(yaw), np.sin(yaw)
            ox, oy = -(cx * cs + cy * sn), -(-cx * sn + cy * cs)
            dx, dy = c[:, 0] * cs + s[:, 0] * sn, -c[:, 0] * sn + s[:, 0] * cs
            with np.errstate(divide="ignore", invalid="ignore"):
                tx0, tx1 = (-hl - ox) / dx, (hl - ox) / dx
                ty0, ty1 = (-hw - oy) / dy, (hw - oy) / dy
            t0 = np.maximum(np.minimum(tx0, tx1), np.minimum(ty0, ty1))
            t1 = np.minimum(np.maximum(tx0, tx1), np.maximum(ty0, ty1))
            if np.any((np.maximum(t0, 0.0) <= np.minimum(t1, rng))):
                visible[o] = 1
    return rng, hit, visible


# ------------------------------------------------------------------------------------------------
def _angle_between(v1, v2):
    v1 = v1 / np.linalg.norm(v1)
    v2 = v2 / np.linalg.norm(v2)
    return np.arccos(np.clip(np.dot(v1, v2), -1.0, 1.0))


def identify_projection_points(ego_xy, corner_points):
    """helper_functions.py:157-176: the corner pair with the largest subtended angle (first maximum)."""
    max_angle = 0
    ret = (corner_points[0], corner_points[0])
    for p1 in corner_points:
        for p2 in corner_points:
            a = _angle_between(p1 - ego_xy, p2 - ego_xy)
            if a > max_angle:
                max_angle = a
                ret = (p1, p2)
    return ret


def _in_convex_quad(P, quad):
    """Closed point-in-convex-polygon test for points [M,2]; orientation independent."""
    q = np.asarray(quad, dtype=np.float64)
    e = np.roll(q, -1, axis=0) - q
    rel = P[:, None, :] - q[None]
    cr = e[None, :, 0] * rel[..., 1] - e[None, :, 1] * rel[..., 0]
    return np.all(cr >= 0, 1) | np.all(cr <= 0, 1)


def reference_point_visible(P, ego, rect, flags, boundary, sensor_radius, fov_deg):
    """Point-wise evaluation of the reference's visible-area construction (see module docstring).
    ``boundary``: [B,4] consecutive exterior-vertex pairs (x1,y1,x2,y2) of road ∩ sector."""
    P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
    ego = np.asarray(ego, dtype=np.float64)
    e = ego[:2]
    d = P - e
    r = np.hypot(d[:, 0], d[:, 1])
    vis = r <= sensor_radius
    if fov_deg < 359.9:                                    # sensor_model.py:201-209 sector
        rel = (np.arctan2(d[:, 1], d[:, 0]) - ego[2] + np.pi) % (2 * np.pi) - np.pi
        vis &= np.abs(rel) <= np.radians(fov_deg) / 2
    if boundary is not None:
        for b in np.asarray(boundary, dtype=np.float64).reshape(-1, 4):
            v1, v2 = b[:2], b[2:]
            quad = [v1, v2, v2 + 100 * (v2 - e), v1 + 100 * (v1 - e)]   # helper_functions.py:90-94
            vis &= ~_in_convex_quad(P, quad)
    rect = np.asarray(rect, dtype=np.float64).reshape(-1, 5)
    cor = rect_corners(rect)
    for o in range(len(rect)):
        if not (flags[o] & RECT_EXISTS) or (flags[o] & RECT_TRANSPARENT):   # sensor_model.py:177
            continue
        c1, c2 = identify_projection_points(e, cor[o])
        u1 = (c1 - e) / np.linalg.norm(c1 - e)
        u2 = (c2 - e) / np.linalg.norm(c2 - e)
        quad = [c1, c2, c2 + u2 * 100, c1 + u1 * 100]                    # helper_functions.py:143-148
        vis &= ~_in_convex_quad(P, quad)
        vis &= ~_in_convex_quad(P, cor[o])                                # obstacle polygon itself
    return vis


# ------------------------------------------------------------------------------------------------
def rollout_cv(x0, y0, v, phi, dt, horizon, var0=0.1, var_factor=1.05):
    """Pedestrian constant-velocity prediction, float64, exactly the reference's arithmetic
    (agent.py:487-503, 520-536, 260-280).  Returns dict of [A,T] arrays."""
    x0, y0, v, phi = (np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in (x0, y0, v, phi))
    n = int(horizon / dt) + 1                                              # agent.py:496
    vx = np.array([round(float(a), 3) for a in v * np.cos(phi)])          # agent.py:492
    vy = np.array([round(float(a), 3) for a in v * np.sin(phi)])          # agent.py:493
    t = np.arange(n)[None, :] * dt                                         # agent.py:499
    k = np.arange(n)
    return {"x": x0[:, None] + t * vx[:, None], "y": y0[:, None] + t * vy[:, None],
            "yaw": np.repeat(phi[:, None], n, 1), "v": np.repeat(v[:, None], n, 1),
            "var": np.repeat((var0 * np.power(var_factor, k))[None], len(v), 0)}


def rollout_path(path, x0, y0, v0, dt, horizon, t1=3.0, var0=0.1, var_factor=1.05):
    """float64 restatement of the path-following vehicle rollout for ONE reference polyline
    (agent.py:364-426 + utils/frenetix_handler.py:66-125; the C++ frenetix arithmetic is restated from
    the published Werling quartic/quintic scheme -- PARITY UNPINNED).  ``path`` is rounded to float32
    like the kernel input.  Returns dict of [T] arrays and the selected sample index."""
    pts = np.asarray(path, dtype=np.float64).astype(np.float32).astype(np.float64).reshape(-1, 2)
    seg = np.diff(pts, axis=0)
    ln = np.hypot(seg[:, 0], seg[:, 1])
    cum = np.concatenate(([0.0], np.cumsum(ln)))
    n = len(pts)
    # projection
    best = (np.inf, 0.0, 0.0)
    for j in range(n - 1):
        if ln[j] <= 0:
            continue
        u = ((x0 - pts[j, 0]) * seg[j, 0] + (y0 - pts[j, 1]) * seg[j, 1]) / ln[j] ** 2
        uc = u if ((j == 0 and u < 0) or (j == n - 2 and u > 1)) else min(max(u, 0.0), 1.0)
        q = pts[j] + uc * seg[j]
        dist = np.hypot(x0 - q[0], y0 - q[1])
        if dist < best[0]:
            best = (dist, cum[j] + uc * ln[j], (seg[j, 0] * (y0 - pts[j, 1]) - seg[j, 1] * (x0 - pts[j, 0])) / ln[j])
    s0, d0 = best[1], best[2]
    T = int(horizon / dt) + 1
    t = np.arange(T) * dt
    head = np.arctan2(seg[:, 1], seg[:, 0])

    def frenet(sd1, d1):
        a3, a4 = (sd1 - v0) / t1 ** 2, (v0 - sd1) / (2 * t1 ** 3)
        tc = np.minimum(t, t1)
        s = s0 + v0 * tc + a3 * tc ** 3 + a4 * tc ** 4 + np.where(t > t1, sd1 * (t - t1), 0.0)
        sd = np.where(t < t1, v0 + 3 * a3 * tc ** 2 + 4 * a4 * tc ** 3, sd1)
        if v0 < 0.5:                     # low-velocity mode (frenetix_handler.py:91-95): lateral quintic over arc length
            s1 = s0 + v0 * t1 + a3 * t1 ** 3 + a4 * t1 ** 4
            span = max(s1 - s0, 1e-9)
            tau = np.clip((s - s0) / span, 0.0, 1.0)
            scale = sd / span
        else:
            tau = tc / t1
            scale = 1.0 / t1
        inside = (t < t1) | (v0 < 0.5)   # low-velocity mode: d follows the covered arc length at every t
        d = np.where(inside, d0 + (d1 - d0) * (10 * tau ** 3 - 15 * tau ** 4 + 6 * tau ** 5), d1)
        dd = np.where(inside, (d1 - d0) * scale * (30 * tau ** 2 - 60 * tau ** 3 + 30 * tau ** 4), 0.0)
        return s, sd, d, dd

    def lookup(s):
        j = np.clip(np.searchsorted(cum, s, side="right") - 1, 0, n - 2)
        kap = np.zeros_like(s)
        ok = j + 2 < n
        jj = np.where(ok, j, 0)
        dh = head[np.minimum(jj + 1, n - 2)] - head[jj]
        dh -= 2 * np.pi * np.rint(dh / (2 * np.pi))
        kap = np.where(ok, dh / (0.5 * (cum[np.minimum(jj + 2, n - 1)] - cum[jj])), 0.0)
        return j, kap

    best_k, best_var, keep = 0, np.inf, None
    for smp in range(9):
        sd1, d1 = v0 * (0.8 + 0.2 * (smp // 3)), -0.5 + 0.5 * (smp % 3)
        s, sd, d, dd = frenet(sd1, d1)
        j, kap = lookup(s)
        vl = sd * (1 - kap * d)
        v = np.sqrt(vl ** 2 + dd ** 2)
        var = max(np.mean(v * v) - np.mean(v) ** 2, 0.0)
        if var < best_var - 1e-15:
            best_k, best_var, keep = smp, var, (s, sd, d, dd, j, kap, vl, v)
    s, sd, d, dd, j, kap, vl, v = keep
    l = np.maximum(cum[j + 1] - cum[j], 1e-12)
    u = (s - cum[j]) / l
    tx, ty = seg[j, 0] / l, seg[j, 1] / l
    k = np.arange(T)
    return {"x": pts[j, 0] + u * seg[j, 0] - d * ty, "y": pts[j, 1] + u * seg[j, 1] + d * tx,
            "yaw": np.arctan2(ty, tx) + np.arctan2(dd, vl), "v": v, "var": var0 * np.power(var_factor, k),
            "sample": best_k, "s0": s0, "d0": d0}


# ------------------------------------------------------------------------------------------------
PT_IN_SENSOR, PT_ON_ROAD, PT_SHADOWED, PT_IN_OBSTACLE, PT_VISIBLE, PT_OCCLUDED, PT_FOCUS_SHADOW, PT_FOCUS_NEAR = \
    1, 2, 4, 8, 16, 32, 64, 128


def points_in_polygon(P, poly):
    """Even-odd rule in float64 for points [M,2] against one open ring [V,2]."""
    P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
    poly = np.asarray(poly, dtype=np.float64).reshape(-1, 2)
    x, y = P[:, 0][:, None], P[:, 1][:, None]
    x1, y1 = poly[:, 0][None], poly[:, 1][None]
    x0, y0 = np.roll(poly[:, 0], 1)[None], np.roll(poly[:, 1], 1)[None]
    cond = (y0 > y) != (y1 > y)
    with np.errstate(divide="ignore", invalid="ignore"):
        xin = (x1 - x0) * (y - y0) / (y1 - y0) + x0
    return (np.sum(cond & (x < xin), axis=1) % 2).astype(bool)


def classify_points(P, ego, rect, flags, boundary, polygons, sensor_radius, fov_deg, occluded_radius, focus=-1,
                    focus_margin=0.0):
    """Point-wise evaluation of the reference's visible / occluded area construction (float64):
    visible_area = road ∩ sector − border-edge shadow quads − obstacles − obstacle shadows
    (sensor_model.py:103-193), occluded_area = (±90° sector of radius 1.5 R) ∩ road − visible_area
    (sensor_model.py:85-93), obstacle_occlusions[id] = shadow(id) − obstacle(id) (sensor_model.py:182-183).
    Returns (flags uint32 [M], lanelet bit mask uint64 [M])."""
    P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
    ego = np.asarray(ego, dtype=np.float64)
    e = ego[:2]
    d = P - e
    r = np.hypot(d[:, 0], d[:, 1])
    rel = (np.arctan2(d[:, 1], d[:, 0]) - ego[2] + np.pi) % (2 * np.pi) - np.pi
    in_sensor = r <= sensor_radius
    if fov_deg < 359.9:
        in_sensor &= np.abs(rel) <= np.radians(fov_deg) / 2
    shadow = np.zeros(len(P), dtype=bool)
    if boundary is not None:
        for b in np.asarray(boundary, dtype=np.float64).reshape(-1, 4):
            v1, v2 = b[:2], b[2:]
            if np.allclose(v1, v2):
                continue
            shadow |= _in_convex_quad(P, [v1, v2, v2 + 100 * (v2 - e), v1 + 100 * (v1 - e)])
    rect = np.asarray(rect, dtype=np.float64).reshape(-1, 5)
    flags = np.asarray(flags).reshape(-1)
    cor = rect_corners(rect)
    in_obst = np.zeros(len(P), dtype=bool)
    focus_shadow = np.zeros(len(P), dtype=bool)
    focus_near = np.zeros(len(P), dtype=bool)
    for o in range(len(rect)):
        if not (flags[o] & RECT_EXISTS) or (flags[o] & RECT_TRANSPARENT):
            continue
        c1, c2 = identify_projection_points(e, cor[o])
        u1 = (c1 - e) / np.linalg.norm(c1 - e)
        u2 = (c2 - e) / np.linalg.norm(c2 - e)
        wedge = _in_convex_quad(P, [c1, c2, c2 + u2 * 100, c1 + u1 * 100])
        inside = _in_convex_quad(P, cor[o])
        shadow |= wedge & ~inside
        in_obst |= inside
        if o == focus:
            focus_shadow = wedge & ~inside
            # distance to the rectangle (0 inside) <= margin  ==  within polygon.buffer(margin) (round joins)
            cs, sn = np.cos(rect[o, 2]), np.sin(rect[o, 2])
            dl = P - rect[o, :2]
            lx, ly = dl[:, 0] * cs + dl[:, 1] * sn, -dl[:, 0] * sn + dl[:, 1] * cs
            focus_near = np.hypot(np.maximum(np.abs(lx) - rect[o, 3], 0), np.maximum(np.abs(ly) - rect[o, 4], 0)) <= focus_margin
    lan = np.zeros(len(P), dtype=np.uint64)
    on_road = np.zeros(len(P), dtype=bool)
    for i, poly in enumerate(polygons):
        m = points_in_polygon(P, poly)
        on_road |= m
        if i < 64:
            lan |= np.where(m, np.uint64(1) << np.uint64(i), np.uint64(0))
    visible = in_sensor & on_road & ~shadow & ~in_obst
    occluded = on_road & ~visible & (np.abs(rel) <= np.pi / 2) & (r <= occluded_radius)
    f = (in_sensor * PT_IN_SENSOR + on_road * PT_ON_ROAD + shadow * PT_SHADOWED + in_obst * PT_IN_OBSTACLE
         + visible * PT_VISIBLE + occluded * PT_OCCLUDED + focus_shadow * PT_FOCUS_SHADOW
         + focus_near * PT_FOCUS_NEAR).astype(np.uint32)
    return f, lan
