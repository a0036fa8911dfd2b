"""Geometry helpers of the hot path (mirror of reference utils/helper_functions.py and of the few
commonroad_dc / commonroad_route_planner utilities spawn_locator.py imports), polygon-library free.

All functions are small host-side numpy; the heavy predicates (visibility / occlusion / road membership of
sampled points) live in the CUDA kernels."""
from __future__ import annotations

import numpy as np


def vector_from_angle(angle_rad):
    """helper_functions.py:72-76."""
    return np.array([np.cos(angle_rad), np.sin(angle_rad)])


def compute_pathlength_from_polyline(polyline) -> np.ndarray:
    """commonroad_dc.geometry.util.compute_pathlength_from_polyline: cumulative chord length, starts at 0."""
    p = np.asarray(polyline, dtype=np.float64)
    return np.concatenate(([0.0], np.cumsum(np.hypot(*np.diff(p, axis=0).T))))


def compute_curvature_from_polyline(polyline) -> np.ndarray:
    """commonroad_dc.geometry.util.compute_curvature_from_polyline (2023.1): second-order numpy gradients."""
    p = np.asarray(polyline, dtype=np.float64)
    x_d, y_d = np.gradient(p[:, 0]), np.gradient(p[:, 1])
    x_dd, y_dd = np.gradient(x_d), np.gradient(y_d)
    return (x_d * y_dd - x_dd * y_d) / ((x_d ** 2 + y_d ** 2) ** 1.5)


def oriented_rectangle_corners(pos, length, width, orientation) -> np.ndarray:
    """helper_functions.py:14-35 ``create_oriented_rectangle``: corner ring [4,2]."""
    dx, dy = 0.5 * length, 0.5 * width
    loc = np.array([[-dx, -dy], [-dx, dy], [dx, dy], [dx, -dy]])
    c, s = np.cos(orientation), np.sin(orientation)
    return loc @ np.array([[c, -s], [s, c]]).T + np.asarray(pos, dtype=np.float64)


def point_in_convex_ring(P, ring) -> np.ndarray:
    """Closed membership of points [M,2] in a convex ring [V,2] (either orientation)."""
    P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
    q = np.asarray(ring, dtype=np.float64)
    e = np.roll(q, -1, axis=0) - q
    rel = P[:, None, :] - q[None]
    cr = e[None, :, 0] * rel[..., 1] - e[None, :, 1] * rel[..., 0]
    return np.all(cr >= 0, 1) | np.all(cr <= 0, 1)


def point_segment_distance(P, a, b) -> np.ndarray:
    P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    e = b - a
    l2 = max(float(e @ e), 1e-18)
    t = np.clip(((P - a) @ e) / l2, 0.0, 1.0)
    q = a + t[:, None] * e
    return np.hypot(*(P - q).T)


def point_ring_distance(P, ring) -> np.ndarray:
    """Distance of points to a convex polygon given by its ring (0 inside)."""
    P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
    ring = np.asarray(ring, dtype=np.float64)
    d = np.full(len(P), np.inf)
    for i in range(len(ring)):
        d = np.minimum(d, point_segment_distance(P, ring[i], ring[(i + 1) % len(ring)]))
    return np.where(point_in_convex_ring(P, ring), 0.0, d)


def point_rings_min_distance(p, rings) -> float:
    """min over convex rings of ``point_ring_distance(p, ring)`` for ONE point (inf without rings): all rings in one
    vectorised pass when they have the same vertex count (obstacle rectangles)."""
    if rings is None or len(rings) == 0:
        return float("inf")
    try:
        R = np.asarray(rings, dtype=np.float64)
    except ValueError:
        R = None
    if R is None or R.ndim != 3:
        return float(min(point_ring_distance(np.asarray(p)[None], ring)[0] for ring in rings))
    p = np.asarray(p, dtype=np.float64).reshape(2)
    e = np.roll(R, -1, axis=1) - R
    rel = p - R
    l2 = np.maximum((e * e).sum(-1), 1e-18)
    t = np.clip((rel * e).sum(-1) / l2, 0.0, 1.0)
    q = R + t[..., None] * e
    d = np.hypot(p[0] - q[..., 0], p[1] - q[..., 1]).min(1)
    cr = e[..., 0] * rel[..., 1] - e[..., 1] * rel[..., 0]
    inside = np.all(cr >= 0, 1) | np.all(cr <= 0, 1)
    return float(np.where(inside, 0.0, d).min())


def _segments_intersect(a, b, c, d) -> bool:
    def orient(p, q, r):
        return (q[0] - p[0]) * (r[1] - p[1]) - (q[1] - p[1]) * (r[0] - p[0])
    o1, o2, o3, o4 = orient(a, b, c), orient(a, b, d), orient(c, d, a), orient(c, d, b)
    return (o1 * o2 <= 0) and (o3 * o4 <= 0) and not (o1 == 0 and o2 == 0 and o3 == 0 and o4 == 0 and
                                                        (max(a[0], b[0]) < min(c[0], d[0]) or max(c[0], d[0]) < min(a[0], b[0])
                                                         or max(a[1], b[1]) < min(c[1], d[1]) or max(c[1], d[1]) < min(a[1], b[1])))


def segment_ring_distance(a, b, ring) -> float:
    """Distance between the segment a-b and a convex polygon ring (0 when they touch or overlap)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ring = np.asarray(ring, dtype=np.float64)
    if point_in_convex_ring(a[None], ring)[0] or point_in_convex_ring(b[None], ring)[0]:
        return 0.0
    best = np.inf
    n = len(ring)
    for i in range(n):
        c, d = ring[i], ring[(i + 1) % n]
        if _segments_intersect(a, b, c, d):
            return 0.0
        best = min(best, float(point_segment_distance(np.stack((a, b)), c, d).min()),
                   float(point_segment_distance(np.stack((c, d)), a, b).min()))
    return best


def disc_samples(center, radius, n_rim=32, n_rings=3) -> np.ndarray:
    """Centre + ``n_rings`` concentric rims (the outermost at ``radius``) sampling a closed disc."""
    c = np.asarray(center, dtype=np.float64)
    ang = np.linspace(0.0, 2.0 * np.pi, n_rim, endpoint=False)
    pts = [c[None]]
    for k in range(1, n_rings + 1):
        r = radius * k / n_rings
        pts.append(c + r * np.stack((np.cos(ang), np.sin(ang)), -1))
    return np.concatenate(pts)


def disc_samples_many(centers, radius, n_rim=32, n_rings=3) -> np.ndarray:
    """``disc_samples`` for K centres at once: [K, 1 + n_rim * n_rings, 2], same sample order per disc."""
    c = np.asarray(centers, dtype=np.float64).reshape(-1, 2)
    ang = np.linspace(0.0, 2.0 * np.pi, n_rim, endpoint=False)
    unit = np.stack((np.cos(ang), np.sin(ang)), -1)
    offs = np.concatenate([np.zeros((1, 2))] + [radius * k / n_rings * unit for k in range(1, n_rings + 1)])
    return c[:, None, :] + offs[None, :, :]


def min_area_rectangle(points):
    """Minimum rotated bounding rectangle of a point set: (area, width, height, angle) by rotating calipers
    over the convex hull edges (shapely ``minimum_rotated_rectangle`` semantics)."""
    from scipy.spatial import ConvexHull, QhullError
    P = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    if len(P) < 3:
        return 0.0, 0.0, 0.0, 0.0
    try:
        hull = P[ConvexHull(P).vertices]
    except QhullError:
        return 0.0, 0.0, 0.0, 0.0
    e = np.roll(hull, -1, axis=0) - hull
    ang = np.unique(np.mod(np.arctan2(e[:, 1], e[:, 0]), np.pi / 2))
    c, s = np.cos(ang), np.sin(ang)
    u = hull[:, :1] * c + hull[:, 1:] * s              # [H, K]: all caliper directions at once
    v = hull[:, 1:] * c - hull[:, :1] * s
    w, h = u.max(0) - u.min(0), v.max(0) - v.min(0)
    k = int(np.argmin(w * h))                          # first minimum in angle order
    return float(w[k] * h[k]), float(w[k]), float(h[k]), float(ang[k])
