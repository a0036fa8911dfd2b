"""float64 convex-polygon predicates standing in for shapely/GEOS (oracle, test infrastructure).

The reference calls ``shapely`` (GEOS) ``Polygon.distance`` (``metrics/dce.py:75-79``) and
``Polygon.intersects`` (``metrics/be.py:181``) on pairs of *rectangles*.  shapely 2.0.2 is not
installable here, so these are restated for convex polygons, where both have an
unambiguous closed form:

* ``intersects``: the closed point sets share at least one point  <=>  no separating axis with a
  strictly positive gap (touching counts, as in GEOS).
* ``distance``: 0 when they intersect (including containment), otherwise the minimum over all
  (vertex, edge) pairs of the point-segment distance.

Two independent formulations are kept so they can be cross-checked in the tests:
``ConvexPolygon`` (generic vertices, used by the oracle-A shim) and ``obb_distance`` /
``obb_intersects`` (vectorised oriented boxes in a local frame, used by oracle B).
"""
from __future__ import annotations

import numpy as np


def rect_vertices(cx, cy, length, width, yaw):
    """Corner ring of an oriented rectangle, commonroad ``Rectangle`` vertex order
    ([-l/2,-w/2], [-l/2,w/2], [l/2,w/2], [l/2,-w/2]), rotated by ``yaw`` then translated."""
    hl, hw = 0.5 * length, 0.5 * width
    loc = np.array([[-hl, -hw], [-hl, hw], [hl, hw], [hl, -hw]], dtype=np.float64)
    c, s = np.cos(yaw), np.sin(yaw)
    rot = np.array([[c, -s], [s, c]], dtype=np.float64)
    return loc @ rot.T + np.array([cx, cy], dtype=np.float64)


class ConvexPolygon:
    """Minimal stand-in for ``shapely.geometry.Polygon`` restricted to convex rings."""

    def __init__(self, vertices):
        v = np.asarray(vertices, dtype=np.float64)
        if len(v) > 1 and np.all(v[0] == v[-1]):
            v = v[:-1]
        self.v = v

    # -- helpers -------------------------------------------------------------------------
    def _axes(self):
        e = np.roll(self.v, -1, axis=0) - self.v
        return np.stack((-e[:, 1], e[:, 0]), axis=1)

    @staticmethod
    def _pt_seg(p, a, b):
        ab = b - a
        l2 = float(ab @ ab)
        if l2 == 0.0:
            return float(np.hypot(*(p - a)))
        r = float((p - a) @ ab) / l2
        if r <= 0.0:
            return float(np.hypot(*(p - a)))
        if r >= 1.0:
            return float(np.hypot(*(p - b)))
        s = ((a[1] - p[1]) * ab[0] - (a[0] - p[0]) * ab[1]) / l2
        return abs(s) * np.sqrt(l2)

    # -- shapely-like API ------------------------------------------------------------------
    def intersects(self, other: "ConvexPolygon") -> bool:
        for ax in np.concatenate((self._axes(), other._axes())):
            pa = self.v @ ax
            pb = other.v @ ax
            if pa.max() < pb.min() or pb.max() < pa.min():
                return False
        return True

    def distance(self, other: "ConvexPolygon") -> float:
        if self.intersects(other):
            return 0.0
        best = np.inf
        for P, Q in ((self.v, other.v), (other.v, self.v)):
            n = len(Q)
            for p in P:
                for k in range(n):
                    d = self._pt_seg(p, Q[k], Q[(k + 1) % n])
                    if d < best:
                        best = d
        return float(best)


# ------------------------------------------------------------------------------------------
# vectorised oriented boxes (oracle B)
# ------------------------------------------------------------------------------------------
def _point_box_dist2(px, py, hx, hy):
    dx = np.maximum(np.abs(px) - hx, 0.0)
    dy = np.maximum(np.abs(py) - hy, 0.0)
    return dx * dx + dy * dy


def obb_intersects(ax, ay, ayaw, ahl, ahw, bx, by, byaw, bhl, bhw):
    """Separating-axis test for two oriented boxes (broadcasting float64 arrays).
    Touching counts as intersecting (GEOS ``intersects`` semantics)."""
    ca, sa = np.cos(ayaw), np.sin(ayaw)
    cb, sb = np.cos(byaw), np.sin(byaw)
    rx, ry = bx - ax, by - ay
    # relative position in A's frame and in B's frame
    rax = rx * ca + ry * sa
    ray = -rx * sa + ry * ca
    rbx = rx * cb + ry * sb
    rby = -rx * sb + ry * cb
    c = np.abs(ca * cb + sa * sb)
    s = np.abs(sb * ca - cb * sa)
    sep = (np.abs(rax) > ahl + bhl * c + bhw * s) | (np.abs(ray) > ahw + bhl * s + bhw * c) | \
          (np.abs(rbx) > bhl + ahl * c + ahw * s) | (np.abs(rby) > bhw + ahl * s + ahw * c)
    return ~sep


def obb_distance(ax, ay, ayaw, ahl, ahw, bx, by, byaw, bhl, bhw):
    """Euclidean distance between two oriented boxes; 0 where they intersect.

    For disjoint convex polygons the minimum distance is attained at a vertex of one polygon
    against the (closed) other polygon, so the minimum over the 8 corner-to-box distances,
    each evaluated in the other box's local frame, is exact."""
    ca, sa = np.cos(ayaw), np.sin(ayaw)
    cb, sb = np.cos(byaw), np.sin(byaw)
    hit = obb_intersects(ax, ay, ayaw, ahl, ahw, bx, by, byaw, bhl, bhw)
    rx, ry = bx - ax, by - ay
    best = None
    for sx in (-1.0, 1.0):
        for sy in (-1.0, 1.0):
            # corner of B in world coordinates relative to A's centre, then in A's frame
            wx = rx + sx * bhl * cb - sy * bhw * sb
            wy = ry + sx * bhl * sb + sy * bhw * cb
            d2 = _point_box_dist2(wx * ca + wy * sa, -wx * sa + wy * ca, ahl, ahw)
            best = d2 if best is None else np.minimum(best, d2)
            # corner of A relative to B's centre, in B's frame
            wx = -rx + sx * ahl * ca - sy * ahw * sa
            wy = -ry + sx * ahl * sa + sy * ahw * ca
            d2 = _point_box_dist2(wx * cb + wy * sb, -wx * sb + wy * cb, bhl, bhw)
            best = np.minimum(best, d2)
    return np.where(hit, 0.0, np.sqrt(best))
