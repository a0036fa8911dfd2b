"""Float64 planar-geometry stand-in for the shapely 2.0.2 / GEOS surface that the reference's
``sensor_model.py``, ``spawn_locator.py``, ``utils/fo_obstacle.py`` and ``utils/helper_functions.py`` use
(``poetry.lock:1379``; shapely cannot be installed in this image).  TEST INFRASTRUCTURE: ``oracle/ref_pipeline.py``
injects these classes as ``shapely.geometry.*`` / ``shapely.ops.unary_union`` so that the reference's own files run
**unmodified**; nothing under ``frenetix_occlusion_b200/`` imports this module and this module imports nothing
of the product.

What is restated is third-party behaviour only (published semantics of the OGC simple-feature operations GEOS
implements), not the library's code:

* ``intersection / union / difference / unary_union`` -- overlay of polygonal point sets: all boundary segments
  are noded (proper crossings, T-junctions, collinear overlaps; vertices closer than ``EPS`` = 1e-9 m are merged), each
  unique noded edge is classified by point-in-region tests ``DELTA`` = 1e-7 m to its left and right for every
  operand, edges with different result values on their two sides form the result boundary, which is traced into
  rings (interior on the left), split at touching vertices into minimal rings, and assembled into polygons with
  holes.  Exact up to the two tolerances; lower-dimensional leftovers (GEOS returns them inside a
  ``GeometryCollection``) are dropped, which is what the reference does with them anyway
  (``sensor_model.py:211-234``, ``helper_functions.py:115-136``).
* ``buffer(d)`` -- Minkowski sum with a disc, arcs approximated with GEOS' default 16 segments per quadrant
  (``join_style=2``: mitred corners with the default mitre limit 5); built as the union of the geometry, one
  rectangle per boundary edge and one fan / mitre kite per vertex.  ``Point.buffer`` reproduces GEOS' 64-gon
  (first vertex at angle 0, clockwise).  ``buffer(0)`` re-normalises.
* predicates ``intersects / within`` (boundary contact counts for ``intersects``; ``Point.within`` needs the
  interior), ``distance`` is not needed; ``area``, ``centroid``, ``minimum_rotated_rectangle`` (rotating calipers
  over the convex hull), ``is_valid`` for single rings (simple ring with non-zero area), ``LineString.intersection``
  with a polygonal geometry (pieces in the direction and order of the line), ``LinearRing.intersection`` with a
  line (points), ``LineString.project / interpolate``.

PARITY NOTE: GEOS is not available, so this stand-in is pinned by its own property tests
(``tests/test_polygon_oracle.py``: random overlays against point sampling, closed-form areas, buffer areas) -- the
reference's *Python* control flow above it is the reference's own.
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-9       # vertices closer than this are the same vertex
DELTA = 1e-7     # offset of the side samples that classify a noded edge
_QUAD_SEGS = 16  # GEOS default: segments per quarter circle
_MITRE_LIMIT = 5.0


# =====================================================================================================================
# low-level helpers
# =====================================================================================================================
def _as_ring(coords) -> np.ndarray:
    """[n, 2] array without the closing duplicate and without consecutive duplicates."""
    a = np.asarray(coords, dtype=np.float64).reshape(-1, 2)
    if len(a) > 1 and np.all(np.abs(a[0] - a[-1]) <= 0.0):
        a = a[:-1]
    if len(a) > 1:
        keep = np.ones(len(a), dtype=bool)
        keep[1:] = np.any(np.abs(np.diff(a, axis=0)) > 0.0, axis=1)
        if np.all(np.abs(a[0] - a[-1]) <= 0.0) and keep.sum() > 1:
            keep[-1] = False
        a = a[keep]
    return a


def _signed_area(r: np.ndarray) -> float:
    if len(r) < 3:
        return 0.0
    x, y = r[:, 0], r[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _ring_segments(r: np.ndarray):
    return r, np.roll(r, -1, axis=0)


def _points_in_rings(pts: np.ndarray, rings) -> np.ndarray:
    """Even-odd rule over all rings (shells and holes alike)."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    inside = np.zeros(len(pts), dtype=bool)
    if not len(pts):
        return inside
    px, py = pts[:, 0][:, None], pts[:, 1][:, None]
    for r in rings:
        if len(r) < 3:
            continue
        a, b = _ring_segments(r)
        ax, ay, bx, by = a[:, 0][None], a[:, 1][None], b[:, 0][None], b[:, 1][None]
        for lo in range(0, len(pts), 4096):           # bounded temporaries
            sl = slice(lo, lo + 4096)
            cond = (ay > py[sl]) != (by > py[sl])
            with np.errstate(divide="ignore", invalid="ignore"):
                xi = ax + (py[sl] - ay) * (bx - ax) / (by - ay)
            inside[sl] ^= (np.sum(cond & (px[sl] < xi), axis=1) % 2).astype(bool)
    return inside


def _dist_points_segments(pts, a, b):
    """[n_pts, n_seg] distances."""
    d = b - a
    l2 = np.maximum(np.sum(d * d, axis=1), 1e-300)
    rel = pts[:, None, :] - a[None, :, :]
    t = np.clip(np.sum(rel * d[None], axis=2) / l2[None], 0.0, 1.0)
    foot = a[None] + t[..., None] * d[None]
    return np.hypot(pts[:, None, 0] - foot[..., 0], pts[:, None, 1] - foot[..., 1]), t


def _on_boundary(pts, rings, tol=EPS):
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    out = np.zeros(len(pts), dtype=bool)
    for r in rings:
        if len(r) < 2:
            continue
        a, b = _ring_segments(r)
        dist, _ = _dist_points_segments(pts, a, b)
        out |= np.any(dist <= tol, axis=1)
    return out


class _Snapper:
    """Merges points closer than EPS (grid hash with neighbour-cell lookup)."""

    def __init__(self):
        self.cells = {}
        self.pts = []

    def add(self, x, y) -> int:
        cx, cy = math.floor(x / (4 * EPS)), math.floor(y / (4 * EPS))
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for idx in self.cells.get((cx + dx, cy + dy), ()):
                    q = self.pts[idx]
                    if abs(q[0] - x) <= EPS and abs(q[1] - y) <= EPS:
                        return idx
        idx = len(self.pts)
        self.pts.append((x, y))
        self.cells.setdefault((cx, cy), []).append(idx)
        return idx


def _node_segments(A: np.ndarray, B: np.ndarray):
    """A, B: [m, 2] segment end points.  Returns (vertex array [v, 2], list of unique undirected edges (i, j), list of the
    input segments each edge is a piece of)."""
    m = len(A)
    if m == 0:
        return np.zeros((0, 2)), [], []
    lo = np.minimum(A, B) - 2 * EPS
    hi = np.maximum(A, B) + 2 * EPS
    splits = [[] for _ in range(m)]                    # (param, x, y)
    d = B - A
    len2 = np.maximum(np.sum(d * d, axis=1), 1e-300)
    # ---- proper crossings (chunked all-pairs with a bounding-box prefilter) --------------------------------------
    cross_pts = []
    step = max(1, int(4_000_000 // max(m, 1)))
    for i0 in range(0, m, step):
        i1 = min(m, i0 + step)
        ov = ((lo[i0:i1, None, 0] <= hi[None, :, 0]) & (hi[i0:i1, None, 0] >= lo[None, :, 0]) &
              (lo[i0:i1, None, 1] <= hi[None, :, 1]) & (hi[i0:i1, None, 1] >= lo[None, :, 1]))
        ii, jj = np.nonzero(ov)
        ii = ii + i0
        keep = ii < jj
        ii, jj = ii[keep], jj[keep]
        if not len(ii):
            continue
        r, s = d[ii], d[jj]
        qp = A[jj] - A[ii]
        den = r[:, 0] * s[:, 1] - r[:, 1] * s[:, 0]
        ok = np.abs(den) > 1e-14 * np.sqrt(len2[ii] * len2[jj])
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (qp[:, 0] * s[:, 1] - qp[:, 1] * s[:, 0]) / den
            u = (qp[:, 0] * r[:, 1] - qp[:, 1] * r[:, 0]) / den
        hit = ok & (t > 0.0) & (t < 1.0) & (u > 0.0) & (u < 1.0)
        for i, j, tt, uu in zip(ii[hit], jj[hit], t[hit], u[hit]):
            x, y = A[i, 0] + tt * d[i, 0], A[i, 1] + tt * d[i, 1]
            splits[i].append((tt, x, y))
            splits[j].append((uu, x, y))
            cross_pts.append((x, y))
    # ---- vertices (end points and crossings) lying on other segments: T-junctions and collinear overlaps ---------
    V = np.concatenate([A, B, np.asarray(cross_pts, dtype=np.float64).reshape(-1, 2)])
    stepv = max(1, int(2_000_000 // max(m, 1)))
    for v0 in range(0, len(V), stepv):
        P = V[v0:v0 + stepv]
        inb = ((P[:, None, 0] >= lo[None, :, 0]) & (P[:, None, 0] <= hi[None, :, 0]) &
               (P[:, None, 1] >= lo[None, :, 1]) & (P[:, None, 1] <= hi[None, :, 1]))
        pi, si = np.nonzero(inb)
        if not len(pi):
            continue
        rel = P[pi] - A[si]
        t = np.sum(rel * d[si], axis=1) / len2[si]
        foot = A[si] + t[:, None] * d[si]
        dist = np.hypot(P[pi, 0] - foot[:, 0], P[pi, 1] - foot[:, 1])
        seglen = np.sqrt(len2[si])
        on = (dist <= EPS) & (t * seglen > EPS) & ((1.0 - t) * seglen > EPS)
        for p, sidx, tt in zip(pi[on], si[on], t[on]):
            splits[sidx].append((tt, P[p, 0], P[p, 1]))
    # ---- snap and cut ---------------------------------------------------------------------------------------------
    snap = _Snapper()
    edges = {}
    for i in range(m):
        chain = [(0.0, A[i, 0], A[i, 1])] + sorted(splits[i]) + [(1.0, B[i, 0], B[i, 1])]
        ids = [snap.add(x, y) for _, x, y in chain]
        for p, q in zip(ids[:-1], ids[1:]):
            if p != q:
                edges.setdefault((p, q) if p < q else (q, p), []).append(i)
    keys = sorted(edges)
    return np.asarray(snap.pts, dtype=np.float64).reshape(-1, 2), keys, [edges[k] for k in keys]


def _trace_rings(verts: np.ndarray, directed):
    """directed: list of (i, j) with the result's interior on the LEFT.  Returns minimal rings (vertex index lists)."""
    out_edges = {}
    for e, (i, j) in enumerate(directed):
        out_edges.setdefault(i, []).append(e)
    ang = [math.atan2(verts[j, 1] - verts[i, 1], verts[j, 0] - verts[i, 0]) for i, j in directed]
    used = [False] * len(directed)
    rings = []
    for e0 in range(len(directed)):
        if used[e0]:
            continue
        path = []
        e = e0
        ok = True
        while True:
            used[e] = True
            path.append(directed[e][0])
            v = directed[e][1]
            cands = [c for c in out_edges.get(v, ()) if not used[c] or c == e0]
            if not cands:
                ok = False
                break
            back = ang[e] + math.pi

            def turn(c):                       # counter-clockwise angle from the reversed incoming direction
                a = (ang[c] - back) % (2 * math.pi)
                return a if a > 1e-12 else 2 * math.pi
            # leftmost turn = smallest clockwise angle from the reversed incoming direction = largest ccw angle
            nxt = max(cands, key=turn)
            if nxt == e0:
                break
            e = nxt
        if not ok or len(path) < 3:
            continue
        # split at repeated vertices into minimal rings
        stack, pos = [], {}
        for v in path:
            if v in pos:
                k = pos[v]
                loop = stack[k:]
                for w in loop:
                    pos.pop(w, None)
                del stack[k:]
                if len(loop) >= 3:
                    rings.append(loop)
            pos[v] = len(stack)
            stack.append(v)
        if len(stack) >= 3:
            rings.append(stack)
    return rings


def _assemble(verts, rings):
    """Minimal rings (interior on the left) -> list of Polygon (shell ccw + holes)."""
    shells, holes = [], []
    for idx in rings:
        r = verts[idx]
        a = _signed_area(r)
        if abs(a) <= 1e-14:
            continue
        (shells if a > 0 else holes).append((abs(a), r))
    polys = [[s, []] for s in sorted(shells, key=lambda t: t[0])]        # small shells first
    for ha, h in holes:
        # a point of the polygon's interior next to the hole: just left of its first edge (hole rings run clockwise,
        # the interior is on their left)
        p, q = h[0], h[1]
        n = np.array([-(q[1] - p[1]), q[0] - p[0]])
        n /= max(np.hypot(*n), 1e-300)
        sample = 0.5 * (p + q) + DELTA * n
        for shell, hl in polys:
            if shell[0] > ha and _points_in_rings(sample[None], [shell[1]])[0]:
                hl.append(h)
                break
    return [Polygon(s[1], [h for h in hl]) for s, hl in polys]


def _edge_sides(mid, nrm, tang, A, B, excl_edge, excl_seg):
    """Inside status of a region (even-odd over the segments A -> B) immediately to the LEFT and RIGHT of noded edges.

    Exact, without an offset sample: the status just beyond the edge is the parity of the region's boundary segments
    crossed by the ray from the edge midpoint along +-normal, not counting the segments the edge is a piece of
    (``excl_edge[k]``, ``excl_seg[k]`` index pairs: after noding nothing else passes through the midpoint).  Nearly
    parallel neighbours at any small angle are ordinary crossings of that ray."""
    n = len(mid)
    left = np.zeros(n, dtype=np.int64)
    right = np.zeros(n, dtype=np.int64)
    if not len(A) or n == 0:
        return left.astype(bool), right.astype(bool)
    order = np.argsort(excl_edge, kind="stable")
    excl_edge, excl_seg = np.asarray(excl_edge)[order], np.asarray(excl_seg)[order]
    step = max(1, int(1_500_000 // max(len(A), 1)))
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        sl = slice(lo, hi)
        ra = A[None, :, :] - mid[sl, None, :]
        rb = B[None, :, :] - mid[sl, None, :]
        ua = ra[..., 0] * nrm[sl, None, 0] + ra[..., 1] * nrm[sl, None, 1]
        ub = rb[..., 0] * nrm[sl, None, 0] + rb[..., 1] * nrm[sl, None, 1]
        wa = ra[..., 0] * tang[sl, None, 0] + ra[..., 1] * tang[sl, None, 1]
        wb = rb[..., 0] * tang[sl, None, 0] + rb[..., 1] * tang[sl, None, 1]
        cross = (wa > 0.0) != (wb > 0.0)
        k0, k1 = np.searchsorted(excl_edge, lo), np.searchsorted(excl_edge, hi)
        cross[excl_edge[k0:k1] - lo, excl_seg[k0:k1]] = False
        with np.errstate(divide="ignore", invalid="ignore"):
            ui = ua - wa * (ub - ua) / (wb - wa)
        left[sl] = np.sum(cross & (ui > 0.0), axis=1)
        right[sl] = np.sum(cross & (ui < 0.0), axis=1)
    return (left % 2).astype(bool), (right % 2).astype(bool)


def _overlay(operands, predicate):
    """operands: list of ring lists; predicate(list of bool arrays) -> bool array.  Returns list of Polygon."""
    segA, segB, owner = [], [], []
    for k, rings in enumerate(operands):
        for r in rings:
            if len(r) >= 3:
                a, b = _ring_segments(r)
                segA.append(a)
                segB.append(b)
                owner.append(np.full(len(a), k))
    if not segA:
        return []
    segA, segB, owner = np.concatenate(segA), np.concatenate(segB), np.concatenate(owner)
    verts, edges, sources = _node_segments(segA, segB)
    if not edges:
        return []
    E = np.asarray(edges, dtype=np.int64)
    p, q = verts[E[:, 0]], verts[E[:, 1]]
    mid = 0.5 * (p + q)
    dvec = q - p
    ln = np.maximum(np.hypot(dvec[:, 0], dvec[:, 1]), 1e-300)
    tang = dvec / ln[:, None]
    nrm = np.stack([-tang[:, 1], tang[:, 0]], axis=1)
    src_e = np.array([e for e, ss in enumerate(sources) for _ in ss], dtype=np.int64)
    src_s = np.array([x for ss in sources for x in ss], dtype=np.int64)
    inL, inR = [], []
    for k in range(len(operands)):
        idx = np.nonzero(owner == k)[0]
        if not len(idx):
            z = np.zeros(len(E), dtype=bool)
            inL.append(z)
            inR.append(z.copy())
            continue
        local = np.full(len(owner), -1, dtype=np.int64)
        local[idx] = np.arange(len(idx))
        sel = local[src_s] >= 0
        l, r = _edge_sides(mid, nrm, tang, segA[idx], segB[idx], src_e[sel], local[src_s[sel]])
        inL.append(l)
        inR.append(r)
    resL, resR = predicate(inL), predicate(inR)
    directed = [(int(i), int(j)) for (i, j), l, r in zip(E, resL, resR) if l and not r]
    directed += [(int(j), int(i)) for (i, j), l, r in zip(E, resL, resR) if r and not l]
    if not directed:
        return []
    return _assemble(verts, _trace_rings(verts, directed))


def _pred_union(ins):
    out = ins[0].copy()
    for a in ins[1:]:
        out |= a
    return out


def _pred_intersection(ins):
    out = ins[0].copy()
    for a in ins[1:]:
        out &= a
    return out


def _pred_difference(ins):
    out = ins[0].copy()
    for a in ins[1:]:
        out &= ~a
    return out


# =====================================================================================================================
# geometry classes (the shapely names the reference imports)
# =====================================================================================================================
class _Coords(list):
    @property
    def xy(self):
        a = np.asarray(self, dtype=np.float64).reshape(-1, 2)
        return a[:, 0].copy(), a[:, 1].copy()


class BaseGeometry:
    geom_type = "GeometryCollection"

    # ---- structure -----------------------------------------------------------------------------------------------
    def _rings(self):
        return []

    def _polys(self):
        return []

    @property
    def is_empty(self):
        return True

    @property
    def area(self):
        return 0.0

    @property
    def is_valid(self):
        return True

    # ---- set operations on polygonal geometries -----------------------------------------------------------------------
    def intersection(self, other):
        if isinstance(other, (LineString,)):
            return other.intersection(self)
        if isinstance(other, Point):
            return other if other.within(self) or _on_boundary([other._p], self._rings())[0] else GeometryCollection()
        if self.is_empty or other.is_empty:
            return Polygon()
        return _collect(_overlay([self._rings(), other._rings()], _pred_intersection))

    def union(self, other):
        if self.is_empty:
            return _collect(other._polys())
        if other.is_empty:
            return _collect(self._polys())
        return _collect(_overlay([self._rings(), other._rings()], _pred_union))

    def difference(self, other):
        if self.is_empty:
            return Polygon()
        if other.is_empty:
            return _collect(self._polys())
        return _collect(_overlay([self._rings(), other._rings()], _pred_difference))

    def buffer(self, distance, quad_segs=_QUAD_SEGS, join_style=1, **_kw):
        return _buffer_polygonal(self, float(distance), join_style)

    # ---- predicates ---------------------------------------------------------------------------------------------------
    def intersects(self, other):
        return _intersects(self, other)

    def within(self, other):
        return _within(self, other)

    @property
    def centroid(self):
        polys = self._polys()
        A = sum(p.area for p in polys)
        if A <= 0.0:
            return Point()
        cx = sum(p._centroid_xy()[0] * p.area for p in polys) / A
        cy = sum(p._centroid_xy()[1] * p.area for p in polys) / A
        return Point(cx, cy)


class GeometryCollection(BaseGeometry):
    geom_type = "GeometryCollection"

    def __init__(self, geoms=None):
        self.geoms = list(geoms or [])

    def _rings(self):
        return [r for g in self.geoms for r in g._rings()]

    def _polys(self):
        return [p for g in self.geoms for p in g._polys()]

    @property
    def is_empty(self):
        return all(g.is_empty for g in self.geoms)

    @property
    def area(self):
        return sum(g.area for g in self.geoms)


class LinearRing:
    geom_type = "LinearRing"

    def __init__(self, coords):
        self._r = _as_ring(coords)

    @property
    def coords(self):
        c = _Coords(map(tuple, self._r))
        if len(c):
            c.append(c[0])
        return c

    @property
    def xy(self):
        return self.coords.xy

    @property
    def is_empty(self):
        return len(self._r) == 0

    def intersection(self, other):
        """Ring x LineString -> Point / MultiPoint / empty (spawn_locator.py:414-415)."""
        if not isinstance(other, LineString):
            raise NotImplementedError("LinearRing.intersection is only defined with a LineString here")
        pts = []
        if len(self._r) >= 2 and len(other._p) >= 2:
            a, b = _ring_segments(self._r)
            c, d = other._p[:-1], other._p[1:]
            for k in range(len(c)):
                r = b - a
                s = d[k] - c[k]
                den = r[:, 0] * s[1] - r[:, 1] * s[0]
                qp = c[k][None] - a
                with np.errstate(divide="ignore", invalid="ignore"):
                    t = (qp[:, 0] * s[1] - qp[:, 1] * s[0]) / den
                    u = (qp[:, 0] * r[:, 1] - qp[:, 1] * r[:, 0]) / den
                tol = 1e-12
                hit = (np.abs(den) > 1e-300) & (t >= -tol) & (t <= 1 + tol) & (u >= -tol) & (u <= 1 + tol)
                for tt, uu, i in sorted(zip(t[hit], u[hit], np.nonzero(hit)[0]), key=lambda z: z[1]):
                    x, y = a[i] + tt * r[i]
                    if not any(abs(x - px) <= EPS and abs(y - py) <= EPS for px, py in pts):
                        pts.append((float(x), float(y)))
        if not pts:
            return LineString()
        if len(pts) == 1:
            return Point(pts[0])
        return MultiPoint([Point(p) for p in pts])


class Polygon(BaseGeometry):
    geom_type = "Polygon"

    def __init__(self, shell=None, holes=None):
        self._shell = _as_ring(shell) if shell is not None and len(shell) else np.zeros((0, 2))
        self._holes = [_as_ring(h) for h in (holes or [])]

    def _rings(self):
        return ([self._shell] if len(self._shell) >= 3 else []) + [h for h in self._holes if len(h) >= 3]

    def _polys(self):
        return [] if self.is_empty else [self]

    @property
    def is_empty(self):
        return len(self._shell) < 3

    @property
    def exterior(self):
        return LinearRing(self._shell)

    @property
    def interiors(self):
        return [LinearRing(h) for h in self._holes]

    @property
    def area(self):
        return abs(_signed_area(self._shell)) - sum(abs(_signed_area(h)) for h in self._holes)

    def _centroid_xy(self):
        num, den = np.zeros(2), 0.0
        for r, sgn in [(self._shell, 1.0)] + [(h, -1.0) for h in self._holes]:
            x, y = r[:, 0], r[:, 1]
            xn, yn = np.roll(x, -1), np.roll(y, -1)
            cr = x * yn - xn * y
            a = 0.5 * np.sum(cr)
            if a == 0.0:
                continue
            c = np.array([np.sum((x + xn) * cr), np.sum((y + yn) * cr)]) / (6.0 * a)
            num += sgn * abs(a) * c
            den += sgn * abs(a)
        return num / den if den else np.array([np.nan, np.nan])

    @property
    def is_valid(self):
        """Simple-ring validity of the shell (the reference only asks it of 4-corner shadow polygons)."""
        r = self._shell
        n = len(r)
        if n < 3 or abs(_signed_area(r)) <= 1e-14:
            return False
        a, b = _ring_segments(r)
        for i in range(n):
            for j in range(i + 1, n):
                adjacent = (j == i + 1) or (i == 0 and j == n - 1)
                p, q, c, d = a[i], b[i], a[j], b[j]
                rr, ss = q - p, d - c
                den = rr[0] * ss[1] - rr[1] * ss[0]
                qp = c - p
                if abs(den) > 1e-300:
                    t = (qp[0] * ss[1] - qp[1] * ss[0]) / den
                    u = (qp[0] * rr[1] - qp[1] * rr[0]) / den
                    if adjacent:
                        continue
                    if -1e-12 <= t <= 1 + 1e-12 and -1e-12 <= u <= 1 + 1e-12:
                        return False
                else:                                   # parallel: overlapping collinear edges make the ring invalid
                    if abs(qp[0] * rr[1] - qp[1] * rr[0]) <= 1e-300 * max(1.0, np.hypot(*rr)):
                        l2 = float(rr @ rr)
                        t0, t1 = float(qp @ rr) / l2, float((d - p) @ rr) / l2
                        lo_, hi_ = min(t0, t1), max(t0, t1)
                        if adjacent:
                            if hi_ > 1e-12 and lo_ < 1 - 1e-12 and (hi_ - max(lo_, 0.0)) > 1e-12 and min(hi_, 1.0) - max(lo_, 0.0) > 1e-12:
                                return False
                        elif hi_ >= -1e-12 and lo_ <= 1 + 1e-12:
                            return False
        return True

    @property
    def minimum_rotated_rectangle(self):
        pts = _convex_hull(np.concatenate(self._rings())) if not self.is_empty else np.zeros((0, 2))
        if len(pts) < 3:
            return Polygon()
        best = None
        for i in range(len(pts)):
            e = pts[(i + 1) % len(pts)] - pts[i]
            ln = np.hypot(*e)
            if ln <= 0:
                continue
            ux = e / ln
            uy = np.array([-ux[1], ux[0]])
            px, py = pts @ ux, pts @ uy
            w, h = px.max() - px.min(), py.max() - py.min()
            if best is None or w * h < best[0]:
                c = [px.min() * ux + py.min() * uy, px.max() * ux + py.min() * uy,
                     px.max() * ux + py.max() * uy, px.min() * ux + py.max() * uy]
                best = (w * h, c)
        return Polygon(best[1])


class MultiPolygon(BaseGeometry):
    geom_type = "MultiPolygon"

    def __init__(self, polygons=None):
        self.geoms = [p for p in (polygons or []) if not p.is_empty]

    def _rings(self):
        return [r for p in self.geoms for r in p._rings()]

    def _polys(self):
        return list(self.geoms)

    @property
    def is_empty(self):
        return not self.geoms

    @property
    def area(self):
        return sum(p.area for p in self.geoms)


def _collect(polys):
    polys = [p for p in polys if not p.is_empty]
    if not polys:
        return Polygon()
    if len(polys) == 1:
        return polys[0]
    return MultiPolygon(polys)


class Point(BaseGeometry):
    geom_type = "Point"

    def __init__(self, *xy):
        if len(xy) == 0:
            self._p = None
        else:
            p = np.asarray(xy[0] if len(xy) == 1 else xy, dtype=np.float64).reshape(-1)
            self._p = np.array([p[0], p[1]])

    @property
    def is_empty(self):
        return self._p is None

    @property
    def x(self):
        return float(self._p[0])

    @property
    def y(self):
        return float(self._p[1])

    @property
    def coords(self):
        return _Coords([] if self._p is None else [(self.x, self.y)])

    def buffer(self, distance, quad_segs=_QUAD_SEGS, **_kw):
        if self._p is None or distance <= 0:
            return Polygon()
        n = 4 * quad_segs
        k = np.arange(n)
        ang = -2.0 * np.pi * k / n                      # GEOS: starts at angle 0 and runs clockwise
        return Polygon(np.stack([self._p[0] + distance * np.cos(ang), self._p[1] + distance * np.sin(ang)], axis=1))


class MultiPoint(BaseGeometry):
    geom_type = "MultiPoint"

    def __init__(self, points=None):
        self.geoms = list(points or [])

    @property
    def is_empty(self):
        return not self.geoms


class LineString(BaseGeometry):
    geom_type = "LineString"

    def __init__(self, coords=None):
        self._p = np.asarray(coords if coords is not None else [], dtype=np.float64).reshape(-1, 2)
        seg = np.diff(self._p, axis=0) if len(self._p) > 1 else np.zeros((0, 2))
        self._len = np.hypot(seg[:, 0], seg[:, 1])
        self._cum = np.concatenate(([0.0], np.cumsum(self._len)))

    @property
    def is_empty(self):
        return len(self._p) < 2

    @property
    def coords(self):
        return _Coords(map(tuple, self._p))

    @property
    def length(self):
        return float(self._cum[-1])

    def project(self, point):
        q = point._p
        seg = np.diff(self._p, axis=0)
        l2 = np.maximum(self._len ** 2, 1e-300)
        u = np.clip(((q - self._p[:-1]) * seg).sum(1) / l2, 0.0, 1.0)
        foot = self._p[:-1] + u[:, None] * seg
        j = int(np.argmin(np.hypot(*(q - foot).T)))
        return float(self._cum[j] + u[j] * self._len[j])

    def interpolate(self, s):
        s = min(max(float(s), 0.0), float(self._cum[-1]))
        j = int(np.clip(np.searchsorted(self._cum, s, side="right") - 1, 0, len(self._len) - 1))
        t = (s - self._cum[j]) / self._len[j] if self._len[j] > 0 else 0.0
        return Point(self._p[j] + t * (self._p[j + 1] - self._p[j]))

    def buffer(self, distance, quad_segs=_QUAD_SEGS, **_kw):
        if self.is_empty or distance <= 0:
            return Polygon()
        pieces = []
        for a, b in zip(self._p[:-1], self._p[1:]):
            e = b - a
            ln = np.hypot(*e)
            if ln <= 0:
                continue
            n = np.array([-e[1], e[0]]) / ln * distance
            pieces.append(Polygon([a + n, a - n, b - n, b + n]))
        for v in self._p:
            pieces.append(Point(v).buffer(distance, quad_segs))
        return unary_union(pieces)

    def intersection(self, other):
        """Part of the line inside a polygonal geometry: LineString / MultiLineString / empty LineString, pieces in the
        order and direction of this line (spawn_locator.py:521-533)."""
        rings = other._rings()
        if self.is_empty or not rings:
            return LineString()
        pieces, cur = [], []
        for a, b in zip(self._p[:-1], self._p[1:]):
            d = b - a
            ts = [0.0, 1.0]
            for r in rings:
                c, e = _ring_segments(r)
                s = e - c
                den = d[0] * s[:, 1] - d[1] * s[:, 0]
                qp = c - a[None]
                with np.errstate(divide="ignore", invalid="ignore"):
                    t = (qp[:, 0] * s[:, 1] - qp[:, 1] * s[:, 0]) / den
                    u = (qp[:, 0] * d[1] - qp[:, 1] * d[0]) / den
                hit = (np.abs(den) > 1e-300) & (t > 0) & (t < 1) & (u >= -1e-12) & (u <= 1 + 1e-12)
                ts.extend(t[hit].tolist())
            ts = sorted(set(ts))
            for t0, t1 in zip(ts[:-1], ts[1:]):
                if t1 - t0 <= 1e-14:
                    continue
                mid = a + 0.5 * (t0 + t1) * d
                if _points_in_rings(mid[None], rings)[0]:
                    p0, p1 = a + t0 * d, a + t1 * d
                    if cur and np.hypot(*(cur[-1] - p0)) <= EPS:
                        cur.append(p1)
                    else:
                        if len(cur) >= 2:
                            pieces.append(cur)
                        cur = [p0, p1]
                else:
                    if len(cur) >= 2:
                        pieces.append(cur)
                    cur = []
        if len(cur) >= 2:
            pieces.append(cur)
        if not pieces:
            return LineString()
        if len(pieces) == 1:
            return LineString(pieces[0])
        return MultiLineString([LineString(p) for p in pieces])


class MultiLineString(BaseGeometry):
    geom_type = "MultiLineString"

    def __init__(self, lines=None):
        self.geoms = list(lines or [])

    @property
    def is_empty(self):
        return not self.geoms


# =====================================================================================================================
# free functions
# =====================================================================================================================
def unary_union(geoms):
    geoms = [g for g in geoms if g is not None and not g.is_empty]
    if not geoms:
        return Polygon()
    if len(geoms) == 1:
        return _collect(geoms[0]._polys())
    return _collect(_overlay([g._rings() for g in geoms], _pred_union))


def _convex_hull(pts):
    pts = np.unique(np.asarray(pts, dtype=np.float64).reshape(-1, 2), axis=0)
    if len(pts) < 3:
        return pts
    pts = pts[np.lexsort((pts[:, 1], pts[:, 0]))]

    def half(seq):
        h = []
        for p in seq:
            while len(h) >= 2 and ((h[-1][0] - h[-2][0]) * (p[1] - h[-2][1]) - (h[-1][1] - h[-2][1]) * (p[0] - h[-2][0])) <= 0:
                h.pop()
            h.append(p)
        return h
    lower, upper = half(pts), half(pts[::-1])
    return np.asarray(lower[:-1] + upper[:-1])


def _arc(center, a0, a1, radius, quad_segs):
    """Points of the arc from angle a0 counter-clockwise to a1 (a1 > a0), GEOS step pi / (2 quad_segs)."""
    step = 0.5 * np.pi / quad_segs
    n = max(1, int(math.ceil((a1 - a0) / step - 1e-9)))
    ang = a0 + (a1 - a0) * np.arange(n + 1) / n
    return np.stack([center[0] + radius * np.cos(ang), center[1] + radius * np.sin(ang)], axis=1)


def _buffer_polygonal(geom, distance, join_style):
    polys = geom._polys()
    if not polys:
        return Polygon()
    if distance == 0.0:
        return unary_union(polys)
    if distance < 0.0:
        raise NotImplementedError("negative buffers are not used by the reference")
    pieces = list(polys)
    for poly in polys:
        for ring in poly._rings():
            # orient so that the region is on the left: shell ccw, holes cw (their outside = polygon interior... the
            # offset goes to the RIGHT of the direction of travel, i.e. away from the region)
            is_shell = ring is poly._shell
            r = ring
            if (_signed_area(r) > 0) != is_shell:
                r = r[::-1]
            n = len(r)
            nxt = np.roll(r, -1, axis=0)
            e = nxt - r
            ln = np.maximum(np.hypot(e[:, 0], e[:, 1]), 1e-300)
            out = np.stack([e[:, 1] / ln, -e[:, 0] / ln], axis=1)          # outward normal (right of travel)
            for i in range(n):
                a, b, nn = r[i], nxt[i], out[i] * distance
                pieces.append(Polygon([a, b, b + nn, a + nn]))
            for i in range(n):                                              # joins at vertex i+1 between edge i and i+1
                v = nxt[i]
                n0, n1 = out[i], out[(i + 1) % n]
                crs = n0[0] * n1[1] - n0[1] * n1[0]
                if crs <= 1e-14:                                            # reflex or straight: nothing to fill
                    continue
                if join_style == 2:
                    bis = n0 + n1
                    bl = np.hypot(*bis)
                    cosh = bl / 2.0
                    mitre = distance / max(cosh, 1e-300)
                    if mitre <= _MITRE_LIMIT * distance:
                        tip = v + bis / bl * mitre
                        pieces.append(Polygon([v, v + n0 * distance, tip, v + n1 * distance]))
                    else:                                                   # limited mitre: cut at the limit distance
                        lim = _MITRE_LIMIT * distance
                        dirb = bis / bl
                        d0 = np.array([-n0[1], n0[0]])                      # direction of travel on the two edges
                        d1 = np.array([-n1[1], n1[0]])
                        p0, p1 = v + n0 * distance, v + n1 * distance
                        t0 = (lim - distance * float(n0 @ dirb)) / max(float(d0 @ dirb), 1e-300)
                        t1 = (lim - distance * float(n1 @ dirb)) / max(float(-d1 @ dirb), 1e-300)
                        pieces.append(Polygon([v, p0, p0 + t0 * d0, p1 - t1 * d1, p1]))
                else:
                    a0 = math.atan2(n0[1], n0[0])
                    a1 = math.atan2(n1[1], n1[0])
                    if a1 < a0:
                        a1 += 2 * np.pi
                    arc = _arc(v, a0, a1, distance, _QUAD_SEGS)
                    pieces.append(Polygon(np.concatenate([v[None], arc])))
    return unary_union(pieces)


def _segments_touch(ringsA, ringsB, openA=False):
    """True when any boundary segment of A meets any boundary segment of B (closed segments)."""
    for ra in ringsA:
        if openA:
            a0, a1 = ra[:-1], ra[1:]
        else:
            a0, a1 = _ring_segments(ra)
        if not len(a0):
            continue
        for rb in ringsB:
            b0, b1 = _ring_segments(rb)
            r = (a1 - a0)[:, None, :]
            s = (b1 - b0)[None, :, :]
            qp = b0[None] - a0[:, None]
            den = r[..., 0] * s[..., 1] - r[..., 1] * s[..., 0]
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (qp[..., 0] * s[..., 1] - qp[..., 1] * s[..., 0]) / den
                u = (qp[..., 0] * r[..., 1] - qp[..., 1] * r[..., 0]) / den
            tol = 1e-12
            if np.any((np.abs(den) > 1e-300) & (t >= -tol) & (t <= 1 + tol) & (u >= -tol) & (u <= 1 + tol)):
                return True
            # parallel pairs: collinear overlap
            par = np.abs(den) <= 1e-300
            if np.any(par):
                ia, ib = np.nonzero(par)
                pts = np.concatenate([a0[ia], a1[ia]])
                d, _ = _dist_points_segments(pts, b0[np.concatenate([ib, ib])], b1[np.concatenate([ib, ib])])
                if np.any(np.diagonal(d) <= EPS):
                    return True
    return False


def _intersects(a, b):
    if a.is_empty or b.is_empty:
        return False
    if isinstance(a, Point):
        a, b = b, a
    if isinstance(b, Point):
        if isinstance(a, Point):
            return bool(np.all(np.abs(a._p - b._p) <= EPS))
        if isinstance(a, LineString):
            d, _ = _dist_points_segments(b._p[None], a._p[:-1], a._p[1:])
            return bool(np.any(d <= EPS))
        rings = a._rings()
        return bool(_points_in_rings(b._p[None], rings)[0] or _on_boundary(b._p[None], rings)[0])
    if isinstance(a, LineString) or isinstance(b, LineString):
        if isinstance(b, LineString) and not isinstance(a, LineString):
            a, b = b, a
        if isinstance(b, LineString):
            return _segments_touch([a._p], [np.concatenate([b._p, b._p[::-1]])[:-1]], openA=True)
        rings = b._rings()
        if _segments_touch([a._p], rings, openA=True):
            return True
        return bool(_points_in_rings(a._p[:1], rings)[0])
    ra, rb = a._rings(), b._rings()
    if _segments_touch(ra, rb):
        return True
    # no boundary contact: one may contain the other
    return bool(_points_in_rings(ra[0][:1], rb)[0] or _points_in_rings(rb[0][:1], ra)[0])


def _within(a, b):
    """a within b: no point of a outside b and the interiors meet."""
    if a.is_empty or b.is_empty:
        return False
    rb = b._rings()
    if isinstance(a, Point):
        return bool(_points_in_rings(a._p[None], rb)[0] and not _on_boundary(a._p[None], rb)[0])
    if isinstance(a, LineString):
        inside = _points_in_rings(a._p, rb) | _on_boundary(a._p, rb)
        if not inside.all():
            return False
        out = a.intersection(b)
        ln = sum(g.length for g in (out.geoms if isinstance(out, MultiLineString) else [out]) if not g.is_empty)
        return abs(ln - a.length) <= 1e-9 * max(1.0, a.length)
    diff = _overlay([a._rings(), rb], _pred_difference)
    return sum(p.area for p in diff) <= 1e-12 * max(1.0, a.area)
