"""Property tests of the float64 polygon stand-in (``oracle/polygon.py``) that lets the reference's ``sensor_model.py`` and
``spawn_locator.py`` run without shapely.  GEOS is not available to compare with, so the overlay is pinned by what a
correct point-set operation must satisfy: membership of random sample points, closed-form areas, inclusion-exclusion."""
import numpy as np

from oracle import polygon as G


def _star(rng, c, r0, n):
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = rng.uniform(0.4, 1.0, n) * r0
    return G.Polygon(np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1))


def test_overlay_membership_on_random_polygons():
    rng = np.random.default_rng(3)
    for _ in range(25):
        a, b, c = (_star(rng, rng.uniform(-1, 1, 2), r, rng.integers(3, 40)) for r in (4, 4, 3))
        ina, inb, inc = (lambda p, g=g: G._points_in_rings(p, g._rings()) for g in (a, b, c))
        cases = [(a.intersection(b), lambda p: ina(p) & inb(p)), (a.union(b), lambda p: ina(p) | inb(p)),
                 (a.difference(b), lambda p: ina(p) & ~inb(p)), (a.difference(b).difference(c), lambda p: ina(p) & ~inb(p) & ~inc(p)),
                 (G.unary_union([a, b, c]), lambda p: ina(p) | inb(p) | inc(p))]
        pts = rng.uniform(-6, 6, (6000, 2))
        for res, pred in cases:
            bad = G._points_in_rings(pts, res._rings()) != pred(pts)
            if bad.any():
                assert G._on_boundary(pts[bad], a._rings() + b._rings() + c._rings(), tol=1e-6).all()
        # inclusion-exclusion
        assert abs(a.union(b).area + a.intersection(b).area - a.area - b.area) < 1e-9


def test_closed_forms_and_degenerate_contacts():
    sq = G.Polygon([(0, 0), (4, 0), (4, 4), (0, 4)])
    assert abs(sq.intersection(G.Polygon([(2, 2), (6, 2), (6, 6), (2, 6)])).area - 4.0) < 1e-12
    hole = sq.difference(G.Polygon([(1, 1), (2, 1), (2, 2), (1, 2)]))
    assert abs(hole.area - 15.0) < 1e-12 and len(hole.interiors) == 1
    assert not G.Point(1.5, 1.5).within(hole) and G.Point(3, 3).within(hole) and not G.Point(0, 1).within(sq)
    # shared edges (every shadow quad of the sensor model shares an edge with the area it is cut from) and a grid of tiles
    q = G.Polygon([(4, 0), (4, 4), (8, 4), (8, 0)])
    assert abs(sq.union(q).area - 32.0) < 1e-12 and abs(sq.difference(q).area - 16.0) < 1e-12
    tiles = [G.Polygon([(i, j), (i + 1, j), (i + 1, j + 1), (i, j + 1)]) for i in range(6) for j in range(6)]
    u = G.unary_union(tiles)
    assert u.geom_type == "Polygon" and abs(u.area - 36.0) < 1e-12
    # corner contact: two polygons, not one ring through the touching vertex
    two = sq.union(G.Polygon([(4, 4), (6, 4), (6, 6), (4, 6)]))
    assert two.geom_type == "MultiPolygon" and len(two.geoms) == 2
    assert sq.intersects(G.Polygon([(4, 4), (6, 4), (6, 6), (4, 6)])) and not sq.intersects(G.Polygon([(5, 5), (6, 5), (6, 6)]))
    # nearly parallel neighbours a few micrometres apart (offset quads of an almost straight boundary)
    v = np.array([[0.0, 0.0], [1.5, 0.0], [1.53, 0.00005], [3.0, 0.003], [3.0, 2.0], [0.0, 2.0]])
    p = G.Polygon(v)
    peri = float(np.hypot(*(np.roll(v, -1, axis=0) - v).T).sum())
    assert abs(p.buffer(0.01, join_style=2).area - (p.area + 0.01 * peri)) < 1e-3      # + four mitred corners of 1e-4 each


def test_buffers_validity_lines_and_rectangles():
    sq = G.Polygon([(0, 0), (4, 0), (4, 4), (0, 4)])
    disc = G.Point(0, 0).buffer(1.0)
    assert len(disc.exterior.coords) == 65 and abs(disc.area - 0.5 * 64 * np.sin(2 * np.pi / 64)) < 1e-12
    assert abs(sq.buffer(1.0).area - (16 + 16 + disc.area)) < 1e-9          # rounded corners = one 64-gon in total
    assert abs(sq.buffer(0.5, join_style=2).area - 25.0) < 1e-9            # mitred corners
    ls = G.LineString([(-1, 2), (10, 2)])
    assert np.allclose(ls.intersection(sq).coords, [(0, 2), (4, 2)]) and ls.intersects(sq)
    parts = G.LineString([(-1, 1), (10, 1)]).intersection(G.MultiPolygon([sq, G.Polygon([(6, 0), (8, 0), (8, 4), (6, 4)])]))
    assert parts.geom_type == "MultiLineString" and np.allclose(parts.geoms[-1].coords[0], (6, 1))
    assert G.LineString([(5, 5), (6, 6)]).intersection(sq).is_empty
    assert abs(ls.buffer(0.5).area - (11 + 0.25 * disc.area)) < 1e-9
    ring = sq.buffer(1.0).exterior.intersection(G.LineString([(2, -5), (2, 10)]))
    assert ring.geom_type == "MultiPoint" and sorted((round(g.x, 9), round(g.y, 9)) for g in ring.geoms) == [(2.0, -1.0), (2.0, 5.0)]
    assert not G.Polygon([(0, 0), (1, 1), (1, 0), (0, 1)]).is_valid and sq.is_valid
    assert not G.Polygon([(0, 0), (1, 0), (2, 0), (3, 0)]).is_valid       # collinear shadow quad (ego on the edge's line)
    assert abs(G.Polygon([(0, 0), (2, 1), (1, 3), (-1, 2)]).minimum_rotated_rectangle.area - 5.0) < 1e-12
    c = sq.difference(G.Polygon([(1, 1), (2, 1), (2, 2), (1, 2)])).centroid
    assert abs(c.x - (16 * 2 - 1.5) / 15) < 1e-12 and abs(c.y - (16 * 2 - 1.5) / 15) < 1e-12
    assert G.Point(2, 2).buffer(0.15).within(sq) and not G.Point(3.9, 2).buffer(0.15).within(sq)
