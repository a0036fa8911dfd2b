"""Oracle A for the host-side leaves of stages 1-2: the reference's OWN ``agent.py`` (FOAgentManager,
OAPPedestrianAgent), ``utils/fo_obstacle.py`` (FOObstacle) and ``utils/helper_functions.py`` run **verbatim** from
``/root/reference`` over stand-ins for the third-party symbols they import.  TEST INFRASTRUCTURE; only usable in the
build container (the reference tree is absent on the GPU box) -- it generates ``tests/golden/stage12_reference.json``.

Stand-ins added on top of ``ref_shims.install()`` (reference import -> semantics):
* ``shapely.geometry.{Point, Polygon, MultiPolygon, LineString}`` (fo_obstacle.py:9-10, helper_functions.py:9):
  plain coordinate holders; ``LineString.project`` / ``.interpolate`` = arc-length projection on the polyline
  (shapely 2.0.2 semantics for a 2-D LineString), used by ``hf.calc_normal_vector_to_curve`` (helper_functions.py:38-57);
* ``shapely.affinity.{rotate, translate}``, ``shapely.ops.unary_union``: importable placeholders (not reached here);
* ``commonroad_route_planner.*``, ``commonroad_dc.pycrccosy``, ``frenetix.*``: importable placeholders so that
  ``route_planner.py`` / ``utils/frenetix_handler.py`` import; vehicle agents (which need them) are NOT run.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from . import ref_shims


class Point:
    def __init__(self, *xy):
        p = np.asarray(xy[0] if len(xy) == 1 else xy, dtype=np.float64).reshape(-1)
        self.x, self.y = float(p[0]), float(p[1])
        self.coords = [(self.x, self.y)]


class Polygon:
    def __init__(self, shell=None, holes=None):
        self.exterior_coords = np.asarray(shell if shell is not None else [], dtype=np.float64).reshape(-1, 2)


class MultiPolygon:
    def __init__(self, polygons=None):
        self.geoms = list(polygons or [])


class LineString:
    def __init__(self, coords):
        self._p = np.asarray(coords, dtype=np.float64).reshape(-1, 2)
        seg = np.diff(self._p, axis=0)
        self._len = np.hypot(seg[:, 0], seg[:, 1])
        self._cum = np.concatenate(([0.0], np.cumsum(self._len)))
        self.coords = [tuple(r) for r in self._p]

    def project(self, point):
        q = np.array([point.x, point.y])
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


_done = False


def install():
    global _done
    ref_shims.install()
    if _done:
        return
    m = ref_shims._module
    sh = m("shapely")
    sh.geometry = m("shapely.geometry", Point=Point, Polygon=Polygon, MultiPolygon=MultiPolygon, LineString=LineString)
    m("shapely.geometry.multipolygon", MultiPolygon=MultiPolygon)
    sh.affinity = m("shapely.affinity", rotate=lambda g, *a, **k: g, translate=lambda g, *a, **k: g)
    sh.ops = m("shapely.ops", unary_union=lambda gs: gs)
    crp = m("commonroad_route_planner")
    crp.route_planner = m("commonroad_route_planner.route_planner", Route=object)
    crp.route = m("commonroad_route_planner.route", RouteType=types.SimpleNamespace(REGULAR=0), Route=object)
    crp.utility = m("commonroad_route_planner.utility")
    crp.utility.route = m("commonroad_route_planner.utility.route", lanelet_orientation_at_position=lambda *a: 0.0)
    dc = sys.modules["commonroad_dc"]
    dc.pycrccosy = m("commonroad_dc.pycrccosy", CurvilinearCoordinateSystem=object)
    fx = m("frenetix")
    fx.trajectory_functions = m("frenetix.trajectory_functions")
    fx.trajectory_functions.feasability_functions = m("frenetix.trajectory_functions.feasability_functions")
    _done = True


def reference_modules():
    """(agent, fo_obstacle, helper_functions) modules of the reference, imported unmodified."""
    install()
    import frenetix_occlusion.agent as ref_agent
    import frenetix_occlusion.utils.fo_obstacle as ref_obst
    import frenetix_occlusion.utils.helper_functions as ref_hf
    return ref_agent, ref_obst, ref_hf
