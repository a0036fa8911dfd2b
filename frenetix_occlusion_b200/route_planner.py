"""Route exploration for vehicle phantoms (mirror of reference route_planner.py:14-93).

Depth-first walk over lanelet successors and same-direction neighbours, ``max_depth = 2``; one
reference polyline per route (the reference delegates this to commonroad_route_planner 2022.3
``Route.reference_path`` -- not installable, PARITY UNPINNED): centre lines of the route's lanelets chained without
the duplicated joint vertices, lanelets that are only crossed by a lane change skipped, resampled at a fixed 2 m
step and smoothed with four Chaikin corner-cutting refinements, as that library does.

The two graph-walk methods ``_find_all_routes`` / ``_explore_routes`` follow reference route_planner.py:54-90
statement by statement (same traversal order decides the order of the predictions); SURVEY.md 2 lists the route
planner as out of scope as code -- it is host glue that only produces the kernel's input polylines."""
from __future__ import annotations

from typing import List

import numpy as np


def lanelet_orientation_at_position(lanelet, pos):
    """commonroad_route_planner.utility.route.lanelet_orientation_at_position (2022.3): heading of the centre-line
    segment that STARTS at the centre vertex closest to ``pos`` (the last vertex is not a candidate)."""
    c = np.asarray(lanelet.center_vertices, dtype=np.float64)
    k = int(np.argmin(np.hypot(*(np.asarray(pos, dtype=np.float64) - c[:-1]).T)))
    return float(np.arctan2(c[k + 1, 1] - c[k, 1], c[k + 1, 0] - c[k, 0]))


def resample_polyline(pts, step=1.0):
    pts = np.asarray(pts, dtype=np.float64)
    keep = np.concatenate(([True], np.hypot(*np.diff(pts, axis=0).T) > 1e-6))
    pts = pts[keep]
    cum = np.concatenate(([0.0], np.cumsum(np.hypot(*np.diff(pts, axis=0).T))))
    n = max(2, int(np.ceil(cum[-1] / step)) + 1)
    s = np.linspace(0.0, cum[-1], n)
    return np.stack((np.interp(s, cum, pts[:, 0]), np.interp(s, cum, pts[:, 1])), -1)


def resample_fixed_step(pts, step=2.0):
    """Points at arc lengths 0, step, 2 step, ... plus the last point (commonroad_route_planner ``resample_polyline``)."""
    pts = np.asarray(pts, dtype=np.float64)
    cum = np.concatenate(([0.0], np.cumsum(np.hypot(*np.diff(pts, axis=0).T))))
    n = max(int(np.floor(cum[-1] / step)), 1)
    s = np.arange(n + 1) * step
    if cum[-1] - s[-1] > 1e-9:
        s = np.concatenate((s, [cum[-1]]))
    return np.stack((np.interp(s, cum, pts[:, 0]), np.interp(s, cum, pts[:, 1])), -1)


def chaikins_corner_cutting(pts, refinements=4):
    """Each refinement replaces every edge by its 1/4 and 3/4 points; the end points stay."""
    p = np.asarray(pts, dtype=np.float64)
    for _ in range(refinements):
        q = np.empty((2 * (len(p) - 1), 2))
        q[0::2] = 0.75 * p[:-1] + 0.25 * p[1:]
        q[1::2] = 0.25 * p[:-1] + 0.75 * p[1:]
        p = np.concatenate((p[:1], q, p[-1:]))
    return p


def route_reference_path(lanelet_network, route):
    """``Route.reference_path`` of commonroad_route_planner for a list of lanelet ids."""
    parts = []
    for a, b in zip(route, list(route[1:]) + [None]):
        la = lanelet_network.find_lanelet_by_id(a)
        if b is not None and b not in la.successor:
            continue                                   # lane change: the neighbour's centre line takes over
        c = np.asarray(la.center_vertices, dtype=np.float64)
        if parts and np.allclose(parts[-1][-1], c[0]):
            c = c[1:]
        parts.append(c)
    return chaikins_corner_cutting(resample_fixed_step(np.concatenate(parts), 2.0), 4)


def path_heading_variance(path) -> float:
    """Variance of the (unwrapped) segment headings of a polyline: how straight a route's reference path goes."""
    d = np.diff(np.asarray(path, dtype=np.float64), axis=0)
    return float(np.var(np.unwrap(np.arctan2(d[:, 1], d[:, 0]))))


class FORoutePlanner:
    def __init__(self, scenario, lanelet_network, visualization=None, debug=False):
        self.cr_scenario = scenario
        self.lanelet_network = lanelet_network
        self.debug = debug
        self.visualization = visualization
        self.lanelet_orientation = None
        self.start_lanelet = None
        self.route_candidates = None
        self.reference_paths = None

    def calc_possible_reference_paths(self, pos) -> List[np.ndarray]:
        ids = self.cr_scenario.lanelet_network.find_lanelet_by_position([pos])[0]
        if not ids:
            raise ValueError("[OAP - Route Planner] position is not on a lanelet")
        start = self.cr_scenario.lanelet_network.find_lanelet_by_id(ids[0])
        self.start_lanelet = start
        self.lanelet_orientation = lanelet_orientation_at_position(start, pos)
        # routes and their reference paths depend on the start lanelet only and the lanelet network is static: computed
        # once per start lanelet and network (the reference re-plans them for every phantom agent of every cycle)
        cache = self.lanelet_network.__dict__.setdefault("_fo_route_cache", {})
        hit = cache.get(ids[0])
        if hit is None:
            routes = [r for r in self._find_all_routes(ids[0], max_depth=2) if r]
            paths = [route_reference_path(self.lanelet_network, route) for route in routes]
            for q in paths:
                q.setflags(write=False)
            hit = cache[ids[0]] = (routes, paths, [path_heading_variance(q) for q in paths])
        self.route_candidates = [list(r) for r in hit[0]]
        self.reference_paths = list(hit[1])
        self.heading_variances = list(hit[2])
        return self.reference_paths

    def _find_all_routes(self, id_lanelet_start, max_depth=2):
        all_routes = []
        self._explore_routes(id_lanelet_start, [], all_routes, 0, max_depth)
        if not all_routes:
            raise ValueError("[OAP - Route Planner] Route Explorer could not find a Route")
        return all_routes

    def _explore_routes(self, id_lanelet_current, route, all_routes, depth, max_depth):
        """route_planner.py:61-90 (successors, then right / left neighbours driving the same way)."""
        lanelet = self.lanelet_network.find_lanelet_by_id(id_lanelet_current)
        route.append(lanelet.lanelet_id)
        successors = []
        if lanelet.successor:
            successors.extend(lanelet.successor)
        if lanelet.adj_right and lanelet.adj_right_same_direction:
            if self.lanelet_network.find_lanelet_by_id(lanelet.adj_right).successor:
                successors.append(lanelet.adj_right)
        if lanelet.adj_left and lanelet.adj_left_same_direction:
            if self.lanelet_network.find_lanelet_by_id(lanelet.adj_left).successor:
                successors.append(lanelet.adj_left)
        if depth >= max_depth:
            successors = []
        if not successors:
            all_routes.append(route.copy())
            return
        for successor in successors:
            self._explore_routes(successor, route, all_routes, depth + 1, max_depth)
            route.pop()
