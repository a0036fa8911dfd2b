"""Route exploration for vehicle phantoms (mirror of reference route_planner.py:14-93).

Depth-first walk over lanelet successors and same-direction neighbours, ``max_depth = 2``; one
reference polyline per route = concatenated lanelet centre lines (the reference delegates this to
commonroad_route_planner 2022.3 ``Route.reference_path`` -- not installable; PARITY UNPINNED for the
smoothing/resampling that library applies)."""
from __future__ import annotations

from typing import List

import numpy as np


def lanelet_orientation_at_position(lanelet, pos):
    """Heading of the centre-line segment closest to ``pos`` (commonroad_route_planner utility)."""
    c = np.asarray(lanelet.center_vertices, dtype=np.float64)
    seg = c[1:] - c[:-1]
    l2 = np.maximum((seg ** 2).sum(1), 1e-12)
    u = np.clip(((np.asarray(pos) - c[:-1]) * seg).sum(1) / l2, 0.0, 1.0)
    q = c[:-1] + u[:, None] * seg
    j = int(np.argmin(np.hypot(*(np.asarray(pos) - q).T)))
    return float(np.arctan2(seg[j, 1], seg[j, 0]))


def resample_polyline(pts, step=1.0):
    pts = np.asarray(pts, dtype=np.float64)
    keep = np.concatenate(([True], np.hypot(*np.diff(pts, axis=0).T) > 1e-6))
    pts = pts[keep]
    cum = np.concatenate(([0.0], np.cumsum(np.hypot(*np.diff(pts, axis=0).T))))
    n = max(2, int(np.ceil(cum[-1] / step)) + 1)
    s = np.linspace(0.0, cum[-1], n)
    return np.stack((np.interp(s, cum, pts[:, 0]), np.interp(s, cum, pts[:, 1])), -1)


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
        self.route_candidates = [r for r in self._find_all_routes(ids[0], max_depth=2) if r]
        self.reference_paths = []
        for route in self.route_candidates:
            pts = np.concatenate([self.lanelet_network.find_lanelet_by_id(i).center_vertices for i in route])
            self.reference_paths.append(resample_polyline(pts, 1.0))
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
