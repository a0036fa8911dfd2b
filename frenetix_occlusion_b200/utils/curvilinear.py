"""Polyline curvilinear coordinate system with the subset of the ``pycrccosy.CurvilinearCoordinateSystem``
interface the hot path calls (spawn_locator.py:229,375,385,398,449,536,549,554,653).

In deployment the planner hands its own ``cosy_cl`` (commonroad-drivability-checker 2023.1) to
``FOInterface.evaluate_scenario``; this class is what the scenario replay harness uses when that library is
absent.  PARITY UNPINNED: pycrccosy resamples the reference path and interpolates segment normals; here the
projection is onto the raw polyline with per-segment normals."""
from __future__ import annotations

import numpy as np


class CurvilinearCoordinateSystem:
    def __init__(self, reference_path, projection_domain_limit: float = 30.0, eps: float = 0.1):
        p = np.asarray(reference_path, dtype=np.float64).reshape(-1, 2)
        keep = np.concatenate(([True], np.hypot(*np.diff(p, axis=0).T) > 1e-9))
        self._p = p[keep]
        if len(self._p) < 2:
            raise ValueError("reference path needs at least two distinct points")
        self._seg = np.diff(self._p, axis=0)
        self._len = np.hypot(self._seg[:, 0], self._seg[:, 1])
        self._t = self._seg / self._len[:, None]
        self._n = np.stack((-self._t[:, 1], self._t[:, 0]), -1)
        self._s = np.concatenate(([0.0], np.cumsum(self._len)))
        self._limit = float(projection_domain_limit)
        self._eps = float(eps)

    def reference_path(self):
        return self._p.copy()

    def length(self):
        return float(self._s[-1])

    def convert_to_curvilinear_coords(self, x, y):
        q = np.array([float(x), float(y)])
        u = np.einsum("ij,ij->i", q - self._p[:-1], self._t)
        uc = np.clip(u, 0.0, self._len)
        foot = self._p[:-1] + uc[:, None] * self._t
        dist = np.hypot(*(q - foot).T)
        j = int(np.argmin(dist))
        if (j == 0 and u[0] < -self._eps) or (j == len(self._len) - 1 and u[-1] > self._len[-1] + self._eps) \
                or dist[j] > self._limit:
            raise ValueError("<CurvilinearCoordinateSystem/convert_to_curvilinear_coords> point outside of the "
                             "projection domain")
        side = float(self._t[j, 0] * (q[1] - self._p[j, 1]) - self._t[j, 1] * (q[0] - self._p[j, 0]))
        d = float(dist[j]) if side >= 0 else -float(dist[j])   # Euclidean distance to the foot, signed by the side
        return np.array([self._s[j] + uc[j], d])

    def convert_to_cartesian_coords(self, s, d):
        s = float(s)
        if s < -self._eps or s > self._s[-1] + self._eps:
            raise ValueError("<CurvilinearCoordinateSystem/convert_to_cartesian_coords> longitudinal coordinate "
                             "outside of the reference path")
        j = int(np.clip(np.searchsorted(self._s, s, side="right") - 1, 0, len(self._len) - 1))
        return self._p[j] + (s - self._s[j]) * self._t[j] + float(d) * self._n[j]

    def convert_array_to_cartesian_coords(self, s, d):
        """Vectorised ``convert_to_cartesian_coords`` (extension used by the spawn locator when available): rows whose
        longitudinal coordinate lies outside the path are NaN."""
        s = np.atleast_1d(np.asarray(s, dtype=np.float64))
        d = np.broadcast_to(np.asarray(d, dtype=np.float64), s.shape)
        j = np.clip(np.searchsorted(self._s, s, side="right") - 1, 0, len(self._len) - 1)
        out = self._p[j] + (s - self._s[j])[:, None] * self._t[j] + d[:, None] * self._n[j]
        out[(s < -self._eps) | (s > self._s[-1] + self._eps)] = np.nan
        return out

    def convert_list_of_points_to_curvilinear_coords(self, points, num_omp_threads=1):
        out = []
        for p in points:
            p = np.asarray(p, dtype=np.float64).reshape(-1)
            out.append(self.convert_to_curvilinear_coords(p[0], p[1]))
        return out

    def cartesian_point_inside_projection_domain(self, x, y) -> bool:
        try:
            self.convert_to_curvilinear_coords(x, y)
            return True
        except ValueError:
            return False
