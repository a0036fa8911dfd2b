"""Oracle A, part 1: stand-ins for the reference's un-installable third-party *leaf* symbols.

TEST INFRASTRUCTURE.  ``install()`` injects minimal stand-ins into ``sys.modules`` for exactly the
symbols the reference's ``frenetix_occlusion/metrics`` package imports, so that
package can then be imported **unmodified** from ``/root/reference`` and its own control flow
(`metrics/metric.py`, `cp.py`, `dce.py`, `ttc.py`, `ttce.py`, `wttc.py`, `hr.py`, `be.py`,
`metrics/utils/*.py`) executed as shipped.  What is restated here is only third-party behaviour:

=====================================================  =============================================
reference import (file:line)                           stand-in semantics (library, pinned version)
=====================================================  =============================================
``commonroad.scenario.obstacle.ObstacleType``          enum with commonroad-io 2023.2's string values
  (harm_model.py:10, convert_dynamic_obstacle.py:13)
``commonroad.scenario.obstacle.DynamicObstacle``       ``occupancy_at_time(t)``: initial-state occupancy at
  (convert_dynamic_obstacle.py:13)                     its own time step, else the prediction's occupancy,
                                                       else ``None`` (commonroad-io 2023.2)
``commonroad.geometry.shape.Rectangle``                length/width/center/orientation, ``shapely_object``
  (convert_dynamic_obstacle.py:10)                     (convex-polygon stand-in), ``rotate_translate_local``
``commonroad.prediction.prediction.TrajectoryPrediction``  occupancy per trajectory state =
  (convert_dynamic_obstacle.py:11)                     ``shape.rotate_translate_local(position, orientation)``
``commonroad.scenario.trajectory.Trajectory``          ``initial_time_step``, ``state_list``
``commonroad.scenario.state.CustomState``              attribute bag + ``translate_rotate(t, angle)``
``commonroad_dc.pycrcc.RectOBB``                       ``center() / r_x() / local_x_axis()`` only
  (collision_probability.py:10,149-156)                (commonroad-drivability-checker 2023.1)
``scipy.stats.mvn.mvnun``                              rectangle probability of a bivariate normal
  (collision_probability.py:11,117)                    (scipy 1.12.0 Fortran ``mvndst``; deterministic BVN).
                                                       Diagonal covariance -> exact Phi product via
                                                       ``scipy.special.ndtr``; otherwise
                                                       ``multivariate_normal.cdf(..., lower_limit=)``.
shapely ``Polygon.distance`` / ``.intersects``         ``oracle.geometry.ConvexPolygon`` (shapely 2.0.2/GEOS)
  (dce.py:75-79, be.py:181)
=====================================================  =============================================
"""
from __future__ import annotations

import enum
import sys
import types

import numpy as np

from .geometry import ConvexPolygon

REFERENCE_ROOT = "/root/reference"


# ---------------------------------------------------------------------------- commonroad-io
class ObstacleType(enum.Enum):
    UNKNOWN = "unknown"
    CAR = "car"
    TRUCK = "truck"
    BUS = "bus"
    BICYCLE = "bicycle"
    PEDESTRIAN = "pedestrian"
    PRIORITY_VEHICLE = "priorityVehicle"
    PARKED_VEHICLE = "parkedVehicle"
    CONSTRUCTION_ZONE = "constructionZone"
    TRAIN = "train"
    ROAD_BOUNDARY = "roadBoundary"
    MOTORCYCLE = "motorcycle"
    TAXI = "taxi"
    BUILDING = "building"
    PILLAR = "pillar"
    MEDIAN_STRIP = "median_strip"


class Rectangle:
    def __init__(self, length, width, center=None, orientation=0.0):
        self.length = float(length)
        self.width = float(width)
        self.center = np.array([0.0, 0.0]) if center is None else np.asarray(center, dtype=np.float64)
        self.orientation = float(orientation)
        self._poly = None

    @property
    def vertices(self):
        hl, hw = 0.5 * self.length, 0.5 * self.width
        loc = np.array([[-hl, -hw], [-hl, hw], [hl, hw], [hl, -hw], [-hl, -hw]])
        c, s = np.cos(self.orientation), np.sin(self.orientation)
        rot = np.array([[c, -s], [s, c]])
        return loc @ rot.T + self.center

    @property
    def shapely_object(self):
        if self._poly is None:
            self._poly = ConvexPolygon(self.vertices)
        return self._poly

    def rotate_translate_local(self, translation, angle):
        return Rectangle(self.length, self.width, self.center + np.asarray(translation, dtype=np.float64),
                         self.orientation + float(angle))


class CustomState:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def translate_rotate(self, translation, angle):
        kw = dict(self.__dict__)
        p = np.asarray(kw["position"], dtype=np.float64) + np.asarray(translation, dtype=np.float64)
        c, s = np.cos(angle), np.sin(angle)
        kw["position"] = np.array([c * p[0] - s * p[1], s * p[0] + c * p[1]])
        kw["orientation"] = kw["orientation"] + angle
        return CustomState(**kw)


InitialState = CustomState


class Trajectory:
    def __init__(self, initial_time_step, state_list):
        self.initial_time_step = initial_time_step
        self.state_list = state_list


class Occupancy:
    def __init__(self, time_step, shape):
        self.time_step = time_step
        self.shape = shape


class TrajectoryPrediction:
    def __init__(self, trajectory, shape):
        self.trajectory = trajectory
        self.shape = shape
        self._occ = {}
        for k, st in enumerate(trajectory.state_list):
            t = trajectory.initial_time_step + k
            self._occ[t] = Occupancy(t, shape.rotate_translate_local(st.position, st.orientation))

    def occupancy_at_time_step(self, time_step):
        return self._occ.get(time_step)


class DynamicObstacle:
    def __init__(self, obstacle_id, obstacle_type, obstacle_shape, initial_state, prediction=None):
        self.obstacle_id = obstacle_id
        self.obstacle_type = obstacle_type
        self.obstacle_shape = obstacle_shape
        self.initial_state = initial_state
        self.prediction = prediction
        self._initial_occ = Occupancy(initial_state.time_step,
                                      obstacle_shape.rotate_translate_local(initial_state.position,
                                                                            initial_state.orientation))

    def occupancy_at_time(self, time_step):
        if time_step == self.initial_state.time_step:
            return self._initial_occ
        if time_step > self.initial_state.time_step and self.prediction is not None:
            return self.prediction.occupancy_at_time_step(time_step)
        return None


# ------------------------------------------------------------------- commonroad-drivability-checker
class RectOBB:
    def __init__(self, r_x, r_y, orientation, x, y):
        self._rx, self._ry, self._o, self._x, self._y = r_x, r_y, orientation, x, y

    def center(self):
        return np.array([self._x, self._y], dtype=np.float64)

    def r_x(self):
        return self._rx

    def r_y(self):
        return self._ry

    def local_x_axis(self):
        return np.array([np.cos(self._o), np.sin(self._o)], dtype=np.float64)


# ------------------------------------------------------------------------------- scipy.stats.mvn
def mvnun(lower, upper, means, covar):
    """Probability mass of N(means, covar) over the axis-aligned box [lower, upper] (2-D)."""
    from scipy.special import ndtr
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    means = np.asarray(means, dtype=np.float64)
    cov = np.asarray(covar, dtype=np.float64)
    if cov[0, 1] == 0.0 and cov[1, 0] == 0.0:
        sx, sy = np.sqrt(cov[0, 0]), np.sqrt(cov[1, 1])
        px = ndtr((upper[0] - means[0]) / sx) - ndtr((lower[0] - means[0]) / sx)
        py = ndtr((upper[1] - means[1]) / sy) - ndtr((lower[1] - means[1]) / sy)
        return float(px * py), 0
    from scipy.stats import multivariate_normal
    p = multivariate_normal(mean=means, cov=cov, allow_singular=False).cdf(upper, lower_limit=lower)
    return float(p), 0


# ------------------------------------------------------------------------------------ installation
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install(reference_root: str = REFERENCE_ROOT):
    """Register the stand-ins and put the reference tree on ``sys.path`` (idempotent)."""
    global _installed
    if _installed:
        return
    import os
    if not os.path.isdir(os.path.join(reference_root, "frenetix_occlusion", "metrics")):
        raise FileNotFoundError(f"reference tree not found at {reference_root} (oracle A only runs in the "
                                "build container; use tests/golden fixtures elsewhere)")
    cr = _module("commonroad")
    cr.scenario = _module("commonroad.scenario")
    cr.geometry = _module("commonroad.geometry")
    cr.prediction = _module("commonroad.prediction")
    cr.scenario.obstacle = _module("commonroad.scenario.obstacle", ObstacleType=ObstacleType,
                                   DynamicObstacle=DynamicObstacle)
    cr.scenario.state = _module("commonroad.scenario.state", CustomState=CustomState, InitialState=InitialState)
    cr.scenario.trajectory = _module("commonroad.scenario.trajectory", Trajectory=Trajectory)
    cr.geometry.shape = _module("commonroad.geometry.shape", Rectangle=Rectangle)
    cr.prediction.prediction = _module("commonroad.prediction.prediction", TrajectoryPrediction=TrajectoryPrediction,
                                       Occupancy=Occupancy)
    dc = _module("commonroad_dc")
    dc.pycrcc = _module("commonroad_dc.pycrcc", RectOBB=RectOBB)
    import scipy.stats
    mvn = _module("scipy.stats.mvn", mvnun=mvnun)
    scipy.stats.mvn = mvn
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    _installed = True
