"""Stand-ins for the WORLD the reference's per-cycle pipeline runs in: commonroad-io scenario objects, the curvilinear
coordinate system of commonroad-drivability-checker, the route of commonroad-route-planner and the C++ ``frenetix``
trajectory sampler -- everything ``interface.py``, ``sensor_model.py``, ``spawn_locator.py``, ``agent.py``,
``route_planner.py`` and ``utils/frenetix_handler.py`` import but this image cannot install.  TEST INFRASTRUCTURE
(``oracle/pipeline_oracle.py`` injects them and then runs those reference files **unmodified**).  Imports nothing
of the product.

=====================================================  ================================================================
reference import (file:line)                           stand-in semantics (library, pinned version) -- all PARITY UNPINNED
=====================================================  ================================================================
``scenario.lanelet_network`` (sensor_model.py:196,      commonroad-io 2023.2 object model read from the compact scene
spawn_locator.py:184-249, agent.py, route_planner.py)   dict of ``tests/golden/scene_*.json``: ``find_lanelet_by_position``
                                                        (closed point-in-polygon), ``find_lanelet_by_id``,
                                                        ``lanelet_polygons[i].shapely_object``, ``intersections``
``commonroad_dc.pycrccosy.CurvilinearCoordinate-``     projection on the nearest segment of the reference polyline,
``System`` (spawn_locator.py:229-653,                   s = arc length, d = signed offset (left positive); raises outside
frenetix_handler.py:33,77)                              a 20 m projection domain or beyond the path ends
``commonroad_dc.geometry.util.compute_pathlength_``    cumulative chord length / ``np.gradient`` curvature
``from_polyline, compute_curvature_from_polyline``      (commonroad-drivability-checker 2023.1)
``commonroad_route_planner...lanelet_orientation_``    direction of the centre-line segment at the closest centre vertex
``at_position`` (spawn_locator.py:14,662)               (commonroad-route-planner 2022.3)
``commonroad_route_planner.route_planner.Route``       ``reference_path``: centre lines of the route resampled at 2 m,
(route_planner.py:42-45)                                four Chaikin corner-cutting refinements (route planner 2022.3)
``frenetix.TrajectoryHandler / CoordinateSystem-``     quartic-in-s / quintic-in-d polynomial samples of the sampling
``Wrapper / trajectory_functions.FillCoordinates``      matrix (Werling et al. 2010), Cartesian map along the polyline;
(frenetix_handler.py:45-125)                            beyond t1 the end velocity is kept; feasibility checks recorded
                                                        but nothing is dropped; order = sampling-matrix order
=====================================================  ================================================================
"""
from __future__ import annotations

import enum
import types

import numpy as np

from . import polygon as G
from .ref_shims import ObstacleType, Rectangle


# ---------------------------------------------------------------------------------------------------------------------
# commonroad-io object model
# ---------------------------------------------------------------------------------------------------------------------
class ObstacleRole(enum.Enum):
    STATIC = "static"
    DYNAMIC = "dynamic"


class LaneletType(enum.Enum):
    URBAN = "urban"
    COUNTRY = "country"
    HIGHWAY = "highway"
    UNKNOWN = "unknown"
    INTERSECTION = "intersection"
    SIDEWALK = "sidewalk"
    CROSSWALK = "crosswalk"
    BICYCLE_LANE = "bicycleLane"
    BUS_LANE = "busLane"
    INTERSTATE = "interstate"
    DRIVE_WAY = "driveWay"
    MAIN_CARRIAGE_WAY = "mainCarriageWay"
    ACCESS_RAMP = "accessRamp"
    EXIT_RAMP = "exitRamp"
    SHOULDER = "shoulder"
    BUS_STOP = "busStop"
    BORDER = "border"
    PARKING = "parking"
    RESTRICTED = "restricted"


class State:
    def __init__(self, position, orientation, velocity, time_step, **kw):
        self.position = np.asarray(position, dtype=np.float64)
        self.orientation = float(orientation)
        self.velocity = float(velocity)
        self.time_step = int(time_step)
        self.__dict__.update(kw)


class _Polygon:
    def __init__(self, vertices):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.shapely_object = G.Polygon(self.vertices)


class Lanelet:
    def __init__(self, e):
        self.lanelet_id = int(e["id"])
        self.left_vertices = np.asarray(e["left"], dtype=np.float64)
        self.right_vertices = np.asarray(e["right"], dtype=np.float64)
        self.center_vertices = 0.5 * (self.left_vertices + self.right_vertices)
        self.predecessor, self.successor = list(e["pred"]), list(e["succ"])
        self.adj_left, self.adj_left_same_direction = e["adj_left"]
        self.adj_right, self.adj_right_same_direction = e["adj_right"]
        self.lanelet_type = {LaneletType(t) for t in e.get("type", [])}
        self.polygon = _Polygon(np.concatenate([self.right_vertices, self.left_vertices[::-1]]))


class LaneletNetwork:
    def __init__(self, lanelets, intersections):
        self.lanelets = lanelets
        self._by_id = {l.lanelet_id: l for l in lanelets}
        self.intersections = intersections

    @property
    def lanelet_polygons(self):
        return [l.polygon for l in self.lanelets]

    def find_lanelet_by_id(self, lanelet_id):
        return self._by_id[lanelet_id]

    def find_lanelet_by_position(self, point_list):
        out = []
        for p in point_list:
            p = np.asarray(p, dtype=np.float64).reshape(1, 2)
            ids = []
            for l in self.lanelets:
                ring = [l.polygon.shapely_object._shell]
                if G._points_in_rings(p, ring)[0] or G._on_boundary(p, ring)[0]:
                    ids.append(l.lanelet_id)
            out.append(ids)
        return out


class Obstacle:
    def __init__(self, o):
        self.obstacle_id = int(o["id"])
        self.obstacle_type = ObstacleType(o["type"])
        self.obstacle_role = ObstacleRole(o["role"])
        sh = o["shape"]
        self.obstacle_shape = Rectangle(sh[0], sh[1], center=np.array([sh[2], sh[3]]), orientation=sh[4])
        mk = lambda v: State(position=[v[0], v[1]], orientation=v[2], velocity=v[3], time_step=v[4])  # noqa: E731
        self.initial_state = mk(o["initial"])
        self.prediction = None
        if o["states"] is not None:
            states = [mk(v) for v in o["states"]]
            self.prediction = types.SimpleNamespace(trajectory=types.SimpleNamespace(
                state_list=states, initial_time_step=states[0].time_step if states else 0))


class Scenario:
    def __init__(self, scene: dict):
        self.dt = float(scene["dt"])
        self.scenario_id = scene.get("scenario_id", "")
        inters = []
        for it in scene.get("intersections", []):
            inc = [types.SimpleNamespace(incoming_id=e["id"], incoming_lanelets=set(e["in"]), successors_right=set(e["right"]),
                                         successors_straight=set(e["straight"]), successors_left=set(e["left"]))
                   for e in it["incomings"]]
            inters.append(types.SimpleNamespace(intersection_id=it["id"], incomings=inc))
        self.lanelet_network = LaneletNetwork([Lanelet(e) for e in scene["lanelets"]], inters)
        self._obstacles = [Obstacle(o) for o in scene["obstacles"]]
        pp = scene.get("planning_problem")
        self.planning_problem = None if pp is None else types.SimpleNamespace(
            planning_problem_id=pp["id"], goal_lanelet=pp["goal_lanelet"],
            initial_state=State(pp["initial"][:2], pp["initial"][2], pp["initial"][3], pp["initial"][4]))

    @property
    def obstacles(self):
        return self._obstacles

    @property
    def dynamic_obstacles(self):
        return [o for o in self._obstacles if o.obstacle_role is ObstacleRole.DYNAMIC]

    @property
    def static_obstacles(self):
        return [o for o in self._obstacles if o.obstacle_role is ObstacleRole.STATIC]

    def add_objects(self, obj):
        # agent.py:250 adds configured real agents (DynamicObstacle of the commonroad shim) to the scenario
        if not hasattr(obj, "obstacle_role"):
            obj.obstacle_role = ObstacleRole.DYNAMIC
        if not hasattr(obj, "obstacle_shape"):
            obj.obstacle_shape = obj.shape
        self._obstacles.append(obj)


# ---------------------------------------------------------------------------------------------------------------------
# commonroad_dc: curvilinear coordinate system and polyline utilities
# ---------------------------------------------------------------------------------------------------------------------
def compute_pathlength_from_polyline(polyline):
    p = np.asarray(polyline, dtype=np.float64)
    d = np.hypot(*np.diff(p, axis=0).T)
    return np.concatenate(([0.0], np.cumsum(d)))


def compute_curvature_from_polyline(polyline):
    p = np.asarray(polyline, dtype=np.float64)
    xd, yd = np.gradient(p[:, 0]), np.gradient(p[:, 1])
    xdd, ydd = np.gradient(xd), np.gradient(yd)
    return (xd * ydd - xdd * yd) / np.power(xd ** 2 + yd ** 2, 1.5)


class CurvilinearCoordinateSystem:
    PROJECTION_DOMAIN = 20.0

    def __init__(self, reference_path, *_a, **_kw):
        self._p = np.asarray(reference_path, dtype=np.float64).reshape(-1, 2)
        if len(self._p) < 2:
            raise ValueError("reference path needs at least two points")
        self._seg = np.diff(self._p, axis=0)
        self._len = np.hypot(self._seg[:, 0], self._seg[:, 1])
        if np.any(self._len <= 0):
            raise ValueError("reference path has duplicate points")
        self._cum = np.concatenate(([0.0], np.cumsum(self._len)))
        self._t = self._seg / self._len[:, None]

    def reference_path(self):
        return self._p.copy()

    def length(self):
        return float(self._cum[-1])

    def convert_to_curvilinear_coords(self, x, y):
        q = np.array([float(x), float(y)])
        u = np.sum((q - self._p[:-1]) * self._seg, axis=1) / self._len ** 2
        uc = np.clip(u, 0.0, 1.0)
        foot = self._p[:-1] + uc[:, None] * self._seg
        dist = np.hypot(q[0] - foot[:, 0], q[1] - foot[:, 1])
        j = int(np.argmin(dist))
        if (j == 0 and u[0] < 0.0) or (j == len(u) - 1 and u[-1] > 1.0) or dist[j] > self.PROJECTION_DOMAIN:
            raise ValueError("point outside the projection domain of the curvilinear coordinate system")
        d = self._t[j, 0] * (q[1] - self._p[j, 1]) - self._t[j, 1] * (q[0] - self._p[j, 0])
        return np.array([self._cum[j] + uc[j] * self._len[j], d])

    def convert_list_of_points_to_curvilinear_coords(self, points, _n_threads=1):
        return [self.convert_to_curvilinear_coords(float(np.asarray(p).reshape(-1)[0]), float(np.asarray(p).reshape(-1)[1]))
                for p in points]

    def convert_to_cartesian_coords(self, s, d):
        s, d = float(s), float(d)
        if s < 0.0 or s > self._cum[-1] or abs(d) > self.PROJECTION_DOMAIN:
            raise ValueError("curvilinear point outside the domain of the coordinate system")
        j = int(np.clip(np.searchsorted(self._cum, s, side="right") - 1, 0, len(self._len) - 1))
        base = self._p[j] + (s - self._cum[j]) * self._t[j]
        return np.array([base[0] - d * self._t[j, 1], base[1] + d * self._t[j, 0]])


# ---------------------------------------------------------------------------------------------------------------------
# commonroad-route-planner
# ---------------------------------------------------------------------------------------------------------------------
def lanelet_orientation_at_position(lanelet, position):
    """Direction from the closest centre vertex (the last one excluded) to its successor."""
    c = lanelet.center_vertices
    k = int(np.argmin(np.hypot(*(np.asarray(position, dtype=np.float64) - c[:-1]).T)))
    return float(np.arctan2(c[k + 1, 1] - c[k, 1], c[k + 1, 0] - c[k, 0]))


def _chaikin(polyline, refinements):
    p = np.asarray(polyline, dtype=np.float64)
    for _ in range(refinements):
        q = 0.75 * p[:-1] + 0.25 * p[1:]
        r = 0.25 * p[:-1] + 0.75 * p[1:]
        mid = np.empty((2 * len(q), 2))
        mid[0::2], mid[1::2] = q, r
        p = np.concatenate((p[:1], mid, p[-1:]))
    return p


def _resample(polyline, step):
    p = np.asarray(polyline, dtype=np.float64)
    cum = compute_pathlength_from_polyline(p)
    n = max(int(np.floor(cum[-1] / step)), 1)
    s = np.concatenate((np.arange(n + 1) * step, [cum[-1]])) if cum[-1] - n * step > 1e-9 else np.arange(n + 1) * step
    return np.stack([np.interp(s, cum, p[:, 0]), np.interp(s, cum, p[:, 1])], axis=1)


class RouteType(enum.Enum):
    REGULAR = "regular"
    SURVIVAL = "survival"


class Route:
    def __init__(self, lanelet_network, list_ids_lanelets, route_type=RouteType.REGULAR, **_kw):
        self.lanelet_network = lanelet_network
        self.list_ids_lanelets = list(list_ids_lanelets)
        pts = []
        ids = self.list_ids_lanelets
        for a, b in zip(ids, ids[1:] + [None]):
            la = lanelet_network.find_lanelet_by_id(a)
            if b is not None and b not in la.successor:
                continue                                   # lane change: the neighbour's centre line takes over
            c = la.center_vertices
            pts.append(c if not pts else c[1:] if np.allclose(pts[-1][-1], c[0]) else c)
        path = np.concatenate(pts)
        self.reference_path = _chaikin(_resample(path, 2.0), 4)


# ---------------------------------------------------------------------------------------------------------------------
# frenetix (C++ trajectory sampler)
# ---------------------------------------------------------------------------------------------------------------------
class CoordinateSystemWrapper:
    def __init__(self, ref_path):
        self.ref_line = np.asarray(ref_path, dtype=np.float64)
        seg = np.diff(self.ref_line, axis=0)
        self.ref_pos = compute_pathlength_from_polyline(self.ref_line)
        th = np.unwrap(np.arctan2(seg[:, 1], seg[:, 0]))
        self.ref_theta = np.concatenate((th, th[-1:]))
        self.ref_curv = np.gradient(self.ref_theta, self.ref_pos)
        self.system = CurvilinearCoordinateSystem(self.ref_line)


class FillCoordinates:
    def __init__(self, lowVelocityMode, initialOrientation, coordinateSystem, horizon):
        self.low_velocity_mode = bool(lowVelocityMode)
        self.initial_orientation = float(initialOrientation)
        self.cs = coordinateSystem
        self.horizon = horizon


def _feasability(name):
    def ctor(**kw):
        return types.SimpleNamespace(name=name, **kw)
    return ctor


def _quartic(t, t0, t1, x0, xd0, xdd0, xd1, xdd1):
    """Quartic with (x, x', x'')(t0) and (x', x'')(t1) given; returns x, x', x'' on t (held at the end values beyond t1)."""
    T = t1 - t0
    c0, c1, c2 = x0, xd0, 0.5 * xdd0
    A = np.array([[3 * T ** 2, 4 * T ** 3], [6 * T, 12 * T ** 2]])
    b = np.array([xd1 - c1 - 2 * c2 * T, xdd1 - 2 * c2])
    c3, c4 = np.linalg.solve(A, b)
    tau = np.clip(t - t0, 0.0, T)
    x = c0 + c1 * tau + c2 * tau ** 2 + c3 * tau ** 3 + c4 * tau ** 4
    xd = c1 + 2 * c2 * tau + 3 * c3 * tau ** 2 + 4 * c4 * tau ** 3
    xdd = 2 * c2 + 6 * c3 * tau + 12 * c4 * tau ** 2
    over = np.maximum(t - t1, 0.0)
    return x + xd1 * over, np.where(t > t1, xd1, xd), np.where(t > t1, 0.0, xdd)


def _quintic(t, t0, t1, x0, xd0, xdd0, x1, xd1, xdd1):
    T = t1 - t0
    c0, c1, c2 = x0, xd0, 0.5 * xdd0
    A = np.array([[T ** 3, T ** 4, T ** 5], [3 * T ** 2, 4 * T ** 3, 5 * T ** 4], [6 * T, 12 * T ** 2, 20 * T ** 3]])
    b = np.array([x1 - c0 - c1 * T - c2 * T ** 2, xd1 - c1 - 2 * c2 * T, xdd1 - 2 * c2])
    c3, c4, c5 = np.linalg.solve(A, b)
    tau = np.clip(t - t0, 0.0, T)
    x = c0 + c1 * tau + c2 * tau ** 2 + c3 * tau ** 3 + c4 * tau ** 4 + c5 * tau ** 5
    xd = c1 + 2 * c2 * tau + 3 * c3 * tau ** 2 + 4 * c4 * tau ** 3 + 5 * c5 * tau ** 4
    return np.where(t > t1, x1, x), np.where(t > t1, 0.0, xd)


class TrajectoryHandler:
    def __init__(self, dt):
        self.dt = float(dt)
        self.feasability_functions, self.functions, self.trajectories = [], [], []

    def add_feasability_function(self, f):
        self.feasability_functions.append(f)

    def add_function(self, f):
        self.functions.append(f)

    def reset_Trajectories(self):
        self.trajectories = []

    def generate_trajectories(self, sampling_matrix, low_vel_mode):
        self._rows = np.asarray(sampling_matrix, dtype=np.float64).reshape(-1, 13)
        self._low = bool(low_vel_mode)

    def evaluate_all_current_functions(self, _calc_all=True):
        fill = [f for f in self.functions if isinstance(f, FillCoordinates)][-1]
        cs = fill.cs
        path, cum, th = cs.ref_line, cs.ref_pos, cs.ref_theta
        n_seg = len(path) - 1
        for k, row in enumerate(self._rows):
            t0, t1, s0, ss0, sss0, ss1, sss1, d0, dd0, ddd0, d1, dd1, ddd1 = row
            steps = int(round(max(float(fill.horizon), t1) / self.dt)) + 1
            t = np.arange(steps) * self.dt
            s, sd, sdd = _quartic(t, t0, t1, s0, ss0, sss0, ss1, sss1)
            if self._low:
                # low-velocity mode: the lateral offset is a function of the covered arc length, not of time
                s_end = float(_quartic(np.array([t1]), t0, t1, s0, ss0, sss0, ss1, sss1)[0][0])
                span = max(s_end - s0, 1e-9)
                d, dprime = _quintic(np.clip(s - s0, 0.0, span), 0.0, span, d0, dd0, ddd0, d1, dd1, ddd1)
                dd = dprime * sd
            else:
                d, dd = _quintic(t, t0, t1, d0, dd0, ddd0, d1, dd1, ddd1)
            sc = np.clip(s, 0.0, cum[-1] - 1e-9)
            j = np.clip(np.searchsorted(cum, sc, side="right") - 1, 0, n_seg - 1)
            tx, ty = np.cos(th[j]), np.sin(th[j])
            # curvature of the polyline at the segment: heading change to the next segment over the mean segment length
            jn = np.minimum(j + 1, n_seg - 1)
            dth = th[jn] - th[j]
            kap = np.where(j + 2 <= n_seg, dth / np.maximum(0.5 * (cum[np.minimum(j + 2, n_seg)] - cum[j]), 1e-12), 0.0)
            vl = sd * (1.0 - kap * d)
            x = path[j, 0] + (sc - cum[j]) * tx - d * ty
            y = path[j, 1] + (sc - cum[j]) * ty + d * tx
            theta = th[j] + np.arctan2(dd, vl)
            v = np.hypot(vl, dd)
            cart = types.SimpleNamespace(x=x, y=y, theta=theta, v=v, a=sdd, kappa=kap)
            cl = types.SimpleNamespace(s=s, ss=sd, sss=sdd, d=d, dd=dd)
            self.trajectories.append(types.SimpleNamespace(cartesian=cart, curvilinear=cl, sampling_parameters=row.copy(),
                                                           feasible=True, valid=True, cost=0.0, uniqueId=k))

    def get_sorted_trajectories(self):
        return sorted(self.trajectories, key=lambda tr: tr.cost)          # stable: sampling-matrix order at equal cost


class FOVisualization:
    """``utils/visualization.py`` needs matplotlib + the commonroad renderer; every method is a no-op here."""

    def __init__(self, *a, **k):
        self.rnd = None

    def __getattr__(self, name):
        return lambda *a, **k: None

    def __bool__(self):
        return False
