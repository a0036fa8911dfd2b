"""Dependency-free CommonRoad 2020a scenario loader and lanelet-network geometry.

The reference works on commonroad-io objects; commonroad-io is not installable here, so this module
parses the XML with the standard library into light objects exposing exactly the attributes this path
reads (duck-type compatible with commonroad-io 2023.2: a real ``commonroad.scenario.scenario.Scenario``
can be passed to ``FOInterface`` instead).  It also derives the *road border* -- the exterior of the
union of all lanelet polygons the reference builds with shapely (``sensor_model.py:195-199``) -- as a
soup of opaque segments for the ray caster.
"""
from __future__ import annotations

import types
import xml.etree.ElementTree as ET
from typing import List, Optional

import numpy as np


class _Enum:
    """Tiny stand-in for the commonroad enums this path touches (``.value`` / ``.name``)."""

    def __init__(self, value, name=None):
        self.value = value
        self.name = name if name is not None else str(value).upper()

    def __eq__(self, other):
        return getattr(other, "value", other) == self.value

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return f"<{self.name}: {self.value!r}>"


class Rectangle:
    """commonroad ``Rectangle``: ``vertices`` is the closed 5-point ring in the shape's own frame
    (including its ``center`` offset and ``orientation``), as ``hf.calc_corner_points`` expects."""

    def __init__(self, length, width, center=None, orientation=0.0):
        self.length = float(length)
        self.width = float(width)
        self.center = np.zeros(2) if center is None else np.asarray(center, dtype=np.float64)
        self.orientation = float(orientation)

    @property
    def vertices(self):
        hl, hw = 0.5 * self.length, 0.5 * self.width
        loc = np.array([[-hl, -hw], [-hl, hw], [hl, hw], [hl, -hw], [-hl, -hw]])
        c, s = np.cos(self.orientation), np.sin(self.orientation)
        return loc @ np.array([[c, -s], [s, c]]).T + self.center


class State:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


class Obstacle:
    def __init__(self, obstacle_id, obstacle_type, role, shape, initial_state, state_list=None):
        self.obstacle_id = obstacle_id
        self.obstacle_type = _Enum(obstacle_type)
        self.obstacle_role = _Enum(role.lower(), role.upper())
        self.obstacle_shape = shape
        self.initial_state = initial_state
        self.prediction = None
        if state_list is not None:
            traj = types.SimpleNamespace(state_list=state_list,
                                         initial_time_step=state_list[0].time_step if state_list else 0)
            self.prediction = types.SimpleNamespace(trajectory=traj, shape=shape)


class Lanelet:
    def __init__(self, lanelet_id, left, right):
        self.lanelet_id = lanelet_id
        self.left_vertices = np.asarray(left, dtype=np.float64)
        self.right_vertices = np.asarray(right, dtype=np.float64)
        self.center_vertices = 0.5 * (self.left_vertices + self.right_vertices)
        self.predecessor: List[int] = []
        self.successor: List[int] = []
        self.adj_left: Optional[int] = None
        self.adj_left_same_direction: Optional[bool] = None
        self.adj_right: Optional[int] = None
        self.adj_right_same_direction: Optional[bool] = None
        self.lanelet_type = set()      # of _Enum, like commonroad's Set[LaneletType]

    @property
    def polygon_vertices(self):
        return np.concatenate((self.left_vertices, self.right_vertices[::-1]))


def _points_in_polygon(P, poly):
    """Even-odd rule, vectorised over points [M,2]; boundary membership is unspecified."""
    x, y = P[:, 0][:, None], P[:, 1][:, None]
    x0, y0 = poly[:, 0][None], poly[:, 1][None]
    x1, y1 = np.roll(poly[:, 0], -1)[None], np.roll(poly[:, 1], -1)[None]
    cond = (y0 > y) != (y1 > y)
    with np.errstate(divide="ignore", invalid="ignore"):
        xin = (x1 - x0) * (y - y0) / (y1 - y0) + x0
    return (np.sum(cond & (x < xin), axis=1) % 2).astype(bool)


class IntersectionIncomingElement:
    def __init__(self, incoming_id, incoming_lanelets, successors_right, successors_straight, successors_left):
        self.incoming_id = incoming_id
        self.incoming_lanelets = set(incoming_lanelets)
        self.successors_right = set(successors_right)
        self.successors_straight = set(successors_straight)
        self.successors_left = set(successors_left)


class Intersection:
    def __init__(self, intersection_id, incomings):
        self.intersection_id = intersection_id
        self.incomings = list(incomings)


class LaneletNetwork:
    def __init__(self, lanelets, intersections=None):
        self.lanelets = list(lanelets)
        self.intersections = list(intersections or [])
        self._by_id = {l.lanelet_id: l for l in self.lanelets}

    def find_lanelet_by_id(self, lanelet_id):
        return self._by_id[lanelet_id]

    @property
    def lanelet_polygons(self):
        return [l.polygon_vertices for l in self.lanelets]

    def _stacked(self):
        """All lanelet rings padded to one [L, V, 2] array (the last vertex repeated: zero-length edges never cross)."""
        if getattr(self, "_stack", None) is None:
            vmax = max(len(l.polygon_vertices) for l in self.lanelets)
            st = np.empty((len(self.lanelets), vmax, 2))
            for k, l in enumerate(self.lanelets):
                pv = l.polygon_vertices
                st[k, :len(pv)] = pv
                st[k, len(pv):] = pv[-1]
            # edge k of a ring: vertex k-1 -> vertex k; the padding closes the ring through repeated last vertices
            self._stack = (np.roll(st, 1, axis=1), st)
            for k, l in enumerate(self.lanelets):          # closing edge: last real vertex -> first vertex
                self._stack[0][k, 0] = l.polygon_vertices[-1]
        return self._stack

    def points_in_lanelets(self, P) -> np.ndarray:
        """Even-odd membership of points [M,2] in every lanelet polygon at once -> bool [M, L]."""
        P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
        a, b = self._stacked()
        x, y = P[:, 0][:, None, None], P[:, 1][:, None, None]
        x0, y0, x1, y1 = a[None, ..., 0], a[None, ..., 1], b[None, ..., 0], b[None, ..., 1]
        cond = (y0 > y) != (y1 > y)
        with np.errstate(divide="ignore", invalid="ignore"):
            xin = (x1 - x0) * (y - y0) / (y1 - y0) + x0
        return (np.sum(cond & (x < xin), axis=2) % 2).astype(bool)

    def find_lanelet_by_position(self, point_list):
        inside = self.points_in_lanelets(np.asarray([np.asarray(p, dtype=np.float64).reshape(2) for p in point_list]))
        return [[self.lanelets[k].lanelet_id for k in np.nonzero(row)[0]] for row in inside]

    def points_on_road(self, P):
        P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
        out = np.zeros(len(P), dtype=bool)
        for lo in range(0, len(P), 4096):                  # bounded temporaries
            out[lo:lo + 4096] = self.points_in_lanelets(P[lo:lo + 4096]).any(1)
        return out

    def road_border_segments(self, piece=1.0, eps=2e-3) -> np.ndarray:
        """Exterior of the union of all lanelet polygons as segments [B,4].  A piece of a lanelet edge
        lies on the union's border iff the road-membership of the two points eps to its left and right
        differs (shared lane borders and edges running through crossing lanelets drop out)."""
        segs = []
        for l in self.lanelets:
            ring = l.polygon_vertices
            a, b = ring, np.roll(ring, -1, axis=0)
            for p0, p1 in zip(a, b):
                n = max(1, int(np.ceil(np.hypot(*(p1 - p0)) / piece)))
                t = np.linspace(0.0, 1.0, n + 1)
                pts = p0[None] + t[:, None] * (p1 - p0)[None]
                segs.append(np.concatenate((pts[:-1], pts[1:]), 1))
        segs = np.concatenate(segs)
        segs = segs[np.hypot(segs[:, 2] - segs[:, 0], segs[:, 3] - segs[:, 1]) > 1e-9]
        mid = 0.5 * (segs[:, :2] + segs[:, 2:])
        d = segs[:, 2:] - segs[:, :2]
        nrm = np.stack((-d[:, 1], d[:, 0]), -1) / np.hypot(d[:, 0], d[:, 1])[:, None]
        left = self.points_on_road(mid + eps * nrm)
        right = self.points_on_road(mid - eps * nrm)
        keep = left != right
        out = segs[keep]
        # merge duplicates produced by coincident borders of neighbouring lanelets
        key = np.round(np.where((out[:, :2] < out[:, 2:]).all(1, keepdims=True) | (out[:, 0:1] < out[:, 2:3]),
                                out, out[:, [2, 3, 0, 1]]), 6)
        _, idx = np.unique(key, axis=0, return_index=True)
        return out[np.sort(idx)]


class Scenario:
    def __init__(self, dt, lanelet_network, obstacles, scenario_id="", planning_problem=None):
        self.dt = dt
        self.lanelet_network = lanelet_network
        self._obstacles = list(obstacles)
        self.scenario_id = scenario_id
        self.planning_problem = planning_problem

    @property
    def obstacles(self):
        return self._obstacles

    @property
    def dynamic_obstacles(self):
        return [o for o in self._obstacles if o.obstacle_role.name == "DYNAMIC"]

    @property
    def static_obstacles(self):
        return [o for o in self._obstacles if o.obstacle_role.name == "STATIC"]

    def add_objects(self, obj):
        self._obstacles.append(obj)

    def obstacle_by_id(self, oid):
        for o in self._obstacles:
            if o.obstacle_id == oid:
                return o
        return None


# ------------------------------------------------------------------------------------------------
def _f(node, path, default=None):
    n = node.find(path)
    return float(n.text) if n is not None and n.text is not None else default


def _points(bound):
    return [[float(p.find("x").text), float(p.find("y").text)] for p in bound.findall("point")]


def _state(node):
    pos = node.find("position/point")
    position = np.array([float(pos.find("x").text), float(pos.find("y").text)]) if pos is not None else None
    t = node.find("time/exact")
    return State(position=position, orientation=_f(node, "orientation/exact", 0.0),
                 velocity=_f(node, "velocity/exact", 0.0), acceleration=_f(node, "acceleration/exact", 0.0),
                 time_step=int(float(t.text)) if t is not None else 0)


def _shape(node):
    r = node.find("rectangle")
    if r is None:
        raise NotImplementedError("only rectangle obstacle shapes are supported")
    c = r.find("center")
    center = None if c is None else np.array([float(c.find("x").text), float(c.find("y").text)])
    return Rectangle(_f(r, "length"), _f(r, "width"), center=center, orientation=_f(r, "orientation", 0.0) or 0.0)


def load_commonroad_xml(path: str) -> Scenario:
    root = ET.parse(path).getroot()
    dt = float(root.attrib.get("timeStepSize", 0.1))
    lanelets = []
    for ln in root.findall("lanelet"):
        l = Lanelet(int(ln.attrib["id"]), _points(ln.find("leftBound")), _points(ln.find("rightBound")))
        l.predecessor = [int(p.attrib["ref"]) for p in ln.findall("predecessor")]
        l.successor = [int(p.attrib["ref"]) for p in ln.findall("successor")]
        al, ar = ln.find("adjacentLeft"), ln.find("adjacentRight")
        if al is not None:
            l.adj_left, l.adj_left_same_direction = int(al.attrib["ref"]), al.attrib.get("drivingDir") == "same"
        if ar is not None:
            l.adj_right, l.adj_right_same_direction = int(ar.attrib["ref"]), ar.attrib.get("drivingDir") == "same"
        l.lanelet_type = {_Enum(t.text.strip()) for t in ln.findall("laneletType") if t.text}
        lanelets.append(l)
    intersections = []
    for it in root.findall("intersection"):
        incs = []
        for inc in it.findall("incoming"):
            refs = lambda tag: [int(n.attrib["ref"]) for n in inc.findall(tag)]  # noqa: E731
            incs.append(IntersectionIncomingElement(int(inc.attrib["id"]), refs("incomingLanelet"), refs("successorsRight"),
                                                    refs("successorsStraight"), refs("successorsLeft")))
        intersections.append(Intersection(int(it.attrib["id"]), incs))
    obstacles = []
    for tag, role in (("staticObstacle", "static"), ("dynamicObstacle", "dynamic")):
        for ob in root.findall(tag):
            states = None
            tr = ob.find("trajectory")
            if tr is not None:
                states = [_state(s) for s in tr.findall("state")]
            obstacles.append(Obstacle(int(ob.attrib["id"]), ob.find("type").text.strip(), role, _shape(ob.find("shape")),
                                      _state(ob.find("initialState")), states))
    pp = None
    ppn = root.find("planningProblem")
    if ppn is not None:
        goal = ppn.find("goalState/position/lanelet")
        pp = types.SimpleNamespace(planning_problem_id=int(ppn.attrib["id"]), initial_state=_state(ppn.find("initialState")),
                                   goal_lanelet=int(goal.attrib["ref"]) if goal is not None else None)
    return Scenario(dt, LaneletNetwork(lanelets, intersections), obstacles, root.attrib.get("benchmarkID", ""), pp)


# ------------------------------------------------------------------------------------------------
def scenario_to_dict(sc: Scenario, ndigits: int = 6) -> dict:
    """Compact JSON-able form of exactly what this path reads from a scenario (test fixtures)."""
    r = lambda a: np.round(np.asarray(a, dtype=np.float64), ndigits).tolist()  # noqa: E731
    ln = sc.lanelet_network
    out = {"dt": sc.dt, "scenario_id": sc.scenario_id, "lanelets": [], "intersections": [], "obstacles": []}
    for l in ln.lanelets:
        out["lanelets"].append({"id": l.lanelet_id, "left": r(l.left_vertices), "right": r(l.right_vertices),
                                "pred": l.predecessor, "succ": l.successor,
                                "adj_left": [l.adj_left, l.adj_left_same_direction],
                                "adj_right": [l.adj_right, l.adj_right_same_direction],
                                "type": sorted(t.value for t in l.lanelet_type)})
    for it in ln.intersections:
        out["intersections"].append({"id": it.intersection_id, "incomings": [
            {"id": e.incoming_id, "in": sorted(e.incoming_lanelets), "right": sorted(e.successors_right),
             "straight": sorted(e.successors_straight), "left": sorted(e.successors_left)} for e in it.incomings]})
    st = lambda s: [float(s.position[0]), float(s.position[1]), float(s.orientation), float(s.velocity), int(s.time_step)]  # noqa: E731
    for o in sc.obstacles:
        sh = o.obstacle_shape
        out["obstacles"].append({"id": o.obstacle_id, "type": o.obstacle_type.value, "role": o.obstacle_role.value,
                                 "shape": [sh.length, sh.width, float(sh.center[0]), float(sh.center[1]), sh.orientation],
                                 "initial": r(st(o.initial_state)),
                                 "states": None if o.prediction is None else r([st(s) for s in o.prediction.trajectory.state_list])})
    pp = sc.planning_problem
    if pp is not None:
        out["planning_problem"] = {"id": pp.planning_problem_id, "initial": r(st(pp.initial_state)), "goal_lanelet": pp.goal_lanelet}
    return out


def scenario_from_dict(d: dict) -> Scenario:
    def mk_state(v):
        return State(position=np.array([v[0], v[1]]), orientation=float(v[2]), velocity=float(v[3]), acceleration=0.0,
                     time_step=int(v[4]))
    lanelets = []
    for e in d["lanelets"]:
        l = Lanelet(e["id"], e["left"], e["right"])
        l.predecessor, l.successor = list(e["pred"]), list(e["succ"])
        l.adj_left, l.adj_left_same_direction = e["adj_left"]
        l.adj_right, l.adj_right_same_direction = e["adj_right"]
        l.lanelet_type = {_Enum(t) for t in e.get("type", [])}
        lanelets.append(l)
    inters = [Intersection(it["id"], [IntersectionIncomingElement(e["id"], e["in"], e["right"], e["straight"], e["left"])
                                      for e in it["incomings"]]) for it in d.get("intersections", [])]
    obstacles = []
    for o in d["obstacles"]:
        sh = o["shape"]
        shape = Rectangle(sh[0], sh[1], center=np.array([sh[2], sh[3]]), orientation=sh[4])
        states = None if o["states"] is None else [mk_state(v) for v in o["states"]]
        obstacles.append(Obstacle(o["id"], o["type"], o["role"], shape, mk_state(o["initial"]), states))
    pp = None
    if "planning_problem" in d:
        p = d["planning_problem"]
        pp = types.SimpleNamespace(planning_problem_id=p["id"], initial_state=mk_state(p["initial"]), goal_lanelet=p["goal_lanelet"])
    return Scenario(d["dt"], LaneletNetwork(lanelets, inters), obstacles, d.get("scenario_id", ""), pp)
