"""Scenario replay harness: stands in for the external planner / scenario handler
(Frenetix-Motion-Planner, cr_scenario_handler -- neither is part of the reference repository) so that the
assessment hot path can be driven over the reference's example scenarios.

It provides (i) an ego reference path from the planning problem's start lanelet to its goal lanelet,
(ii) a deterministic Frenet trajectory fan in the layout the planner hands over (``[N, T, 5]`` =
x, y, theta, v, a at the rear axle), (iii) an open-loop ego that follows the reference path at its initial speed,
and (iv) ``replay()`` which calls ``FOInterface.evaluate_scenario`` + ``assess_bundle`` per planning cycle.
None of this is on the measured path; it only produces inputs."""
from __future__ import annotations

import types
from collections import deque

import numpy as np

from .route_planner import resample_polyline
from .utils.curvilinear import CurvilinearCoordinateSystem

# BMW 320i parameters the reference's planner configuration uses (vehicle_params duck type)
DEFAULT_VEHICLE = types.SimpleNamespace(length=4.508, width=1.61, mass=1093.3, wb_rear_axle=1.4227, a_max=11.5)


def deployment_config(agents="default", activated_metrics=None, thresholds=None) -> dict:
    """The configuration the planner passes in deployment: package defaults with the values of the reference's
    ``configurations/simulation/occlusion.yaml`` (harm threshold 0.1, bicycle entering at time step 6)."""
    import os
    import yaml
    with open(os.path.join(os.path.dirname(__file__), "config", "config.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["metrics"]["metric_thresholds"]["harm"] = 0.1
    cfg["agents"][1]["timestep"] = 6
    if agents != "default":
        cfg["agents"] = agents
    if activated_metrics is not None:
        cfg["metrics"]["activated_metrics"] = list(activated_metrics)
    if thresholds is not None:
        cfg["metrics"]["metric_thresholds"].update(thresholds)
    return cfg


def _chaikin(p, iterations=2):
    p = np.asarray(p, dtype=np.float64)
    for _ in range(iterations):
        q = 0.75 * p[:-1] + 0.25 * p[1:]
        r = 0.25 * p[:-1] + 0.75 * p[1:]
        mid = np.empty((2 * len(q), 2))
        mid[0::2], mid[1::2] = q, r
        p = np.concatenate((p[:1], mid, p[-1:]))
    return p


def find_route(lanelet_network, start_id, goal_id):
    """Breadth-first search over successors and same-direction neighbours."""
    prev = {start_id: None}
    dq = deque([start_id])
    while dq:
        cur = dq.popleft()
        if cur == goal_id:
            break
        l = lanelet_network.find_lanelet_by_id(cur)
        nxt = list(l.successor)
        if l.adj_left is not None and l.adj_left_same_direction:
            nxt.append(l.adj_left)
        if l.adj_right is not None and l.adj_right_same_direction:
            nxt.append(l.adj_right)
        for n in nxt:
            if n not in prev:
                prev[n] = cur
                dq.append(n)
    if goal_id not in prev:
        raise ValueError(f"no route from lanelet {start_id} to lanelet {goal_id}")
    route = [goal_id]
    while prev[route[-1]] is not None:
        route.append(prev[route[-1]])
    return route[::-1]


def reference_path_for(scenario, step=1.0):
    """Centre lines of the route start -> goal, corner-cut and resampled (the planner's route planner does the
    same kind of smoothing; the exact path is the planner's business, not this package's)."""
    pp = scenario.planning_problem
    ln = scenario.lanelet_network
    start_ids = ln.find_lanelet_by_position([pp.initial_state.position])[0]
    if not start_ids:
        raise ValueError("planning problem does not start on a lanelet")
    route = None
    for sid in start_ids:
        try:
            route = find_route(ln, sid, pp.goal_lanelet)
            break
        except ValueError:
            continue
    if route is None:
        raise ValueError("goal lanelet not reachable")
    # successive lanelets are chained by centre line; lane changes to a neighbour keep only the later lanelet
    pts = []
    for a, b in zip(route, route[1:] + [None]):
        la = ln.find_lanelet_by_id(a)
        if b is not None and b not in la.successor:
            continue
        pts.append(la.center_vertices)
    path = np.concatenate(pts)
    path = resample_polyline(path, 2.0)
    path = _chaikin(path, 3)
    return resample_polyline(path, step), route


def frenet_fan(cosy, s0, d0, v0, dt=0.1, horizon=3.0, speed_factors=None, lateral_targets=None, accel0=0.0):
    """Deterministic stand-in for the planner's sampled bundle: quartic velocity keeping in s, quintic in d.
    Returns float64 ``[N, T, 5]`` (x, y, theta, v, a), N = len(speed_factors) * len(lateral_targets)."""
    speed_factors = np.asarray(speed_factors if speed_factors is not None else np.linspace(0.0, 1.3, 14))
    lateral_targets = np.asarray(lateral_targets if lateral_targets is not None else np.linspace(-1.5, 1.5, 7))
    T = int(round(horizon / dt)) + 1
    t = np.arange(T) * dt
    t1 = horizon
    out = []
    path = cosy.reference_path()
    seg = np.diff(path, axis=0)
    cum = np.concatenate(([0.0], np.cumsum(np.hypot(seg[:, 0], seg[:, 1]))))
    head = np.unwrap(np.arctan2(seg[:, 1], seg[:, 0]))
    for f in speed_factors:
        v1 = max(v0 * f, 0.0)
        a3, a4 = (v1 - v0) / t1 ** 2, (v0 - v1) / (2 * t1 ** 3)
        s = s0 + v0 * t + a3 * t ** 3 + a4 * t ** 4
        sd = v0 + 3 * a3 * t ** 2 + 4 * a4 * t ** 3
        sdd = 6 * a3 * t + 12 * a4 * t ** 2
        for d1 in lateral_targets:
            tau = t / t1
            d = d0 + (d1 - d0) * (10 * tau ** 3 - 15 * tau ** 4 + 6 * tau ** 5)
            dd = (d1 - d0) / t1 * (30 * tau ** 2 - 60 * tau ** 3 + 30 * tau ** 4)
            sc = np.clip(s, 0.0, cum[-1] - 1e-6)
            j = np.clip(np.searchsorted(cum, sc, side="right") - 1, 0, len(seg) - 1)
            tx, ty = np.cos(head[j]), np.sin(head[j])
            x = path[j, 0] + (sc - cum[j]) * tx - d * ty
            y = path[j, 1] + (sc - cum[j]) * ty + d * tx
            theta = head[j] + np.arctan2(dd, np.maximum(sd, 1e-3))
            v = np.hypot(sd, dd)
            out.append(np.stack((x, y, theta, v, sdd), -1))
    return np.stack(out)


class OpenLoopEgo:
    """Ego that follows the reference path at constant speed (the planner is out of scope)."""

    def __init__(self, scenario, dt=None):
        self.scenario = scenario
        self.dt = scenario.dt if dt is None else dt
        self.reference_path, self.route = reference_path_for(scenario)
        self.cosy = CurvilinearCoordinateSystem(self.reference_path)
        st = scenario.planning_problem.initial_state
        self.s0, self.d0 = self.cosy.convert_to_curvilinear_coords(st.position[0], st.position[1])
        self.v0 = float(st.velocity)

    def state(self, timestep):
        s = min(self.s0 + self.v0 * self.dt * timestep, self.cosy.length() - 1.0)
        pos = self.cosy.convert_to_cartesian_coords(s, self.d0)
        ahead = self.cosy.convert_to_cartesian_coords(min(s + 0.5, self.cosy.length()), self.d0)
        heading = float(np.arctan2(ahead[1] - pos[1], ahead[0] - pos[0]))
        return {"pos": np.asarray(pos), "orientation": heading, "pos_cl": np.array([s, self.d0]), "v": self.v0}


def replay(interface, ego: OpenLoopEgo, timesteps, fan_kwargs=None, collect_detail=False):
    """Drive ``interface`` (an ``FOInterface``) over ``timesteps`` planning cycles.  Returns one record per cycle."""
    def _np(t):
        return t.detach().cpu().numpy() if hasattr(t, "detach") else (None if t is None else np.asarray(t))

    records = []
    for ts in timesteps:
        st = ego.state(ts)
        predictions = {}
        interface.evaluate_scenario(predictions, st["pos"], st["orientation"], st["pos_cl"], st["v"], ts, ego.cosy)
        fan = frenet_fan(ego.cosy, st["pos_cl"][0], st["pos_cl"][1], st["v"], dt=ego.dt, **(fan_kwargs or {}))
        res = interface.assess_bundle(fan, want_pair=collect_detail)
        rec = {"timestep": int(ts), "ego": st, "fan": fan,
               "spawn_points": [{"position": np.asarray(sp.position, dtype=np.float64), "agent_type": sp.agent_type,
                                 "source": sp.source, "orientation": sp.orientation} for sp in interface.spawn_points],
               "visible_obstacles": list(interface.sensor_model.visible_objects_timestep),
               "prediction_ids": list(interface.agent_manager.predictions.keys()),
               "predictions": {k: {kk: (np.asarray(vv) if kk != "shape" else dict(vv)) for kk, vv in p.items()}
                               for k, p in interface.agent_manager.predictions.items()},
               "agent_types": {k: interface.agent_manager.agent_by_prediction_id(k).agent_type
                               for k in interface.agent_manager.predictions},
               "valid": _np(res.valid).astype(bool), "summary": _np(getattr(res, "summary", None)),
               "flags": _np(getattr(res, "flags", None)), "result": res}
        if collect_detail and getattr(res, "pair", None) is not None:
            rec["pair"] = _np(res.pair)
        records.append(rec)
    return records
