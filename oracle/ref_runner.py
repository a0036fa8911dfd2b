"""Oracle A, part 2: run the reference's own ``Metric.evaluate_metrics`` on a plain-array *case*.

TEST INFRASTRUCTURE; build container only (needs ``/root/reference``).  Control flow executed is
the reference's, unmodified: ``frenetix_occlusion/metrics/metric.py:35-100`` and every plugin it
dispatches to.  Only the leaves listed in ``ref_shims`` are stand-ins.

Case format (shared by oracle A, oracle B, the golden fixtures and the CUDA parity tests)::

    {"dt": float,
     "vehicle": {"length","width","mass","wb_rear_axle","a_max"},
     "ego": float64 [N, T, 5]  (x, y, theta, v, a) -- rear-axle reference point, as
             ``trajectory.cartesian.*`` in the reference,
     "agents": [ {"agent_type": "Pedestrian"|"Bicycle"|"Car"|"Truck",
                  "length","width"        raw agent.shape (agent.py:216),
                  "buf_length","buf_width" prediction['shape'] (agent.py:402-407,525-526),
                  "pos": [Ta,2], "yaw": [Ta], "v": [Ta], "var": [Ta]  (cov = var*I, agent.py:261-280)} ],
     "activated_metrics": [...], "thresholds": {harm,risk,be,cp,ttc,wttc,ttce,dce}}
"""
from __future__ import annotations

import copy
import types

import numpy as np

from . import ref_shims


class _Cartesian:
    def __init__(self, row):
        self.x = np.array(row[:, 0], dtype=np.float64)
        self.y = np.array(row[:, 1], dtype=np.float64)
        self.theta = np.array(row[:, 2], dtype=np.float64)
        self.v = np.array(row[:, 3], dtype=np.float64)
        self.a = np.array(row[:, 4], dtype=np.float64)


class TrajectoryStub:
    """Duck type of frenetix ``TrajectorySample`` as used by the reference."""

    def __init__(self, row):
        self.cartesian = _Cartesian(np.asarray(row, dtype=np.float64))
        self.costMap = {}
        self.feasabilityMap = {}
        self.cost = 0.0
        self.feasible = True
        self.sampling_parameters = np.empty(0)
        self.uniqueId = 0
        self.valid = True


def build_agent_manager(case):
    """Stand-in for ``FOAgentManager`` (agent.py:27-199) holding phantom agents + predictions with the
    reference's dict layout (agent.py:420-424, 530-534) and id scheme (agent.py:179-183)."""
    ref_shims.install()
    am = types.SimpleNamespace()
    am.dt = float(case["dt"])
    am.visualization = None
    am.phantom_agents = []
    am.predictions = {}
    for k, ag in enumerate(case["agents"]):
        agent_id = 10000 + k
        agent = types.SimpleNamespace()
        agent.agent_id = agent_id
        agent.agent_type = ag["agent_type"]
        agent.obstacle_type = ref_shims.ObstacleType(ag["agent_type"].lower())
        agent.shape = ref_shims.Rectangle(ag["length"], ag["width"], center=np.array([0.0, 0.0]), orientation=0.0)
        am.phantom_agents.append(agent)
        var = np.asarray(ag["var"], dtype=np.float64)
        pred = {"orientation_list": np.asarray(ag["yaw"], dtype=np.float64),
                "v_list": np.asarray(ag["v"], dtype=np.float64),
                "pos_list": np.asarray(ag["pos"], dtype=np.float64).reshape(-1, 2),
                "shape": {"length": float(ag["buf_length"]), "width": float(ag["buf_width"])},
                "cov_list": np.array([[[v, 0.0], [0.0, v]] for v in var])}
        am.predictions[int(str(agent_id) + "0")] = pred

    def agent_by_prediction_id(prediction_id):
        if not am.phantom_agents:
            return None
        aid = int(str(prediction_id)[:5])
        for agent in am.phantom_agents:
            if agent.agent_id == aid:
                return agent

    am.agent_by_prediction_id = agent_by_prediction_id
    return am


def run_reference_metrics(case):
    """Returns a list (one entry per ego trajectory) of ``(results, safety_check)`` exactly as
    ``FOInterface.trajectory_safety_assessment`` would (interface.py:216-219).  A trajectory on
    which the reference raises (BE's ``interp1d`` bounds error, be.py:117-124) yields
    ``({"error": "<ExceptionType>"}, None)``."""
    ref_shims.install()
    from frenetix_occlusion.metrics.metric import Metric  # the reference's own module

    am = build_agent_manager(case)
    vp = types.SimpleNamespace(**{k: float(v) for k, v in case["vehicle"].items()})
    config = {"activated_metrics": list(case["activated_metrics"]),
              "metric_thresholds": copy.deepcopy(case["thresholds"])}
    metric = Metric(config, vp, am)
    out = []
    ego = np.asarray(case["ego"], dtype=np.float64)
    for n in range(ego.shape[0]):
        try:
            with np.errstate(all="ignore"):
                res, ok = metric.evaluate_metrics(TrajectoryStub(ego[n]))
        except ValueError as e:  # interp1d out-of-range in BE
            out.append(({"error": type(e).__name__}, None))
            continue
        out.append((res, bool(ok)))
    return out, list(metric.metrics.keys())


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return [_jsonable(v) for v in o.tolist()]
    if isinstance(o, (np.floating, float)):
        f = float(o)
        if np.isinf(f):
            return "inf" if f > 0 else "-inf"
        if np.isnan(f):
            return "nan"
        return f
    if isinstance(o, (np.integer, int)):
        return int(o)
    if isinstance(o, (np.bool_, bool)):
        return bool(o)
    return o


def results_to_json(results):
    return _jsonable(results)
