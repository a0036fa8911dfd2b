"""CPU oracle of the whole per-planning-cycle pipeline (visibility -> spawn points -> phantom predictions ->
dense metric core).  TEST INFRASTRUCTURE: only ``tests/``, ``oracle/make_scenario_golden.py`` and
``__graft_entry__.smoke()`` use it.

PARITY UNPINNED for stages 1-2 and the spawn locator: the reference needs shapely / commonroad / frenetix, none
of which can be installed, and ships no test vectors.  What this oracle pins is the DEVICE side of the product:
every call the host classes make into ``libfo_b200.so`` -- ray casting, point classification, CV / path rollouts,
the metric bundle -- is replaced by the float64 numpy restatement of the reference's construction
(``oracle/visibility_oracle.py``, ``oracle/metric_oracle.py``), while the host bookkeeping of
``frenetix_occlusion_b200`` (thresholds, sorting, raster components: no arithmetic of the path) is shared.
The scenario goldens under ``tests/golden/scene_*.json`` are produced with it on the CPU."""
from __future__ import annotations

import types

import numpy as np

from frenetix_occlusion_b200 import agent as A
from frenetix_occlusion_b200 import sensor_model as SM
from frenetix_occlusion_b200.interface import FOInterface

from . import metric_oracle as MO
from . import visibility_oracle as VO


class _HostFrame:
    def __init__(self, origin, heading, rect, flags, boundary, polygons, radius, fov):
        self.origin = np.asarray(origin, dtype=np.float64)
        self.ego = np.array([self.origin[0], self.origin[1], float(heading)])
        # the device path sees float32 values relative to the ego: round the same way so both sides
        # consume identical numbers
        f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)  # noqa: E731
        r = np.array(rect, dtype=np.float64).reshape(-1, 5)
        r[:, :2] -= self.origin
        self.rect = f32(r)
        self.flags = np.asarray(flags, dtype=np.uint8).reshape(-1)
        b = np.zeros((0, 4)) if boundary is None else np.asarray(boundary, dtype=np.float64).reshape(-1, 4)
        self.boundary = f32(b - np.tile(self.origin, 2))
        self.polygons = [f32(np.asarray(p, dtype=np.float64) - self.origin) for p in polygons]
        self.radius, self.fov = float(radius), float(fov)
        self.n_obstacles = len(self.rect)
        self.heading = float(np.float32(heading))


class OracleSensorModel(SM.SensorModel):
    def _build_frame(self, rect, flags, border):
        return _HostFrame(self.ego_pos, self.ego_orientation, rect, flags, border, self.lanelet_polygons,
                          self.sensor_radius, self.sensor_angle)

    def _raycast(self):
        f = self._frame
        rng, hit, vis = VO.raycast(np.array([0.0, 0.0, f.heading]), f.rect, f.flags, f.boundary, f.radius, f.fov,
                                   self.n_rays)
        return rng, hit, vis

    def _classify(self, points, focus=-1, focus_margin=0.0):
        f = self._frame
        P = np.asarray(points, dtype=np.float64).reshape(-1, 2) - f.origin
        P = P.astype(np.float32).astype(np.float64)
        flags, lan = VO.classify_points(P, np.array([0.0, 0.0, f.heading]), f.rect, f.flags, f.boundary, f.polygons,
                                        f.radius, f.fov, 1.5 * f.radius, focus=focus, focus_margin=focus_margin)
        return flags, np.full(len(P), VO.HIT_NONE, dtype=np.int32), lan


class _OracleRollouts:
    def _rollout_cv(self, pos, velocity, phi, var_factor):
        r = VO.rollout_cv([pos[0]], [pos[1]], [velocity], [phi], self.dt, self.horizon, 0.1, var_factor)
        return {k: r[k].astype(np.float32).astype(np.float64) for k in ("x", "y", "yaw", "v", "var")}

    def _rollout_path(self, paths, pos, velocity, var_factor):
        rows = [VO.rollout_path(p, pos[0], pos[1], velocity, self.dt, self.horizon, 3.0, 0.1, var_factor) for p in paths]
        out = {k: np.stack([r[k] for r in rows]).astype(np.float32).astype(np.float64) for k in ("x", "y", "yaw", "v", "var")}
        out["sample"] = np.array([r["sample"] for r in rows])
        return out


class OraclePedestrianAgent(_OracleRollouts, A.OAPPedestrianAgent):
    pass


class OracleVehicleAgent(_OracleRollouts, A.OAPVehicleAgent):
    pass


class OracleAgentManager(A.FOAgentManager):
    pedestrian_cls = OraclePedestrianAgent
    vehicle_cls = OracleVehicleAgent


def case_from_predictions(agent_manager, vehicle_params, ego, activated_metrics, thresholds):
    """Plain-array case (the format of ``metric_oracle.evaluate_bundle``) from an agent manager."""
    agents = []
    for pid, pred in agent_manager.predictions.items():
        ag = agent_manager.agent_by_prediction_id(pid)
        agents.append({"agent_type": ag.agent_type, "length": ag.shape.length, "width": ag.shape.width,
                       "buf_length": pred["shape"]["length"], "buf_width": pred["shape"]["width"],
                       "pos": np.asarray(pred["pos_list"]), "yaw": np.asarray(pred["orientation_list"]),
                       "v": np.asarray(pred["v_list"]), "var": np.asarray(pred["cov_list"])[:, 0, 0]})
    g = (lambda k: float(vehicle_params[k])) if isinstance(vehicle_params, dict) else (lambda k: float(getattr(vehicle_params, k)))
    return {"dt": float(agent_manager.dt), "vehicle": {k: g(k) for k in ("length", "width", "mass", "wb_rear_axle", "a_max")},
            "ego": np.asarray(ego, dtype=np.float64), "agents": agents,
            "activated_metrics": list(activated_metrics), "thresholds": dict(thresholds)}


class OracleMetric:
    """Batched stand-in for ``metrics.metric.Metric`` on the float64 oracle."""

    def __init__(self, config, vehicle_params, agent_manager):
        self.config = config
        self.metric_thresholds = config["metric_thresholds"]
        self.vehicle_params = vehicle_params
        self.agent_manager = agent_manager

    def evaluate_bundle(self, trajectories, want_pair=False, want_step=False):
        ego = np.asarray(trajectories, dtype=np.float64)
        case = case_from_predictions(self.agent_manager, self.vehicle_params, ego, self.config["activated_metrics"],
                                     self.metric_thresholds)
        out = MO.evaluate_bundle(case, want_detail=False)
        return types.SimpleNamespace(valid=out["valid"], out=out, case=case)


class OracleFOInterface(FOInterface):
    def _make_sensor_model(self):
        return OracleSensorModel(lanelet_network=self.lanelet_network, ref_path=self.ego_reference_path,
                                 sensor_radius=self.sensor_radius, sensor_angle=self.sensor_angle, debug=self.debug)

    def _make_agent_manager(self):
        return OracleAgentManager(scenario=self.cr_scenario, reference_path=self.ego_reference_path,
                                  config=self.config["agent_manager"], timestep=self.timestep, dt=self.dt,
                                  debug=self.debug, fo_obstacles=self.fo_obstacles)

    def _make_metrics(self):
        return OracleMetric(self.config["metrics"], self.vehicle_params, self.agent_manager)
