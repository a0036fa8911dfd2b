"""Agent manager and phantom/real agents (mirror of reference agent.py:27-537).

Same public surface (``FOAgentManager.add_agent / reset / update_real_agents /
agent_by_prediction_id``, ``predictions`` dict in the reference's layout, id scheme
``int(str(agent_id) + str(i))``); the rollouts run in the CUDA kernels ``fo_rollout_cv`` /
``fo_rollout_path``.  Predictions are brought back to the host as float64 numpy arrays because the
reference's plugin protocol hands them to arbitrary user metrics; the dense core re-uploads them once
per planning cycle (a few kB)."""
from __future__ import annotations

import types
from random import randint

import numpy as np
import torch

from .prediction import rollout_cv, rollout_path
from .route_planner import FORoutePlanner
from .scenario import Obstacle, Rectangle, State


def calc_normal_vector_to_curve(curve, pos):
    """helper_functions.py:38-57: unit vector from ``pos`` to the closest point of the polyline."""
    c = np.asarray(curve, dtype=np.float64)
    pos = np.asarray(pos, dtype=np.float64)
    seg = c[1:] - c[:-1]
    l2 = np.maximum((seg ** 2).sum(1), 1e-18)
    u = np.clip(((pos - c[:-1]) * seg).sum(1) / l2, 0.0, 1.0)
    q = c[:-1] + u[:, None] * seg
    j = int(np.argmin(np.hypot(*(pos - q).T)))
    v = q[j] - pos
    return v / np.linalg.norm(v)


def angle_between_positive(v1, v2):
    """helper_functions.py:60-69."""
    v1 = np.asarray(v1, dtype=np.float64) / np.linalg.norm(v1)
    v2 = np.asarray(v2, dtype=np.float64) / np.linalg.norm(v2)
    a = np.arctan2(v1[0] * v2[1] - v1[1] * v2[0], float(np.dot(v1, v2)))
    return a + 2 * np.pi if a < 0 else a


class OAPAgent:
    def __init__(self, pos, velocity, agent_type, agent_params, scenario, config, dt=0.1, horizon=3.0,
                 visualization=None, debug=False, device="cuda:0"):
        self.cr_scenario = scenario
        self.dt = dt
        self.horizon = horizon
        self.agent_id = agent_params["agent_id"]
        self.visualization = visualization
        self.debug = debug
        self.initial_position = np.asarray(pos, dtype=np.float64)
        self.initial_velocity = velocity
        self.agent_type = agent_type
        self.obstacle_type = types.SimpleNamespace(value=agent_type.lower(), name=agent_type.upper())
        self.agent_params = agent_params
        self.shape = Rectangle(agent_params["length"], agent_params["width"], center=np.array([0.0, 0.0]), orientation=0.0)
        self.initial_state = None
        self.predictions = None
        self.config = config
        self.device = device
        self.commonroad_dynamic_obstacle = None

    def _buffered_shape(self):
        """agent.py:402-407 / 525-526: Bicycle uses the *_l factors, everything else *_s."""
        p = self.config["prediction"]
        if self.agent_type == "Bicycle":
            return {"length": self.shape.length * p["size_factor_length_l"], "width": self.shape.width * p["size_factor_width_l"]}
        return {"length": self.shape.length * p["size_factor_length_s"], "width": self.shape.width * p["size_factor_width_s"]}

    # ---- device calls (the only places the agents touch the C ABI); host float64 arrays come back -------------
    # The kernels store float32: positions are rolled out RELATIVE to the agent's start (origin = initial position,
    # subtracted and added back in float64 on the host), so world coordinates of 1e3 .. 1e4 m keep sub-micrometre
    # resolution instead of the 6e-5 .. 1e-3 m of an absolute float32 coordinate.
    @staticmethod
    def _host_rollout(packed, origins, sample=None):
        """ONE read-back of a packed [6, J, T] rollout (x, y, yaw, v, var_x, var_y); origins [J, 2] added in float64."""
        h = packed.cpu().numpy().astype(np.float64)                    # .cpu() synchronises
        org = np.asarray(origins, dtype=np.float64).reshape(-1, 2)
        out = {"x": h[0] + org[:, :1], "y": h[1] + org[:, 1:], "yaw": h[2], "v": h[3], "var": h[4]}
        if sample is not None:
            out["sample"] = sample.cpu().numpy()
        return out

    def _rollout_cv(self, pos, velocity, phi, var_factor):
        org = (float(pos[0]), float(pos[1]))
        ro = rollout_cv([pos[0]], [pos[1]], [velocity], [phi], self.dt, self.horizon, 0.1, var_factor, origin=org,
                        device=self.device)
        return self._host_rollout(ro["packed"], [org])

    def _rollout_path(self, paths, pos, velocity, var_factor):
        J = len(paths)
        org = (float(pos[0]), float(pos[1]))
        ro = rollout_path(paths, [pos[0]] * J, [pos[1]] * J, [velocity] * J, self.dt, self.horizon, 3.0, 0.1, var_factor,
                          origin=org, device=self.device)
        return self._host_rollout(ro["packed"], [org] * J, ro["sample"])

    def _prediction_dict(self, ro, j, n):
        x, y = ro["x"][j, :n], ro["y"][j, :n]
        var = ro["var"][j, :n]
        cov = np.zeros((len(var), 2, 2))
        cov[:, 0, 0] = cov[:, 1, 1] = var
        return {"orientation_list": np.array(ro["yaw"][j, :n]), "v_list": np.array(ro["v"][j, :n]),
                "pos_list": np.column_stack((x, y)), "shape": self._buffered_shape(), "cov_list": cov}

    def add_to_commonroad_scenario(self, timestep=0):
        """agent.py:225-252: the agent becomes a dynamic obstacle whose trajectory starts at timestep + 1.  With
        commonroad-io installed (the planner's environment) real commonroad objects are built, exactly as the
        reference does; otherwise the light stand-ins of ``scenario.py`` (replay harness)."""
        pred = self._full_prediction
        n = len(pred["pos_list"])
        try:
            if not type(self.cr_scenario).__module__.startswith("commonroad"):
                raise ImportError("scenario is not a commonroad-io object")
            from commonroad.geometry.shape import Rectangle as CRRectangle
            from commonroad.prediction.prediction import TrajectoryPrediction
            from commonroad.scenario.obstacle import DynamicObstacle, ObstacleType
            from commonroad.scenario.state import CustomState, InitialState
            from commonroad.scenario.trajectory import Trajectory
        except ImportError:
            init = State(position=self.initial_state.position, orientation=self.initial_state.orientation,
                         velocity=self.initial_state.velocity, time_step=timestep)
            states = [State(position=pred["pos_list"][i], orientation=pred["orientation_list"][i],
                            velocity=pred["v_list"][i], time_step=timestep + i) for i in range(1, n)]
            self.commonroad_dynamic_obstacle = Obstacle(self.agent_id, self.agent_type.lower(), "dynamic", self.shape, init,
                                                        states)
        else:
            shape = CRRectangle(self.shape.length, self.shape.width, center=np.array([0.0, 0.0]), orientation=0.0)
            init = InitialState(position=np.asarray(self.initial_state.position), orientation=self.initial_state.orientation,
                                velocity=self.initial_state.velocity, time_step=timestep)
            states = [CustomState(position=np.asarray(pred["pos_list"][i]), orientation=float(pred["orientation_list"][i]),
                                  velocity=float(pred["v_list"][i]), time_step=timestep + i) for i in range(1, n)]
            trajectory = Trajectory(initial_time_step=timestep + 1, state_list=states)
            self.commonroad_dynamic_obstacle = DynamicObstacle(
                obstacle_id=self.agent_id, obstacle_type=ObstacleType(self.agent_type.lower()), obstacle_shape=shape,
                initial_state=init, prediction=TrajectoryPrediction(trajectory=trajectory, shape=shape))
        self.cr_scenario.add_objects(self.commonroad_dynamic_obstacle)


class OAPPedestrianAgent(OAPAgent):
    def __init__(self, pos, velocity, agent_type, agent_params, scenario, config, dt=0.1, horizon=3.0, ref_path=None,
                 visualization=None, debug=False, mode="ref_path", orientation=None, device="cuda:0", defer=False):
        super().__init__(pos, velocity, agent_type, agent_params, scenario, config, dt, horizon, visualization, debug, device)
        self.ego_reference_path = ref_path
        self.reference_curve = None
        phi = self._initial_orientation(mode, orientation)
        self.initial_state = State(position=self.initial_position, orientation=phi, velocity=self.initial_velocity, time_step=0)
        # constant-velocity rollout on the device (agent.py:487-503); ``defer``: the manager rolls out all agents of a
        # planning cycle in one launch and calls ``_finish``
        self._job = ("cv", float(pos[0]), float(pos[1]), float(velocity), float(phi))
        if not defer:
            self._finish(self._rollout_cv(pos, velocity, phi, config["prediction"]["variance_factor"]), 0)

    n_jobs = 1

    def _finish(self, ro, j0):
        n = int(self.horizon / self.dt) + 1
        self._full_prediction = self._prediction_dict(ro, j0, n)
        self.trajectory = self._full_prediction["pos_list"]
        self.predictions = self._create_cr_predictions(0)

    def _initial_orientation(self, mode, orientation):
        """agent.py:453-484: towards the lane centre / the ego reference path unless the spawn point fixes it."""
        if mode == "lane_center":
            ids = self.cr_scenario.lanelet_network.find_lanelet_by_position([self.initial_position])[0]
            curve = np.array(self.cr_scenario.lanelet_network.find_lanelet_by_id(ids[0]).center_vertices) if ids \
                else self.ego_reference_path            # (the reference falls through to None here, agent.py:463-465)
        elif mode == "ref_path":
            curve = self.ego_reference_path
        else:
            raise NotImplementedError(f'Selected mode "{mode}" is not implemented: use "ref_path" or "lane_center"!')
        if orientation is not None:
            return float(orientation)
        self.reference_curve = curve
        return float(angle_between_positive(np.array([1, 0]), calc_normal_vector_to_curve(curve, self.initial_position)))

    def _create_cr_predictions(self, timestep) -> list:
        """agent.py:520-536: the covariances are rebuilt for the SLICED position list (create_cov_matrix on
        pos_list[timestep:]), i.e. the variance restarts at 0.1 -- the first len(slice) entries of the full list."""
        p = self._full_prediction
        m = len(p["pos_list"][timestep:])
        return [{"orientation_list": p["orientation_list"][timestep:], "v_list": p["v_list"][timestep:],
                 "pos_list": p["pos_list"][timestep:], "shape": p["shape"], "cov_list": p["cov_list"][:m]}]


class OAPVehicleAgent(OAPAgent):
    def __init__(self, pos, velocity, agent_type, agent_params, scenario, config, dt=0.1, horizon=3.0,
                 visualization=None, debug=False, device="cuda:0", defer=False):
        super().__init__(pos, velocity, agent_type, agent_params, scenario, config, dt, horizon, visualization, debug, device)
        self.route_planner = FORoutePlanner(scenario, scenario.lanelet_network, visualization, debug)
        self.reference_paths = self.route_planner.calc_possible_reference_paths(pos)
        self.initial_state = State(position=self.initial_position, orientation=self.route_planner.lanelet_orientation,
                                   velocity=self.initial_velocity, time_step=0)
        self._windows = [self._path_window(p, pos, velocity, horizon) for p in self.reference_paths]
        self.n_jobs = len(self._windows)
        self._job = ("path", float(pos[0]), float(pos[1]), float(velocity))
        if not defer:
            self._finish(self._rollout_path(self._windows, pos, velocity, config["prediction"]["variance_factor"]), 0)

    def _finish(self, ro, j0):
        J = self.n_jobs
        n = int(self.horizon / self.dt) + 1
        smp = ro["sample"]
        valid = [j for j in range(J) if smp[j0 + j] >= 0]
        self._all_predictions = [self._prediction_dict(ro, j0 + j, n) for j in valid]
        # the trajectory a REAL agent drives (agent.py:348-362 with handler=None): the route whose reference path goes
        # straightest, i.e. the smallest variance of the path heading
        self._full_prediction = None
        if valid:
            hv = self.route_planner.heading_variances
            best = min(range(len(valid)), key=lambda q: hv[valid[q]])
            self._full_prediction = self._all_predictions[best]
        self.predictions = self._create_cr_predictions(0)

    @staticmethod
    def _path_window(path, pos, velocity, horizon, max_points=1024):
        """The rollout kernel stages at most 1024 polyline points per job: keep the part of a long route the agent can
        reach within its horizon (10 m behind the projection of its start, 1.3 x the distance at the fastest sampled
        end speed plus 20 m ahead)."""
        path = np.asarray(path, dtype=np.float64)
        if len(path) <= max_points:
            return path
        cum = np.concatenate(([0.0], np.cumsum(np.hypot(*np.diff(path, axis=0).T))))
        k0 = int(np.argmin(np.hypot(*(path - np.asarray(pos, dtype=np.float64)).T)))
        lo = int(np.searchsorted(cum, cum[k0] - 10.0))
        hi = int(np.searchsorted(cum, cum[k0] + 1.3 * 1.2 * float(velocity) * float(horizon) + 20.0)) + 1
        lo = max(min(lo, len(path) - 2), 0)
        hi = min(max(hi, lo + 2), len(path), lo + max_points)
        return path[lo:hi]

    def _create_cr_predictions(self, timestep) -> list:
        """agent.py:398-426: one prediction per route reference path; covariances restart at 0.1 for the sliced list
        (create_cov_matrix(pos_list[timestep:]), agent.py:414-416)."""
        return [{"orientation_list": p["orientation_list"][timestep:], "v_list": p["v_list"][timestep:],
                 "pos_list": p["pos_list"][timestep:], "shape": p["shape"],
                 "cov_list": p["cov_list"][:len(p["pos_list"][timestep:])]}
                for p in self._all_predictions]

    @staticmethod
    def get_lanelet_velocity(pos, scenario):
        """agent.py:324-346: speed from the type of the lanelet the agent stands on."""
        import warnings
        lanelet_id = scenario.lanelet_network.find_lanelet_by_position([pos])[0][0]
        lanelet = scenario.lanelet_network.find_lanelet_by_id(lanelet_id)
        lanelet_type = list(lanelet.lanelet_type)[0].value
        if lanelet_type == "urban":
            return 30 / 3.6
        if lanelet_type == "country":
            return 80 / 3.6
        if lanelet_type == "highway":
            return 100 / 3.6
        warnings.warn(f'OAPAgent: Lanelet type "{lanelet_type}" is not implemented yet! Default value of 50 km/h will be used!')
        return 50 / 3.6


class FOAgentManager:
    pedestrian_cls = OAPPedestrianAgent
    vehicle_cls = OAPVehicleAgent

    def __init__(self, scenario, reference_path, config, timestep, visualization=None, dt=0.1, fo_obstacles=None,
                 debug=False, device="cuda:0"):
        self.scenario = scenario
        self.timestep = timestep
        self.reference_path = reference_path
        self.config = config
        self.visualization = visualization
        self.fo_obstacles = fo_obstacles
        self.dt = dt
        self.debug = debug
        self.device = device
        self.phantom_agents = []
        self.real_agents = []
        self.predictions = {}
        self.version = 0
        self._get_all_ids()

    def reset(self):
        self.phantom_agents = []
        self.predictions = {}
        self.version += 1

    _PARAMS = {"bicycle": ("Bicycle", 4, 11.5), "car": ("Car", 7.32, 11.5), "truck": ("Truck", 7.32, 7)}

    def add_agent(self, pos, velocity="default", agent_type="Car", add_to_scenario=False, timestep=0, horizon=3.0,
                  mode="ref_path", orientation=None, _defer=False):
        """agent.py:48-157 (same defaults, same errors)."""
        if self.timestep != timestep:
            return
        agent_id = self._create_id()
        key = agent_type.lower()
        if key in self._PARAMS:
            conf = self.config[key]
            name, sw, amax = self._PARAMS[key]
            agent_params = {"agent_id": agent_id, "type": name, "width": conf["width"], "length": conf["length"],
                            "delta_max": 1.066, "wheelbase": conf["wheelbase"], "switching_velocity": sw,
                            "max_acceleration": amax, "velocity_delta_max": 0.4}
            if velocity == "default":
                velocity = conf["default_velocity"]
            elif velocity == "lanelet":
                if key == "bicycle":
                    raise NotImplementedError
                velocity = OAPVehicleAgent.get_lanelet_velocity(pos=pos, scenario=self.scenario)
            elif not isinstance(velocity, (float, int)):
                raise ValueError('Only "default", "lanelet", int or float is allowed!')
        elif key == "pedestrian":
            conf = self.config["pedestrian"]
            agent_params = {"agent_id": agent_id, "type": "Pedestrian", "width": conf["width"], "length": conf["length"]}
        else:
            raise NotImplementedError(f'OAPManager: Agent type "{agent_type}" is not implemented!')

        if agent_type == "Pedestrian":
            if velocity == "default":
                velocity = conf["default_velocity"]
            elif velocity == "lanelet":
                raise NotImplementedError
            elif not isinstance(velocity, (float, int)):
                raise ValueError('Only "default", int or float is allowed!')
            agent = self.pedestrian_cls(pos=pos, velocity=velocity, agent_type=agent_type, agent_params=agent_params,
                                       scenario=self.scenario, dt=self.dt, horizon=horizon, ref_path=self.reference_path,
                                       visualization=self.visualization, debug=self.debug, mode=mode,
                                       orientation=orientation, config=self.config, device=self.device, defer=_defer)
        else:
            agent = self.vehicle_cls(pos=pos, velocity=velocity, agent_type=agent_type, agent_params=agent_params,
                                    scenario=self.scenario, dt=self.dt, horizon=horizon, visualization=self.visualization,
                                    debug=self.debug, config=self.config, device=self.device, defer=_defer)
        if _defer:
            self.phantom_agents.append(agent)
            return agent
        if add_to_scenario:
            self.real_agents.append(agent)
            agent.add_to_commonroad_scenario(timestep=timestep)
            if self.fo_obstacles is not None:
                self.fo_obstacles.add(agent.commonroad_dynamic_obstacle)
        else:
            self.phantom_agents.append(agent)
            self._add_prediction(agent)
        return agent

    def add_agents(self, specs):
        """Phantom agents of one planning cycle at once: the same agents, ids and predictions as one ``add_agent`` call
        per entry (``specs``: dicts of its keyword arguments), but ONE ``fo_rollout_cv`` launch for all pedestrians, ONE
        ``fo_rollout_path`` launch for all vehicle routes and one read-back each instead of a launch and a blocking
        copy per agent."""
        specs = [dict(s) for s in specs]
        if any(s.get("add_to_scenario") for s in specs) or len({float(s.get("horizon", 3.0)) for s in specs}) > 1:
            return [self.add_agent(**s) for s in specs]
        agents = [self.add_agent(_defer=True, **s) for s in specs]
        live = [a for a in agents if a is not None]
        vf = self.config["prediction"]["variance_factor"]
        peds = [a for a in live if a._job[0] == "cv"]
        vehs = [a for a in live if a._job[0] == "path" and a.n_jobs > 0]
        ro_p = ro_v = None
        if peds:
            j = np.array([a._job[1:] for a in peds])
            ro_p = rollout_cv(j[:, 0], j[:, 1], j[:, 2], j[:, 3], self.dt, peds[0].horizon, 0.1, vf, origin=j[:, :2], device=self.device)
        if vehs:
            paths = [w for a in vehs for w in a._windows]
            j = np.array([a._job[1:] for a in vehs for _ in range(a.n_jobs)])
            ro_v = rollout_path(paths, j[:, 0], j[:, 1], j[:, 2], self.dt, vehs[0].horizon, 3.0, 0.1, vf, origin=j[:, :2],
                                device=self.device)
            host_v = OAPAgent._host_rollout(ro_v["packed"], j[:, :2], ro_v["sample"])
        if peds:
            host_p = OAPAgent._host_rollout(ro_p["packed"], np.array([a._job[1:3] for a in peds]))
        jp = jv = 0
        for a in live:                                  # finish in creation order: prediction ids keep their order
            if a._job[0] == "cv":
                a._finish(host_p, jp)
                jp += 1
            else:
                a._finish(host_v if a.n_jobs else {"sample": np.zeros(0, np.int32)}, jv)
                jv += a.n_jobs
            self._add_prediction(a)
        return agents

    def agent_by_prediction_id(self, prediction_id):
        if not self.phantom_agents:
            return
        agent_id = int(str(prediction_id)[:5])
        for agent in self.phantom_agents:
            if agent.agent_id == agent_id:
                return agent

    def update_real_agents(self, cr_scenario_predictions):
        """agent.py:171-177: overwrite the planner's prediction of real pedestrians (mutates the caller's dict)."""
        for agent in self.real_agents:
            if agent.agent_type.lower() == "pedestrian":
                if agent.agent_id in cr_scenario_predictions:
                    agent.predictions = agent._create_cr_predictions(self.timestep)
                    cr_scenario_predictions[agent.agent_id] = agent.predictions[0]

    def _add_prediction(self, agent):
        for i, prediction in enumerate(agent.predictions):
            self.predictions[int(str(agent.agent_id) + str(i))] = prediction
        self.version += 1

    def _get_all_ids(self):
        self.all_obstacle_id = [o.obstacle_id for o in self.scenario.obstacles]

    def _create_id(self):
        agent_id = randint(10000, 11000)
        if agent_id in self.all_obstacle_id:
            agent_id = self._create_id()
        else:
            self.all_obstacle_id.append(agent_id)
        return agent_id
