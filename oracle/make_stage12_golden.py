"""Generates tests/golden/stage12_reference.json by running the reference's own agent.py / fo_obstacle.py /
helper_functions.py VERBATIM (oracle/ref_stage12.py) on the committed scenario1 scene.  TEST INFRASTRUCTURE.

    python -m oracle.make_stage12_golden            # needs /root/reference (build container only)
    python -m oracle.make_stage12_golden --check
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from frenetix_occlusion_b200 import replay as R  # noqa: E402
from frenetix_occlusion_b200.scenario import scenario_from_dict  # noqa: E402
from oracle import ref_stage12  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
OUT = os.path.join(GOLDEN, "stage12_reference.json")

# (position, velocity, mode, orientation, horizon) of pedestrian phantoms on the scenario1 scene
PEDESTRIANS = [([14.4, 12.4], "default", "ref_path", None, 3.0), ([22.5, 0.6], 1.4, "ref_path", None, 3.0),
               ([10.0, 0.3], 2.25, "lane_center", None, 3.0), ([31.95, -2.24], "default", "ref_path", 1.234567, 3.0),
               ([20.0, 3.0], 0.7, "lane_center", -2.5, 5.0), ([-8.0, 1.5], 1.4, "ref_path", None, 5.0),
               ([12.0, -3.0], 1.4, "ref_path", float(np.pi / 2), 3.0)]
OBSTACLE_STEPS = [0, 1, 2, 5, 50, 146, 147, 148, 149, 400]


def _r(a, nd=9):
    return np.round(np.asarray(a, dtype=np.float64), nd).tolist()


def run():
    ref_agent, ref_obst, ref_hf = ref_stage12.reference_modules()
    with open(os.path.join(GOLDEN, "scene_scenario1.json")) as f:
        scene = json.load(f)["scene"]
    sc = scenario_from_dict(scene)
    ego = R.OpenLoopEgo(sc)
    cfg = R.deployment_config()["agent_manager"]
    out = {"source": "reference agent.py / utils/fo_obstacle.py / utils/helper_functions.py run verbatim over "
                     "oracle/ref_stage12.py on tests/golden/scene_scenario1.json", "pedestrians": [], "obstacles": [],
           "projection": [], "helpers": []}
    # ---- P1 / P3: FOAgentManager.add_agent -> OAPPedestrianAgent (agent.py:48-157, 428-536) ---------------------
    random.seed(11)
    am = ref_agent.FOAgentManager(scenario=sc, reference_path=ego.reference_path, config=cfg, timestep=0, dt=sc.dt)
    for pos, vel, mode, ori, hor in PEDESTRIANS:
        n_before = len(am.predictions)
        am.add_agent(pos=np.array(pos), velocity=vel, agent_type="Pedestrian", timestep=0, horizon=hor, mode=mode,
                     orientation=ori)
        pid = list(am.predictions.keys())[n_before]
        p = am.predictions[pid]
        ag = am.agent_by_prediction_id(pid)
        assert ag is am.phantom_agents[-1] and pid == int(str(ag.agent_id) + "0")
        out["pedestrians"].append({"pos": pos, "velocity": vel, "mode": mode, "orientation": ori, "horizon": hor,
                                   "pos_list": _r(p["pos_list"]), "v_list": _r(p["v_list"]),
                                   "orientation_list": _r(p["orientation_list"]), "cov_diag": _r(np.asarray(p["cov_list"])[:, 0, 0], 12),
                                   "cov_offdiag_max": float(np.abs(np.asarray(p["cov_list"])[:, 0, 1]).max()),
                                   "shape": p["shape"], "agent_shape": [ag.shape.length, ag.shape.width]})
    # a configured REAL pedestrian enters the scenario as a dynamic obstacle (agent.py:142-151, 225-252, 507-518)
    am.add_agent(pos=np.array([3.0, 3.0]), velocity=1.0, agent_type="Pedestrian", add_to_scenario=True, timestep=0,
                 horizon=5.0, orientation=0.5)
    dyn = am.real_agents[-1].commonroad_dynamic_obstacle
    st = dyn.prediction.trajectory.state_list
    out["real_pedestrian"] = {"initial_time_step": dyn.prediction.trajectory.initial_time_step, "n_states": len(st),
                              "first": _r(st[0].position), "last": _r(st[-1].position),
                              "time_steps": [int(st[0].time_step), int(st[-1].time_step)]}
    # ---- V4: FOObstacle.update_at_timestep / hf.calc_corner_points (fo_obstacle.py:79-116) ------------------------
    sc2 = scenario_from_dict(scene)
    for ob in sc2.obstacles:
        fo = ref_obst.FOObstacle(ob)
        rows = []
        for ts in OBSTACLE_STEPS:
            fo.update_at_timestep(ts)
            rows.append(None if fo.current_pos is None else
                        {"pos": _r(fo.current_pos), "orientation": float(fo.current_orientation),
                         "corners": _r(fo.current_corner_points)})
        out["obstacles"].append({"id": ob.obstacle_id, "steps": rows})
    # ---- V3: shadow corner pair / shadow quads (helper_functions.py:79-96, 139-176) -----------------------------------
    rng = np.random.default_rng(5)
    for _ in range(40):
        ego_pos = rng.uniform(-5, 5, 2)
        c = rng.uniform(-30, 30, 2)
        while np.hypot(*(c - ego_pos)) < 6.0:
            c = rng.uniform(-30, 30, 2)
        yaw, hl, hw = rng.uniform(-np.pi, np.pi), rng.uniform(0.2, 5.0), rng.uniform(0.2, 1.5)
        loc = np.array([[-hl, -hw], [-hl, hw], [hl, hw], [hl, -hw]])
        cs, sn = np.cos(yaw), np.sin(yaw)
        corners = loc @ np.array([[cs, -sn], [sn, cs]]).T + c
        poly, c1, c2 = ref_hf.get_polygon_from_obstacle_occlusion(ego_pos, corners)
        quad = ref_hf.create_polygon_from_vertices(corners[0], corners[1], ego_pos)
        out["projection"].append({"ego": _r(ego_pos), "corners": _r(corners), "c1": _r(c1), "c2": _r(c2),
                                  "shadow": _r(poly.exterior_coords), "edge_shadow": _r(quad.exterior_coords)})
    for _ in range(20):
        v1, v2 = rng.normal(size=2), rng.normal(size=2)
        curve = np.cumsum(rng.uniform(0.2, 2.0, (12, 2)), axis=0)
        pos = rng.uniform(0, 12, 2)
        out["helpers"].append({"v1": _r(v1), "v2": _r(v2), "angle_between_positive": float(ref_hf.angle_between_positive(v1, v2)),
                               "curve": _r(curve), "pos": _r(pos),
                               "normal_vector": _r(ref_hf.calc_normal_vector_to_curve(curve, pos)),
                               "vector_from_angle": _r(ref_hf.vector_from_angle(float(v1[0])))})
    # ---- P2: the 3 x 3 Frenet sampling matrix of a vehicle phantom (frenetix_handler.py:78-105, utils/sampling.py) ----
    from frenetix_occlusion.utils.sampling import SamplingHandler, generate_sampling_matrix
    sh = SamplingHandler(dt=0.1, max_sampling_number=1, t_min=2.0, horizon=3.0, delta_d_max=0.5, delta_d_min=-0.5)
    v0, s0, d0 = 10.0, 12.5, 0.3
    sm = generate_sampling_matrix(t0_range=0.0, t1_range=3.0, s0_range=s0, ss0_range=v0, sss0_range=0,
                                  ss1_range=np.array([v0 * 0.8, v0, v0 * 1.2]), sss1_range=0, d0_range=d0, dd0_range=0,
                                  ddd0_range=0, d1_range=np.array(list(sh.d_sampling.to_range(0))), dd1_range=0.0,
                                  ddd1_range=0.0)
    out["sampling_matrix"] = {"v0": v0, "s0": s0, "d0": d0, "rows": _r(sm)}
    return out


def main():
    got = json.loads(json.dumps(run()))
    if "--check" in sys.argv:
        with open(OUT) as f:
            assert got == json.load(f), "the reference run no longer reproduces tests/golden/stage12_reference.json"
        print("stage12 golden ok")
        return
    with open(OUT, "w") as f:
        json.dump(got, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
