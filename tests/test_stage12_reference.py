"""Stages 1-2 against the reference's OWN code: ``tests/golden/stage12_reference.json`` was produced by running the
reference's agent.py / utils/fo_obstacle.py / utils/helper_functions.py verbatim (oracle/ref_stage12.py,
oracle/make_stage12_golden.py).  CPU: the float64 oracle restatements and the host mirrors reproduce it.  GPU: the
product's agent manager (CUDA rollout kernel) does."""
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN

from oracle import visibility_oracle as VO


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "stage12_reference.json")) as f:
        return json.load(f)


def _scene():
    from frenetix_occlusion_b200.scenario import scenario_from_dict
    with open(os.path.join(GOLDEN, "scene_scenario1.json")) as f:
        return scenario_from_dict(json.load(f)["scene"])


def test_cv_oracle_equals_reference_pedestrian_agent(gold):
    """VO.rollout_cv == OAPPedestrianAgent._create_ped_trajectory + _create_cr_predictions (agent.py:451-536)."""
    for p in gold["pedestrians"]:
        pos = np.asarray(p["pos_list"])
        phi, v = p["orientation_list"][0], p["v_list"][0]
        r = VO.rollout_cv([pos[0, 0]], [pos[0, 1]], [v], [phi], 0.1, p["horizon"], 0.1, 1.05)
        assert r["x"].shape[1] == len(pos) == int(p["horizon"] / 0.1) + 1
        np.testing.assert_allclose(np.stack((r["x"][0], r["y"][0]), -1), pos, rtol=0, atol=1e-9)
        np.testing.assert_allclose(r["var"][0], p["cov_diag"], rtol=0, atol=1e-11)      # golden stored to 12 decimals
        assert p["cov_offdiag_max"] == 0.0
        assert np.allclose(p["shape"]["length"], p["agent_shape"][0] * 1.2) and np.allclose(p["shape"]["width"], p["agent_shape"][1] * 1.3)


def test_shadow_construction_equals_reference(gold):
    """identify_projection_points / the shadow quads used by the point-wise visibility oracle == helper_functions.py."""
    for c in gold["projection"]:
        ego, cor = np.asarray(c["ego"]), np.asarray(c["corners"])
        c1, c2 = VO.identify_projection_points(ego, cor)
        assert np.allclose(c1, c["c1"], atol=1e-9) and np.allclose(c2, c["c2"], atol=1e-9)
        u1, u2 = (c1 - ego) / np.linalg.norm(c1 - ego), (c2 - ego) / np.linalg.norm(c2 - ego)
        np.testing.assert_allclose(np.array([c1, c2, c2 + u2 * 100, c1 + u1 * 100]), np.asarray(c["shadow"])[:4], atol=1e-7)
        v1, v2 = cor[0], cor[1]
        quad = np.array([v1, v2, v2 + 100 * (v2 - ego), v1 + 100 * (v1 - ego)])       # as in VO.classify_points
        np.testing.assert_allclose(quad, np.asarray(c["edge_shadow"])[:4], atol=1e-7)


def test_host_helpers_equal_reference(gold):
    from frenetix_occlusion_b200.agent import angle_between_positive, calc_normal_vector_to_curve
    from frenetix_occlusion_b200.utils.helper_functions import vector_from_angle
    for h in gold["helpers"]:        # the golden stores its inputs to 9 decimals -> compare to 1e-7
        assert abs(angle_between_positive(h["v1"], h["v2"]) - h["angle_between_positive"]) < 1e-7
        np.testing.assert_allclose(calc_normal_vector_to_curve(h["curve"], h["pos"]), h["normal_vector"], atol=1e-7)
        np.testing.assert_allclose(vector_from_angle(h["v1"][0]), h["vector_from_angle"], atol=1e-7)


def test_obstacle_wrapper_equals_reference(gold):
    """FOObstacle.update_at_timestep / calc_corner_points (fo_obstacle.py:79-116, helper_functions.py:99-112)."""
    from frenetix_occlusion_b200.utils.fo_obstacle import FOObstacle
    sc = _scene()
    steps = [0, 1, 2, 5, 50, 146, 147, 148, 149, 400]
    for ob, g in zip(sc.obstacles, gold["obstacles"]):
        assert ob.obstacle_id == g["id"]
        fo = FOObstacle(ob)
        for ts, row in zip(steps, g["steps"]):
            fo.update_at_timestep(ts)
            if row is None:
                assert fo.current_pos is None and fo.current_corner_points is None
            else:
                np.testing.assert_allclose(fo.current_pos, row["pos"], atol=1e-9)
                assert abs(fo.current_orientation - row["orientation"]) < 1e-12
                np.testing.assert_allclose(fo.current_corner_points, row["corners"], atol=1e-9)
                cx, cy, yaw, hl, hw = fo.as_rect()                 # the rectangle handed to the kernels
                rect_ring = VO.rect_corners(np.array([[cx, cy, yaw, hl, hw]]))[0]
                np.testing.assert_allclose(rect_ring, row["corners"], atol=1e-9)


def test_vehicle_sample_enumeration_equals_reference(gold):
    """The 9 Frenet samples of a vehicle phantom in the reference's order (itertools.product: end speed major, lateral
    target minor; frenetix_handler.py:78-105) -- the order decides the first-minimum selection (agent.py:364-375)."""
    g = gold["sampling_matrix"]
    rows = np.asarray(g["rows"])
    assert rows.shape == (9, 13)
    for smp in range(9):        # enumeration used by VO.rollout_path and fo_rollout_path
        sd1, d1 = g["v0"] * (0.8 + 0.2 * (smp // 3)), -0.5 + 0.5 * (smp % 3)
        assert np.isclose(rows[smp, 5], sd1) and np.isclose(rows[smp, 10], d1)
        assert rows[smp, 1] == 3.0 and rows[smp, 3] == g["v0"] and rows[smp, 2] == g["s0"] and rows[smp, 7] == g["d0"]
        assert rows[smp, 4] == 0 and rows[smp, 6] == 0 and rows[smp, 8] == 0 and rows[smp, 11] == 0


@pytest.mark.gpu
def test_cuda_agent_manager_equals_reference(gold, cuda_device):
    """FOAgentManager.add_agent -> OAPPedestrianAgent on the CUDA rollout kernel == the reference's own classes."""
    from frenetix_occlusion_b200 import replay as R
    from frenetix_occlusion_b200.agent import FOAgentManager
    sc = _scene()
    ego = R.OpenLoopEgo(sc)
    random.seed(11)
    am = FOAgentManager(scenario=sc, reference_path=ego.reference_path, config=R.deployment_config()["agent_manager"],
                        timestep=0, dt=sc.dt)
    for k, p in enumerate(gold["pedestrians"]):
        am.add_agent(pos=np.array(p["pos"]), velocity=p["velocity"], agent_type="Pedestrian", timestep=0,
                     horizon=p["horizon"], mode=p["mode"], orientation=p["orientation"])
        pid = list(am.predictions.keys())[k]
        ag = am.agent_by_prediction_id(pid)
        assert ag is am.phantom_agents[-1] and pid == int(str(ag.agent_id) + "0")
        q = am.predictions[pid]
        assert set(q.keys()) == {"orientation_list", "v_list", "pos_list", "shape", "cov_list"}
        np.testing.assert_allclose(q["pos_list"], p["pos_list"], rtol=0, atol=4e-6)          # float32 table
        np.testing.assert_allclose(q["orientation_list"], p["orientation_list"], rtol=0, atol=3e-7)
        np.testing.assert_allclose(q["v_list"], p["v_list"], rtol=1e-7)
        np.testing.assert_allclose(np.asarray(q["cov_list"])[:, 0, 0], p["cov_diag"], rtol=1e-6)
        assert np.all(np.asarray(q["cov_list"])[:, 0, 1] == 0) and np.all(np.asarray(q["cov_list"])[:, 1, 0] == 0)
        assert abs(q["shape"]["length"] - p["shape"]["length"]) < 1e-12 and abs(q["shape"]["width"] - p["shape"]["width"]) < 1e-12
    # a real pedestrian becomes a dynamic obstacle whose trajectory starts one step later (agent.py:225-252)
    am.add_agent(pos=np.array([3.0, 3.0]), velocity=1.0, agent_type="Pedestrian", add_to_scenario=True, timestep=0,
                 horizon=5.0, orientation=0.5)
    dyn = am.real_agents[-1].commonroad_dynamic_obstacle
    st = dyn.prediction.trajectory.state_list
    g = gold["real_pedestrian"]
    assert dyn.prediction.trajectory.initial_time_step == g["initial_time_step"] and len(st) == g["n_states"]
    assert [int(st[0].time_step), int(st[-1].time_step)] == g["time_steps"]
    np.testing.assert_allclose(st[0].position, g["first"], atol=4e-6)
    np.testing.assert_allclose(st[-1].position, g["last"], atol=4e-6)
    # update_real_agents (agent.py:171-177, 520-534): the sliced prediction gets covariances REBUILT for the slice
    # (create_cov_matrix(pos_list[timestep:])), i.e. the variance restarts at 0.1 instead of continuing at 0.1 * vf^t
    real = am.real_agents[-1]
    am.timestep = 20
    preds = {real.agent_id: None}
    am.update_real_agents(preds)
    q = preds[real.agent_id]
    n = len(q["pos_list"])
    assert n == 51 - 20 and len(q["cov_list"]) == n and len(q["v_list"]) == n
    vf = R.deployment_config()["agent_manager"]["prediction"]["variance_factor"]
    np.testing.assert_allclose(np.asarray(q["cov_list"])[:, 0, 0], 0.1 * vf ** np.arange(n), rtol=1e-6)
    np.testing.assert_allclose(q["pos_list"][0], real._full_prediction["pos_list"][20], atol=0)


@pytest.mark.gpu
def test_rollouts_keep_float64_resolution_far_from_the_origin(cuda_device):
    """Predictions are rolled out relative to the agent's start and shifted back in float64: at world coordinates of
    1e4 m an absolute float32 position would be off by up to 1e-3 m (the rounding grid of np.round(dce, 3))."""
    from frenetix_occlusion_b200 import replay as R
    from frenetix_occlusion_b200.agent import FOAgentManager
    sc = _scene()
    ego = R.OpenLoopEgo(sc)
    random.seed(3)
    am = FOAgentManager(scenario=sc, reference_path=ego.reference_path, config=R.deployment_config()["agent_manager"],
                        timestep=0, dt=sc.dt)
    p0 = np.array([12345.678901, -9876.543211])
    am.add_agent(pos=p0, velocity=1.4, agent_type="Pedestrian", timestep=0, horizon=3.0, orientation=0.7)
    q = list(am.predictions.values())[0]
    exp = VO.rollout_cv(p0[0], p0[1], 1.4, 0.7, sc.dt, 3.0)
    np.testing.assert_allclose(q["pos_list"], np.column_stack((exp["x"][0], exp["y"][0])), rtol=0, atol=2e-6)
    np.testing.assert_array_equal(q["pos_list"][0], p0)
