"""Generates tests/golden/scene_scenario{1,2,3}.json: compact scene (what the path reads from the reference's
example_scenarios/*.xml) + the CPU oracle pipeline's results over a few planning cycles.  TEST INFRASTRUCTURE.

    python -m oracle.make_scenario_golden            # needs /root/reference (build container only)
    python -m oracle.make_scenario_golden --check    # re-run the oracle on the committed scenes and compare
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
from frenetix_occlusion_b200.scenario import load_commonroad_xml, scenario_from_dict, scenario_to_dict  # noqa: E402
from oracle.pipeline_oracle import OracleFOInterface  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SCENARIOS = "/root/reference/example_scenarios"
# scenario -> (time steps replayed, configured real agents); the truck/bicycle of occlusion.yaml:72-82 are
# placed for scenario1's intersection and enter at time steps 0 and 6
PLAN = {"scenario1": ([0, 6, 12, 18, 24], "default"), "scenario2": ([0, 10, 20, 30], None),
        "scenario3": ([0, 5, 10, 20], None), "parked_car": ([0, 5, 10, 15], None)}
FAN = {"speed_factors": np.linspace(0.0, 1.3, 14).tolist(), "lateral_targets": np.linspace(-1.5, 1.5, 7).tolist()}


def parked_car_scene() -> dict:
    """Synthetic scene (not from the reference) that exercises the static-obstacle finder, which none of the three
    example scenarios reaches: a straight two-lane road, a car parked half on the right lane 20 m ahead."""
    xs = np.linspace(-20, 100, 61)
    right = {"id": 1, "left": [[float(x), 0.0] for x in xs], "right": [[float(x), -3.5] for x in xs], "pred": [], "succ": [],
             "adj_left": [2, False], "adj_right": [None, None], "type": ["urban"]}
    left = {"id": 2, "left": [[float(x), 0.0] for x in xs[::-1]], "right": [[float(x), 3.5] for x in xs[::-1]], "pred": [],
            "succ": [], "adj_left": [1, False], "adj_right": [None, None], "type": ["urban"]}
    car = {"id": 7, "type": "parkedVehicle", "role": "static", "shape": [4.5, 1.8, 0.0, 0.0, 0.0],
           "initial": [20.0, -2.9, 0.0, 0.0, 0], "states": None}
    return {"dt": 0.1, "scenario_id": "parked_car", "lanelets": [right, left], "intersections": [], "obstacles": [car],
            "planning_problem": {"id": 1, "initial": [0.0, -1.75, 0.0, 8.0, 0], "goal_lanelet": 1}}


def run_oracle(scene: dict, timesteps, agents):
    random.seed(7)
    sc = scenario_from_dict(scene)
    ego = R.OpenLoopEgo(sc)
    cfg = R.deployment_config(agents=agents)
    fo = OracleFOInterface(sc, ego.reference_path, R.DEFAULT_VEHICLE, sc.dt, config_path=cfg)
    recs = R.replay(fo, ego, timesteps, fan_kwargs=FAN)
    out = []
    for r in recs:
        o = r["result"].out
        preds = [{"agent_type": r["agent_types"][k], "pos0": np.round(p["pos_list"][0], 5).tolist(),
                  "pos_end": np.round(p["pos_list"][-1], 5).tolist(), "n": len(p["pos_list"]),
                  "yaw0": round(float(p["orientation_list"][0]), 6), "v0": round(float(p["v_list"][0]), 5)}
                 for k, p in r["predictions"].items()]
        out.append({"timestep": r["timestep"], "ego": [float(r["ego"]["pos"][0]), float(r["ego"]["pos"][1]), r["ego"]["orientation"]],
                    "visible_obstacles": [int(v) if v < 10000 else "real_agent" for v in r["visible_obstacles"]],
                    "spawn_points": [{"agent_type": s["agent_type"], "source": s["source"],
                                      "position": np.round(s["position"], 4).tolist(),
                                      "orientation": None if s["orientation"] is None else round(float(s["orientation"]), 6)}
                                     for s in r["spawn_points"]],
                    "predictions": preds,
                    "valid": "".join("1" if v else "0" for v in r["valid"]),
                    "max_obst_harm_with_cp_all": (np.round(o["max_obst_harm_with_cp_all"], 9).tolist()
                                                  if "max_obst_harm_with_cp_all" in o else None)})
    return out


def main():
    check = "--check" in sys.argv
    for name, (timesteps, agents) in PLAN.items():
        path = os.path.join(GOLDEN, f"scene_{name}.json")
        if check:
            with open(path) as f:
                doc = json.load(f)
            got = run_oracle(doc["scene"], doc["timesteps"], doc["agents"])
            assert json.loads(json.dumps(got)) == doc["cycles"], f"{name}: oracle no longer reproduces the golden"
            print(name, "ok")
            continue
        if name == "parked_car":
            scene = scenario_to_dict(scenario_from_dict(parked_car_scene()))
        else:
            scene = scenario_to_dict(load_commonroad_xml(os.path.join(REF_SCENARIOS, name + ".xml")))
        scene = json.loads(json.dumps(scene))
        cycles = run_oracle(scene, timesteps, agents)
        doc = {"source": ("synthetic scene of oracle/make_scenario_golden.py" if name == "parked_car" else
                          f"example_scenarios/{name}.xml of the reference, via oracle/make_scenario_golden.py"),
               "timesteps": timesteps, "agents": agents, "fan": FAN, "scene": scene, "cycles": cycles}
        with open(path, "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        print(name, os.path.getsize(path), "bytes;", [(c["timestep"], len(c["spawn_points"]), c["valid"].count("1")) for c in cycles])


if __name__ == "__main__":
    main()
