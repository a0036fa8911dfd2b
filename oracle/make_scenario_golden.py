"""Generates tests/golden/scene_scenario{1,2,3}.json and scene_parked_car.json.  TEST INFRASTRUCTURE.

    python -m oracle.make_scenario_golden            # needs /root/reference (build container only)
    python -m oracle.make_scenario_golden --check    # re-run the reference on the committed inputs and compare

Every fixture holds
* ``scene``   -- what the path reads from the reference's ``example_scenarios/*.xml`` (lanelets, intersections, obstacles,
                 planning problem), compacted by the stdlib-XML loader;
* ``inputs``  -- what the external planner hands over per planning cycle: ego reference path, ego pose / curvilinear
                 position / speed per cycle.  The planner is not part of the reference repository; these come from the
                 replay harness (``frenetix_occlusion_b200/replay.py``: route centre line, open-loop ego), which produces
                 INPUTS only -- the Frenet fan that is assessed is regenerated from them by the same harness function at
                 test time;
* ``cycles``  -- the OUTPUTS of the reference's own ``FOInterface`` (``oracle/pipeline_oracle.py``: ``interface.py``,
                 ``sensor_model.py``, ``spawn_locator.py``, ``agent.py``, ``route_planner.py``, ``frenetix_handler.py``, ``metrics/*``
                 run unmodified over third-party stand-ins, configured by the reference's
                 ``configurations/simulation/occlusion.yaml``): visible obstacles, spawn points with their integer indices,
                 phantom predictions, and the validity mask of ``trajectory_safety_assessment`` over the fan.
No class of ``frenetix_occlusion_b200`` takes part in computing ``cycles``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pipeline_oracle as PO  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SCENARIOS = "/root/reference/example_scenarios"
# scenario -> (time steps replayed, configured real agents); the truck/bicycle of occlusion.yaml:72-82 are
# placed for scenario1's intersection and enter at time steps 0 and 6
PLAN = {"scenario1": ([0, 6, 12, 18, 24], "default"), "scenario2": ([0, 10, 20, 30], None),
        "scenario3": ([0, 5, 10, 20], None), "parked_car": ([0, 5, 10, 15], None)}
FAN = {"speed_factors": np.linspace(0.0, 1.3, 14).tolist(), "lateral_targets": np.linspace(-1.5, 1.5, 7).tolist()}
VEHICLE = {"length": 4.508, "width": 1.61, "mass": 1093.3, "wb_rear_axle": 1.4227, "a_max": 11.5}


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


def spawn_indices(position, reference_path, obstacle_positions):
    """Integer indices of a spawn point: the sample of the ego reference path closest to it and the
    scenario obstacle (by id) closest to it -- bit-comparable between implementations."""
    p = np.asarray(position, dtype=np.float64)
    ref_idx = int(np.argmin(np.hypot(*(np.asarray(reference_path) - p).T)))
    obst = None
    if obstacle_positions:
        ids = list(obstacle_positions.keys())
        obst = ids[int(np.argmin([np.hypot(*(np.asarray(obstacle_positions[i]) - p)) for i in ids]))]
    return ref_idx, obst


def obstacle_positions_at(scene: dict, timestep: int) -> dict:
    """id -> position of the scenario's own obstacles at ``timestep`` (fo_obstacle.py:79-116 indexing)."""
    out = {}
    for o in scene["obstacles"]:
        rel = timestep - int(o["initial"][4])
        if o["role"] == "static" or rel == 0:
            out[int(o["id"])] = o["initial"][:2]
        elif rel >= 1 and o["states"] is not None and rel - 1 < len(o["states"]):
            out[int(o["id"])] = o["states"][rel - 1][:2]
    return out


def planner_inputs(scene: dict, timesteps):
    """Planner-side inputs from the replay harness (the only use of product code here: INPUT generation)."""
    from frenetix_occlusion_b200 import replay as R
    from frenetix_occlusion_b200.scenario import scenario_from_dict
    sc = scenario_from_dict(scene)
    ego = R.OpenLoopEgo(sc)
    cycles = []
    for ts in timesteps:
        st = ego.state(ts)
        cycles.append({"timestep": int(ts), "ego_pos": [float(st["pos"][0]), float(st["pos"][1])],
                       "ego_orientation": float(st["orientation"]), "ego_pos_cl": [float(st["pos_cl"][0]), float(st["pos_cl"][1])],
                       "ego_v": float(st["v"])})
    return {"reference_path": np.round(ego.reference_path, 9).tolist(), "cycles": cycles}


def fans_for(scene: dict, inputs: dict):
    from frenetix_occlusion_b200 import replay as R
    from frenetix_occlusion_b200.utils.curvilinear import CurvilinearCoordinateSystem
    cosy = CurvilinearCoordinateSystem(np.asarray(inputs["reference_path"]))
    return [R.frenet_fan(cosy, c["ego_pos_cl"][0], c["ego_pos_cl"][1], c["ego_v"], dt=scene["dt"], **FAN) for c in inputs["cycles"]]


def run_reference(scene: dict, inputs: dict, agents):
    cycles = [dict(c) for c in inputs["cycles"]]
    for c, fan in zip(cycles, fans_for(scene, inputs)):
        c["fan"] = fan
    cfg = PO.REFERENCE_CONFIG
    if agents is None:                                   # the reference's deployment file without its configured agents
        import tempfile
        import yaml
        with open(PO.REFERENCE_CONFIG) as f:
            doc = yaml.safe_load(f)
        doc["agents"] = None
        tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
        yaml.safe_dump(doc, tmp)
        tmp.close()
        cfg = tmp.name
    recs = PO.run_cycles(scene, inputs["reference_path"], VEHICLE, cycles, config_path=cfg)
    out = []
    for r in recs:
        obst = obstacle_positions_at(scene, r["timestep"])
        sps = []
        for s in r["spawn_points"]:
            ref_idx, ob = spawn_indices(s["position"], inputs["reference_path"], obst)
            sps.append({"agent_type": s["agent_type"], "source": s["source"], "position": np.round(s["position"], 4).tolist(),
                        "orientation": None if s["orientation"] is None else round(float(s["orientation"]), 6),
                        "ref_index": ref_idx, "obstacle": ob})
        preds = [{"agent_type": p["agent_type"], "pos0": np.round(p["pos_list"][0], 5).tolist(),
                  "pos_end": np.round(p["pos_list"][-1], 5).tolist(), "n": len(p["pos_list"]),
                  "yaw0": round(float(p["orientation_list"][0]), 6), "v0": round(float(p["v_list"][0]), 5)}
                 for p in r["predictions"].values()]
        out.append({"timestep": r["timestep"],
                    "visible_obstacles": [int(v) if v < 10000 else "real_agent" for v in r["visible_obstacles"]],
                    "visible_area_m2": round(r["visible_area"], 3), "occluded_area_m2": round(r["occluded_area"], 3),
                    "spawn_points": sps, "predictions": preds,
                    "valid": "".join("1" if v else "0" for v in r["valid"]),
                    "max_obst_harm_with_cp_all": [None if h is None else round(h, 9) for h in r["max_obst_harm_with_cp_all"]]})
    return out


def main():
    check = "--check" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name, (timesteps, agents) in PLAN.items():
        if only and name not in only:
            continue
        path = os.path.join(GOLDEN, f"scene_{name}.json")
        if check:
            with open(path) as f:
                doc = json.load(f)
            got = run_reference(doc["scene"], doc["inputs"], doc["agents"])
            assert json.loads(json.dumps(got)) == doc["cycles"], f"{name}: the reference run no longer reproduces the golden"
            print(name, "ok")
            continue
        if name == "parked_car":
            scene = parked_car_scene()
        elif os.path.exists(path):
            with open(path) as f:
                scene = json.load(f)["scene"]           # the compact scene is an input; keep it stable across regenerations
        else:
            from frenetix_occlusion_b200.scenario import load_commonroad_xml, scenario_to_dict
            scene = scenario_to_dict(load_commonroad_xml(os.path.join(REF_SCENARIOS, name + ".xml")))
        scene = json.loads(json.dumps(scene))
        inputs = planner_inputs(scene, timesteps)
        cycles = run_reference(scene, inputs, agents)
        doc = {"source": ("synthetic scene of oracle/make_scenario_golden.py" if name == "parked_car" else
                          f"example_scenarios/{name}.xml of the reference") +
                         "; cycles = outputs of the reference's own FOInterface run over third-party stand-ins "
                         "(oracle/pipeline_oracle.py)",
               "timesteps": timesteps, "agents": agents, "fan": FAN, "vehicle": VEHICLE, "scene": scene, "inputs": inputs,
               "cycles": cycles}
        with open(path, "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        print(name, os.path.getsize(path), "bytes;",
              [(c["timestep"], c["visible_obstacles"], [(s["agent_type"], s["source"]) for s in c["spawn_points"]],
                c["valid"].count("1")) for c in cycles])


if __name__ == "__main__":
    main()
