"""Drive the assessment hot path over one of the scene fixtures (needs a B200 and the built library).

    python examples/replay_scene.py [scene_scenario1.json] [--xml /path/to/commonroad_scenario.xml]

Per planning cycle it prints what the reference's FOInterface would hand back to the planner: the visible
obstacles, the phantom agents spawned in the occluded regions, and how many candidate trajectories of a sampled
Frenet fan survive the occlusion-aware safety assessment (occlusion.yaml thresholds: harm 0.1, risk 1)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from frenetix_occlusion_b200 import replay as R  # noqa: E402
from frenetix_occlusion_b200.interface import FOInterface  # noqa: E402
from frenetix_occlusion_b200.scenario import load_commonroad_xml, scenario_from_dict  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--xml" in sys.argv:
        scenario = load_commonroad_xml(sys.argv[sys.argv.index("--xml") + 1])
        timesteps, agents = [0, 5, 10, 15, 20], None
    else:
        name = args[0] if args else "scene_scenario1.json"
        with open(os.path.join(ROOT, "tests", "golden", name)) as f:
            doc = json.load(f)
        scenario, timesteps, agents = scenario_from_dict(doc["scene"]), doc["timesteps"], doc["agents"]
    ego = R.OpenLoopEgo(scenario)                                   # stands in for the planner's ego state
    fo = FOInterface(scenario, ego.reference_path, R.DEFAULT_VEHICLE, scenario.dt,
                     config_path=R.deployment_config(agents=agents))
    for ts in timesteps:
        st = ego.state(ts)
        fo.evaluate_scenario({}, st["pos"], st["orientation"], st["pos_cl"], st["v"], ts, ego.cosy)
        fan = R.frenet_fan(ego.cosy, st["pos_cl"][0], st["pos_cl"][1], st["v"],
                           speed_factors=np.linspace(0, 1.3, 40), lateral_targets=np.linspace(-1.5, 1.5, 25))
        res = fo.assess_bundle(fan)                                  # one launch for the whole bundle
        valid = res.valid.cpu().numpy().astype(bool)
        print(f"t={ts:3d}  ego=({st['pos'][0]:7.2f},{st['pos'][1]:7.2f})  visible obstacles {fo.sensor_model.visible_objects_timestep}")
        for sp in fo.spawn_points:
            print(f"        phantom {sp.agent_type:10s} at ({sp.position[0]:7.2f},{sp.position[1]:7.2f})  [{sp.source}]")
        print(f"        {valid.sum()} of {len(valid)} candidate trajectories pass; "
              f"max harm x cp over the bundle {float(res.summary[:, 5].max()):.3f}")


if __name__ == "__main__":
    main()
