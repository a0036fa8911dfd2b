"""cProfile of FOInterface.evaluate_scenario over the scenario1 scene fixture (host-side hot spots)."""
import cProfile
import json
import os
import pstats
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from frenetix_occlusion_b200 import replay as R  # noqa: E402
from frenetix_occlusion_b200.interface import FOInterface  # noqa: E402
from frenetix_occlusion_b200.scenario import scenario_from_dict  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "scene_scenario1.json"
doc = json.load(open(os.path.join(ROOT, "tests", "golden", name)))
random.seed(7)
sc = scenario_from_dict(doc["scene"])
ego = R.OpenLoopEgo(sc)
fo = FOInterface(sc, ego.reference_path, R.DEFAULT_VEHICLE, sc.dt, config_path=R.deployment_config(agents=doc["agents"]))
st = ego.state(0)
fo.evaluate_scenario({}, st["pos"], st["orientation"], st["pos_cl"], st["v"], 0, ego.cosy)   # warm-up
pr = cProfile.Profile()
pr.enable()
for ts in doc["timesteps"][1:]:
    st = ego.state(ts)
    fo.evaluate_scenario({}, st["pos"], st["orientation"], st["pos_cl"], st["v"], ts, ego.cosy)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
