"""Where the FIRST planning cycle after FOInterface construction spends its host time (cProfile, cumulative).
usage: python scripts/profile_first_cycle.py"""
import cProfile
import io
import json
import os
import pstats
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frenetix_occlusion_b200 import replay as R  # noqa: E402
from frenetix_occlusion_b200.interface import FOInterface  # noqa: E402
from frenetix_occlusion_b200.scenario import scenario_from_dict  # noqa: E402

doc = json.load(open(os.path.join(ROOT, "tests", "golden", "scene_scenario1.json")))
random.seed(7)
sc = scenario_from_dict(doc["scene"])
ego = R.OpenLoopEgo(sc)
t0 = time.perf_counter()
fo = FOInterface(sc, ego.reference_path, R.DEFAULT_VEHICLE, sc.dt, config_path=R.deployment_config(agents=doc["agents"]))
print("construction (incl. warm-up) %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
recs = R.replay(fo, ego, doc["timesteps"][:1], fan_kwargs=doc["fan"])
t1 = time.perf_counter()
pr.disable()
print("first cycle (evaluate_scenario + assessment of the fan) %.1f ms" % ((t1 - t0) * 1e3))
t0 = time.perf_counter()
R.replay(fo, ego, doc["timesteps"][1:2], fan_kwargs=doc["fan"])
print("second cycle %.1f ms" % ((time.perf_counter() - t0) * 1e3))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40)
print(s.getvalue()[:9000])
