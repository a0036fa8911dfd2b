"""Host-side profile (cProfile) of FOInterface.evaluate_scenario on scenario1: where the planning cycle's wall time goes.
usage (GPU box): python scripts/profile_cycle_host.py [repeats]"""
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
import torch  # noqa: E402

from frenetix_occlusion_b200 import replay as R  # noqa: E402
from frenetix_occlusion_b200.interface import FOInterface  # noqa: E402
from frenetix_occlusion_b200.scenario import scenario_from_dict  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with open(os.path.join(ROOT, "tests", "golden", "scene_scenario1.json")) as f:
    doc = json.load(f)
random.seed(7)
sc = scenario_from_dict(doc["scene"])
ego = R.OpenLoopEgo(sc)
fo = FOInterface(sc, ego.reference_path, R.DEFAULT_VEHICLE, sc.dt, config_path=R.deployment_config())
steps = [0, 6, 12, 18, 24]
for ts in steps:                                       # warm-up pass (first-call costs, real agents inserted)
    st = ego.state(ts)
    fo.evaluate_scenario({}, st["pos"], st["orientation"], st["pos_cl"], st["v"], ts, ego.cosy)
torch.cuda.synchronize()
plain = {ts: [] for ts in steps}
for _ in range(reps):
    for ts in steps:
        st = ego.state(ts)
        t0 = time.perf_counter()
        fo.evaluate_scenario({}, st["pos"], st["orientation"], st["pos_cl"], st["v"], ts, ego.cosy)
        plain[ts].append((time.perf_counter() - t0) * 1e3)
print({ts: round(min(v), 3) for ts, v in plain.items()}, "ms (min over", reps, "repeats, no profiler)")
times = {ts: [] for ts in steps}
pr = cProfile.Profile()
for _ in range(reps):
    for ts in steps:
        st = ego.state(ts)
        t0 = time.perf_counter()
        pr.enable()
        fo.evaluate_scenario({}, st["pos"], st["orientation"], st["pos_cl"], st["v"], ts, ego.cosy)
        pr.disable()
        times[ts].append((time.perf_counter() - t0) * 1e3)
print({ts: round(min(v), 3) for ts, v in times.items()}, "ms (min over", reps, "repeats, under cProfile)")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:6000])
