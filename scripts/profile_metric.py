"""One or two launches of the dense metric kernel on a reduced C-sweep bundle, for `ncu --set full`.
usage: python scripts/profile_metric.py [n_traj] [n_agents] [n_states] [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from frenetix_occlusion_b200.engine import AgentSet, MetricEngine  # noqa: E402

n, a, t, reps = (int(x) for x in (sys.argv[1:5] + ["200000", "256", "51", "2"][len(sys.argv) - 1:]))
case = S.make_case(n, a, t)
eng = MetricEngine(case["vehicle"], case["dt"], case["activated_metrics"], case["thresholds"])
eng.set_agents(AgentSet.from_case(case["agents"]))
ego = torch.from_numpy(case["ego"].astype("float32")).cuda()
for _ in range(reps):
    r = eng.assess(ego)
torch.cuda.synchronize()
print("valid fraction", float(r.valid.float().mean()))
