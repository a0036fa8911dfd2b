"""Kernel time of the summary path for all seven metrics vs the default six (no BE): how much of the BE share of the
instruction count comes back as time.  usage: python scripts/time_metric_sets.py [n_traj]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from frenetix_occlusion_b200.engine import AgentSet, MetricEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
case = S.make_case(n, 256, 51)
ego = torch.from_numpy(case["ego"].astype("float32")).cuda()
for name, metrics in (("all7", S.ALL_METRICS), ("default6", S.DEFAULT_METRICS)):
    eng = MetricEngine(case["vehicle"], case["dt"], metrics, case["thresholds"])
    eng.set_agents(AgentSet.from_case(case["agents"]))
    for _ in range(2):
        eng.assess(ego)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        r = eng.assess(ego)
    b.record()
    torch.cuda.synchronize()
    print(name, "ms per launch %.3f" % (a.elapsed_time(b) / 5), "per 1M trajectories %.2f" % (a.elapsed_time(b) / 5 * 1e6 / n))
