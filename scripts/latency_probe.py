"""p50 device latency (CUDA-graph replay) of the summary path for a few bundle shapes / metric sets.
usage: python scripts/latency_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from frenetix_occlusion_b200.engine import AgentSet, MetricEngine  # noqa: E402


def probe(n, a, t, metrics, iters=300):
    case = S.make_case(n, a, t)
    eng = MetricEngine(case["vehicle"], case["dt"], metrics, case["thresholds"])
    eng.set_agents(AgentSet.from_case(case["agents"]))
    ego = torch.from_numpy(case["ego"].astype("float32")).cuda()
    g, out = eng.capture(ego)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.percentile(ts, 50)), float(np.percentile(ts, 95))


if __name__ == "__main__":
    for (n, a, t) in [(1000, 32, 31), (1000, 8, 31), (1000, 3, 31), (100, 32, 31), (4000, 32, 31)]:
        for name, m in (("all7", S.ALL_METRICS), ("default6", S.DEFAULT_METRICS)):
            p50, p95 = probe(n, a, t, list(m))
            print(f"N={n:5d} A={a:3d} T={t} {name:9s} p50 {p50:7.1f} us  p95 {p95:7.1f} us", flush=True)
