"""Timing of the detail kernel (pair / step outputs) on a device-resident bundle.
usage: python scripts/bench_detail.py [n_traj] [n_agents] [n_states]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from frenetix_occlusion_b200.engine import AgentSet, MetricEngine  # noqa: E402

n, a, t = (int(x) for x in (sys.argv[1:4] + ["20000", "64", "51"][len(sys.argv) - 1:]))
case = S.make_case(n, a, t)
eng = MetricEngine(case["vehicle"], case["dt"], case["activated_metrics"], case["thresholds"])
eng.set_agents(AgentSet.from_case(case["agents"]))
ego = torch.from_numpy(case["ego"].astype("float32")).cuda()
r = eng.assess(ego, want_pair=True, want_step=True)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ms = []
for _ in range(5):
    ev[0].record()
    r = eng.assess(ego, want_pair=True, want_step=True, out=r)
    ev[1].record()
    torch.cuda.synchronize()
    ms.append(ev[0].elapsed_time(ev[1]))
ms.sort()
evals = n * a * (t - 1)
out_bytes = r.pair.numel() * 4 + r.step.numel() * 4
print(f"detail kernel {n}x{a}x{t - 1}: {ms[2]:.3f} ms, {evals / ms[2] / 1e6:.1f} G evals/s, "
      f"{out_bytes / ms[2] / 1e6:.1f} GB/s written, checksum {float(r.summary.nan_to_num(posinf=0).sum()):.6e}")
