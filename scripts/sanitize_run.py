"""Small invocations of every kernel for compute-sanitizer (memcheck / racecheck / synccheck).
usage: compute-sanitizer --tool racecheck python scripts/sanitize_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from frenetix_occlusion_b200.engine import AgentSet, MetricEngine  # noqa: E402
from frenetix_occlusion_b200.prediction import rollout_cv, rollout_path  # noqa: E402
from frenetix_occlusion_b200.visibility import FrameGeometry, raycast_frames  # noqa: E402

for (n, a, t) in [(40, 32, 31), (300, 8, 31), (12, 300, 51), (9000, 3, 31), (5, 40, 128)]:
    case = S.make_case(n, a, t, seed=n)
    eng = MetricEngine(case["vehicle"], case["dt"], case["activated_metrics"], case["thresholds"])
    eng.set_agents(AgentSet.from_case(case["agents"]))
    r = eng.assess(case["ego"])
    r2 = eng.assess(case["ego"][: min(n, 6)], want_pair=True, want_step=True)
    st = eng.work_stats(case["ego"][: min(n, 50)])
    torch.cuda.synchronize()
    print("metric", n, a, t, int(r.valid.sum()), st["cp"])
os.environ["FO_TEAM_WARPS"] = "1"      # one-warp shape: window filter, queue flush; 6000 trajectories: claim counter
for (n, a, t) in [(60, 40, 51), (30, 300, 31), (6000, 20, 12), (20, 33, 97)]:
    case = S.make_case(n, a, t, seed=n)
    eng = MetricEngine(case["vehicle"], case["dt"], case["activated_metrics"], case["thresholds"])
    eng.set_agents(AgentSet.from_case(case["agents"]))
    r = eng.assess(case["ego"])
    st = eng.work_stats(case["ego"][: min(n, 50)])
    torch.cuda.synchronize()
    print("metric/window filter", n, a, t, int(r.valid.sum()), st["windows_kept"], st["windows"])
del os.environ["FO_TEAM_WARPS"]
rect = torch.from_numpy(S.obstacle_frames(3, 40)).cuda()
flags = torch.ones((3, 40), dtype=torch.uint8, device="cuda")
flags[:, ::7] |= 2
ego = torch.zeros((3, 3), dtype=torch.float32, device="cuda")
ang = np.linspace(0, 2 * np.pi, 41)
ring = np.concatenate((45 * np.stack((np.cos(ang), np.sin(ang)), -1)[:-1], 45 * np.stack((np.cos(ang), np.sin(ang)), -1)[1:]), 1)
res = raycast_frames(ego, rect, flags, torch.from_numpy(ring.astype(np.float32)).cuda(), 50.0, 360.0, 700)
fr = FrameGeometry([1.0, 2.0], 0.3, rect[0].cpu().numpy().astype(np.float64) + np.array([1.0, 2.0, 0, 0, 0]), flags[0].cpu().numpy(),
                   ring + np.array([1.0, 2.0, 1.0, 2.0]), [np.array([[-30, -5], [30, -5], [30, 5], [-30, 5.0]]) + np.array([1.0, 2.0])],
                   50.0, 360.0)
f, b, l = fr.classify(np.random.default_rng(0).uniform(-40, 40, (3000, 2)), focus_obstacle=1, focus_margin=1.0)
from frenetix_occlusion_b200 import _lib as L  # noqa: E402
rng_h, hit_h, vis_h, road_h = fr.raycast_host(700, road_hits=(0.3 - np.pi, 2 * np.pi / 700))
cnt, cen, inside, ncomp, handle = fr.spawn_region(np.array([9.0, 3.0]), 12.0, 0.1, 240, 1, L.PT_OCCLUDED | L.PT_VISIBLE, L.PT_FOCUS_NEAR,
                                                  12.0, np.array([12.0, 2.0]), 1, 1.0)
boxes = fr.spawn_rects(handle, [(np.array([12.0, 2.0]), 5.5, 2.5, False), (np.array([12.0, 2.0]), 2.0, 1.0, True)], 0.2, 0.025)
print("spawn", cnt, ncomp, [b[0] for b in boxes], int(road_h.sum()))
rc = rollout_cv([0.0, 1.0], [0.0, 2.0], [1.4, 2.0], [0.1, 2.0], 0.1, 3.0)
path = np.stack((np.linspace(-10, 100, 56), np.zeros(56)), -1)
rp = rollout_path([path, path + 1.0], [5.0, 6.0], [0.4, 1.2], [10.0, 8.0], 0.1, 5.0)
torch.cuda.synchronize()
print("ok", int(res.hit.max()), int(f.sum()), float(rc["x"].sum()), int(rp["sample"][0]))
