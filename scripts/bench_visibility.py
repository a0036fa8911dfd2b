"""C-vis: 4096 rays x 512 rectangles per frame over F frames (BASELINE.json configs[4]).
Prints one JSON line: ray x edge tests / s (brute-force count R * 4 * O per frame), kernel ms."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from frenetix_occlusion_b200 import _lib as L  # noqa: E402
from frenetix_occlusion_b200.visibility import raycast_frames  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ring = len(sys.argv) > 2 and sys.argv[2] == "ring"
R, O = S.C_VIS["n_rays"], S.C_VIS["n_obstacles"]
rect = torch.from_numpy(S.obstacle_frames(F, O)).cuda()
flags = torch.ones((F, O), dtype=torch.uint8, device="cuda")
ego = torch.zeros((F, 3), dtype=torch.float32, device="cuda")
boundary = None
if ring:
    ang = np.linspace(0, 2 * np.pi, 401)
    pts = np.stack((45.0 * np.cos(ang), 45.0 * np.sin(ang)), -1)
    boundary = torch.from_numpy(np.concatenate((pts[:-1], pts[1:]), 1).astype(np.float32)).cuda()
res = raycast_frames(ego, rect, flags, boundary, 50.0, 360.0, R)
for _ in range(3):
    raycast_frames(ego, rect, flags, boundary, 50.0, 360.0, R, out=res)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    raycast_frames(ego, rect, flags, boundary, 50.0, 360.0, R, out=res)
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.mean(ts))
cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
raycast_frames(ego, rect, flags, boundary, 50.0, 360.0, R, out=res, stats=cnt)
torch.cuda.synchronize()
tested, skipped, staged, listed = cnt.cpu().tolist()
n_edges = 4 * O + (400 if ring else 0)
tests = F * R * n_edges
# CPU baseline leg: the float64 port of the ray cast on a few frames (the only use of oracle/ here)
sys.path.insert(0, ROOT)
from oracle import visibility_oracle as VO  # noqa: E402
rect_h, t0 = rect[:4].cpu().numpy(), time.perf_counter()
for f in range(4):
    VO.raycast(np.zeros(3), rect_h[f], np.ones(O, np.uint8), None if boundary is None else boundary.cpu().numpy(), 50.0, 360.0, R)
cpu_s = (time.perf_counter() - t0) / 4
print(json.dumps({"workload": "C-vis", "frames": F, "rays": R, "obstacles": O, "boundary_edges": 400 if ring else 0,
                  "kernel_ms": ms, "per_frame": {"tests_executed": tested / F, "warp_edge_skips": skipped / F,
                                                  "staged_edges": staged / F, "fan_list_entries": listed / F}, "frames_per_s": F / (ms * 1e-3), "ray_edge_tests_per_s": tests / (ms * 1e-3),
                  "algorithmic_bytes": F * (O * 21 + R * 8 + O), "hbm_gbs": F * (O * 21 + R * 8 + O) / (ms * 1e-3) / 1e9,
                  "visible_fraction": float(res.visible.float().mean()), "mean_range": float(res.range.mean()),
                  "cpu_port_frames_per_s_1core": 1.0 / cpu_s}))
