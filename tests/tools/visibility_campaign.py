"""Randomised CUDA-vs-oracle campaign for stage 1: ray cast (range / hit / visible) and point classification
(FO_PT_* bits, lanelet masks) over seeded frames.  Run on the GPU box; prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from frenetix_occlusion_b200.visibility import FrameGeometry, raycast_frames  # noqa: E402
from oracle import visibility_oracle as VO  # noqa: E402
from test_visibility import _frame  # noqa: E402

f32 = lambda a: np.asarray(a, np.float64).astype(np.float32).astype(np.float64)  # noqa: E731
rng = np.random.default_rng(77)
agg = {"ray_frames": 0, "rays": 0, "range_mismatch": 0, "hit_mismatch": 0, "visible_flag_mismatch": 0, "obstacles": 0,
       "point_frames": 0, "points": 0, "bit_mismatch": {}, "lanelet_mask_mismatch": 0}
for case in range(40):
    n_obst = int(rng.choice([0, 5, 40, 200, 512, 1500]))
    fov = float(rng.choice([360.0, 360.0, 200.0, 120.0, 60.0]))
    n_rays = int(rng.choice([64, 333, 1024, 4096]))
    ring = bool(rng.integers(0, 2))
    ego, rect, flags, boundary = _frame(1000 + case, n_obst, ring, transparent_every=int(rng.choice([0, 3, 9])), half_extent=50.0)
    ego, rect = f32(ego), f32(rect)
    boundary = None if boundary is None else f32(boundary)
    res = raycast_frames(ego[None], rect[None], flags[None], boundary, 50.0, fov, n_rays)
    torch.cuda.synchronize()
    o_rng, o_hit, o_vis = VO.raycast(ego, rect, flags, boundary, 50.0, fov, n_rays)
    g_rng, g_hit = res.range[0].cpu().numpy(), res.hit[0].cpu().numpy()
    agg["ray_frames"] += 1
    agg["rays"] += n_rays
    agg["range_mismatch"] += int((~np.isclose(g_rng, o_rng, rtol=2e-5, atol=2e-4)).sum())
    agg["hit_mismatch"] += int((g_hit != o_hit).sum())
    if n_obst:
        agg["visible_flag_mismatch"] += int((res.visible[0].cpu().numpy() != o_vis).sum())
        agg["obstacles"] += n_obst
    # point classification on the same frame
    polys = [np.array([[-60, -6], [60, -6], [60, 6], [-60, 6.0]]), np.array([[-5, -60], [7, -60], [7, 60], [-5, 60.0]]),
             np.array([[10, 10], [45, 20], [40, 35], [5, 25.0]])]
    P = f32(rng.uniform(-70, 70, (8000, 2)))
    focus = int(np.argmax(flags == 1)) if n_obst and (flags == 1).any() else -1
    fr = FrameGeometry([0.0, 0.0], ego[2], rect, flags, boundary, polys, 50.0, fov)
    gf, _, gl = fr.classify(P, focus_obstacle=focus, focus_margin=1.0)
    of, ol = VO.classify_points(P, np.array([0.0, 0.0, ego[2]]), rect, flags, boundary, polys, 50.0, fov, 75.0, focus=focus,
                                focus_margin=1.0)
    inside = (of & VO.PT_IN_OBSTACLE) != 0
    for bit, name in ((1, "in_sensor"), (2, "on_road"), (4, "shadowed"), (8, "in_obstacle"), (16, "visible"), (32, "occluded"),
                      (64, "focus_shadow"), (128, "focus_near")):
        bad = ((gf & bit) != 0) != ((of & bit) != 0)
        if bit == 4:
            bad &= ~inside          # the oracle's shadow excludes the obstacle's own interior by construction
        agg["bit_mismatch"][name] = agg["bit_mismatch"].get(name, 0) + int(bad.sum())
    agg["lanelet_mask_mismatch"] += int((gl != ol).sum())
    agg["point_frames"] += 1
    agg["points"] += len(P)
print(json.dumps(agg, indent=1))
