"""Randomised CUDA-vs-oracle campaign: many seeded bundles, both kernels (detail and summary), aggregated tie
statistics and hard failures.  Run on the GPU box; prints one JSON object (committed under profiles/)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from oracle import metric_oracle as MO  # noqa: E402
import parity  # noqa: E402

rng = np.random.default_rng(int(os.environ.get("FO_CAMPAIGN_SEED", "2024")))     # FO_CAMPAIGN_SEED: another draw of bundles
shapes = [(400, 32, 31), (150, 64, 51), (600, 6, 31), (80, 256, 51), (250, 20, 31), (1000, 3, 31)]
agg = {"cases": 0, "pairs": 0, "trajectories": 0, "evaluations": 0, "hard_failures": [], "mask_mismatch": 0,
       "ties_detail": {}, "ties_summary": {}, "ties_summary_one_warp_window_filter": {}, "ties_summary_float64_tie_variant": {},
       "cp_max_rel": 0.0}
for rep_i in range(4):
    for (n, a, t) in shapes:
        seed = int(rng.integers(1, 1 << 30))
        case = S.make_case(n, a, t, seed=seed)
        out = MO.evaluate_bundle(case)
        arms = [(True, "ties_detail", None, False), (False, "ties_summary", None, False)]
        if a >= 17:      # the throughput shape (one warp per trajectory, window filter); small bundles need it forced
            arms.append((False, "ties_summary_one_warp_window_filter", "1", False))
        # the float64 tie variant of the summary kernel (FO_EXACT_DCE=1: what armed dce / ttc / be thresholds select)
        arms.append((False, "ties_summary_float64_tie_variant", "1" if a >= 17 else None, True))
        for detail, key, team, exact in arms:
            if team is None:
                os.environ.pop("FO_TEAM_WARPS", None)
            else:
                os.environ["FO_TEAM_WARPS"] = team
            if exact:
                os.environ["FO_EXACT_DCE"] = "1"
            else:
                os.environ.pop("FO_EXACT_DCE", None)
            res, _ = parity.run_gpu(case, want_pair=detail, want_step=detail)
            rep = parity.compare_bundle(out, res, case)
            for k, v in rep["ties"].items():
                agg[key][k] = agg[key].get(k, 0) + v
            agg["mask_mismatch"] += rep.get("mask_mismatch", 0)
            agg["cp_max_rel"] = max(agg["cp_max_rel"], rep.get("cp_max_rel", 0.0))
            if rep["fail"]:
                agg["hard_failures"].append({"seed": seed, "shape": [n, a, t], "arm": key, "fail": rep["fail"][:3]})
        agg["cases"] += 1
        agg["pairs"] += n * a
        agg["trajectories"] += n
        agg["evaluations"] += n * a * (t - 1)
print(json.dumps(agg, indent=1))
