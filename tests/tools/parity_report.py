"""Print CUDA-vs-oracle mismatch / tie statistics for a synthetic bundle (run on the GPU box)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from oracle import metric_oracle as MO  # noqa: E402
import parity  # noqa: E402

n, a, t = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1000, 32, 31)))
case = S.make_case(n, a, t)
t0 = time.time()
out = MO.evaluate_bundle(case)
t1 = time.time()
res, _ = parity.run_gpu(case)
rep = parity.compare_bundle(out, res, case)
rep["oracle_seconds"] = t1 - t0
rep["gate_fraction"] = out.get("gate_fraction")
rep["be_pair_fraction"] = float((np.isfinite(out["ttc"]) & (out["ttc"] > 0)).mean())
rep["be_error_traj"] = int(out["be_error"].sum())
rep["collision_traj"] = int(np.isfinite(out["wttc"]).sum())
print(json.dumps(rep, indent=1))
