"""Times oracle A -- the reference's OWN ``metrics/metric.py:35-100`` and every plugin it dispatches to, imported unmodified
from /root/reference over the leaf stand-ins of ``oracle/ref_shims.py`` -- on samples of the benchmark workloads.
Only possible in the build container (the reference tree does not travel to the GPU box), so the result is committed as
``profiles/oracle_a_cpu_r2.json`` and quoted by bench.py with this provenance.

    python scripts/time_oracle_a.py > profiles/oracle_a_cpu_r2.json
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402  (workload generator only)


def _case(n_traj, n_agents, n_states, seed):
    c = S.make_case(n_traj, n_agents, n_states, seed=seed)
    c["ego"] = np.asarray(c["ego"], dtype=np.float64)
    return c


def _run(args):
    n_traj, n_agents, n_states, seed = args
    from oracle import ref_runner
    case = _case(n_traj, n_agents, n_states, seed)
    t0 = time.perf_counter()
    out, order = ref_runner.run_reference_metrics(case)
    return time.perf_counter() - t0, sum(1 for r, ok in out if ok)


def main():
    cores = os.cpu_count() or 1
    res = {"what": "reference frenetix_occlusion/metrics (metric.py:35-100 + cp, dce, ttc, hr, be, ttce, wttc plugins) run "
                   "verbatim over oracle/ref_shims.py leaf stand-ins, all 7 metrics, float64",
           "host": {"cpu": platform.processor() or platform.machine(), "cores": cores,
                    "where": "build container (the reference tree is absent on the GPU box)"},
           "workloads": {}}
    for name, (n_agents, n_states, n_traj_1, n_traj_all) in {"C-lat": (32, 31, 4, 2), "C-sweep": (256, 51, 1, 1)}.items():
        steps = n_states - 1
        dt1, _ = _run((n_traj_1, n_agents, n_states, 11))
        ev1 = n_traj_1 * n_agents * steps
        with mp.get_context("fork").Pool(cores) as pool:
            t0 = time.perf_counter()
            pool.map(_run, [(n_traj_all, n_agents, n_states, 100 + k) for k in range(cores)])
            dta = time.perf_counter() - t0
        eva = cores * n_traj_all * n_agents * steps
        res["workloads"][name] = {"one_core": {"pairs": n_traj_1 * n_agents, "evals": ev1, "seconds": dt1, "evals_per_s": ev1 / dt1},
                                  "all_cores": {"processes": cores, "pairs": cores * n_traj_all * n_agents, "evals": eva,
                                                "seconds": dta, "evals_per_s": eva / dta}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
