"""Generate ``tests/golden/metric_*.json`` by running the REFERENCE's own metric modules
(``/root/reference/frenetix_occlusion/metrics``) verbatim over the stand-ins in ``ref_shims``.

Build-container only (the reference tree does not travel to the GPU box).  Usage::

    python -m oracle.make_golden            # regenerate all fixtures
    python -m oracle.make_golden --check    # re-run the reference and diff against the committed files

Each fixture holds the *case* (inputs, float32-representable), the reference's evaluation order,
and per trajectory the reference's ``(results, safety_check)`` exactly as
``FOInterface.trajectory_safety_assessment`` returns them (interface.py:216-219).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from frenetix_occlusion_b200 import synthetic as S  # noqa: E402
from oracle import metric_oracle as MO  # noqa: E402
from oracle import ref_runner as RR  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
f32 = S._f32


def _agent(kind, x0, y0, yaw, n_states, dt=0.1, v=None, round_v=True):
    d = S.AGENT_DEFAULTS[kind]
    v = d["v"] if v is None else v
    yaw = float(f32(yaw))
    k = np.arange(n_states)
    vx, vy = v * np.cos(yaw), v * np.sin(yaw)
    if round_v:  # pedestrian rollout semantics, agent.py:492-493
        vx, vy = round(vx, 3), round(vy, 3)
    pos = np.stack((x0 + k * dt * vx, y0 + k * dt * vy), -1)
    return {"agent_type": kind, "length": float(f32(d["length"])), "width": float(f32(d["width"])),
            "buf_length": float(f32(d["length"] * d["fl"])), "buf_width": float(f32(d["width"] * d["fw"])),
            "pos": f32(pos), "yaw": np.full(n_states, yaw), "v": f32(np.full(n_states, v)),
            "var": f32(0.1 * np.power(1.05, k))}


def _straight_ego(v0, n_states, dt=0.1, acc=0.0, heading=0.0, x0=0.0, y0=0.0):
    t = np.arange(n_states) * dt
    t_stop = v0 / -acc if acc < 0 else np.inf
    tc = np.minimum(t, t_stop)
    s = v0 * tc + 0.5 * acc * tc * tc
    v = np.maximum(v0 + acc * tc, 0.0)
    a = np.where(t < t_stop, acc, 0.0)
    return f32(np.stack((x0 + s * np.cos(heading), y0 + s * np.sin(heading), np.full(n_states, heading), v, a), -1))


def build_cases():
    veh = {k: float(f32(v)) for k, v in S.VEHICLE.items()}
    cases = {}

    # 1. SURVEY.md §8(c) end-to-end KAT: straight ego at 8 m/s, one crossing pedestrian
    cases["kat_pedestrian"] = {
        "dt": 0.1, "vehicle": veh, "ego": _straight_ego(8.0, 31)[None],
        "agents": [_agent("Pedestrian", 12.0, -3.0, np.pi / 2, 31)],
        "activated_metrics": list(S.ALL_METRICS), "thresholds": dict(S.DEFAULT_THRESHOLDS)}

    # 2. mixed bundle, all seven metrics, agents close to the ego fan
    rng = np.random.default_rng(11)
    ego = S.ego_bundle(14, 31, rng=rng)
    agents = S.agent_table(9, 31, rng=rng, area=((2.0, 30.0), (-6.0, 6.0)))
    cases["mixed_all_metrics"] = {"dt": 0.1, "vehicle": veh, "ego": ego, "agents": agents,
                                  "activated_metrics": list(S.ALL_METRICS),
                                  "thresholds": dict(S.DEFAULT_THRESHOLDS)}

    # 3. default activated list (no BE) and package-default thresholds, different seed
    rng = np.random.default_rng(12)
    ego = S.ego_bundle(10, 31, rng=rng)
    agents = S.agent_table(8, 31, rng=rng, area=((2.0, 35.0), (-8.0, 8.0)))
    cases["default_metrics"] = {"dt": 0.1, "vehicle": veh, "ego": ego, "agents": agents,
                                "activated_metrics": list(S.DEFAULT_METRICS),
                                "thresholds": {**S.DEFAULT_THRESHOLDS, "harm": 1}}

    # 4. ragged prediction lengths: shorter, longer and single-state predictions; a Truck
    rng = np.random.default_rng(13)
    ego = S.ego_bundle(8, 31, rng=rng)
    agents = [_agent("Pedestrian", 10.0, -2.0, np.pi / 2, 12),
              _agent("Bicycle", 18.0, 3.0, -2.0, 51, round_v=False),
              _agent("Car", 25.0, -1.0, np.pi, 20, round_v=False),
              _agent("Truck", 30.0, 2.5, -3.0, 31, round_v=False),
              _agent("Pedestrian", 6.0, 1.0, 0.3, 1),
              _agent("Car", 8.0, 4.0, -1.2, 2, round_v=False)]
    cases["ragged_lengths"] = {"dt": 0.1, "vehicle": veh, "ego": ego, "agents": agents,
                               "activated_metrics": ["hr", "ttc", "ttce", "dce", "wttc", "cp"],
                               "thresholds": dict(S.DEFAULT_THRESHOLDS)}

    # 5. every threshold armed (be/cp/ttc/dce as well), accelerating egos so BE stays in range
    rng = np.random.default_rng(14)
    ego = np.concatenate([_straight_ego(v0, 31, acc=a, heading=h)[None]
                          for v0, a, h in [(6.0, 0.5, 0.0), (9.0, 0.0, 0.05), (12.0, 1.0, -0.05), (4.0, 1.5, 0.1),
                                           (10.0, 0.2, 0.0), (7.0, 0.0, -0.1), (14.0, 0.0, 0.02), (3.0, 2.0, 0.0)]])
    agents = S.agent_table(10, 31, rng=rng, area=((4.0, 32.0), (-5.0, 5.0)))
    cases["all_thresholds"] = {"dt": 0.1, "vehicle": veh, "ego": ego, "agents": agents,
                               "activated_metrics": list(S.ALL_METRICS),
                               "thresholds": {"harm": 0.3, "risk": 0.05, "be": 0.3, "cp": 0.2, "ttc": 1.0,
                                              "wttc": None, "ttce": None, "dce": 0.5}}

    # 6. braking egos: the reference's BE raises from interp1d on trajectories that stop (be.py:117-124)
    ego = np.concatenate([_straight_ego(v0, 31, acc=a)[None]
                          for v0, a in [(8.0, -1.0), (8.0, -3.0), (3.0, -4.0), (12.0, -2.0), (2.0, -1.0), (15.0, -4.0)]])
    agents = [_agent("Pedestrian", 9.0, -2.5, np.pi / 2, 31), _agent("Pedestrian", 14.0, 3.0, -np.pi / 2, 31),
              _agent("Bicycle", 20.0, -6.0, 1.7, 31, round_v=False), _agent("Car", 30.0, 0.5, np.pi, 31, round_v=False)]
    cases["braking_be"] = {"dt": 0.1, "vehicle": veh, "ego": ego, "agents": agents,
                           "activated_metrics": list(S.ALL_METRICS), "thresholds": dict(S.DEFAULT_THRESHOLDS)}

    # 7. longer horizon (T = 51, the "real agent" horizon) with dt = 0.1, curved egos
    rng = np.random.default_rng(15)
    ego = S.ego_bundle(6, 51, rng=rng)
    agents = S.agent_table(7, 51, rng=rng, area=((2.0, 45.0), (-8.0, 8.0)))
    cases["horizon51"] = {"dt": 0.1, "vehicle": veh, "ego": ego, "agents": agents,
                          "activated_metrics": ["hr", "ttc", "ttce", "dce", "wttc", "cp"],
                          "thresholds": dict(S.DEFAULT_THRESHOLDS)}

    # 8. SURVEY closed-form corner cases: agent exactly on the 5 m gate, touching boxes, containment
    ego = _straight_ego(0.0, 4)[None]  # standing ego at the origin (centre at wb, 0)
    wb, L, W = veh["wb_rear_axle"], veh["length"], veh["width"]
    ag_gate = _agent("Pedestrian", 5.0, 0.0, np.pi / 2, 4, v=0.0)       # centre point exactly 5.0 from ego point
    ag_in = _agent("Pedestrian", float(f32(wb)), 0.0, 0.0, 4, v=0.0)    # contained in the ego rectangle
    ag_gap = _agent("Car", float(f32(wb + L / 2 + 2.4 + 1.25)), 0.0, 0.0, 4, v=0.0, round_v=False)  # 1.25 m gap
    cases["closed_form"] = {"dt": 0.1, "vehicle": veh, "ego": ego, "agents": [ag_gate, ag_in, ag_gap],
                            "activated_metrics": ["hr", "ttc", "ttce", "dce", "wttc", "cp"],
                            "thresholds": dict(S.DEFAULT_THRESHOLDS)}
    return cases


def run_case(case):
    out, order = RR.run_reference_metrics(case)
    return {"case": MO.case_to_json(case), "order": order,
            "reference": [{"results": RR.results_to_json(res), "safety_check": ok} for res, ok in out]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    bad = 0
    for name, case in build_cases().items():
        path = os.path.join(GOLDEN, f"metric_{name}.json")
        doc = run_case(case)
        text = json.dumps(doc, separators=(",", ":"))
        if args.check:
            same = os.path.exists(path) and open(path).read() == text
            print(f"{name}: {'OK' if same else 'DIFFERS'}")
            bad += not same
        else:
            with open(path, "w") as f:
                f.write(text)
            n_err = sum(r["safety_check"] is None for r in doc["reference"])
            print(f"{name}: N={len(doc['reference'])} A={len(case['agents'])} order={doc['order']} "
                  f"raised={n_err} bytes={len(text)}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
