"""Oracle B (vectorised float64 restatement) against the golden vectors produced by the
reference's own metric modules (oracle A, ``oracle/make_golden.py``), plus closed-form KATs
(SURVEY.md §8c).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_files
from oracle import metric_oracle as MO
from oracle.compare import compare_results
from oracle.geometry import ConvexPolygon, obb_distance, obb_intersects, rect_vertices


@pytest.mark.parametrize("fname", golden_files("metric_"))
def test_oracle_b_matches_reference_golden(fname):
    case, reference, order = MO.load_case_json(os.path.join(GOLDEN, fname))
    out = MO.evaluate_bundle(case)
    assert out["order"] == order
    for n, ref in enumerate(reference):
        if ref["safety_check"] is None:           # the reference raised (BE interp1d bounds)
            assert out["be_error"][n], f"{fname}[{n}]: reference raised but oracle B did not flag it"
            continue
        assert not out["be_error"][n]
        got = MO.to_reference_dict(out, n, case)
        problems = compare_results(ref["results"], got, rtol=1e-12, atol=1e-15)
        assert not problems, f"{fname}[{n}]:\n" + "\n".join(problems[:10])
        assert bool(out["valid"][n]) == ref["safety_check"], f"{fname}[{n}] safety_check"


def test_kat_values_from_survey():
    case, reference, _ = MO.load_case_json(os.path.join(GOLDEN, "metric_kat_pedestrian.json"))
    out = MO.evaluate_bundle(case)
    assert out["dce"][0, 0] == 0.0 and out["time_dce"][0, 0] == 15
    assert out["ttc"][0, 0] == 1.5 and out["ttce"][0, 0] == 1.5 and out["wttc"][0] == 1.5
    assert out["be_rcd"][0, 0] == 4.609375
    assert abs(out["max_collision_probability_all"][0] - 0.4925777927) < 1e-6
    assert abs(out["max_obst_harm_with_cp_all"][0] - 0.2738569749) < 1e-6
    assert not out["valid"][0]                    # harm 0.1 threshold -> rejected (README.md:19-25)


def test_closed_form_geometry():
    # axis-aligned gap g -> round(g, 3); overlap -> 0; containment -> 0
    d = obb_distance(0.0, 0.0, 0.0, 2.0, 1.0, 5.25, 0.0, 0.0, 1.0, 1.0)
    assert abs(float(d) - 2.25) < 1e-15
    assert float(obb_distance(0.0, 0.0, 0.0, 2.0, 1.0, 0.2, 0.1, 0.7, 0.3, 0.2)) == 0.0
    assert bool(obb_intersects(0.0, 0.0, 0.0, 2.0, 1.0, 3.0, 0.0, 0.0, 1.0, 1.0))   # touching counts
    assert not bool(obb_intersects(0.0, 0.0, 0.0, 2.0, 1.0, 3.0 + 1e-9, 0.0, 0.0, 1.0, 1.0))


def test_two_geometry_formulations_agree():
    rng = np.random.default_rng(5)
    for _ in range(300):
        ax, ay, bx, by = rng.uniform(-6, 6, 4)
        aw, bw = rng.uniform(-np.pi, np.pi, 2)
        al, awd, bl, bwd = rng.uniform(0.3, 5.0, 4)
        pa = ConvexPolygon(rect_vertices(ax, ay, al, awd, aw))
        pb = ConvexPolygon(rect_vertices(bx, by, bl, bwd, bw))
        d_box = float(obb_distance(ax, ay, aw, al / 2, awd / 2, bx, by, bw, bl / 2, bwd / 2))
        assert abs(pa.distance(pb) - d_box) < 1e-12
        assert pa.intersects(pb) == bool(obb_intersects(ax, ay, aw, al / 2, awd / 2, bx, by, bw, bl / 2, bwd / 2))


def test_mass_and_lr4s_boundaries():
    assert abs(MO.obstacle_mass("car", 4.8 * 1.2 * 2.0 * 1.3) - (-1333.5 + 526.9 * 14.976 ** 0.8)) < 1e-9
    assert MO.obstacle_mass("pedestrian", 1.0) == 75 and MO.obstacle_mass("bicycle", 1.0) == 90
    assert MO.check_required_metrics(["hr", "ttc", "ttce", "dce", "wttc", "cp"]) == \
        ["cp", "dce", "ttc", "hr", "ttce", "wttc"]
    assert MO.check_required_metrics(["hr", "ttc", "be", "ttce", "dce", "wttc", "cp"]) == \
        ["cp", "dce", "ttc", "hr", "be", "ttce", "wttc"]


def test_no_agents_is_valid():
    case, _, _ = MO.load_case_json(os.path.join(GOLDEN, "metric_kat_pedestrian.json"))
    case["agents"] = []
    out = MO.evaluate_bundle(case)
    assert out["valid"].all()                     # metric.py:44-45
