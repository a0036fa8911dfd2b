"""Oracle of the whole per-planning-cycle pipeline: the REFERENCE'S OWN ``FOInterface`` (``interface.py``) with its own
``sensor_model.py``, ``spawn_locator.py``, ``agent.py``, ``route_planner.py``, ``utils/{fo_obstacle,helper_functions,
frenetix_handler,sampling}.py`` and ``metrics/*`` imported **unmodified** from ``/root/reference`` and run over stand-ins for
the third-party packages this image cannot install.  TEST INFRASTRUCTURE: only ``oracle/make_scenario_golden.py``
(build container) uses it; it imports nothing of ``frenetix_occlusion_b200``.

What is the reference's: every threshold, ordering, filter chain and piece of control flow of the visibility
computation (``sensor_model.py:41-193``), the three spawn-point finders (``spawn_locator.py:80-578``), the agent manager and
agents (``agent.py``), the vehicle sampling matrix and selection (``frenetix_handler.py``, ``agent.py:364-426``), the metric
plugins and the validity clauses (``metrics/metric.py:35-100``).  What is restated (PARITY UNPINNED at these leaves, listed
with their pinned versions in the headers of the stand-in modules):

* shapely 2.0.2 / GEOS                          -> ``oracle/polygon.py`` (float64 overlay, buffers, predicates)
* commonroad-io 2023.2 object model             -> ``oracle/ref_world.py`` (scene dict of the fixtures) + ``oracle/ref_shims.py``
* commonroad-drivability-checker 2023.1         -> ``ref_world.CurvilinearCoordinateSystem``, polyline utilities, ``RectOBB``
* commonroad-route-planner 2022.3               -> ``ref_world.Route``, ``lanelet_orientation_at_position``
* frenetix (C++)                                -> ``ref_world.TrajectoryHandler`` etc.
* scipy 1.12 ``mvn.mvnun``                      -> ``ref_shims.mvnun``
* matplotlib visualisation                      -> no-op ``ref_world.FOVisualization``
"""
from __future__ import annotations

import random
import sys
import types

import numpy as np

from . import polygon as G
from . import ref_shims
from . import ref_world as W

REFERENCE_CONFIG = "/root/reference/configurations/simulation/occlusion.yaml"
_done = False


def _rotate(geom, angle, origin=(0, 0), use_radians=False):
    a = float(angle) if use_radians else np.radians(float(angle))
    c, s = np.cos(a), np.sin(a)
    o = np.asarray(origin, dtype=np.float64)
    rot = lambda r: (np.asarray(r) - o) @ np.array([[c, s], [-s, c]]) + o  # noqa: E731
    return G.Polygon(rot(geom._shell), [rot(h) for h in geom._holes])


def _translate(geom, xoff=0.0, yoff=0.0):
    t = np.array([float(xoff), float(yoff)])
    return G.Polygon(geom._shell + t, [h + t for h in geom._holes])


def install():
    """Register every stand-in and put the reference tree on ``sys.path``."""
    global _done
    if _done:
        return
    ref_shims.install()
    m = ref_shims._module
    sh = m("shapely")
    sh.geometry = m("shapely.geometry", Point=G.Point, Polygon=G.Polygon, MultiPolygon=G.MultiPolygon, LineString=G.LineString,
                    MultiPoint=G.MultiPoint, MultiLineString=G.MultiLineString, GeometryCollection=G.GeometryCollection)
    m("shapely.geometry.multipolygon", MultiPolygon=G.MultiPolygon)
    sh.affinity = m("shapely.affinity", rotate=_rotate, translate=_translate)
    sh.ops = m("shapely.ops", unary_union=G.unary_union)
    dc = sys.modules["commonroad_dc"]
    dc.pycrccosy = m("commonroad_dc.pycrccosy", CurvilinearCoordinateSystem=W.CurvilinearCoordinateSystem)
    dc.geometry = m("commonroad_dc.geometry")
    dc.geometry.util = m("commonroad_dc.geometry.util", compute_pathlength_from_polyline=W.compute_pathlength_from_polyline,
                         compute_curvature_from_polyline=W.compute_curvature_from_polyline)
    crp = m("commonroad_route_planner")
    crp.route_planner = m("commonroad_route_planner.route_planner", Route=W.Route)
    crp.route = m("commonroad_route_planner.route", RouteType=W.RouteType, Route=W.Route)
    crp.utility = m("commonroad_route_planner.utility")
    crp.utility.route = m("commonroad_route_planner.utility.route",
                          lanelet_orientation_at_position=W.lanelet_orientation_at_position)
    fx = m("frenetix", TrajectoryHandler=W.TrajectoryHandler, CoordinateSystemWrapper=W.CoordinateSystemWrapper)
    fx.trajectory_functions = m("frenetix.trajectory_functions", FillCoordinates=W.FillCoordinates)
    fx.trajectory_functions.feasability_functions = m(
        "frenetix.trajectory_functions.feasability_functions",
        CheckYawRateConstraint=W._feasability("yaw_rate"), CheckAccelerationConstraint=W._feasability("acceleration"),
        CheckCurvatureConstraint=W._feasability("curvature"), CheckCurvatureRateConstraint=W._feasability("curvature_rate"))
    # utils/visualization.py pulls in matplotlib (TkAgg) and the commonroad renderer: replaced as a whole module
    import frenetix_occlusion.utils as ref_utils          # the reference's package (plain directory import)
    vis = m("frenetix_occlusion.utils.visualization", FOVisualization=W.FOVisualization)
    ref_utils.visualization = vis
    _done = True


def reference_interface():
    install()
    import frenetix_occlusion.interface as ref_interface
    return ref_interface


class TrajectoryStub:
    """The planner's trajectory duck type: ``.cartesian.{x,y,theta,v,a}``."""

    def __init__(self, arr):
        a = np.asarray(arr, dtype=np.float64)
        self.cartesian = types.SimpleNamespace(x=a[:, 0].copy(), y=a[:, 1].copy(), theta=a[:, 2].copy(), v=a[:, 3].copy(),
                                               a=a[:, 4].copy())
        self.costMap, self.feasabilityMap = {}, {}
        self.cost, self.feasible, self.valid, self.uniqueId, self.sampling_parameters = 0.0, True, True, 0, None


def run_cycles(scene: dict, reference_path, vehicle_params, cycles, config_path=REFERENCE_CONFIG, seed=7):
    """Drive the reference's ``FOInterface`` over planning cycles.

    ``cycles``: list of dicts with the planner-side INPUTS of one cycle -- ``timestep``, ``ego_pos``, ``ego_orientation``,
    ``ego_pos_cl``, ``ego_v`` and optionally ``fan`` ([N, T, 5] trajectories to assess).  Returns one record per cycle
    with the reference's own outputs."""
    ref_interface = reference_interface()
    random.seed(seed)                                     # agent ids: randint(10000, 11000), agent.py:190
    sc = W.Scenario(scene)
    path = np.asarray(reference_path, dtype=np.float64)
    vp = types.SimpleNamespace(**vehicle_params) if isinstance(vehicle_params, dict) else vehicle_params
    fo = ref_interface.FOInterface(sc, path, vp, sc.dt, config_path=config_path)
    cosy = W.CurvilinearCoordinateSystem(path)
    out = []
    for c in cycles:
        predictions = {}
        fo.evaluate_scenario(predictions, np.asarray(c["ego_pos"], dtype=np.float64), float(c["ego_orientation"]),
                             np.asarray(c["ego_pos_cl"], dtype=np.float64), float(c["ego_v"]), int(c["timestep"]), cosy)
        am = fo.agent_manager
        rec = {"timestep": int(c["timestep"]),
               "visible_obstacles": list(fo.sensor_model.visible_objects_timestep),
               "visible_area": float(fo.sensor_model.visible_area.area), "occluded_area": float(fo.sensor_model.occluded_area.area),
               "spawn_points": [{"position": np.asarray(sp.position, dtype=np.float64), "agent_type": sp.agent_type,
                                 "source": sp.source, "orientation": sp.orientation} for sp in fo.spawn_points],
               "predictions": {pid: {"agent_type": am.agent_by_prediction_id(pid).agent_type,
                                     "pos_list": np.asarray(p["pos_list"], dtype=np.float64),
                                     "orientation_list": np.asarray(p["orientation_list"], dtype=np.float64),
                                     "v_list": np.asarray(p["v_list"], dtype=np.float64),
                                     "cov_list": np.asarray(p["cov_list"], dtype=np.float64), "shape": dict(p["shape"])}
                               for pid, p in am.predictions.items()}}
        if c.get("fan") is not None:
            valid, harm = [], []
            for tr in np.asarray(c["fan"], dtype=np.float64):
                res, ok = fo.trajectory_safety_assessment(TrajectoryStub(tr))
                valid.append(bool(ok))
                harm.append(float(res["hr"]["max_obst_harm_with_cp_all"]) if "hr" in res else None)
            rec["valid"], rec["max_obst_harm_with_cp_all"] = valid, harm
        out.append(rec)
    return out


def case_from_predictions(agent_manager, vehicle_params, ego, activated_metrics, thresholds):
    """Plain-array case (the format of ``metric_oracle.evaluate_bundle``) from any agent manager with the reference's
    ``predictions`` layout."""
    agents = []
    for pid, pred in agent_manager.predictions.items():
        ag = agent_manager.agent_by_prediction_id(pid)
        agents.append({"agent_type": ag.agent_type, "length": ag.shape.length, "width": ag.shape.width,
                       "buf_length": pred["shape"]["length"], "buf_width": pred["shape"]["width"],
                       "pos": np.asarray(pred["pos_list"]), "yaw": np.asarray(pred["orientation_list"]),
                       "v": np.asarray(pred["v_list"]), "var": np.asarray(pred["cov_list"])[:, 0, 0]})
    g = (lambda k: float(vehicle_params[k])) if isinstance(vehicle_params, dict) else (lambda k: float(getattr(vehicle_params, k)))
    return {"dt": float(agent_manager.dt), "vehicle": {k: g(k) for k in ("length", "width", "mass", "wb_rear_axle", "a_max")},
            "ego": np.asarray(ego, dtype=np.float64), "agents": agents,
            "activated_metrics": list(activated_metrics), "thresholds": dict(thresholds)}
