"""Sensor model (mirror of reference sensor_model.py:18-234) on the CUDA visibility kernels.

Same constructor arguments and products (``visible_area``, ``occluded_area``, ``obstacle_occlusions``,
``road_polygon``, ``visible_objects_timestep``, per-obstacle ``current_visible`` / ``last_visible_at_ts``).
The reference materialises these as shapely polygons through one GEOS ``difference`` per road-border edge and
obstacle; here they are *region objects* answering the predicates the spawn locator asks (``contains`` for
sampled points) through ``fo_visibility_points``, plus the polar first-hit map of ``fo_visibility_raycast``
(``VisibleArea.ranges`` / ``.exterior``) for the planner and the visible-obstacle list."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .visibility import FrameGeometry, ray_angle_params


class _Region:
    """A planar region known through a point-membership predicate evaluated on the device."""
    flag = 0

    def __init__(self, sensor_model):
        self._sm = sensor_model

    def contains(self, points):
        """bool [M] for points [M,2] (a single point gives a python bool)."""
        P = np.asarray(points, dtype=np.float64).reshape(-1, 2)
        f, _, _ = self._sm._classify(P)
        out = (f & self.flag) != 0
        return out if out.size != 1 or np.ndim(points) > 1 else bool(out[0])

    def intersects_points(self, points) -> bool:
        return bool(np.any(np.atleast_1d(self.contains(points))))


class VisibleArea(_Region):
    """``SensorModel.visible_area``: star-shaped region around the ego.  ``ranges[r]`` is the first-hit distance
    along ``angles[r]``; ``contains`` is exact (segment ego -> point against every opaque edge)."""
    flag = L.PT_VISIBLE

    def __init__(self, sensor_model, ego_pos, angle0, dangle, ranges, hit, full_circle, sensor_radius):
        super().__init__(sensor_model)
        self.ego_pos = np.asarray(ego_pos, dtype=np.float64)
        self.angle0, self.dangle = float(angle0), float(dangle)
        self.ranges = np.asarray(ranges, dtype=np.float64)
        self.hit = np.asarray(hit)
        self.full_circle = bool(full_circle)
        self.sensor_radius = float(sensor_radius)

    @property
    def angles(self):
        return self.angle0 + self.dangle * np.arange(len(self.ranges))

    @property
    def exterior(self):
        """Polar fan polygon (what the reference returns to the planner as a shapely polygon, clipped to the
        road there; here the un-clipped star polygon)."""
        a = self.angles
        ring = self.ego_pos + np.stack((self.ranges * np.cos(a), self.ranges * np.sin(a)), -1)
        if not self.full_circle:
            ring = np.concatenate((self.ego_pos[None], ring, self.ego_pos[None]))
        return ring

    @property
    def area(self):
        s = np.sin(self.dangle)
        a = 0.5 * np.sum(self.ranges[:-1] * self.ranges[1:]) * s
        if self.full_circle:
            a += 0.5 * self.ranges[-1] * self.ranges[0] * s
        return float(a)

    @property
    def is_empty(self):
        return not bool(np.any(self.ranges > 0))


class OccludedArea(_Region):
    """(+-90 deg sector of radius 1.5 R) ∩ road − visible area (sensor_model.py:85-93)."""
    flag = L.PT_OCCLUDED


class RoadArea(_Region):
    """Union of all lanelet polygons (sensor_model.py:195-199)."""
    flag = L.PT_ON_ROAD


class ObstacleShadow(_Region):
    """``obstacle_occlusions[id]`` = shadow of one obstacle minus the obstacle itself (sensor_model.py:182-183)."""
    flag = L.PT_FOCUS_SHADOW

    def __init__(self, sensor_model, index, ray_ids):
        super().__init__(sensor_model)
        self.index = int(index)
        self.ray_ids = np.asarray(ray_ids, dtype=int)

    def contains(self, points):
        P = np.asarray(points, dtype=np.float64).reshape(-1, 2)
        f, _, _ = self._sm._classify(P, focus=self.index)
        out = (f & self.flag) != 0
        return out if out.size != 1 or np.ndim(points) > 1 else bool(out[0])


class SensorModel:
    def __init__(self, lanelet_network, ref_path=None, sensor_radius=50, sensor_angle=360, visualization=None,
                 debug=False, n_rays=4096, device="cuda:0"):
        self.lanelet_network = lanelet_network
        self.ref_path = ref_path
        self.sensor_radius = sensor_radius
        self.sensor_angle = sensor_angle
        self.visualization = visualization
        self.debug = debug
        self.n_rays = int(n_rays)
        self.device = device
        self.timestep = None
        self.ego_pos = None
        self.ego_orientation = None
        self.visible_area = None
        self.occluded_area = None
        self.sensor_sector = None
        self.obstacle_occlusions = {}
        self.all_obstacle_occlusions_polygon = None
        self.visible_objects_timestep = []
        # road polygon = union of all lanelet polygons; its exterior is what casts the border shadows
        # (sensor_model.py:131-155, 195-199)
        self.lanelet_polygons = [np.asarray(p, dtype=np.float64) for p in _lanelet_polygons(lanelet_network)]
        self.road_border = lanelet_network.road_border_segments() if hasattr(lanelet_network, "road_border_segments") \
            else _border_from_polygons(self.lanelet_polygons)
        self.road_polygon = RoadArea(self)
        self._frame = None
        self._obstacle_index = {}

    # ---- device calls (the only places this class touches the C ABI) -----------------------------------
    def _build_frame(self, rect, flags, border):
        return FrameGeometry(self.ego_pos, self.ego_orientation, rect, flags, border, self.lanelet_polygons,
                             self.sensor_radius, self.sensor_angle, device=self.device)

    def _raycast(self, road_hits=None):
        return self._frame.raycast_host(self.n_rays, road_hits=road_hits)

    def _hits_on_road(self, rng, hit, vis, ray_lists, angles):
        """Host restatement of ``fo_visibility_hits_on_road`` (the tests hold the kernel to it): the end points of the
        rays that hit a seen obstacle, classified point by point."""
        O = len(ray_lists)
        out = np.zeros(O, dtype=np.uint8)
        need = [k for k in range(O) if bool(vis[k]) and len(ray_lists[k])]
        if need:
            rays = np.concatenate([ray_lists[k] for k in need])
            pts = self.ego_pos + (rng[rays] - 1e-3)[:, None] * np.stack((np.cos(angles[rays]), np.sin(angles[rays])), -1)
            f, _, _ = self._classify(pts)
            road = (f & L.PT_ON_ROAD) != 0
            at = 0
            for k in need:
                out[k] = bool(road[at:at + len(ray_lists[k])].any())
                at += len(ray_lists[k])
        return out

    def _classify(self, points, focus=-1, focus_margin=0.0):
        if self._frame is None:
            raise RuntimeError("calc_visible_and_occluded_area has not been called yet")
        return self._frame.classify(points, focus_obstacle=focus, focus_margin=focus_margin)

    # ---- sensor_model.py:41-101 ----------------------------------------------------------------------------
    def calc_visible_and_occluded_area(self, timestep, ego_pos, ego_orientation, obstacles):
        self.ego_pos = np.asarray(ego_pos, dtype=np.float64)
        self.ego_orientation = float(ego_orientation)
        self.visible_objects_timestep = []
        self.obstacle_occlusions.clear()
        obs = [o for o in obstacles if o.current_pos is not None]
        O = len(obs)
        rect = np.zeros((O, 5))
        flags = np.zeros(O, dtype=np.uint8)
        for k, o in enumerate(obs):
            rect[k] = o.as_rect()
            flags[k] = L.RECT_EXISTS | (L.RECT_TRANSPARENT if o.cr_obstacle.obstacle_type.value == "bicycle" else 0)
        border = self.road_border
        if len(border):     # only segments that can matter for this frame (occluded area reaches 1.5 R)
            rel = border - np.tile(self.ego_pos, 2)
            near = _segment_distance_to_origin(rel) <= 1.5 * self.sensor_radius + 1.0
            border = border[near]
        self._frame = self._build_frame(rect, flags, border)
        self._obstacle_index = {o.cr_obstacle.obstacle_id: k for k, o in enumerate(obs)}
        a0, da = ray_angle_params(self.ego_orientation, self.sensor_angle, self.n_rays)
        # ray cast + "which obstacles are hit on the road" in one stream-ordered pass and one read-back
        rng, hit, vis, road = self._raycast(road_hits=(float(a0), float(da)))
        self.visible_area = VisibleArea(self, self.ego_pos, float(a0), float(da), rng, hit, self.sensor_angle >= 359.9,
                                        self.sensor_radius)
        # visible obstacles: the reference intersects each obstacle polygon with the (road-clipped) visible area
        # buffered by 1 cm (sensor_model.py:59-76); here: some ray ends on the obstacle at a point of the road
        ray_lists = [np.nonzero(hit == k)[0] for k in range(O)]
        on_road = {k: bool(road[k]) for k in range(O) if bool(vis[k]) and len(ray_lists[k])}
        for k, o in enumerate(obs):
            rays = ray_lists[k]
            seen = on_road[k] if k in on_road else bool(vis[k])
            if seen:
                self.visible_objects_timestep.append(o.cr_obstacle.obstacle_id)
                o.current_visible = True
                o.last_visible_at_ts = timestep
            if len(rays):                                                # sensor_model.py:180-183
                self.obstacle_occlusions[o.cr_obstacle.obstacle_id] = ObstacleShadow(self, k, rays)
        self.occluded_area = OccludedArea(self)
        return self.visible_area


def _lanelet_polygons(lanelet_network):
    if hasattr(lanelet_network, "lanelet_polygons") and not callable(getattr(lanelet_network, "lanelet_polygons")):
        polys = lanelet_network.lanelet_polygons
        if polys and hasattr(polys[0], "vertices"):           # commonroad-io Polygon objects
            return [np.asarray(p.vertices)[:-1] if np.allclose(p.vertices[0], p.vertices[-1]) else np.asarray(p.vertices)
                    for p in polys]
        return polys
    return [np.concatenate((l.left_vertices, l.right_vertices[::-1])) for l in lanelet_network.lanelets]


def _segment_distance_to_origin(seg):
    a, b = seg[:, :2], seg[:, 2:]
    e = b - a
    l2 = np.maximum((e ** 2).sum(1), 1e-18)
    t = np.clip(-(a * e).sum(1) / l2, 0.0, 1.0)
    p = a + t[:, None] * e
    return np.hypot(p[:, 0], p[:, 1])


def _border_from_polygons(polys):
    from .scenario import LaneletNetwork, Lanelet
    lanelets = []
    for i, p in enumerate(polys):
        h = len(p) // 2
        lanelets.append(Lanelet(i, p[:h], p[h:][::-1]))
    return LaneletNetwork(lanelets).road_border_segments()
