"""Sensor model on the ray-cast visibility map (mirror of reference sensor_model.py:18-234).

Same constructor arguments and products (``visible_area``, ``occluded_area``, ``obstacle_occlusions``,
``visible_objects_timestep``, per-obstacle ``current_visible`` / ``last_visible_at_ts``), but instead of
shapely polygons the products are small objects over the polar first-hit map computed by
``fo_visibility_raycast``."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .visibility import raycast_frames, ray_angle_params


class VisibleArea:
    """Star-shaped visible region around the ego: ``range[r]`` along ``angle0 + r * dangle``."""

    def __init__(self, ego_pos, angle0, dangle, ranges, hit, full_circle, sensor_radius):
        self.ego_pos = np.asarray(ego_pos, dtype=np.float64)
        self.angle0, self.dangle = float(angle0), float(dangle)
        self.ranges = np.asarray(ranges, dtype=np.float64)
        self.hit = np.asarray(hit)
        self.full_circle = bool(full_circle)
        self.sensor_radius = float(sensor_radius)

    @property
    def angles(self):
        return self.angle0 + self.dangle * np.arange(len(self.ranges))

    def ray_index(self, points):
        """Fractional ray index of the direction ego -> point, and whether it lies inside the fan."""
        d = np.asarray(points, dtype=np.float64).reshape(-1, 2) - self.ego_pos
        rel = np.arctan2(d[:, 1], d[:, 0]) - self.angle0
        rel = np.mod(rel, 2.0 * np.pi)
        idx = rel / self.dangle
        n = len(self.ranges)
        inside = np.ones(len(d), dtype=bool) if self.full_circle else idx <= n - 1
        return idx, inside, np.hypot(d[:, 0], d[:, 1])

    def contains(self, points, margin=0.0):
        """Point(s) visible: inside the fan and closer than the first hit of both neighbouring rays."""
        idx, inside, dist = self.ray_index(points)
        n = len(self.ranges)
        i0 = np.floor(idx).astype(int) % n
        i1 = (i0 + 1) % n if self.full_circle else np.minimum(i0 + 1, n - 1)
        rng = np.minimum(self.ranges[i0], self.ranges[i1])
        out = inside & (dist <= rng + margin)
        return out if out.size > 1 else bool(out[0])

    @property
    def exterior(self):
        a = self.angles
        ring = self.ego_pos + np.stack((self.ranges * np.cos(a), self.ranges * np.sin(a)), -1)
        if not self.full_circle:
            ring = np.concatenate((self.ego_pos[None], ring, self.ego_pos[None]))
        return ring

    @property
    def area(self):
        return float(0.5 * np.sum(self.ranges[:-1] * self.ranges[1:] * np.sin(self.dangle))
                     + (0.5 * self.ranges[-1] * self.ranges[0] * np.sin(self.dangle) if self.full_circle else 0.0))

    @property
    def is_empty(self):
        return not bool(np.any(self.ranges > 0))


class OccludedArea:
    """(+-90 deg sector of radius 1.5 R) ∩ road - visible area (sensor_model.py:85-93) as a predicate."""

    def __init__(self, visible_area, ego_orientation, lanelet_network, factor=1.5):
        self.visible_area = visible_area
        self.ego_orientation = float(ego_orientation)
        self.lanelet_network = lanelet_network
        self.radius = factor * visible_area.sensor_radius

    def contains(self, points):
        P = np.asarray(points, dtype=np.float64).reshape(-1, 2)
        d = P - self.visible_area.ego_pos
        rel = np.mod(np.arctan2(d[:, 1], d[:, 0]) - self.ego_orientation + np.pi, 2 * np.pi) - np.pi
        out = (np.abs(rel) <= np.pi / 2) & (np.hypot(d[:, 0], d[:, 1]) <= self.radius)
        out &= self.lanelet_network.points_on_road(P)
        out &= ~np.atleast_1d(self.visible_area.contains(P))
        return out if out.size > 1 else bool(out[0])


class ObstacleShadow:
    """Shadow of one obstacle (``obstacle_occlusions[id]``, sensor_model.py:182-183): the fan of rays whose
    first hit is this obstacle, from the hit outwards."""

    def __init__(self, visible_area, ray_ids):
        self.visible_area = visible_area
        self.ray_ids = np.asarray(ray_ids, dtype=int)

    @property
    def angular_interval(self):
        a = self.visible_area.angles[self.ray_ids]
        return float(a.min()), float(a.max())

    def contains(self, points):
        idx, inside, dist = self.visible_area.ray_index(points)
        n = len(self.visible_area.ranges)
        i0 = np.round(idx).astype(int) % n
        mask = np.zeros(n, dtype=bool)
        mask[self.ray_ids] = True
        out = inside & mask[i0] & (dist > self.visible_area.ranges[i0]) & (dist <= self.visible_area.ranges[i0] + 100.0)
        return out if out.size > 1 else bool(out[0])


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
        # road border = exterior of the union of all lanelet polygons (sensor_model.py:195-199)
        self.road_polygon = lanelet_network
        self.road_border = lanelet_network.road_border_segments() if hasattr(lanelet_network, "road_border_segments") \
            else np.zeros((0, 4))

    def calc_visible_and_occluded_area(self, timestep, ego_pos, ego_orientation, obstacles):
        """sensor_model.py:41-101 on the polar map: one launch of the ray-cast kernel for this frame."""
        self.ego_pos = np.asarray(ego_pos, dtype=np.float64)
        self.ego_orientation = float(ego_orientation)
        self.visible_objects_timestep = []
        self.obstacle_occlusions.clear()
        obs = [o for o in obstacles if o.current_pos is not None]
        O = len(obs)
        rect = np.zeros((1, max(O, 1), 5))
        flags = np.zeros((1, max(O, 1)), dtype=np.uint8)
        for k, o in enumerate(obs):
            rect[0, k] = o.as_rect()
            flags[0, k] = L.RECT_EXISTS | (L.RECT_TRANSPARENT if o.cr_obstacle.obstacle_type.value == "bicycle" else 0)
        # everything relative to the ego position (float64 shift before the float32 cast)
        rect[0, :, 0] -= self.ego_pos[0]
        rect[0, :, 1] -= self.ego_pos[1]
        border = self.road_border - np.tile(self.ego_pos, 2) if len(self.road_border) else None
        if border is not None:      # only segments that can matter for this frame
            near = np.minimum(np.hypot(border[:, 0], border[:, 1]), np.hypot(border[:, 2], border[:, 3])) \
                <= self.sensor_radius + 5.0
            border = border[near]
        ego = np.array([[0.0, 0.0, self.ego_orientation]])
        res = raycast_frames(ego, rect[:, :O], flags[:, :O], border, self.sensor_radius, self.sensor_angle, self.n_rays,
                             device=self.device)
        torch.cuda.current_stream(torch.device(self.device)).synchronize()
        rng = res.range[0].cpu().numpy()
        hit = res.hit[0].cpu().numpy()
        vis = res.visible[0].cpu().numpy() if O else np.zeros(0, np.uint8)
        a0, da = ray_angle_params(self.ego_orientation, self.sensor_angle, self.n_rays)
        self.visible_area = VisibleArea(self.ego_pos, float(a0), float(da), rng, hit, self.sensor_angle >= 359.9,
                                        self.sensor_radius)
        for k, o in enumerate(obs):
            if vis[k]:                                                   # sensor_model.py:68-76
                self.visible_objects_timestep.append(o.cr_obstacle.obstacle_id)
                o.current_visible = True
                o.last_visible_at_ts = timestep
            rays = np.nonzero(hit == k)[0]
            if len(rays):                                                # sensor_model.py:180-183
                self.obstacle_occlusions[o.cr_obstacle.obstacle_id] = ObstacleShadow(self.visible_area, rays)
        self.occluded_area = OccludedArea(self.visible_area, self.ego_orientation, self.lanelet_network)
        return self.visible_area
