"""Phantom-agent spawn points (mirror of reference spawn_locator.py:18-752) without a polygon library.

The reference answers every geometric question with shapely set operations on the polygons the sensor model
built.  Here the same questions are asked about *sampled geometry* -- raster windows, lines, disc rims -- whose
samples are classified on the GPU by ``fo_visibility_points`` (visible / occluded / on-road / behind obstacle
k / lanelet membership, evaluated exactly per point from the reference's own shadow construction).  What
remains on the host is bookkeeping: thresholds, sorting, connected components and rectangle metrics of small
rasters.  Class, method and attribute names follow the reference; resolution parameters (``raster_cell``,
``line_step``) bound the deviation from exact polygon clipping (documented ties)."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .route_planner import lanelet_orientation_at_position
from .utils import helper_functions as hf


class SpawnPoint:
    """spawn_locator.py:18-27."""

    def __init__(self, pos, agent_type, pos_cl=None, source=None, orientation=None):
        self.position = pos
        self.agent_type = agent_type
        self.cl_pos = pos_cl
        self.source = source
        self.orientation = orientation


class SpawnLocator:
    def __init__(self, agent_manager, ref_path, cosy_cl, sensor_model, fo_obstacles, config, visualization=None,
                 debug=False):
        self.agent_manager = agent_manager
        self.scenario = agent_manager.scenario
        self.sensor_model = sensor_model
        self.fo_obstacles = fo_obstacles
        self.visualization = visualization
        self.config = config["spawn_locator"]
        self.ref_path = np.asarray(ref_path, dtype=np.float64)
        self.cosy_cl = cosy_cl

        self.reference = None
        self.s = None
        self.reference_s = None
        self.ego_pos = None
        self.ego_orientation = None
        self.ego_cl = None
        self.spawn_points = []

        self.spawn_point_behind_turn = self.config["spawn_points_behind_turn"]
        self.spawn_point_behind_dynamic_obstacle = self.config["spawn_point_behind_dynamic_obstacle"]
        self.spawn_point_behind_static_obstacle = self.config["spawn_point_behind_static_obstacle"]
        self.max_dynamic_spawn_points = self.config["max_dynamic_spawn_points"]
        self.max_static_spawn_points = self.config["max_static_spawn_points"]

        # spawn_locator.py:63-78 (same constants)
        self.ped_width = config["agent_manager"]["pedestrian"]["width"]
        self.ped_length = config["agent_manager"]["pedestrian"]["length"]
        self.s_threshold_time = 4
        self.min_s_threshold = 25
        self.s_threshold = None
        self.tolerance_same_direction = np.radians(20)
        self.max_distance_to_other_obstacle = 30
        self.buffer_around_vehicle_from_side = 12
        self.min_area_threshold = 10
        self.agent_area_limits = {"Car": 9, "Bicycle": 1.7}
        self.min_distance_between_pedestrians = 5
        self.debug = debug
        self.offset_ref_path = {"left turn": 3, "right turn": 0}
        self.phantom_offset_s = {"left turn": -0.5, "right turn": 0}
        self.phantom_offset_d = {"left turn": 1, "right turn": -1}

        # sampling resolutions of this implementation
        self.raster_cell = 0.1        # m, occluded-area raster behind dynamic obstacles
        self.rect_cell = 0.025        # m, raster of the candidate phantom rectangles
        self.path_refine = 1024       # subdivisions of the turn finder's 5 cm bracket around the path / occlusion crossing
        self.line_step = 0.01         # m, sampling of the lines perpendicular to the reference path
        self.path_step = 0.05         # m, sampling of the (shifted) reference path

    # ---------------------------------------------------------------------------------------------------
    def find_spawn_points(self, ego_pos, ego_orientation, ego_cl, ego_v):
        """spawn_locator.py:80-139."""
        self.spawn_points.clear()
        self.ego_pos = np.asarray(ego_pos, dtype=np.float64)
        self.ego_orientation = ego_orientation
        self.ego_cl = ego_cl
        self.s_threshold = ego_cl[0] + max(ego_v * self.s_threshold_time, self.min_s_threshold)
        self.reference, self.reference_s = self._prepare_reference_path(ego_cl)
        ego_intention = self._find_ego_intention(self.reference)
        if self.debug:
            print(ego_intention)
        if self.spawn_point_behind_dynamic_obstacle and ego_intention in ("straight ahead", "left turn"):
            self._append_spawn_point(self._find_spawn_point_behind_dynamic_obstacle())
        if self.spawn_point_behind_static_obstacle:
            self._append_spawn_point(self._find_spawn_point_behind_static_obstacle())
        if self.spawn_point_behind_turn and ego_intention in ("left turn", "right turn"):
            self._append_spawn_point(self._find_spawn_point_behind_turn(ego_intention))
        return self.spawn_points

    # ---- behind dynamic obstacles (spawn_locator.py:145-317) ---------------------------------------------
    def _find_spawn_point_behind_dynamic_obstacle(self):
        spawn_points = []
        visible_dyn_obst = [o for o in self.fo_obstacles
                            if o.current_visible and o.cr_obstacle.obstacle_role.name == "DYNAMIC"]
        visible_dyn_obst = sorted(visible_dyn_obst, key=lambda o: np.linalg.norm(self.ego_pos - o.current_pos))
        if not visible_dyn_obst:
            return
        ego_lanelet = self._find_lanelet_by_position(self.ego_pos)
        all_intersections = self.scenario.lanelet_network.intersections
        intersection, incoming_lanelets, intersection_lanelets = self._find_relevant_intersection(all_intersections,
                                                                                                   ego_lanelet)
        if intersection:
            relevant_lanelets = set(incoming_lanelets)
            relevant_lanelets.update(intersection_lanelets)
            relevant_lanelets.remove(ego_lanelet.lanelet_id)
        else:
            lanelets_along_reference_path = self._find_lanelets_along_reference(self.reference, step=5)
            relevant_lanelets = [lanelet.adj_left for lanelet in lanelets_along_reference_path]

        for dyn_obst in visible_dyn_obst:
            otype = dyn_obst.cr_obstacle.obstacle_type.value
            if otype == "bicycle" or otype == "pedestrian":
                continue
            if len(spawn_points) > self.max_dynamic_spawn_points:
                break
            if np.linalg.norm(self.ego_pos - dyn_obst.current_pos) > self.max_distance_to_other_obstacle:
                continue
            dyn_obst_lanelet_ids = self.scenario.lanelet_network.find_lanelet_by_position([dyn_obst.current_pos])[0]
            if not any(e in relevant_lanelets for e in dyn_obst_lanelet_ids):
                continue
            try:
                dyn_obstacle_cl = self.cosy_cl.convert_to_curvilinear_coords(dyn_obst.current_pos[0], dyn_obst.current_pos[1])
            except Exception:
                continue
            if dyn_obstacle_cl[0] < self.ego_cl[0] + 3 or abs(dyn_obstacle_cl[1]) > 15:
                continue

            possible_ids = [i for i in dyn_obst_lanelet_ids if i in relevant_lanelets]
            if intersection and all(i in intersection_lanelets for i in dyn_obst_lanelet_ids):
                first = self.scenario.lanelet_network.find_lanelet_by_id(possible_ids[0])
                possible_ids = possible_ids + [first.predecessor[0]]

            orientation_diff = abs(dyn_obst.current_orientation - self.ego_orientation) % (2 * np.pi)
            opposite = np.pi - self.tolerance_same_direction <= orientation_diff <= np.pi + self.tolerance_same_direction
            if self.debug:
                print("vehicles are on lanelets with opposite direction" if opposite
                      else "obstacle is coming from other direction")

            dx, dy = hf.vector_from_angle(dyn_obst.current_orientation)
            probe = np.array([dyn_obst.current_pos[0] + 4 * dx, dyn_obst.current_pos[1] + 4 * dy])
            region = self._occluded_region_raster(dyn_obst, possible_ids, opposite, probe)
            if region is None or region["area"] < self.min_area_threshold:
                continue
            center_pos = region["centroid"]
            centroid_lanelet = self.scenario.lanelet_network.find_lanelet_by_position([center_pos])[0]
            if not any(e in relevant_lanelets for e in centroid_lanelet):
                continue
            if region["contains_probe"]:
                continue

            rectangles = self._find_matching_rectangle(center_pos, region, dyn_obst, possible_ids, opposite)
            for key in rectangles:
                if rectangles[key]["area"] >= self.agent_area_limits[key] and rectangles[key]["jaccard_similarity"] > 0.98:
                    spawn_points.append(SpawnPoint(pos=rectangles[key]["centroid"], agent_type=key, pos_cl=None,
                                                   source="behind_dynamic_obstacle"))
        return spawn_points

    def _lanelet_index(self, lanelet_id):
        for k, l in enumerate(self.scenario.lanelet_network.lanelets):
            if l.lanelet_id == lanelet_id:
                return k
        raise KeyError(lanelet_id)

    def _region_predicate(self, P, dyn_obst, possible_ids, opposite):
        """Membership of points in  possible_polygon ∩ (obstacle shadow | occluded area) ∩ disc(12 m) − obstacle.buffer(1)
        (spawn_locator.py:254-275)."""
        k = self.sensor_model._obstacle_index[dyn_obst.cr_obstacle.obstacle_id]
        flags, _, lan = self.sensor_model._classify(P, focus=k, focus_margin=1.0)
        mask_bits = np.uint64(0)
        host_ids = []
        for lid in possible_ids:
            idx = self._lanelet_index(lid)
            if idx < 64:
                mask_bits |= np.uint64(1) << np.uint64(idx)
            else:
                host_ids.append(lid)
        inside = (lan & mask_bits) != 0
        for lid in host_ids:          # networks with more than 64 lanelets: remaining polygons on the host
            from .scenario import _points_in_polygon
            inside |= _points_in_polygon(P, self.scenario.lanelet_network.find_lanelet_by_id(lid).polygon_vertices)
        want = L.PT_FOCUS_SHADOW if opposite else L.PT_OCCLUDED
        inside &= (flags & want) != 0
        inside &= np.hypot(*(P - dyn_obst.current_pos).T) <= self.buffer_around_vehicle_from_side
        inside &= (flags & L.PT_FOCUS_NEAR) == 0          # - current_polygon.buffer(1), spawn_locator.py:275
        return inside

    def _lanelet_mask(self, possible_ids):
        """Bit mask of the lanelet polygons ``possible_ids`` for the device predicate; None when one of them lies beyond
        the 64 polygons the kernel reports (then the host path below is used)."""
        bits = 0
        for lid in possible_ids:
            idx = self._lanelet_index(lid)
            if idx >= 64:
                return None
            bits |= 1 << idx
        return bits

    def _occluded_region_raster(self, dyn_obst, possible_ids, opposite, probe):
        """Largest connected part of the relevant occluded area on a ``raster_cell`` grid (spawn_locator.py:270-287):
        generated, classified, labelled and reduced on the device (``fo_spawn_region``)."""
        c = self.raster_cell
        half = self.buffer_around_vehicle_from_side
        n = int(np.ceil(2 * half / c))
        bits = self._lanelet_mask(possible_ids)
        if bits is None:
            return self._occluded_region_raster_host(dyn_obst, possible_ids, opposite, probe)
        k = self.sensor_model._obstacle_index[dyn_obst.cr_obstacle.obstacle_id]
        want = L.PT_FOCUS_SHADOW if opposite else L.PT_OCCLUDED
        count, centroid, inside, _, handle = self.sensor_model._frame.spawn_region(
            dyn_obst.current_pos, half, c, n, bits, want, L.PT_FOCUS_NEAR, half, probe, k, 1.0)
        if count == 0:
            return None
        return {"area": float(count) * c * c, "centroid": centroid, "contains_probe": inside, "handle": handle}

    def _occluded_region_raster_host(self, dyn_obst, possible_ids, opposite, probe):
        """The same on host rasters (networks with more than 64 lanelet polygons)."""
        from scipy import ndimage
        c = self.raster_cell
        half = self.buffer_around_vehicle_from_side
        n = int(np.ceil(2 * half / c))
        ax = (np.arange(n) + 0.5) * c - half
        gx, gy = np.meshgrid(ax + dyn_obst.current_pos[0], ax + dyn_obst.current_pos[1], indexing="ij")
        P = np.stack((gx.ravel(), gy.ravel()), -1)
        inside = self._region_predicate(P, dyn_obst, possible_ids, opposite).reshape(n, n)
        if not inside.any():
            return None
        lab, cnt = ndimage.label(inside)
        sizes = ndimage.sum(inside, lab, index=np.arange(1, cnt + 1))
        best = int(np.argmax(sizes)) + 1
        comp = lab == best
        pts = P.reshape(n, n, 2)[comp]
        origin = np.array([dyn_obst.current_pos[0] - half, dyn_obst.current_pos[1] - half])
        ij = np.floor((np.asarray(probe, dtype=np.float64) - origin) / c).astype(int)
        inside_probe = bool(0 <= ij[0] < n and 0 <= ij[1] < n and comp[ij[0], ij[1]])
        return {"area": float(comp.sum()) * c * c, "centroid": pts.mean(0), "mask": comp, "origin": origin, "n": n,
                "contains_probe": inside_probe}

    def _find_matching_rectangle(self, position, region, dyn_obst, possible_ids, opposite):
        """spawn_locator.py:695-726: clip a 5.5 x 2.5 m box (car) and a 2 x 1 m box (bicycle) with the allowed area
        and rate how rectangular the remainder is (Jaccard index against its minimum rotated rectangle)."""
        _, orientation = self._find_orientation_at_position(position)
        if "handle" in region:       # both boxes in one device pass (``fo_spawn_rect``), one read-back
            c = self.rect_cell
            res = self.sensor_model._frame.spawn_rects(region["handle"], [(position, 5.5, 2.5, False), (position, 2.0, 1.0, True)],
                                                       orientation, c)
            out, centre = {}, np.asarray(position, dtype=np.float64)
            for key, (count, centroid, outline) in zip(("Car", "Bicycle"), res):
                area = float(count) * c * c
                if count < 3:
                    out[key] = {"area": area, "area_ratio": 0.0, "jaccard_similarity": 0.0, "centroid": centre}
                else:
                    _, w, h, ang = hf.min_area_rectangle(outline)
                    grow = c * (abs(np.cos(ang - orientation)) + abs(np.sin(ang - orientation)))
                    mbr_area = (w + grow) * (h + grow)
                    ratio = min(area / mbr_area, 1.0) if mbr_area > 0 else 0.0
                    out[key] = {"area": area, "area_ratio": ratio, "jaccard_similarity": ratio, "centroid": centroid}
                    centre = centroid          # the bicycle box is centred on what is left of the car box
            return out
        vehicle = self._clipped_rectangle_metrics(position, 5.5, 2.5, orientation, region, dyn_obst, possible_ids, opposite)
        bike_center = vehicle["centroid"] if vehicle["area"] > 0 else position
        bike = self._clipped_rectangle_metrics(bike_center, 2.0, 1.0, orientation, region, dyn_obst, possible_ids, opposite)
        return {"Car": vehicle, "Bicycle": bike}

    def _clipped_rectangle_metrics(self, center, length, width, orientation, region, dyn_obst, possible_ids, opposite):
        c = self.rect_cell
        nx, ny = int(round(length / c)), int(round(width / c))
        lx = (np.arange(nx) + 0.5) * c - 0.5 * length
        ly = (np.arange(ny) + 0.5) * c - 0.5 * width
        gx, gy = np.meshgrid(lx, ly, indexing="ij")
        cs, sn = np.cos(orientation), np.sin(orientation)
        P = np.stack((center[0] + gx.ravel() * cs - gy.ravel() * sn, center[1] + gx.ravel() * sn + gy.ravel() * cs), -1)
        inside = self._region_predicate(P, dyn_obst, possible_ids, opposite)
        # restrict to the selected connected part of the allowed area (coarse raster, one cell of slack)
        if "mask_dilated" not in region:
            from scipy import ndimage
            region["mask_dilated"] = ndimage.binary_dilation(region["mask"], iterations=1)
        comp = region["mask_dilated"]
        ij = np.floor((P - region["origin"]) / self.raster_cell).astype(int)
        ok = (ij[:, 0] >= 0) & (ij[:, 0] < region["n"]) & (ij[:, 1] >= 0) & (ij[:, 1] < region["n"])
        sel = np.zeros(len(P), dtype=bool)
        sel[ok] = comp[ij[ok, 0], ij[ok, 1]]
        inside &= sel
        area = float(inside.sum()) * c * c
        if inside.sum() < 3:
            return {"area": area, "area_ratio": 0.0, "jaccard_similarity": 0.0, "centroid": np.asarray(center, float)}
        pts = P[inside]
        # the minimum rotated rectangle only depends on the outline: cells with a missing 4-neighbour
        m = inside.reshape(nx, ny)
        pad = np.pad(m, 1)
        core = pad[:-2, 1:-1] & pad[2:, 1:-1] & pad[1:-1, :-2] & pad[1:-1, 2:]
        _, w, h, ang = hf.min_area_rectangle(P[(m & ~core).ravel()])
        # the samples are cell centres: the covered set extends half a cell beyond them on every side
        grow = c * (abs(np.cos(ang - orientation)) + abs(np.sin(ang - orientation)))
        mbr_area = (w + grow) * (h + grow)
        ratio = min(area / mbr_area, 1.0) if mbr_area > 0 else 0.0
        # polygon ⊂ its minimum rotated rectangle, so intersection/union = area ratio (spawn_locator.py:719-724)
        return {"area": area, "area_ratio": ratio, "jaccard_similarity": ratio, "centroid": pts.mean(0)}

    # ---- behind static obstacles (spawn_locator.py:323-476) ------------------------------------------------
    def _find_spawn_point_behind_static_obstacle(self):
        spawn_points = []
        s_positions = []
        visible_stat_obst = [o for o in self.fo_obstacles
                             if o.current_visible and o.cr_obstacle.obstacle_role.name == "STATIC"]
        visible_stat_obst = sorted(visible_stat_obst, key=lambda o: np.linalg.norm(self.ego_pos - o.current_pos))
        if not visible_stat_obst:
            return
        for stat_obst in visible_stat_obst:
            if len(spawn_points) > self.max_static_spawn_points:
                break
            if np.linalg.norm(self.ego_pos - stat_obst.current_pos) > self.max_distance_to_other_obstacle:
                continue
            try:
                stat_obstacle_cl = self.cosy_cl.convert_to_curvilinear_coords(stat_obst.current_pos[0], stat_obst.current_pos[1])
            except Exception:
                continue
            # (sic) s_threshold already contains the ego's s, spawn_locator.py:113,380
            if self.ego_cl[0] + self.s_threshold < stat_obstacle_cl[0] or stat_obstacle_cl[0] < self.ego_cl[0] + 3:
                continue
            list_of_corner_points = [np.array([[x], [y]]) for x, y in stat_obst.current_corner_points]
            list_of_corner_points_cl = np.array(self.cosy_cl.convert_list_of_points_to_curvilinear_coords(list_of_corner_points, 4))
            offset = 0.8
            s_min, s_max = np.min(list_of_corner_points_cl[:, 0]) - offset, np.max(list_of_corner_points_cl[:, 0]) + offset
            d_min, d_max = np.min(list_of_corner_points_cl[:, 1]) - offset, np.max(list_of_corner_points_cl[:, 1]) + offset
            lines_cl = [np.array([[s_min, d_min], [s_min, d_max]]), np.array([[s_max, d_min], [s_max, d_max]])]

            for line_cl in lines_cl:
                try:
                    line = np.array([self.cosy_cl.convert_to_cartesian_coords(p[0], p[1]) for p in line_cl])
                except Exception:
                    continue
                band = self._line_band(line)
                if band is None:
                    continue
                if not band["occluded_on_line"].any() or not band["visible_on_line"].any():
                    continue
                others = self.fo_obstacles.visible_obstacle_multipolygon or []
                if any(hf.segment_ring_distance(line[0], line[1], ring) <= self.ped_width / 2 for ring in others):
                    continue
                # intersection of the line with the outline of visible_area.buffer(ped_length / 2 * 1.3)
                cand = band["crossings"](self.ped_length / 2 * 1.3)
                if len(cand) == 0:
                    continue
                if len(cand) > 1:
                    obst_lanelet = self._find_lanelet_by_position(stat_obst.current_pos)
                    left_vertices_point = obst_lanelet.left_vertices[0]
                    cand = sorted(cand, key=lambda q: np.linalg.norm(left_vertices_point - q))
                    spawn_pos = None
                    for q in cand:
                        if self.sensor_model.occluded_area.contains(q):
                            spawn_pos = q
                            break
                    if spawn_pos is None:
                        continue
                else:
                    spawn_pos = cand[0]
                disc = hf.disc_samples(spawn_pos, 0.15)
                if self.sensor_model.visible_area.intersects_points(disc):
                    continue
                if not bool(np.all(self.sensor_model.road_polygon.contains(disc))):
                    continue
                spawn_pos_cl = self.cosy_cl.convert_to_curvilinear_coords(spawn_pos[0], spawn_pos[1])
                source = "behind static obstacle " + str(stat_obst.cr_obstacle.obstacle_id)
                if any(abs(s - spawn_pos_cl[0]) <= self.min_distance_between_pedestrians for s in s_positions):
                    continue
                _, orientation = self._find_orientation_at_position(stat_obst.current_pos)
                orientation = orientation + np.pi / 2
                spawn_points.append(SpawnPoint(pos=spawn_pos, agent_type="Pedestrian", pos_cl=spawn_pos_cl,
                                               source=source, orientation=orientation))
                s_positions.append(spawn_pos_cl[0])
                break
        return spawn_points

    def _line_band(self, line, margin=0.3):
        """Raster of a narrow band around the segment ``line`` (cells of ``line_step``), classified on the device.
        Gives the visible / occluded samples on the line itself and the points of the line at a given distance
        from the visible area (outline of ``visible_area.buffer(r)`` ∩ line, spawn_locator.py:414-415)."""
        from scipy import ndimage
        a, b = np.asarray(line[0], dtype=np.float64), np.asarray(line[1], dtype=np.float64)
        length = float(np.hypot(*(b - a)))
        if length < 1e-9:
            return None
        t = (b - a) / length
        nrm = np.array([-t[1], t[0]])
        h = self.line_step
        m = int(round(margin / h))
        nu = int(np.floor(length / h)) + 1
        u = np.arange(-m, nu + m) * h
        w = np.arange(-m, m + 1) * h
        P = a[None, None] + u[:, None, None] * t[None, None] + w[None, :, None] * nrm[None, None]
        flags, _, _ = self.sensor_model._classify(P.reshape(-1, 2))
        flags = flags.reshape(len(u), len(w))
        vis = (flags & L.PT_VISIBLE) != 0
        occ = (flags & L.PT_OCCLUDED) != 0
        on_line = slice(m, m + nu)
        dist = ndimage.distance_transform_edt(~vis) * h if vis.any() else np.full(vis.shape, np.inf)
        dline = dist[on_line, m]
        uline = u[on_line]

        def crossings(r):
            g = dline - r
            out = []
            for i in np.nonzero(np.sign(g[:-1]) * np.sign(g[1:]) < 0)[0]:
                f = g[i] / (g[i] - g[i + 1])
                out.append(a + (uline[i] + f * h) * t)
            for i in np.nonzero(g == 0)[0]:
                out.append(a + uline[i] * t)
            return out

        return {"visible_on_line": vis[on_line, m], "occluded_on_line": occ[on_line, m], "crossings": crossings}

    # ---- behind turns (spawn_locator.py:481-578) ---------------------------------------------------------------
    def _find_spawn_point_behind_turn(self, ego_intention):
        if ego_intention == "left turn":
            curvilinear_list = np.column_stack((self.reference_s, np.full(self.reference_s.shape,
                                                                           self.offset_ref_path["left turn"])))
            path = self._convert_curvilinear_list_to_cartesian_coordinates(curvilinear_list)
        else:
            path = np.asarray(self.reference, dtype=np.float64)
        if len(path) < 2:
            return
        # sample the path and find the runs lying in the occluded area (LineString ∩ occluded_area)
        cum = hf.compute_pathlength_from_polyline(path)
        ss = np.arange(0.0, cum[-1], self.path_step)
        P = np.stack((np.interp(ss, cum, path[:, 0]), np.interp(ss, cum, path[:, 1])), -1)
        occ = np.atleast_1d(self.sensor_model.occluded_area.contains(P))
        if not occ.any():
            return
        starts = np.nonzero(occ & ~np.concatenate(([False], occ[:-1])))[0]
        if len(starts) > 1 and self.debug:
            print("MultiLineString in spawn point processing detected")
        k = int(starts[-1] if len(starts) > 1 else starts[0])
        intersection = P[k]
        if k > 0:
            # the reference intersects the path with the occluded polygon exactly (spawn_locator.py:521-529): refine the
            # 5 cm bracket [not occluded, occluded] with ONE more classified batch of 1025 points -> 0.05 mm
            lo_s, hi_s = ss[k - 1], ss[k]
            cs = np.linspace(lo_s, hi_s, self.path_refine + 1)
            Pc = np.stack((np.interp(cs, cum, path[:, 0]), np.interp(cs, cum, path[:, 1])), -1)
            oc = np.atleast_1d(self.sensor_model.occluded_area.contains(Pc))
            oc[-1] = True
            j = int(np.argmax(oc))
            hi_s = cs[j]
            intersection = np.array([np.interp(hi_s, cum, path[:, 0]), np.interp(hi_s, cum, path[:, 1])])
        s_intersection = self.cosy_cl.convert_to_curvilinear_coords(intersection[0], intersection[1])[0]
        s_phantom = s_intersection + self.phantom_offset_s[ego_intention]
        if s_phantom > self.s_threshold or s_phantom < self.ego_cl[0] + 3:
            return
        d_offset = self.phantom_offset_d[ego_intention] + self.offset_ref_path[ego_intention]
        # walk s in +0.5 m steps until a 0.5 m disc no longer touches the visible area (spawn_locator.py:552-554);
        # the candidate discs are classified in batches of 32 steps (one device call each) instead of one by one
        phantom_pos = None
        vec = getattr(self.cosy_cl, "convert_array_to_cartesian_coords", None)
        for k0 in range(0, 416, 32):
            if vec is not None:          # vectorised conversion (the harness' coordinate system); NaN = outside the path
                cand = vec(s_phantom + 0.5 * np.arange(k0, k0 + 32), d_offset)
                bad = np.nonzero(np.isnan(cand[:, 0]))[0]
                cand = cand[:int(bad[0])] if len(bad) else cand
            else:
                cand = []
                for k in range(k0, k0 + 32):
                    try:
                        cand.append(np.asarray(self.cosy_cl.convert_to_cartesian_coords(s_phantom + 0.5 * k, d_offset)))
                    except Exception:
                        break           # the reference raises out of the coordinate system here
                cand = np.asarray(cand, dtype=np.float64).reshape(-1, 2)
            if not len(cand):
                return
            discs = hf.disc_samples_many(cand, 0.5).reshape(-1, 2)
            vis = np.atleast_1d(self.sensor_model.visible_area.contains(discs)).reshape(len(cand), -1).any(1)
            free = np.nonzero(~vis)[0]
            if len(free):
                s_phantom = s_phantom + 0.5 * (k0 + int(free[0]))
                phantom_pos = cand[int(free[0])]
                break
            if len(cand) < 32:
                return
        if phantom_pos is None:      # the reference would keep walking until the coordinate system raises
            return
        others = self.fo_obstacles.visible_obstacle_multipolygon or []
        if hf.point_rings_min_distance(phantom_pos, others) <= 0.5:
            return
        _, ego_lanelet_orientation = self._find_orientation_at_position(self.ego_pos)
        _, phantom_lanelet_orientation = self._find_orientation_at_position(phantom_pos)
        orientation_diff = abs(phantom_lanelet_orientation - ego_lanelet_orientation) % (2 * np.pi)
        if orientation_diff < np.radians(45):
            return
        return SpawnPoint(pos=phantom_pos, agent_type="Pedestrian",
                          pos_cl=[s_phantom, self.phantom_offset_d[ego_intention]], source=ego_intention)

    # ---- helpers (spawn_locator.py:583-752) ------------------------------------------------------------------------
    @staticmethod
    def _find_intersection_lanelets(intersection):
        incoming_lanelets, intersection_lanelets = set(), set()
        for el in intersection.incomings:
            incoming_lanelets.update(el.incoming_lanelets)
            intersection_lanelets.update(el.successors_left)
            intersection_lanelets.update(el.successors_right)
            intersection_lanelets.update(el.successors_straight)
        return incoming_lanelets, intersection_lanelets

    def _find_intersection_outgoings(self, intersection_incomings):
        return [self.scenario.lanelet_network.find_lanelet_by_id(i).adj_left for i in intersection_incomings]

    def _find_relevant_intersection(self, intersections, ego_lanelet):
        for intersection in intersections:
            incoming_lanelets, intersection_lanelets = self._find_intersection_lanelets(intersection)
            if ego_lanelet.lanelet_id in incoming_lanelets or ego_lanelet.lanelet_id in intersection_lanelets:
                return intersection, incoming_lanelets, intersection_lanelets
        return None, None, None

    def _find_lanelets_along_reference(self, reference, step=5):
        lanelets = []
        for i in range(0, len(reference), step):
            lanelet = self._find_lanelet_by_position(reference[i])
            if lanelet is not None and lanelet not in lanelets:
                lanelets.append(lanelet)
        return lanelets

    def _append_spawn_point(self, spawn_point):
        if spawn_point is not None:
            if type(spawn_point) is list:
                self.spawn_points.extend([p for p in spawn_point if p is not None])
            else:
                self.spawn_points.append(spawn_point)

    def _convert_curvilinear_list_to_cartesian_coordinates(self, curvilinear_list):
        vec = getattr(self.cosy_cl, "convert_array_to_cartesian_coords", None)
        if vec is not None:
            cl = np.asarray(curvilinear_list, dtype=np.float64).reshape(-1, 2)
            out = vec(cl[:, 0], cl[:, 1])
            if np.isnan(out).any():
                raise ValueError("longitudinal coordinate outside of the reference path")
            return out
        return np.array([self.cosy_cl.convert_to_cartesian_coords(item[0], item[1]) for item in curvilinear_list])

    def _find_orientation_at_position(self, pos):
        lanelet = self._find_lanelet_by_position(pos)
        return lanelet, lanelet_orientation_at_position(lanelet, pos)

    def _find_lanelet_by_position(self, pos):
        lanelet_id = self.scenario.lanelet_network.find_lanelet_by_position([pos])
        if not lanelet_id[0]:
            return None
        return self.scenario.lanelet_network.find_lanelet_by_id(lanelet_id[0][0])

    def _prepare_reference_path(self, ego_cl, distance=40):
        self.s = hf.compute_pathlength_from_polyline(self.ref_path)
        index_start = self._find_nearest_index(self.s, ego_cl[0])
        index_end = self._find_nearest_index(self.s, ego_cl[0] + distance)
        return self.ref_path[index_start:index_end], self.s[index_start:index_end]

    @staticmethod
    def _find_ego_intention(reference):
        curvature = hf.compute_curvature_from_polyline(reference)
        if max(curvature) > 0.10:
            return "left turn"
        elif min(curvature) < -0.10:
            return "right turn"
        return "straight ahead"

    @staticmethod
    def _find_nearest_index(path_s, current_s):
        return int(np.argmin(np.abs(path_s - current_s)))
