"""Obstacle wrappers (mirror of reference utils/fo_obstacle.py:18-116), polygon-library free."""
from __future__ import annotations

import numpy as np


def calc_corner_points(pos, orientation, obstacle_shape):
    """helper_functions.py:99-112: ``obstacle_shape.vertices[0:4]`` rotated by the orientation, plus pos."""
    v = np.asarray(obstacle_shape.vertices, dtype=np.float64)[0:4]
    c, s = np.cos(orientation), np.sin(orientation)
    rot = np.array([[c, -s], [s, c]])
    return v @ rot.T + np.asarray(pos, dtype=np.float64)


class FOObstacle:
    def __init__(self, obst):
        self.cr_obstacle = obst
        self.initial_timestep = obst.initial_state.time_step
        self.global_timestep = None
        self.relative_time_step = None
        self.current_pos = None
        self.current_pos_point = None
        self.current_orientation = None
        self.current_corner_points = None
        self.current_polygon = None
        self.current_visible = False
        self.last_visible_at_ts = None
        if self.cr_obstacle.obstacle_role.name == "STATIC":
            self._get_values_from_initial_state()

    def update_at_timestep(self, timestep):
        """fo_obstacle.py:79-94: relative step 0 -> initial state, k >= 1 -> state_list[k-1], beyond -> None."""
        self.global_timestep = timestep
        self.relative_time_step = timestep - self.initial_timestep
        self.current_visible = False
        if self.cr_obstacle.obstacle_role.name == "DYNAMIC":
            if self.relative_time_step == 0:
                self._get_values_from_initial_state()
            elif self.relative_time_step >= 1:
                idx = self.relative_time_step - 1
                if idx < len(self.cr_obstacle.prediction.trajectory.state_list):
                    self._set(self.cr_obstacle.prediction.trajectory.state_list[idx])
                else:
                    self._set_all_values_to_none()

    def _get_values_from_initial_state(self):
        self._set(self.cr_obstacle.initial_state)

    def _set(self, state):
        self.current_pos = np.asarray(state.position, dtype=np.float64)
        self.current_pos_point = self.current_pos
        self.current_orientation = float(state.orientation)
        self.current_corner_points = calc_corner_points(self.current_pos, self.current_orientation,
                                                        self.cr_obstacle.obstacle_shape)
        self.current_polygon = self.current_corner_points

    def _set_all_values_to_none(self):
        self.current_pos = None
        self.current_orientation = None
        self.current_corner_points = None
        self.current_polygon = None
        self.current_pos_point = None

    def as_rect(self):
        """(cx, cy, yaw, half_length, half_width) of the current corner ring (the shape's own centre offset
        and orientation are folded in, so scenario2's off-centre rectangle is handled)."""
        c = self.current_corner_points
        ctr = c.mean(0)
        e = c[3] - c[0]        # (-l,-w) -> (+l,-w): length axis
        yaw = np.arctan2(e[1], e[0])
        return (ctr[0], ctr[1], yaw, 0.5 * np.hypot(*e), 0.5 * np.hypot(*(c[1] - c[0])))


class FOObstacles:
    def __init__(self, cr_obstacles):
        self.cr_obstacles = cr_obstacles
        self.fo_obstacles = [FOObstacle(o) for o in cr_obstacles]
        self.visible_obstacle_multipolygon = None

    def __iter__(self):
        return iter(self.fo_obstacles)

    def __len__(self):
        return len(self.fo_obstacles)

    def add(self, cr_obstacle):
        self.cr_obstacles.append(cr_obstacle) if cr_obstacle not in self.cr_obstacles else None
        self.fo_obstacles.append(FOObstacle(cr_obstacle))

    def update(self, timestep):
        for o in self.fo_obstacles:
            o.update_at_timestep(timestep)

    def update_multipolygon(self):
        """fo_obstacle.py:45-47: corner rings of the currently visible obstacles."""
        self.visible_obstacle_multipolygon = [o.current_polygon for o in self.fo_obstacles if o.current_visible]
