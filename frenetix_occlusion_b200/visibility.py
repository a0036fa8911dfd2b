"""Host wrapper of the visibility ray-cast kernel (stage 1) -- device tensors in, device tensors out."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib as L


@dataclass
class VisibilityResult:
    range: torch.Tensor      # float32 [F, R]
    hit: torch.Tensor        # int32 [F, R]   obstacle index | HIT_NONE | HIT_BOUNDARY
    visible: torch.Tensor    # uint8 [F, O]
    angles0: np.ndarray      # [F] first ray angle (float64, host)
    dangle: float            # angular step


def ray_angle_params(heading, fov_deg: float, n_rays: int):
    """Same convention as the kernel / oracle: full circle when fov >= 359.9 deg (sensor_model.py:119-120)."""
    heading = np.asarray(heading, dtype=np.float64)
    if fov_deg >= 359.9:
        return heading - np.pi, 2.0 * np.pi / n_rays
    fov = np.radians(fov_deg)
    return heading - 0.5 * fov, fov / max(n_rays - 1, 1)


def _dev(t, dtype, device):
    if isinstance(t, torch.Tensor):
        return t.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(t), dtype=dtype).to(device)


def raycast_frames(ego, rect, rect_flags, boundary, sensor_radius: float, sensor_angle_deg: float, n_rays: int,
                   device="cuda:0", out: Optional[VisibilityResult] = None) -> VisibilityResult:
    """``ego`` [F,3] (x, y, heading); ``rect`` [F,O,5] (cx, cy, yaw, half_len, half_wid); ``rect_flags``
    [F,O] uint8 (RECT_EXISTS | RECT_TRANSPARENT); ``boundary`` [B,4] opaque segments or None.
    Asynchronous on the current stream."""
    if not torch.cuda.is_available():
        raise RuntimeError("frenetix_occlusion_b200 needs a CUDA device (no CPU fallback)")
    device = torch.device(device)
    ego_d = _dev(ego, torch.float32, device).reshape(-1, 3)
    F = ego_d.shape[0]
    rect_d = _dev(rect, torch.float32, device).reshape(F, -1, 5)
    O = rect_d.shape[1]
    flags_d = _dev(rect_flags, torch.uint8, device).reshape(F, O)
    bnd_d = None if boundary is None or len(boundary) == 0 else _dev(boundary, torch.float32, device).reshape(-1, 4)
    with torch.cuda.device(device):
        if out is None:
            out = VisibilityResult(torch.empty((F, n_rays), dtype=torch.float32, device=device),
                                   torch.empty((F, n_rays), dtype=torch.int32, device=device),
                                   torch.zeros((F, max(O, 1)), dtype=torch.uint8, device=device)[:, :O], None, 0.0)
        a = L.FoVisibilityArgs()
        a.n_frames, a.n_rays, a.n_obstacles = F, n_rays, O
        a.n_boundary = 0 if bnd_d is None else bnd_d.shape[0]
        a.ego, a.rect, a.rect_flags = ego_d.data_ptr(), rect_d.data_ptr() if O else None, flags_d.data_ptr() if O else None
        a.boundary = None if bnd_d is None else bnd_d.data_ptr()
        a.sensor_radius, a.sensor_angle_deg = float(sensor_radius), float(sensor_angle_deg)
        a.range, a.hit = out.range.data_ptr(), out.hit.data_ptr()
        a.visible = out.visible.data_ptr() if O else None
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        for f0 in range(0, F, 65535):                      # grid.y limit
            f1 = min(F, f0 + 65535)
            b = L.FoVisibilityArgs.from_buffer_copy(a)
            b.n_frames = f1 - f0
            b.ego = ego_d[f0:].data_ptr()
            b.rect = rect_d[f0:].data_ptr() if O else None
            b.rect_flags = flags_d[f0:].data_ptr() if O else None
            b.range, b.hit = out.range[f0:].data_ptr(), out.hit[f0:].data_ptr()
            b.visible = out.visible[f0:].data_ptr() if O else None
            L.check(L.lib.fo_visibility_raycast(C.byref(b), stream), "fo_visibility_raycast")
    out._keepalive = (ego_d, rect_d, flags_d, bnd_d)
    if isinstance(ego, torch.Tensor):
        out.angles0, out.dangle = None, 0.0
    else:
        a0, da = ray_angle_params(np.asarray(ego, dtype=np.float64).reshape(-1, 3)[:, 2], sensor_angle_deg, n_rays)
        out.angles0, out.dangle = a0, da
    return out


class FrameGeometry:
    """One sensor frame resident on the device: ego pose, obstacle rectangles, opaque road-border segments and
    the lanelet polygons -- everything ``fo_visibility_raycast`` and ``fo_visibility_points`` read.  All
    coordinates are relative to ``origin`` (subtracted in float64 before the float32 cast)."""

    def __init__(self, origin, heading, rect, rect_flags, boundary, polygons, sensor_radius, sensor_angle_deg,
                 occluded_factor=1.5, device="cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("frenetix_occlusion_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.origin = np.asarray(origin, dtype=np.float64).reshape(2)
        self.heading = float(heading)
        self.sensor_radius, self.sensor_angle_deg = float(sensor_radius), float(sensor_angle_deg)
        self.occluded_radius = float(occluded_factor) * float(sensor_radius)
        rect = np.array(rect, dtype=np.float64).reshape(-1, 5)
        rect[:, 0] -= self.origin[0]
        rect[:, 1] -= self.origin[1]
        self.n_obstacles = len(rect)
        bnd = np.zeros((0, 4)) if boundary is None else np.asarray(boundary, dtype=np.float64).reshape(-1, 4)
        bnd = bnd - np.tile(self.origin, 2)
        self.n_boundary = len(bnd)
        polys = [np.asarray(p, dtype=np.float64).reshape(-1, 2) - self.origin for p in (polygons or [])]
        self.n_polygons = len(polys)
        off = np.concatenate(([0], np.cumsum([len(p) for p in polys]))).astype(np.int32)
        xy = np.concatenate(polys) if polys else np.zeros((0, 2))
        self.host = {"ego": np.array([0.0, 0.0, self.heading]), "rect": rect, "flags": np.asarray(rect_flags, np.uint8).reshape(-1),
                     "boundary": bnd, "polygons": polys}
        with torch.cuda.device(self.device):
            self.ego_d = torch.tensor([[0.0, 0.0, self.heading]], dtype=torch.float32, device=self.device)
            self.rect_d = torch.from_numpy(rect.astype(np.float32)).to(self.device)
            self.flags_d = torch.from_numpy(self.host["flags"].copy()).to(self.device)
            self.bnd_d = torch.from_numpy(bnd.astype(np.float32)).to(self.device)
            self.poly_xy_d = torch.from_numpy(xy.astype(np.float32)).to(self.device)
            self.poly_off_d = torch.from_numpy(off).to(self.device)

    def raycast(self, n_rays: int) -> VisibilityResult:
        O = self.n_obstacles
        return raycast_frames(self.ego_d, self.rect_d.reshape(1, O, 5), self.flags_d.reshape(1, O),
                              self.bnd_d if self.n_boundary else None, self.sensor_radius, self.sensor_angle_deg,
                              n_rays, device=self.device)

    def classify(self, points, focus_obstacle: int = -1, focus_margin: float = 0.0):
        """Classify world-frame points [M,2]: returns host arrays (flags uint32, blocker int32, lanelets uint64)."""
        P = np.asarray(points, dtype=np.float64).reshape(-1, 2) - self.origin
        M = len(P)
        if M == 0:
            return np.zeros(0, np.uint32), np.zeros(0, np.int32), np.zeros(0, np.uint64)
        dev = self.device
        with torch.cuda.device(dev):
            pts = torch.from_numpy(P.astype(np.float32)).to(dev)
            flags = torch.empty(M, dtype=torch.int32, device=dev)
            blocker = torch.empty(M, dtype=torch.int32, device=dev)
            lan = torch.empty(M, dtype=torch.int64, device=dev)
            a = L.FoPointQueryArgs()
            a.n_points, a.n_obstacles, a.n_boundary, a.n_polygons = M, self.n_obstacles, self.n_boundary, self.n_polygons
            a.ego, a.points = self.ego_d.data_ptr(), pts.data_ptr()
            a.rect = self.rect_d.data_ptr() if self.n_obstacles else None
            a.rect_flags = self.flags_d.data_ptr() if self.n_obstacles else None
            a.boundary = self.bnd_d.data_ptr() if self.n_boundary else None
            a.poly_xy = self.poly_xy_d.data_ptr() if self.n_polygons else None
            a.poly_off = self.poly_off_d.data_ptr() if self.n_polygons else None
            a.sensor_radius, a.sensor_angle_deg = self.sensor_radius, self.sensor_angle_deg
            a.occluded_radius, a.focus_obstacle = self.occluded_radius, int(focus_obstacle)
            a.focus_margin = float(focus_margin)
            a.flags, a.blocker, a.lanelets = flags.data_ptr(), blocker.data_ptr(), lan.data_ptr()
            L.check(L.lib.fo_visibility_points(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                    "fo_visibility_points")
            out = (flags.cpu().numpy().view(np.uint32), blocker.cpu().numpy(), lan.cpu().numpy().view(np.uint64))
        return out
