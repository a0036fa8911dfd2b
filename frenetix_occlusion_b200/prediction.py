"""Host wrapper of the phantom-agent rollout kernels (stage 2)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def rollout_cv(x0, y0, v, phi, dt: float, horizon: float, var0: float = 0.1, var_factor: float = 1.05,
               origin=(0.0, 0.0), device="cuda:0"):
    """Constant-velocity pedestrian predictions (reference agent.py:451-536) for A agents at once.
    Returns float32 device tensors ``x, y, yaw, v, var`` of shape [A, T], T = int(horizon/dt)+1,
    positions relative to ``origin`` (subtracted in float64)."""
    if not torch.cuda.is_available():
        raise RuntimeError("frenetix_occlusion_b200 needs a CUDA device (no CPU fallback)")
    device = torch.device(device)
    org = np.asarray(origin, dtype=np.float64)          # (2,) shared, or [A, 2] one origin per agent
    x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64)) - org[..., 0]
    y0 = np.atleast_1d(np.asarray(y0, dtype=np.float64)) - org[..., 1]
    v = np.atleast_1d(np.asarray(v, dtype=np.float64))
    phi = np.atleast_1d(np.asarray(phi, dtype=np.float64))
    A = len(x0)
    T = int(horizon / dt) + 1
    with torch.cuda.device(device):
        inp = torch.from_numpy(np.stack([x0, y0, v, phi])).to(device)
        out = torch.empty((6, A, T), dtype=torch.float32, device=device)
        a = L.FoRolloutCvArgs()
        a.n_agents, a.n_states, a.t_stride, a.dt = A, T, T, float(dt)
        a.var0, a.var_factor = float(var0), float(var_factor)
        a.x0, a.y0, a.v, a.phi = [inp[i].data_ptr() for i in range(4)]
        a.x, a.y, a.yaw, a.vel, a.var_x, a.var_y = [out[i].data_ptr() for i in range(6)]
        L.check(L.lib.fo_rollout_cv(C.byref(a), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
                "fo_rollout_cv")
    return {"x": out[0], "y": out[1], "yaw": out[2], "v": out[3], "var": out[4], "packed": out, "_keepalive": inp}


def rollout_path(paths, x0, y0, v0, dt: float, horizon: float, t1: float = 3.0, var0: float = 0.1,
                 var_factor: float = 1.05, origin=(0.0, 0.0), device="cuda:0"):
    """Path-following vehicle predictions (reference agent.py:283-426, utils/frenetix_handler.py:66-125):
    one job per reference polyline in ``paths`` (list of [P,2] arrays, 2 <= P <= 1024).  Returns float32
    device tensors ``x, y, yaw, v, var`` [J, T] plus ``sample`` [J] (selected Frenet sample, -1 = invalid)."""
    if not torch.cuda.is_available():
        raise RuntimeError("frenetix_occlusion_b200 needs a CUDA device (no CPU fallback)")
    device = torch.device(device)
    J = len(paths)
    T = int(horizon / dt) + 1
    org = np.asarray(origin, dtype=np.float64)          # (2,) shared, or [J, 2] one origin per job
    orgs = np.broadcast_to(org, (J, 2))
    pts = [np.asarray(p, dtype=np.float64).reshape(-1, 2) - orgs[j] for j, p in enumerate(paths)]
    if any(len(p) > 1024 for p in pts):
        raise ValueError("reference paths are limited to 1024 points (resample coarser)")
    off = np.concatenate(([0], np.cumsum([len(p) for p in pts]))).astype(np.int32)
    xy = np.concatenate(pts).astype(np.float32) if J else np.zeros((0, 2), np.float32)
    st = np.stack([np.atleast_1d(np.asarray(x0, dtype=np.float64)) - orgs[:, 0],
                   np.atleast_1d(np.asarray(y0, dtype=np.float64)) - orgs[:, 1],
                   np.atleast_1d(np.asarray(v0, dtype=np.float64))])
    with torch.cuda.device(device):
        d_xy, d_off, d_st = (torch.from_numpy(a).to(device) for a in (xy, off, st))
        out = torch.empty((6, J, T), dtype=torch.float32, device=device)
        smp = torch.empty(J, dtype=torch.int32, device=device)
        a = L.FoRolloutPathArgs()
        a.n_jobs, a.n_states, a.t_stride, a.dt, a.t1 = J, T, T, float(dt), float(t1)
        a.var0, a.var_factor = float(var0), float(var_factor)
        a.path_xy, a.path_off = d_xy.data_ptr(), d_off.data_ptr()
        a.x0, a.y0, a.v0 = [d_st[i].data_ptr() for i in range(3)]
        a.x, a.y, a.yaw, a.vel, a.var_x, a.var_y = [out[i].data_ptr() for i in range(6)]
        a.sample = smp.data_ptr()
        L.check(L.lib.fo_rollout_path(C.byref(a), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
                "fo_rollout_path")
    return {"x": out[0], "y": out[1], "yaw": out[2], "v": out[3], "var": out[4], "sample": smp, "packed": out,
            "_keepalive": (d_xy, d_off, d_st)}
