"""Planner-side bundle adapter: the planner's sampled trajectories -> one ``[N, T, 5]`` array
(x, y, theta, v, a at the rear axle), the layout ``FOInterface.assess_bundle`` / ``fo_metric_bundle`` consume.

The reference receives ONE trajectory object per call (interface.py:216-219) and re-reads
``trajectory.cartesian.{x,y,theta,v,a}`` inside every metric.  Here the whole bundle is packed once per planning
cycle into a reusable page-locked buffer, so the upload is a single asynchronous copy."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

_FIELDS = ("x", "y", "theta", "v", "a")


class BundlePacker:
    """Reusable page-locked staging buffer for bundles of up to ``capacity`` trajectories of ``n_states`` states."""

    def __init__(self, capacity: int, n_states: int, pin: bool = True):
        self.capacity, self.n_states = int(capacity), int(n_states)
        self.host = torch.empty((self.capacity, self.n_states, 5), dtype=torch.float32)
        if pin and torch.cuda.is_available():
            self.host = self.host.pin_memory()
        self._np = self.host.numpy()

    def pack(self, trajectories: Sequence, origin=(0.0, 0.0)) -> torch.Tensor:
        """Trajectory objects (duck type: ``.cartesian.{x,y,theta,v,a}``) -> float32 view ``[N, T, 5]`` of the buffer.
        ``origin`` is subtracted from x / y in float64 before the float32 store."""
        n = len(trajectories)
        if n > self.capacity:
            raise ValueError(f"bundle of {n} trajectories exceeds the packer's capacity {self.capacity}")
        buf = self._np
        ox, oy = float(origin[0]), float(origin[1])
        for k, tr in enumerate(trajectories):
            c = tr.cartesian
            if len(c.x) != self.n_states:
                raise ValueError("all trajectories of a bundle must have the packer's number of states")
            row = buf[k]
            row[:, 0] = np.asarray(c.x, dtype=np.float64) - ox
            row[:, 1] = np.asarray(c.y, dtype=np.float64) - oy
            row[:, 2] = c.theta
            row[:, 3] = c.v
            row[:, 4] = c.a
        return self.host[:n]

    def pack_columns(self, x, y, theta, v, a, origin=(0.0, 0.0)) -> torch.Tensor:
        """Already-stacked ``[N, T]`` arrays (e.g. the planner's own SoA export) -> ``[N, T, 5]`` view."""
        x = np.asarray(x, dtype=np.float64)
        n = x.shape[0]
        if n > self.capacity or x.shape[1] != self.n_states:
            raise ValueError("shape does not fit the packer")
        buf = self._np[:n]
        buf[..., 0] = x - float(origin[0])
        buf[..., 1] = np.asarray(y, dtype=np.float64) - float(origin[1])
        buf[..., 2], buf[..., 3], buf[..., 4] = theta, v, a
        return self.host[:n]


def bundle_from_trajectories(trajectories: Sequence, packer: Optional[BundlePacker] = None) -> np.ndarray:
    """One-shot form: float64 ``[N, T, 5]`` array (no staging buffer)."""
    if packer is not None:
        return packer.pack(trajectories).numpy()
    return np.stack([np.stack([np.asarray(getattr(t.cartesian, f), dtype=np.float64) for f in _FIELDS], -1)
                     for t in trajectories])
