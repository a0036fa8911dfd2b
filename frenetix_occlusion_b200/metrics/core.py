"""Shared device evaluation behind the metric plugins.

One ``MetricCore`` serves all plugin objects of a ``Metric``: the first plugin asked about a
trajectory triggers ONE launch of the dense kernel (pair + step detail for a 1-trajectory bundle),
the others read the cached device results.  Agents are re-packed whenever the agent manager's
``predictions`` change (new planning cycle).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from ..engine import AgentSet, BundleResult, MetricEngine

_ALL = ["hr", "ttc", "be", "ttce", "dce", "wttc", "cp"]


def trajectory_to_array(trajectory) -> np.ndarray:
    """``trajectory.cartesian.{x,y,theta,v,a}`` -> float64 [T, 5] (SURVEY.md 8b duck type)."""
    c = trajectory.cartesian
    return np.stack([np.asarray(c.x, dtype=np.float64), np.asarray(c.y, dtype=np.float64),
                     np.asarray(c.theta, dtype=np.float64), np.asarray(c.v, dtype=np.float64),
                     np.asarray(c.a, dtype=np.float64)], axis=-1)


class MetricCore:
    def __init__(self, vehicle_params, agent_manager, activated_metrics=None, thresholds=None, harm_coeffs=None,
                 device="cuda:0"):
        self.agent_manager = agent_manager
        self.vehicle_params = vehicle_params
        thresholds = thresholds or {}
        self.engine = MetricEngine(vehicle_params, agent_manager.dt, list(activated_metrics or _ALL), thresholds,
                                   harm_coeffs=harm_coeffs, device=device)
        self._fingerprint = None
        self._cache_key = None
        self._cache = None

    # ---- agents --------------------------------------------------------------------------------
    def _agents_fingerprint(self):
        preds = self.agent_manager.predictions
        ver = getattr(self.agent_manager, "version", None)
        return (id(preds), ver, tuple((k, id(v.get("pos_list")), len(v.get("pos_list"))) for k, v in preds.items()))

    def sync_agents(self, origin=None):
        fp = self._agents_fingerprint()
        if fp == self._fingerprint:
            return
        preds = self.agent_manager.predictions
        agents = AgentSet.from_predictions(preds, self.agent_manager.agent_by_prediction_id)
        if origin is None and agents.n_agents:
            origin = (float(agents.x[0, 0]), float(agents.y[0, 0]))
        self.engine.set_agents(agents, origin=origin)
        self._fingerprint = fp
        self._cache_key = None

    # ---- one trajectory, full detail --------------------------------------------------------------
    def detail(self, trajectory):
        self.sync_agents()
        key = (id(trajectory), id(trajectory.cartesian.x))
        if key == self._cache_key and self._cache is not None:
            return self._cache
        ego = trajectory_to_array(trajectory)
        A = self.engine.n_agents
        T = ego.shape[0]
        # all outputs of the 1-trajectory launch live in ONE device buffer -> one device-to-host copy
        n_pair, n_step = A * L.FO_PAIR_K, A * max(T - 1, 0) * L.FO_STEP_K
        off_flags, off_sum = 16, 32
        off_pair = off_sum + 4 * ((L.FO_SUMMARY_K + 3) // 4 * 4)
        off_step = off_pair + 4 * n_pair
        total = off_step + 4 * n_step
        dev = self.engine.device
        buf = torch.empty(max(total, 64), dtype=torch.uint8, device=dev)
        out = BundleResult(buf[0:1], buf[off_sum:off_sum + 4 * L.FO_SUMMARY_K].view(torch.float32).view(1, L.FO_SUMMARY_K),
                           buf[off_flags:off_flags + 4].view(torch.int32))
        out.pair = buf[off_pair:off_pair + 4 * n_pair].view(torch.float32).view(1, A, L.FO_PAIR_K)
        out.step = buf[off_step:off_step + 4 * n_step].view(torch.float32).view(1, A, max(T - 1, 0), L.FO_STEP_K)
        self.engine.assess(ego[None], out=out)
        host = buf.cpu().numpy()          # synchronises the stream
        f32 = lambda lo, n: host[lo:lo + 4 * n].view(np.float32).astype(np.float64)  # noqa: E731
        d = {"valid": bool(host[0]), "flags": int(host[off_flags:off_flags + 4].view(np.uint32)[0]),
             "summary": f32(off_sum, L.FO_SUMMARY_K),
             "pair": f32(off_pair, n_pair).reshape(A, L.FO_PAIR_K) if A else np.zeros((0, L.FO_PAIR_K)),
             "step": f32(off_step, n_step).reshape(A, T - 1, L.FO_STEP_K) if (A and T > 1) else np.zeros((A, 0, L.FO_STEP_K)),
             "ids": list(self.engine.agents.ids) if self.engine.agents is not None else [],
             "n_states": self.engine.agents.n_states if self.engine.agents is not None else np.zeros(0, int),
             "T": T}
        self._cache_key, self._cache = key, d
        return d

    # ---- a whole bundle, masks + summaries only ------------------------------------------------------
    def bundle(self, ego, want_pair=False, want_step=False):
        self.sync_agents()
        return self.engine.assess(ego, want_pair=want_pair, want_step=want_step)


_cores = {}


def shared_core(vehicle_params, agent_manager) -> MetricCore:
    """Core used by plugin objects constructed stand-alone (all seven metrics enabled)."""
    key = (id(vehicle_params), id(agent_manager))
    core = _cores.get(key)
    if core is None or core.agent_manager is not agent_manager:
        core = MetricCore(vehicle_params, agent_manager)
        _cores.clear()
        _cores[key] = core
    return core
