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
from ..engine import AgentSet, MetricEngine

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
        r = self.engine.assess(ego[None], want_pair=True, want_step=True)
        torch.cuda.current_stream(self.engine.device).synchronize()
        A = self.engine.n_agents
        T = ego.shape[0]
        d = {"valid": bool(r.valid[0].item()), "flags": int(r.flags[0].item()) & 0xffffffff,
             "summary": r.summary[0].cpu().numpy().astype(np.float64),
             "pair": r.pair[0].cpu().numpy().astype(np.float64) if A else np.zeros((0, L.FO_PAIR_K)),
             "step": r.step[0].cpu().numpy().astype(np.float64) if (A and T > 1) else np.zeros((A, 0, L.FO_STEP_K)),
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
