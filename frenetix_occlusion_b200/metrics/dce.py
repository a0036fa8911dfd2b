"""Distance-to-closest-encounter metric (reference frenetix_occlusion/metrics/dce.py:30-99)."""
import numpy as np

from .core import shared_core


class DCE:
    def __init__(self, vehicle_params, agent_manager, core=None):
        self.vehicle_params = vehicle_params
        self.agent_manager = agent_manager
        self._core = core

    def __repr__(self):
        return "<'Distance to Closest Encounter Metric': {}.{} object at {}>".format(
            self.__class__.__module__, self.__class__.__name__, hex(id(self)))

    def evaluate(self, trajectory, results) -> dict:
        """``{prediction_id: {'dce': float (rounded to 1e-3), 'time_dce': int}}`` (dce.py:41-50)."""
        core = self._core or shared_core(self.vehicle_params, self.agent_manager)
        d = core.detail(trajectory)
        out = {}
        for k, pid in enumerate(d["ids"]):
            dce = d["pair"][k, 0]
            out[pid] = {"dce": np.round(dce, 3) if np.isfinite(dce) else np.inf, "time_dce": int(d["pair"][k, 1])}
        return out
