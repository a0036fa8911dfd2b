"""Collision-probability metric (reference frenetix_occlusion/metrics/cp.py:25-42)."""
import numpy as np

from .core import shared_core


class CP:
    def __init__(self, vehicle_params, agent_manager, core=None):
        self.vehicle_params = vehicle_params
        self.agent_manager = agent_manager
        self._core = core

    def __repr__(self):
        return "<'Collision Probability Metric': {}.{} object at {}>".format(
            self.__class__.__module__, self.__class__.__name__, hex(id(self)))

    def evaluate(self, trajectory, results) -> dict:
        """``{prediction_id: ndarray[T-1]}``; entry j is the CP of ego step j+1
        (collision_probability.py:69-124)."""
        core = self._core or shared_core(self.vehicle_params, self.agent_manager)
        d = core.detail(trajectory)
        return {pid: np.array(d["step"][k, :, 0]) for k, pid in enumerate(d["ids"])}
