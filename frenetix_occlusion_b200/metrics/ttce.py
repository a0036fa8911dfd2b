"""Time-to-closest-encounter metric (reference frenetix_occlusion/metrics/ttce.py:28-43)."""
import numpy as np


class TTCE:
    def __init__(self, agent_manager):
        self.agent_manager = agent_manager

    def __repr__(self):
        return "<'Time to Closest Encounter Metric': {}.{} object at {}>".format(
            self.__class__.__module__, self.__class__.__name__, hex(id(self)))

    def evaluate(self, trajectory, results) -> dict:
        if "dce" in results:
            return {key: np.round(v["time_dce"] * self.agent_manager.dt, 3) for key, v in results["dce"].items()}
        raise ValueError("DCE is not available in results, but is needed to evaluate TTCE metric!")
