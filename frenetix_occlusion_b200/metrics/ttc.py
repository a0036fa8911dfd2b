"""Time-to-collision metric (reference frenetix_occlusion/metrics/ttc.py:28-49)."""
import numpy as np


class TTC:
    def __init__(self, agent_manager):
        self.agent_manager = agent_manager

    def __repr__(self):
        return "<'Time to Collision Metric': {}.{} object at {}>".format(
            self.__class__.__module__, self.__class__.__name__, hex(id(self)))

    def evaluate(self, trajectory, results) -> dict:
        if "dce" in results:
            ttc = {}
            for key, dce in results["dce"].items():
                ttc[key] = np.inf
                if np.isclose(dce["dce"], 0.0):
                    ttc[key] = np.round(dce["time_dce"] * self.agent_manager.dt, 3)
            return ttc
        raise ValueError("DCE is not available in results, but is needed to evaluate TTC metric!")
