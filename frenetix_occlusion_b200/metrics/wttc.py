"""Worst-time-to-collision metric (reference frenetix_occlusion/metrics/wttc.py:28-44)."""
import numpy as np


class WTTC:
    def __init__(self):
        pass

    def __repr__(self):
        return "<'Worst Time to Collision  Metric': {}.{} object at {}>".format(
            self.__class__.__module__, self.__class__.__name__, hex(id(self)))

    @staticmethod
    def evaluate(trajectory, results) -> float:
        if "ttc" in results:
            wttc = np.inf
            for v in results["ttc"].values():
                wttc = min(wttc, v)
            return wttc
        raise ValueError("TTC is not available in results, but is needed to evaluate WTTC metric!")
