"""Brake-evaluation metric (reference frenetix_occlusion/metrics/be.py:31-193): minimum constant
deceleration (bisection, <= 10 probes) that avoids the predicted collision, and the brake threat
number.  The bisection runs warp-cooperatively inside the dense kernel."""
from .. import _lib as L
from .core import shared_core


class BE:
    def __init__(self, vehicle_params, agent_manager, core=None):
        self.vehicle_params = vehicle_params
        self.agent_manager = agent_manager
        self._core = core

    def __repr__(self):
        return "<'Break Evaluation Metric': {}.{} object at {}>".format(
            self.__class__.__module__, self.__class__.__name__, hex(id(self)))

    def evaluate(self, trajectory, results) -> dict:
        if "dce" in results:
            results["ttc"]  # KeyError when 'ttc' was not evaluated, as be.py:39
            core = self._core or shared_core(self.vehicle_params, self.agent_manager)
            d = core.detail(trajectory)
            if d["flags"] & L.F_BE_RANGE:
                # scipy.interpolate.interp1d(bounds_error=True) in the reference, be.py:117-124
                raise ValueError("A value in x_new is above the interpolation range: the re-timed braking "
                                 "trajectory overruns the original path")
            return {pid: {"required_constant_deceleration": float(d["pair"][k, 9]),
                          "break_threat_number": float(d["pair"][k, 10])} for k, pid in enumerate(d["ids"])}
        raise ValueError("DCE is not available in results, but is needed to evaluate BE metric!")
