"""Harm-and-risk metric (reference frenetix_occlusion/metrics/hr.py:43-116, utils/harm_model.py,
utils/logistic_regression.py)."""
import json
import os

import numpy as np

from .core import shared_core


class HR:
    def __init__(self, vehicle_params, agent_manager, core=None):
        self.risk_params = self._load_param("risk_params")
        self.harm_params = self._load_param("harm_coefficients")
        self.vehicle_params = vehicle_params
        self.agent_manager = agent_manager
        self._core = core
        if self.risk_params["harm_mode"] != "log_reg":   # harm_model.py:125,152-153
            raise ValueError("Please select a valid mode for harm estimation (log_reg)")

    def __repr__(self):
        return "<'Risk and Harm Metric': {}.{} object at {}>".format(
            self.__class__.__module__, self.__class__.__name__, hex(id(self)))

    @staticmethod
    def _load_param(name):
        path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "config", name + ".json")
        with open(path, "r") as f:
            return json.load(f)

    def evaluate(self, trajectory, result) -> dict:
        cp = result["cp"]                                   # KeyError without 'cp', as hr.py:61
        core = self._core or shared_core(self.vehicle_params, self.agent_manager)
        d = core.detail(trajectory)
        T = d["T"]
        out = {}
        for k, pid in enumerate(d["ids"]):
            P = min(T - 1, int(d["n_states"][k]))           # harm_model.py:67-70: pair skipped when 0
            if P == 0:
                continue
            eh = np.array(d["step"][k, :P, 1])
            oh = np.array(d["step"][k, :P, 2])
            c = np.asarray(cp[pid], dtype=np.float64)
            er = list(eh * c[:P])                           # hr.py:78-79 (lists of numpy floats, as the reference builds)
            orr = list(oh * c[:P])
            p = d["pair"][k]
            out[pid] = {"max_ego_risk": p[2], "max_obst_risk": p[3], "max_obst_harm_with_cp": p[5],
                        "max_obst_risk_index": int(p[4]), "max_ego_harm": p[6], "max_obst_harm": p[7],
                        "ego_risk_traj": er, "obst_risk_traj": orr, "ego_harm_traj": eh, "obst_harm_traj": oh,
                        "collision_probability": c, "max_collision_probability": p[8]}
        s = d["summary"]
        out["max_ego_risk_all"] = s[0]
        out["max_obst_risk_all"] = s[1]
        out["max_ego_harm_all"] = s[2]
        out["max_obst_harm_all"] = s[3]
        out["max_collision_probability_all"] = s[4]
        out["max_obst_harm_with_cp_all"] = s[5]
        return out
