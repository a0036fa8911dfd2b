"""Host side of the dense metric core: packs phantom-agent predictions and trajectory bundles into
device tensors and calls the sm_100a kernels through the C ABI (``include/fo_b200.h``).

PyTorch is used for device memory, streams and (in ``parallel.py``) ``torch.distributed`` only; all
arithmetic of the path happens in ``libfo_b200.so``.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L

# harm coefficients read on this path; same keys/values as the reference's
# frenetix_occlusion/config/harm_params.json (log_reg.reduced_sym_angle_areas, log_reg.ignore_angle,
# pedestrian), see config/harm_coefficients.json
DEFAULT_HARM = {"rs_const": -4.457, "rs_speed": 0.177, "rs_side": 0.244, "rs_rear": -0.431,
                "ia_const": -4.591, "ia_speed": 0.185, "ped_const": 3.164, "ped_speed": 0.288}


def harm_from_reference_json(coeffs: dict) -> dict:
    """Map the reference's ``harm_params.json`` schema onto the kernel's coefficient block."""
    rs = coeffs["log_reg"]["reduced_sym_angle_areas"]
    ia = coeffs["log_reg"]["ignore_angle"]
    ped = coeffs["pedestrian"]
    return {"rs_const": rs["const"], "rs_speed": rs["speed"], "rs_side": rs["side"], "rs_rear": rs["rear"],
            "ia_const": ia["const"], "ia_speed": ia["speed"], "ped_const": ped["const"], "ped_speed": ped["speed"]}


def check_required_metrics(metric_names: Sequence[str]) -> list:
    """Dependency ordering of the activated metrics (reference metrics/metric.py:125-147)."""
    m = list(metric_names)
    if "wttc" in m:
        if "ttc" in m:
            m.remove("ttc")
        m.insert(0, "ttc")
    if "ttc" in m or "ttce" in m or "be" in m:
        if "dce" in m:
            m.remove("dce")
        m.insert(0, "dce")
    if "hr" in m:
        if "cp" in m:
            m.remove("cp")
        m.insert(0, "cp")
    return [x for x in m if x in L.M_BITS]


@dataclass
class AgentSet:
    """Phantom-agent predictions as padded SoA host arrays (what ``FOAgentManager.predictions`` holds,
    reference agent.py:420-424, 530-534, plus the owning agent's type / unbuffered shape)."""
    ids: list
    x: np.ndarray
    y: np.ndarray
    yaw: np.ndarray
    v: np.ndarray
    var_x: np.ndarray
    var_y: np.ndarray
    n_states: np.ndarray
    kind: np.ndarray
    length: np.ndarray
    width: np.ndarray
    buf_length: np.ndarray
    buf_width: np.ndarray
    agent_types: list = field(default_factory=list)

    @property
    def n_agents(self) -> int:
        return len(self.ids)

    @property
    def t_stride(self) -> int:
        return int(self.x.shape[1]) if self.x.ndim == 2 else 0

    @staticmethod
    def _alloc(A, Tp):
        f = lambda: np.zeros((A, Tp), dtype=np.float64)  # noqa: E731
        return f(), f(), f(), f(), f(), f()

    @classmethod
    def from_case(cls, agents: Sequence[dict]) -> "AgentSet":
        """From the plain-array case format used by the tests / bench (see oracle/ref_runner.py)."""
        A = len(agents)
        Tp = max([len(np.asarray(a["yaw"])) for a in agents], default=0)
        x, y, yaw, v, vx, vy = cls._alloc(A, Tp)
        ns = np.zeros(A, dtype=np.int32)
        kind = np.zeros(A, dtype=np.int32)
        dims = np.zeros((4, A), dtype=np.float64)
        types = []
        for k, a in enumerate(agents):
            n = len(np.asarray(a["yaw"]))
            ns[k] = n
            pos = np.asarray(a["pos"], dtype=np.float64).reshape(-1, 2)
            x[k, :n], y[k, :n] = pos[:, 0], pos[:, 1]
            yaw[k, :n], v[k, :n] = a["yaw"], a["v"]
            if "cov" in a:
                cov = np.asarray(a["cov"], dtype=np.float64)
                if np.any(cov[:, 0, 1] != 0) or np.any(cov[:, 1, 0] != 0):
                    raise ValueError("only diagonal prediction covariances are supported on the GPU path")
                vx[k, :n], vy[k, :n] = cov[:, 0, 0], cov[:, 1, 1]
            else:
                vx[k, :n] = vy[k, :n] = a["var"]
            kind[k] = L.KINDS[a["agent_type"].lower()]
            types.append(a["agent_type"])
            dims[:, k] = (a["length"], a["width"], a["buf_length"], a["buf_width"])
        return cls(list(range(A)), x, y, yaw, v, vx, vy, ns, kind, dims[0], dims[1], dims[2], dims[3], types)

    @classmethod
    def from_predictions(cls, predictions: dict, agent_by_prediction_id) -> "AgentSet":
        """From the reference's own structures: ``agent_manager.predictions`` and
        ``agent_manager.agent_by_prediction_id`` (agent.py:159-183)."""
        agents = []
        ids = []
        for pid, pred in predictions.items():
            ag = agent_by_prediction_id(pid)
            cov = np.asarray(pred["cov_list"], dtype=np.float64)
            agents.append({"agent_type": ag.agent_type, "length": ag.shape.length, "width": ag.shape.width,
                           "buf_length": pred["shape"]["length"], "buf_width": pred["shape"]["width"],
                           "pos": pred["pos_list"], "yaw": pred["orientation_list"], "v": pred["v_list"],
                           "cov": cov})
            ids.append(pid)
        out = cls.from_case(agents)
        out.ids = ids
        return out


@dataclass
class BundleResult:
    valid: torch.Tensor                # uint8 [N]
    summary: torch.Tensor              # float32 [N, FO_SUMMARY_K]
    flags: torch.Tensor                # int32 [N]
    pair: Optional[torch.Tensor] = None   # float32 [N, A, FO_PAIR_K]
    step: Optional[torch.Tensor] = None   # float32 [N, A, T-1, FO_STEP_K]
    peer_delta: Optional[Sequence[int]] = None   # byte offsets to the peer-mapped copies of valid / summary / flags (parallel.py)


class MetricEngine:
    """Owns the device-side agent table and launches ``fo_metric_bundle``."""

    def __init__(self, vehicle_params, dt: float, activated_metrics: Sequence[str], thresholds: dict,
                 harm_coeffs: Optional[dict] = None, device="cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("frenetix_occlusion_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        g = (lambda k: float(vehicle_params[k])) if isinstance(vehicle_params, dict) else \
            (lambda k: float(getattr(vehicle_params, k)))
        self.vehicle = L.FoVehicle(g("length"), g("width"), g("mass"), g("wb_rear_axle"), g("a_max"))
        self.dt = float(dt)
        self.order = check_required_metrics(activated_metrics)
        if "be" in self.order and "ttc" not in self.order:
            raise KeyError("ttc")      # reference: be.py:39 reads results['ttc']
        self.metric_mask = 0
        for m in self.order:
            self.metric_mask |= L.M_BITS[m]
        self.thresholds = dict(thresholds)
        self.threshold_mask = 0
        for name, bit in L.T_BITS.items():
            if self.thresholds.get(name) is not None:
                self.threshold_mask |= bit
        h = dict(DEFAULT_HARM if harm_coeffs is None else harm_coeffs)
        self.harm = L.FoHarmCoeffs(*[float(h[k]) for k in ("rs_const", "rs_speed", "rs_side", "rs_rear", "ia_const",
                                                           "ia_speed", "ped_const", "ped_speed")])
        self.origin = np.zeros(2)
        self.agents: Optional[AgentSet] = None
        self._table = None
        self._raw_dev = None
        self.n_agents = 0
        self.t_stride = 0

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_agents(self, agents: AgentSet, origin=None):
        """Upload + pack the phantom predictions (once per planning cycle).  ``origin`` (x, y) is
        subtracted in float64 from agent and ego positions before the cast to float32."""
        self.agents = agents
        self.origin = np.zeros(2) if origin is None else np.asarray(origin, dtype=np.float64)
        A, Tp = agents.n_agents, agents.t_stride
        self.n_agents, self.t_stride = A, Tp
        if A == 0 or Tp == 0:
            self._table = None
            self.n_agents = 0
            return
        if Tp > L.FO_MAX_STATES:
            raise ValueError(f"predictions longer than {L.FO_MAX_STATES} states are not supported")
        fl = np.stack([agents.x - self.origin[0], agents.y - self.origin[1], agents.yaw, agents.v,
                       agents.var_x, agents.var_y]).astype(np.float32)
        pa = np.stack([agents.length, agents.width, agents.buf_length, agents.buf_width]).astype(np.float32)
        ia = np.stack([agents.n_states, agents.kind]).astype(np.int32)
        with torch.cuda.device(self.device):
            d_fl = torch.from_numpy(fl).to(self.device, non_blocking=False)
            d_pa = torch.from_numpy(pa).to(self.device)
            d_ia = torch.from_numpy(ia).to(self.device)
            raw = L.FoAgentsRaw()
            raw.n_agents, raw.t_stride = A, Tp
            raw.x, raw.y, raw.yaw, raw.v, raw.var_x, raw.var_y = [d_fl[i].data_ptr() for i in range(6)]
            raw.n_states, raw.kind = d_ia[0].data_ptr(), d_ia[1].data_ptr()
            raw.length, raw.width, raw.buf_length, raw.buf_width = [d_pa[i].data_ptr() for i in range(4)]
            nbytes = L.lib.fo_agent_table_bytes(A, Tp)
            self._table = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            L.check(L.lib.fo_agents_pack(C.byref(raw), C.byref(self.vehicle), C.c_void_p(self._table.data_ptr()),
                                         nbytes, self._stream()), "fo_agents_pack")
            self._raw_dev = (d_fl, d_pa, d_ia)  # keep alive until the stream work is done

    # ------------------------------------------------------------------------------------------
    def to_device_bundle(self, ego) -> torch.Tensor:
        """[N, T, 5] -> contiguous float32 CUDA tensor (origin-shifted when given as host float64)."""
        if isinstance(ego, torch.Tensor) and ego.is_cuda:
            t = ego
            if np.any(self.origin != 0.0):
                # device tensors are in the caller's frame: shift them like the agents (float64 offset vector)
                off = torch.tensor([self.origin[0], self.origin[1], 0.0, 0.0, 0.0], dtype=torch.float64, device=t.device)
                t = (t.to(torch.float64) - off).to(torch.float32)
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(torch.float32).contiguous()
            return t
        arr = np.array(ego, dtype=np.float64, copy=True)
        if arr.ndim == 2:
            arr = arr[None]
        arr[..., 0] -= self.origin[0]
        arr[..., 1] -= self.origin[1]
        host = torch.from_numpy(arr.astype(np.float32))
        if host.numel() >= (1 << 18):          # page-locking only pays for bundles of a megabyte and more
            return host.pin_memory().to(self.device, non_blocking=True)
        return host.to(self.device)

    def _args(self, t, out):
        N, T = int(t.shape[0]), int(t.shape[1])
        a = L.FoMetricArgs()
        a.ego, a.n_traj, a.n_states = t.data_ptr(), N, T
        a.agent_table = self._table.data_ptr() if self._table is not None else None
        a.n_agents, a.t_stride = self.n_agents, max(self.t_stride, 1)
        a.vehicle, a.harm, a.dt = self.vehicle, self.harm, self.dt
        a.metric_mask, a.threshold_mask = self.metric_mask, self.threshold_mask
        for name in L.T_BITS:
            v = self.thresholds.get(name)
            setattr(a, "thr_" + name, float(v) if v is not None else 0.0)
        a.valid, a.summary, a.flags = out.valid.data_ptr(), out.summary.data_ptr(), out.flags.data_ptr()
        deltas = out.peer_delta                        # parallel.PeerResultGatherer: fused result exchange
        if deltas:
            a.n_peers = len(deltas)
            for i, d in enumerate(deltas):
                a.peer_delta[i] = int(d)
        return a

    def work_stats(self, ego) -> dict:
        """How much of the algorithmic work survives the exact bounds (``fo_metric_stats``): counts of visited
        (trajectory, agent, step) evaluations, exact box distances, LR4S logits, CP evaluations, BE bisections and
        probes for one pass over ``ego``.  Used by bench.py's flop model; synchronises."""
        t = self.to_device_bundle(ego)
        N = int(t.shape[0])
        dev = self.device
        with torch.cuda.device(dev):
            out = BundleResult(torch.empty(N, dtype=torch.uint8, device=dev),
                               torch.empty((N, L.FO_SUMMARY_K), dtype=torch.float32, device=dev),
                               torch.empty(N, dtype=torch.int32, device=dev))
            a = self._args(t, out)
            cnt = torch.zeros(8, dtype=torch.int64, device=dev)
            L.check(L.lib.fo_metric_stats(C.byref(a), C.c_void_p(cnt.data_ptr()), self._stream()), "fo_metric_stats")
            c = cnt.cpu().tolist()
        return {"visited": c[0], "obb": c[1], "lr4s": c[2], "cp": c[3], "be": c[4], "be_probes": c[5],
                "windows": c[6], "windows_kept": c[7]}

    def assess(self, ego, want_pair: bool = False, want_step: bool = False, out: Optional[BundleResult] = None
               ) -> BundleResult:
        """One pass of the dense core over a trajectory bundle (asynchronous on the current stream)."""
        t = self.to_device_bundle(ego)
        if t.ndim != 3 or t.shape[2] != 5:
            raise ValueError("trajectory bundle must have shape [N, T, 5] (x, y, theta, v, a)")
        N, T = int(t.shape[0]), int(t.shape[1])
        if T > L.FO_MAX_STATES:
            raise ValueError(f"trajectories longer than {L.FO_MAX_STATES} states are not supported")
        A = self.n_agents
        dev = self.device
        with torch.cuda.device(dev):
            if out is None:
                out = BundleResult(torch.empty(N, dtype=torch.uint8, device=dev),
                                   torch.empty((N, L.FO_SUMMARY_K), dtype=torch.float32, device=dev),
                                   torch.empty(N, dtype=torch.int32, device=dev))
                if want_pair:
                    out.pair = torch.empty((N, A, L.FO_PAIR_K), dtype=torch.float32, device=dev)
                if want_step:
                    out.step = torch.empty((N, A, max(T - 1, 0), L.FO_STEP_K), dtype=torch.float32, device=dev)
            a = self._args(t, out)
            a.pair = out.pair.data_ptr() if (out.pair is not None and A > 0) else None
            a.step = out.step.data_ptr() if (out.step is not None and A > 0 and T > 1) else None
            L.check(L.lib.fo_metric_bundle(C.byref(a), self._stream()), "fo_metric_bundle")
        out._keepalive = t
        return out

    def capture(self, ego_dev: torch.Tensor, out: Optional[BundleResult] = None):
        """CUDA-graph form of ``assess`` for the per-planning-step latency path: the launch on a device-resident
        bundle (fixed address and shape; refill it in place before every replay) is captured once, ``replay()``
        re-issues it without any per-call host work.  Returns ``(graph, out)``."""
        if not (isinstance(ego_dev, torch.Tensor) and ego_dev.is_cuda and ego_dev.dtype == torch.float32
                and ego_dev.is_contiguous()):
            raise ValueError("capture() needs a contiguous float32 CUDA tensor [N, T, 5]")
        N = int(ego_dev.shape[0])
        dev = self.device
        with torch.cuda.device(dev):
            if out is None:
                out = BundleResult(torch.empty(N, dtype=torch.uint8, device=dev),
                                   torch.empty((N, L.FO_SUMMARY_K), dtype=torch.float32, device=dev),
                                   torch.empty(N, dtype=torch.int32, device=dev))
            self.assess(ego_dev, out=out)                  # warm-up outside the capture (first-use attribute calls)
            torch.cuda.current_stream(dev).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.assess(ego_dev, out=out)
        return g, out
