"""Trajectory sharding across GPUs: the bundle is split by trajectory over the ranks, the (tiny) agent
table is replicated, and ONE all-gather moves the per-trajectory result vectors.

One process per GPU; ``torch.distributed`` (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests) is
plumbing only -- there is no other exchange step on this path.

* Shards are INTERLEAVED blocks (block b of ``block`` trajectories belongs to rank ``b mod world``): the cost of a
  trajectory varies several-fold with the number of near agents and braking pairs, and contiguous shards of a sorted
  or clustered bundle made the slowest rank 4 % slower than the mean (round 1).  ``shard_bounds`` (contiguous) stays
  for callers that need one span per rank.
* The gather is in place and in the results' own dtypes: every rank's slice of ONE byte buffer holds
  ``[valid u8 | flags i32 | summary f32]`` of its shard, the kernel writes its outputs straight into the views of this
  rank's slice (``local_result()``), and a single ``all_gather_into_tensor`` fills the other slices.  No pack / unpack
  kernels, no float round trip of the flags.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_size(n_total: int, world: int) -> int:
    return (n_total + world - 1) // world


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the trajectory axis owned by ``rank`` (last shards may be short/empty)."""
    per = shard_size(n_total, world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def shard_blocks(n_total: int, world: int, rank: int, block: int = 1024) -> List[Tuple[int, int]]:
    """Global [lo, hi) ranges of the interleaved blocks owned by ``rank``, in increasing order."""
    out = []
    b = rank
    while b * block < n_total:
        out.append((b * block, min(n_total, (b + 1) * block)))
        b += world
    return out


def shard_indices(n_total: int, world: int, rank: int, block: int = 1024) -> np.ndarray:
    """Global trajectory indices of ``rank``'s interleaved shard (local order)."""
    parts = [np.arange(lo, hi) for lo, hi in shard_blocks(n_total, world, rank, block)]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


class ResultGatherer:
    """Pre-allocated buffer for the single collective of the path.

    ``capacity`` = the largest shard over all ranks (every rank's slice has the same size, as the collective needs);
    ``counts[r]`` = trajectories rank r really owns."""

    def __init__(self, n_total: int, summary_k: int, device, group=None, block: Optional[int] = 1024):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total, self.k, self.block = n_total, summary_k, block
        if block is None:
            spans = [shard_bounds(n_total, self.world, r) for r in range(self.world)]
            self.index = [np.arange(lo, hi) for lo, hi in spans]
        else:
            self.index = [shard_indices(n_total, self.world, r, block) for r in range(self.world)]
        self.counts = [len(i) for i in self.index]
        self.capacity = max(self.counts + [1])
        cap = self.capacity
        self.off_flags = _align(cap)
        self.off_summary = self.off_flags + _align(4 * cap)
        self.slice_bytes = self.off_summary + _align(4 * summary_k * cap)
        self.full = self._alloc(self.slice_bytes * self.world, device)

    def _alloc(self, nbytes: int, device):
        return torch.zeros(nbytes, dtype=torch.uint8, device=device)

    # ---- views ------------------------------------------------------------------------------------------------
    def _views(self, r: int):
        base = self.full[r * self.slice_bytes:(r + 1) * self.slice_bytes]
        n = self.counts[r]
        valid = base[:self.capacity][:n]
        flags = base[self.off_flags:self.off_flags + 4 * self.capacity].view(torch.int32)[:n]
        summary = base[self.off_summary:self.off_summary + 4 * self.k * self.capacity].view(torch.float32) \
            .view(self.capacity, self.k)[:n]
        return valid, summary, flags

    def local_result(self):
        """``engine.BundleResult`` whose tensors are this rank's slice of the gather buffer: pass it as ``out=`` to
        ``MetricEngine.assess`` and the kernel's stores ARE the collective's send data."""
        from .engine import BundleResult
        v, s, f = self._views(self.rank)
        return BundleResult(v, s, f)

    def current_result(self):
        """The views the last step was evaluated into (never moves on to another buffer)."""
        return self.local_result()

    def slice_of(self, result, lo: int, hi: int):
        """``result[lo:hi]`` as a ``BundleResult`` of its own (chunked evaluation into one gather buffer)."""
        from .engine import BundleResult
        return BundleResult(result.valid[lo:hi], result.summary[lo:hi], result.flags[lo:hi], peer_delta=result.peer_delta)

    def rank_result(self, r: int):
        """(valid, summary, flags) views of rank ``r``'s shard after ``gather()`` (local order of ``index[r]``)."""
        return self._views(r)

    # ---- the collective -----------------------------------------------------------------------------------------
    def gather(self, valid=None, summary=None, flags=None):
        """One all-gather of the result vectors.  With no arguments the local slice is taken as already written (the
        in-place form); tensors, when given, are copied into it first (callers that evaluated elsewhere)."""
        if valid is not None:
            v, s, f = self._views(self.rank)
            n = self.counts[self.rank]
            v.copy_(valid[:n])
            s.copy_(summary[:n])
            if flags is not None:
                f.copy_(flags[:n])
        if self.world > 1:
            mine = self.full[self.rank * self.slice_bytes:(self.rank + 1) * self.slice_bytes]
            if self.full.is_cuda:
                dist.all_gather_into_tensor(self.full, mine, group=self.group)       # in place: input is a slice of output
            else:   # gloo: list form (the CPU tests)
                parts = list(self.full.view(self.world, self.slice_bytes).unbind(0))
                dist.all_gather(parts, mine.clone(), group=self.group)
        return self

    def assembled(self):
        """Global-order (valid[N] u8, summary[N, K] f32, flags[N] i32): scatters the per-rank blocks back (a copy; the
        timed step does not need it)."""
        dev = self.full.device
        valid = torch.empty(self.n_total, dtype=torch.uint8, device=dev)
        flags = torch.empty(self.n_total, dtype=torch.int32, device=dev)
        summary = torch.empty((self.n_total, self.k), dtype=torch.float32, device=dev)
        for r in range(self.world):
            if self.counts[r] == 0:
                continue
            idx = torch.from_numpy(self.index[r]).to(dev)
            v, s, f = self._views(r)
            valid[idx], summary[idx], flags[idx] = v, s, f
        return valid, summary, flags


class _DeviceBytes:
    """Raw device memory as a ``__cuda_array_interface__`` object (zero-copy ``torch.as_tensor``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerResultGatherer(ResultGatherer):
    """The exchange step FUSED into the metric kernel (optional form): no all-gather at all.

    Every rank's gather buffer is CUDA-IPC device memory (``fo_peer_alloc``) mapped into every other rank's process
    (``fo_peer_open``); ``local_result()`` carries the address differences as ``peer_delta``, and the summary kernel's
    epilogue stores valid / summary / flags of each trajectory into this rank's slice of ALL buffers -- the remote
    copies travel over NVLink / NVSwitch while the rest of the bundle is still being evaluated.  What is left of the
    collective is a completion handshake (``gather()``: one 4-byte all-reduce, stream-ordered behind the kernel).

    Two buffers alternate between steps: rank A's kernel of step k + 1 may start while rank B still reads the results
    of step k (its D2H copy, a consumer kernel) -- it writes the other buffer; by the time step k + 2 reuses the first
    one, every rank has passed the handshake of step k + 1, which its own stream ordered behind its reads of step k."""

    def __init__(self, n_total: int, summary_k: int, device, group=None, block: Optional[int] = 1024):
        from . import _lib as L
        import ctypes as C
        self._L, self._C = L, C
        self._own, self._mapped = [], []
        self._phase, self._gathered = 0, False
        self._device = torch.device(device)
        super().__init__(n_total, summary_k, device, group=group, block=block)   # allocates buffer 0 (self.full)
        self._bufs = [self.full, self._alloc(self.slice_bytes * self.world, device)]
        # exchange the handles, map the peers' buffers, keep the address differences
        mine = [bytes(h.bytes) for _, h in self._own]
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self._delta = [[], []]
        for r, handles in enumerate(everyone):
            if r == self.rank:
                continue
            for b, raw in enumerate(handles):
                h = L.FoPeerHandle()
                C.memmove(h.bytes, raw, 64)
                ptr = C.c_void_p()
                L.check(L.lib.fo_peer_open(C.byref(h), C.byref(ptr)), "fo_peer_open")
                self._mapped.append(ptr.value)
                self._delta[b].append(ptr.value - self._own[b][0])
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._nccl = dist.get_backend(self.group) == "nccl"
        self._results = [None, None]

    def _alloc(self, nbytes: int, device):
        L, C = self._L, self._C
        ptr, h = C.c_void_p(), L.FoPeerHandle()
        with torch.cuda.device(self._device):
            L.check(L.lib.fo_peer_alloc(nbytes, C.byref(ptr), C.byref(h)), "fo_peer_alloc")
            t = torch.as_tensor(_DeviceBytes(ptr.value, nbytes), device=self._device)
        self._own.append((ptr.value, h))
        return t

    def local_result(self):
        """This step's output views (+ ``peer_delta``).  The first call after a ``gather()`` moves on to the other buffer."""
        if self._gathered:
            self._phase ^= 1
            self._gathered = False
        self.full = self._bufs[self._phase]
        if self._results[self._phase] is None:
            r = super().local_result()
            r.peer_delta = list(self._delta[self._phase])
            self._results[self._phase] = r
        return self._results[self._phase]

    def current_result(self):
        self.full = self._bufs[self._phase]
        return self._results[self._phase] if self._results[self._phase] is not None else self.local_result()

    def gather(self, valid=None, summary=None, flags=None):
        """Completion handshake: when it has passed on this rank's stream, every rank's kernel of the step has finished
        and with it the stores into this rank's buffer."""
        if valid is not None:
            raise ValueError("the fused exchange has no copy-in form: evaluate into local_result()")
        if self.world > 1:
            if self._nccl:
                dist.all_reduce(self._flag, group=self.group)
            else:       # host-level handshake (gloo in the tests)
                torch.cuda.synchronize(self._device)
                dist.barrier(group=self.group)
        self._gathered = True
        return self

    def close(self):
        L = self._L
        torch.cuda.synchronize(self._device)
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)       # nobody still writes into a buffer that is about to go away
        for p in self._mapped:
            L.lib.fo_peer_close(p)
        self._mapped = []
        self._results, self._bufs, self.full = [None, None], [], None
        for p, _ in self._own:
            L.lib.fo_peer_free(p)
        self._own = []
