"""Fused result exchange of the trajectory-sharded sweep (optional form): the summary kernel's epilogue
stores valid / summary / flags into every rank's peer-mapped gather buffer (``fo_peer_alloc`` / ``fo_peer_open`` /
``FoMetricArgs.peer_delta``); no all-gather follows.  Two ranks share the one GPU of the test box (CUDA IPC maps a
buffer of another PROCESS, same or different device), the handshake runs over gloo; on the 8-GPU box the same code
crosses NVLink (``bench.py --gpus N``)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    try:
        from frenetix_occlusion_b200 import _lib as L
        from frenetix_occlusion_b200 import synthetic as S
        from frenetix_occlusion_b200.engine import AgentSet, MetricEngine
        from frenetix_occlusion_b200.parallel import PeerResultGatherer
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        n_total, block = 3000, 256
        g = PeerResultGatherer(n_total, L.FO_SUMMARY_K, dev, block=block)
        report = []
        for step in range(3):                       # three steps: both buffers, and the first one reused
            case = S.make_case(n_total, 24, 31, seed=100 + step)
            eng = MetricEngine(case["vehicle"], case["dt"], case["activated_metrics"], case["thresholds"], device=dev)
            eng.set_agents(AgentSet.from_case(case["agents"]))
            ego = torch.from_numpy(case["ego"].astype(np.float32)).to(dev)
            ref = eng.assess(ego)                   # the whole bundle on this rank: what the exchange must reproduce
            mine = torch.from_numpy(g.index[rank]).to(dev)
            out = g.local_result()
            assert out.peer_delta and len(out.peer_delta) == world - 1
            half = len(mine) // 2                   # two launches into slices of the same buffer (bench.py's e2e chunks)
            eng.assess(ego[mine[:half]].contiguous(), out=g.slice_of(out, 0, half))
            eng.assess(ego[mine[half:]].contiguous(), out=g.slice_of(out, half, len(mine)))
            g.gather()
            v, s, f = g.assembled()
            torch.cuda.synchronize()
            report.append((bool(torch.equal(v, ref.valid)), bool(torch.equal(s.view(torch.int32), ref.summary.view(torch.int32))),
                           bool(torch.equal(f, ref.flags)), int(ref.valid.sum())))
        dist.barrier()
        g.close()
        q.put((rank, report, None))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc() + repr(e)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.gpu
def test_kernel_epilogue_fills_every_ranks_gather_buffer(cuda_device):
    import socket
    with socket.socket() as sock:                       # a free rendezvous port
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, report, err in res:
        assert err is None, err
        assert all(a and b and c for a, b, c, _ in report), (rank, report)
        assert all(0 < nv < 3000 for *_, nv in report), report      # non-trivial masks


def test_peer_exports_are_declared():
    sys.path.insert(0, ROOT)
    from frenetix_occlusion_b200 import _lib as L
    for name in ("fo_peer_alloc", "fo_peer_open", "fo_peer_close", "fo_peer_free"):
        assert name in L.EXPORTED_SYMBOLS and hasattr(L.lib, name)
    assert L.FoMetricArgs.peer_delta.size == 8 * L.FO_MAX_PEERS
