"""World-size-2 gloo test (CPU) of the N>1 host logic: contiguous trajectory shards + the single
all-gather of the result vectors reassemble the global result exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_results(idx, k):
    idx = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
    valid = (idx % 3 == 0).to(torch.uint8)
    summary = idx.to(torch.float32)[:, None] * 0.5 + torch.arange(k, dtype=torch.float32)[None]
    flags = ((idx % 7 == 0).to(torch.int32)) * 0x01000001          # beyond 2**24: must survive the gather bit for bit
    return valid, summary, flags


def _worker(rank, world, n_total, block, port, q):
    sys.path.insert(0, ROOT)
    from frenetix_occlusion_b200.parallel import ResultGatherer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = ResultGatherer(n_total, 10, torch.device("cpu"), block=block)
        # the "kernel" writes straight into this rank's slice of the gather buffer
        out = g.local_result()
        v, s, f = _fake_results(g.index[rank], 10)
        out.valid.copy_(v), out.summary.copy_(s), out.flags.copy_(f)
        gv, gs, gf = g.gather().assembled()
        ev, es, ef = _fake_results(np.arange(n_total), 10)
        ok = bool(torch.equal(gv, ev) and torch.equal(gs, es) and torch.equal(gf, ef))
        q.put((rank, ok, sorted(g.index[rank].tolist())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total,block", [(10, 3), (11, None), (1, 4), (4099, 512)])
def test_two_rank_gather(n_total, block):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sock:                       # a free rendezvous port
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, n_total, block, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    owned = sorted(i for _, _, idx in res for i in idx)
    assert owned == list(range(n_total))                       # every trajectory on exactly one rank


def test_interleaved_blocks_balance_a_clustered_bundle():
    """Blocks of the bundle go round-robin to the ranks: a bundle whose expensive trajectories are clustered (here the
    second half costs 4x) still gives every rank the same cost within one block."""
    from frenetix_occlusion_b200.parallel import shard_blocks, shard_indices
    n, world, block = 1_000_000, 8, 1024
    cost = np.where(np.arange(n) < n // 2, 1.0, 4.0)
    per_rank = [cost[shard_indices(n, world, r, block)].sum() for r in range(world)]
    assert max(per_rank) / min(per_rank) < 1.02          # granularity: one block of ~122 per rank
    assert shard_blocks(10, 2, 1, 3) == [(3, 6), (9, 10)]
    assert sum(len(shard_indices(n, world, r, block)) for r in range(world)) == n


def test_shard_bounds_cover_everything():
    from frenetix_occlusion_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1000, 1_000_000):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert all(0 <= hi - lo <= (n + w - 1) // w for lo, hi in spans)


def test_bundle_packer_round_trip():
    import types
    import numpy as np
    from frenetix_occlusion_b200.adapter import BundlePacker, bundle_from_trajectories
    rng = np.random.default_rng(0)
    trajs = []
    for _ in range(7):
        c = types.SimpleNamespace(**{f: rng.normal(size=31) for f in ("x", "y", "theta", "v", "a")})
        trajs.append(types.SimpleNamespace(cartesian=c))
    ref = bundle_from_trajectories(trajs)
    assert ref.shape == (7, 31, 5) and np.array_equal(ref[3, :, 2], trajs[3].cartesian.theta)
    pk = BundlePacker(16, 31, pin=False)
    got = pk.pack(trajs, origin=(1.5, -2.0)).numpy()
    assert got.shape == (7, 31, 5) and got.dtype == np.float32
    exp = ref.copy()
    exp[..., 0] -= 1.5
    exp[..., 1] += 2.0
    assert np.array_equal(got, exp.astype(np.float32))
    cols = pk.pack_columns(ref[..., 0], ref[..., 1], ref[..., 2], ref[..., 3], ref[..., 4]).numpy()
    assert np.array_equal(cols, ref.astype(np.float32))
    import pytest
    with pytest.raises(ValueError):
        pk.pack(trajs * 3)
