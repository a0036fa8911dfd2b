"""Stage 1 / stage 2: oracle self-consistency on CPU, CUDA parity on the GPU."""
import numpy as np
import pytest

from frenetix_occlusion_b200 import synthetic as S
from oracle import visibility_oracle as VO


def _frame(seed, n_obst, with_ring=False, transparent_every=0, half_extent=40.0):
    rng = np.random.default_rng(seed)
    rect = S.obstacle_frames(1, n_obst, seed=seed, half_extent=half_extent)[0].astype(np.float64)
    # keep the ego outside every rectangle
    keep = np.hypot(rect[:, 0], rect[:, 1]) > 3.5
    flags = np.where(keep, VO.RECT_EXISTS, 0).astype(np.uint8)
    if transparent_every:
        flags[::transparent_every] |= VO.RECT_TRANSPARENT
    ego = np.array([0.0, 0.0, rng.uniform(-np.pi, np.pi)])
    boundary = None
    if with_ring:
        ang = np.linspace(0, 2 * np.pi, 101)
        rad = 35.0 + 6.0 * np.sin(3 * ang)
        pts = np.stack((rad * np.cos(ang), rad * np.sin(ang)), -1)
        boundary = np.concatenate((pts[:-1], pts[1:]), 1)
    return ego, rect, flags, boundary


@pytest.mark.parametrize("seed,n_obst,ring,fov", [(1, 12, False, 360.0), (2, 30, True, 360.0), (3, 20, False, 120.0),
                                                  (4, 8, True, 200.0)])
def test_raycast_oracle_agrees_with_reference_shadow_construction(seed, n_obst, ring, fov):
    """Polar first-hit map == point-wise evaluation of the reference's shadow-quad construction,
    outside a band of one angular step around shadow borders."""
    ego, rect, flags, boundary = _frame(seed, n_obst, ring)
    R, n_rays = 50.0, 4096
    rng_, hit, _ = VO.raycast(ego, rect, flags, boundary, R, fov, n_rays)
    ang = VO.ray_angles(ego[2], fov, n_rays)
    prng = np.random.default_rng(seed + 100)
    pr = np.sqrt(prng.uniform(0.04, 1.0, 6000)) * R * 0.995
    k = prng.integers(1, n_rays - 2, 6000)
    frac = prng.uniform(0.2, 0.8, 6000)
    pa = ang[k] + frac * (ang[1] - ang[0])
    P = ego[:2] + np.stack((pr * np.cos(pa), pr * np.sin(pa)), -1)
    ref = VO.reference_point_visible(P, ego, rect, flags, boundary, R, fov)
    lo = np.minimum(rng_[k], rng_[k + 1])
    hi = np.maximum(rng_[k], rng_[k + 1])
    tol = 0.02
    must_vis = pr < lo - tol
    must_occ = pr > hi + tol
    # the reference removes the obstacle polygon itself as well; points inside an obstacle are beyond the hit
    assert ref[must_vis].all(), f"{(~ref[must_vis]).sum()} points the ray map calls visible are shadowed in the reference"
    assert (~ref[must_occ]).all(), f"{ref[must_occ].sum()} points the ray map calls occluded are visible in the reference"
    assert must_vis.sum() > 500 and must_occ.sum() > 500


def test_rollout_oracle_kat():
    r = VO.rollout_cv([12.0], [-3.0], [1.4], [np.pi / 2], 0.1, 3.0)
    assert r["x"].shape == (1, 31)
    assert np.allclose(r["x"][0], 12.0) and np.isclose(r["y"][0, 30], -3.0 + 3.0 * 1.4)
    assert np.isclose(r["var"][0, 30], 0.1 * 1.05 ** 30)
    r = VO.rollout_cv([0.0], [0.0], [1.4], [0.3], 0.1, 3.0)
    assert np.isclose(r["x"][0, 10], round(1.4 * np.cos(0.3), 3))      # round(1.4 cos phi, 3)


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_obst,ring,fov,n_rays,transp", [(11, 64, False, 360.0, 4096, 0), (12, 512, False, 360.0, 4096, 0),
                                                                  (13, 40, True, 360.0, 1000, 5), (14, 25, True, 90.0, 333, 3),
                                                                  (15, 300, False, 200.0, 2048, 7), (16, 0, True, 360.0, 512, 0),
                                                                  (17, 2000, False, 360.0, 256, 0)])
def test_cuda_raycast_matches_oracle(seed, n_obst, ring, fov, n_rays, transp, cuda_device):
    from frenetix_occlusion_b200.visibility import raycast_frames
    import torch
    frames = []
    for f in range(3):
        frames.append(_frame(seed * 10 + f, n_obst, ring, transp, half_extent=50.0))
    ego = np.stack([fr[0] for fr in frames])
    ego[:, :2] += np.array([[3.0, -2.0], [0.0, 0.0], [-7.5, 4.25]])      # non-trivial ego positions
    rect = np.stack([fr[1] for fr in frames]).reshape(3, n_obst, 5)
    rect[..., :2] += ego[:, None, :2]
    flags = np.stack([fr[2] for fr in frames]).reshape(3, n_obst)
    boundary = frames[0][3]
    R = 50.0
    # inputs rounded to float32 so both sides see identical numbers
    ego, rect = ego.astype(np.float32).astype(np.float64), rect.astype(np.float32).astype(np.float64)
    if boundary is not None:
        boundary = boundary.astype(np.float32).astype(np.float64)
    res = raycast_frames(ego, rect, flags, boundary, R, fov, n_rays)
    torch.cuda.synchronize()
    g_rng, g_hit, g_vis = res.range.cpu().numpy(), res.hit.cpu().numpy(), res.visible.cpu().numpy()
    n_tie = 0
    for f in range(3):
        o_rng, o_hit, o_vis = VO.raycast(ego[f], rect[f], flags[f], boundary, R, fov, n_rays)
        bad = ~np.isclose(g_rng[f], o_rng, rtol=2e-5, atol=2e-4)
        # grazing ties: a ray that clips a corner within float32 resolution may or may not register it
        n_tie += int(bad.sum())
        same = g_hit[f] == o_hit
        n_tie += int((~same & ~bad).sum())
        assert bad.mean() < 0.004 and (~same).mean() < 0.006, (bad.sum(), (~same).sum())
        vis_bad = g_vis[f] != o_vis
        assert vis_bad.sum() <= max(1, 0.01 * max(n_obst, 1)), vis_bad.sum()
    assert n_tie <= 0.006 * 3 * n_rays + 2


@pytest.mark.gpu
def test_cuda_rollout_cv_matches_oracle(cuda_device):
    from frenetix_occlusion_b200.prediction import rollout_cv
    import torch
    rng = np.random.default_rng(3)
    A = 37
    x0, y0 = rng.uniform(-50, 50, A), rng.uniform(-50, 50, A)
    v, phi = rng.uniform(0.5, 12, A), rng.uniform(-np.pi, np.pi, A)
    for horizon in (3.0, 5.0):
        r = rollout_cv(x0, y0, v, phi, 0.1, horizon)
        torch.cuda.synchronize()
        o = VO.rollout_cv(x0, y0, v, phi, 0.1, horizon)
        for key in ("x", "y", "yaw", "v", "var"):
            g = r[key].cpu().numpy()
            assert g.shape == o[key].shape
            assert np.array_equal(g, o[key].astype(np.float32)), key     # correctly rounded image of the float64 values


def _arc_path(radius, n=200, length=120.0):
    s = np.linspace(0, length, n)
    return np.stack((radius * np.sin(s / radius), radius * (1 - np.cos(s / radius))), -1)


def test_path_rollout_oracle_invariants():
    """Invariants that hold regardless of the un-vendored frenetix arithmetic: on a straight path with zero
    lateral offset the constant-speed, d1 = 0 sample has zero speed variance and reproduces CV exactly."""
    path = np.stack((np.linspace(-10, 100, 56), np.zeros(56)), -1)
    r = VO.rollout_path(path, 5.0, 0.0, 10.0, 0.1, 3.0)
    assert r["sample"] == 4                                   # speed factor 1.0, d1 = 0
    assert np.allclose(r["x"], 5.0 + 10.0 * np.arange(31) * 0.1, atol=1e-9) and np.allclose(r["y"], 0.0, atol=1e-12)
    assert np.allclose(r["v"], 10.0) and np.allclose(r["yaw"], 0.0)
    # lateral offset 0.4 m: nearest target offset is +0.5, speed stays ~constant, ends at d = 0.5
    r = VO.rollout_path(path, 5.0, 0.4, 10.0, 0.1, 3.0)
    assert r["sample"] == 5 and abs(r["y"][-1] - 0.5) < 1e-9 and abs(r["d0"] - 0.4) < 1e-12
    # circular arc: radius preserved for d = 0
    arc = _arc_path(40.0)
    r = VO.rollout_path(arc, arc[10, 0], arc[10, 1], 8.0, 0.1, 3.0)
    rad = np.hypot(r["x"], r["y"] - 40.0)
    assert np.all(np.abs(rad - 40.0) < 0.01)
    # low-velocity mode (v0 < 0.5 m/s, frenetix_handler.py:91-95): the lateral offset is a function of the covered arc
    # length -- an agent that stands still (v0 = 0: every sample has end speed 0) does not move sideways either
    r = VO.rollout_path(path, 5.0, 0.4, 0.0, 0.1, 3.0)
    assert np.allclose(r["x"], 5.0) and np.allclose(r["y"], 0.4) and np.allclose(r["v"], 0.0)
    r = VO.rollout_path(path, 5.0, 0.4, 0.3, 0.1, 3.0)
    assert abs(r["y"][0] - 0.4) < 1e-12 and np.all(np.diff(r["x"]) > 0) and np.all(np.abs(r["y"] - 0.4) <= 0.9 + 1e-9)


@pytest.mark.gpu
def test_cuda_path_rollout_matches_oracle(cuda_device):
    from frenetix_occlusion_b200.prediction import rollout_path
    import torch
    rng = np.random.default_rng(8)
    paths = [np.stack((np.linspace(-10, 100, 56), np.zeros(56)), -1), _arc_path(40.0), _arc_path(-25.0, 300, 150.0),
             np.stack((np.linspace(0, 60, 31), 3.0 * np.sin(np.linspace(0, 60, 31) / 9.0)), -1)]
    paths.append(paths[0])                                  # low-velocity mode: v0 < 0.5 m/s
    x0 = [5.0, 9.0, 3.0, 12.0, 7.0]
    y0 = [0.4, 1.5, -0.3, 2.2, 0.3]
    v0 = [10.0, 8.0, 5.0, 10.0, 0.3]
    for horizon in (3.0, 5.0):
        g = rollout_path(paths, x0, y0, v0, 0.1, horizon)
        torch.cuda.synchronize()
        for j, p in enumerate(paths):
            o = VO.rollout_path(p, x0[j], y0[j], v0[j], 0.1, horizon)
            assert int(g["sample"][j]) == o["sample"], (j, int(g["sample"][j]), o["sample"])
            for key in ("x", "y", "yaw", "v", "var"):
                np.testing.assert_allclose(g[key][j].cpu().numpy(), o[key], rtol=2e-6, atol=2e-5, err_msg=f"{key} job {j}")


@pytest.mark.gpu
def test_full_size_visibility_sweep_properties(cuda_device):
    """BASELINE.json configs[4] at FULL size (10,000 frames x 4096 rays x 512 rectangles): frame-sharding and
    obstacle-order invariance of the ray cast, and two frames against the float64 oracle."""
    import torch
    from frenetix_occlusion_b200.visibility import raycast_frames
    F, R, O = S.C_VIS["n_frames"], S.C_VIS["n_rays"], S.C_VIS["n_obstacles"]
    rect = torch.from_numpy(S.obstacle_frames(F, O)).cuda()
    flags = torch.ones((F, O), dtype=torch.uint8, device="cuda")
    ego = torch.zeros((F, 3), dtype=torch.float32, device="cuda")
    ego[:, 2] = torch.linspace(-3.0, 3.0, F, device="cuda")
    whole = raycast_frames(ego, rect, flags, None, 50.0, 360.0, R)
    torch.cuda.synchronize()
    h = F // 2
    a = raycast_frames(ego[:h], rect[:h], flags[:h], None, 50.0, 360.0, R)
    b = raycast_frames(ego[h:], rect[h:], flags[h:], None, 50.0, 360.0, R)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat((a.range, b.range)).view(torch.int32), whole.range.view(torch.int32))
    # first-hit DISTANCES are a minimum and bit-identical; where two obstacles' edges lie within one float32 ulp along a
    # ray (crossing rectangles), either may be reported as the owner -- the staging order of a frame's edges is not fixed
    assert (torch.cat((a.hit, b.hit)) != whole.hit).float().mean().item() < 1e-5
    assert (torch.cat((a.visible, b.visible)) != whole.visible).float().mean().item() < 1e-5
    # the instrumented twin (fo_visibility_stats) writes the same outputs, and its counters nest as the culls do
    cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
    c = raycast_frames(ego, rect, flags, None, 50.0, 360.0, R, stats=cnt)
    torch.cuda.synchronize()
    assert torch.equal(c.range.view(torch.int32), whole.range.view(torch.int32))
    tested, skipped, staged, listed = cnt.cpu().tolist()
    assert 0 < staged <= F * 4 * O and 0 < listed <= staged * (R // 256)
    assert tested + 32 * skipped <= listed * 256 and 0 < tested < F * R * 4 * O
    perm = torch.randperm(O, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    p = raycast_frames(ego, rect[:, perm].contiguous(), flags[:, perm].contiguous(), None, 50.0, 360.0, R)
    torch.cuda.synchronize()
    assert torch.equal(p.range.view(torch.int32), whole.range.view(torch.int32))          # first-hit distance: a min
    assert torch.equal(p.visible, whole.visible[:, perm])
    mapped = torch.where(p.hit >= 0, perm[p.hit.clamp(min=0).long()].to(torch.int32), p.hit)
    assert (mapped != whole.hit).float().mean().item() < 1e-4       # equal-distance corner hits may pick the other owner
    for f in (0, F - 1):
        o_rng, o_hit, o_vis = VO.raycast(ego[f].cpu().numpy().astype(np.float64), rect[f].cpu().numpy().astype(np.float64),
                                         np.ones(O, np.uint8), None, 50.0, 360.0, R)
        g = whole.range[f].cpu().numpy()
        assert (~np.isclose(g, o_rng, rtol=2e-5, atol=2e-4)).mean() < 0.004
        assert (whole.hit[f].cpu().numpy() != o_hit).mean() < 0.006
