"""Vehicle model, crash test and fused rollout on the GPU (csrc/car.cu) against the oracle
(oracle/car_oracle.c, itself bit-exact against the unmodified reference Car) and the golden
vectors produced by the reference Car.

Tolerances: the crash test and edge distances are exact.  The fp64 dynamics use CUDA's
cos/sin/tan, which may differ from the host libm in the last ulp, so states are compared to
1e-11 relative (+1e-13 absolute) after up to 200 steps; poses narrowed to fp32 must then be
identical for >= 99.99 % of entries and crash indices for >= 99.9 % of cars."""
import numpy as np
import pytest

from pyracecarsimulator_b200 import maps, range_libc
from pyracecarsimulator_b200.racecar import BatchedCar, car_params

pytestmark = pytest.mark.gpu

FOV, NRAYS, DIST_TO_BASE = 4.71, 1080, 0.275


@pytest.fixture(scope="module")
def car():
    c = BatchedCar()
    c.setCarEdgeDistances(NRAYS, -FOV / 2.0, FOV / NRAYS, DIST_TO_BASE)
    return c


@pytest.fixture(scope="module")
def col(orc, colombia, colombia_scan):
    binar = np.where(colombia_scan["grid"] > 0, 255, 0).ravel()
    omap = range_libc.PyOMap(maps.OccupancyGrid.make(binar, 435, 350, colombia["resolution"], colombia["origin"]))
    dist = orc.edt_float(colombia_scan["occ"])
    return dict(rm=range_libc.PyRayMarchingGPU(omap, 300), orc=orc.Marcher(dist, 300, colombia["resolution"], colombia["origin"]),
                dist=dist, res=colombia["resolution"], origin=colombia["origin"])


def close_states(a, b):
    return np.allclose(a, b, rtol=1e-11, atol=1e-13)


def test_params_follow_the_reference_constructor_order(orc):
    assert np.array_equal(car_params(), orc.car_params().as_array())


def test_edge_distances_exact(orc, car, car_golden):
    want = orc.car_edge_distances(orc.car_params(), NRAYS, -FOV / 2.0, FOV / NRAYS, DIST_TO_BASE)
    assert np.array_equal(car.edge_distances(), want)
    # and consistent with the reference's own isCrashed boundary (golden, from the real Car)
    safe = car_golden["first_safe_ray"]
    assert np.all(safe.astype(np.float64) - want >= 0.001)


def test_step_matches_golden_reference_trajectories(car, car_golden):
    import torch
    g = car_golden
    n, steps = g["init"].shape[0], g["states"].shape[1]
    st = torch.from_numpy(g["init"].copy()).cuda()
    for i in range(steps):
        a = torch.from_numpy(np.ascontiguousarray(g["actions"][:, i // 10])).cuda()
        car.step(st, a[:, 0].contiguous(), a[:, 1].contiguous(), 0.01)
        assert close_states(st.cpu().numpy(), g["states"][:, i]), i
    pose = BatchedCar.scan_pose(st, DIST_TO_BASE).cpu().numpy()
    assert np.allclose(pose, g["scan_poses"][:, -1], rtol=1e-11, atol=1e-13)


def test_step_random_states_vs_oracle(orc, car):
    import torch
    rng = np.random.default_rng(1)
    n = 5000
    s0 = np.zeros((n, 11))
    s0[:, :2] = rng.uniform(-5, 5, (n, 2))
    s0[:, 2] = rng.uniform(-np.pi, np.pi, n)
    s0[:, 3] = rng.uniform(-1, 7.5, n)
    s0[:, 4] = rng.uniform(-0.45, 0.45, n)
    s0[:, 5] = rng.uniform(-1, 1, n)
    s0[:, 6] = rng.uniform(-0.2, 0.2, n)
    s0[:, 7] = rng.integers(0, 2, n)
    speed, steer = rng.uniform(0, 7, n), rng.uniform(-0.4189, 0.4189, n)
    st = torch.from_numpy(s0.copy()).cuda()
    car.step(st, torch.from_numpy(speed).cuda(), torch.from_numpy(steer).cuda(), 0.01)
    p = orc.car_params()
    want = s0.copy()
    for i in range(n):
        orc.car_step(p, want[i], speed[i], steer[i], 0.01)
    got = st.cpu().numpy()
    assert close_states(got, want)
    assert np.array_equal(got[:, 7], want[:, 7]) and np.array_equal(got[:, 10], want[:, 10])


def test_is_crashed_exact_boundary(orc, car, car_golden):
    import torch
    edge = car.edge_distances()
    rng = np.random.default_rng(2)
    groups, per = 40, 6
    rays = rng.uniform(0.3, 10.0, (groups, per, NRAYS)).astype(np.float32)
    safe = car_golden["first_safe_ray"]
    for g in range(groups):                     # plant exact-boundary values from the reference
        k, j = rng.integers(per), rng.integers(NRAYS)
        rays[g, k, j] = safe[j] if g % 2 else np.nextafter(safe[j], np.float32(-1))
    rays[3] = 5.0
    got = car.is_crashed_many(torch.from_numpy(rays).cuda().reshape(-1), groups, per).cpu().numpy()
    want = np.array([orc.car_is_crashed(rays[g].ravel(), edge, NRAYS, per, 0.001) for g in range(groups)])
    assert np.array_equal(got, want)
    assert got[3] == -(per + 1)
    # upstream signature
    assert car.isCrashed(rays[5].ravel(), NRAYS, per) == want[5]
    clear = np.full(3 * NRAYS, 5.0, np.float32)
    assert car.isCrashed(clear, NRAYS, 3) == -4
    clear[NRAYS + 500] = 0.05
    assert car.isCrashed(clear, NRAYS, 3) == 1


def test_scan_crash_matches_scan_then_is_crashed(orc, car, col):
    import torch
    edge = car.edge_distances()
    rng = np.random.default_rng(3)
    groups, per = 24, 10
    poses = maps.sample_free_poses(col["dist"], groups * per, 77, col["res"], col["origin"], min_clear_px=1.0)
    want_ranges = col["orc"].calc_range_fan(poses, NRAYS, FOV)
    want = np.array([orc.car_is_crashed(want_ranges[g * per * NRAYS:(g + 1) * per * NRAYS], edge, NRAYS, per, 0.001)
                     for g in range(groups)])
    assert (want >= 0).any() and (want < 0).any()
    dp = torch.from_numpy(poses).cuda()
    first, ranges = car.scan_crash(col["rm"], dp, groups, per, FOV, want_ranges=True)
    assert np.array_equal(ranges.cpu().numpy(), want_ranges)
    assert np.array_equal(first.cpu().numpy(), want)
    first2, none = car.scan_crash(col["rm"], dp, groups, per, FOV, want_ranges=False)
    assert none is None and np.array_equal(first2.cpu().numpy(), want)


def oracle_rollout(orc, marcher, edge, s0, actions, steps, lidar_pose):
    p = orc.car_params()
    st = s0.copy()
    poses = np.zeros((steps, 3), np.float32)
    v = np.zeros(steps)
    for i in range(steps):
        sp, sa = actions[i // 10]
        orc.car_step(p, st, sp, sa, 0.01)
        src = orc.car_scan_pose(st, DIST_TO_BASE) if lidar_pose else st[:3]
        poses[i] = src            # f64 -> f32, where the reference narrows (scripts/mcts.py:229-231)
        v[i] = st[3]
    ranges = marcher.calc_range_fan(poses, NRAYS, FOV)
    idx = orc.car_is_crashed(ranges, edge, NRAYS, steps, 0.001)
    reward = v.sum() if idx < 0 else v[:idx].sum()
    return st, poses, idx, reward


@pytest.mark.parametrize("lidar_pose", [False, True])
def test_fused_rollout_vs_oracle(orc, car, col, lidar_pose):
    import torch
    rng = np.random.default_rng(4)
    n, steps = 96, 50
    start = maps.sample_free_poses(col["dist"], n, 99, col["res"], col["origin"], min_clear_px=6.0)
    s0 = np.zeros((n, 11))
    s0[:, :3] = start
    s0[:, 3] = 2.0
    actions = np.stack([rng.uniform(0, 7.0, (n, 5)), rng.uniform(-0.4189, 0.4189, (n, 5))], axis=2)
    st = torch.from_numpy(s0.copy()).cuda()
    out = car.rollout(col["rm"], st, torch.from_numpy(actions).cuda(), steps, FOV, lidar_pose=lidar_pose,
                      scan_dist_to_base=DIST_TO_BASE)
    edge = car.edge_distances()
    got_idx, got_rew = out["crash_index"].cpu().numpy(), out["reward"].cpu().numpy()
    got_poses = out["poses"].cpu().numpy()          # (steps, n, 3), step-major
    same_idx, same_pose = 0, 0
    for c in range(n):
        w_st, w_poses, w_idx, w_rew = oracle_rollout(orc, col["orc"], edge, s0[c], actions[c], steps, lidar_pose)
        assert close_states(st[c].cpu().numpy(), w_st), c
        same_pose += int((got_poses[:, c] == w_poses).sum())
        if got_idx[c] == w_idx:
            same_idx += 1
            assert got_rew[c] == pytest.approx(w_rew, rel=1e-11, abs=1e-12)
    assert same_pose >= 0.9999 * n * steps * 3
    assert same_idx >= 0.999 * n
    assert (got_idx >= 0).any() and (got_idx < 0).any()     # the case exercises both outcomes
    assert np.all(got_idx[got_idx < 0] == -(steps + 1))


def test_device_action_schedule_is_the_oracles_bit_for_bit(orc, car):
    """rl_rollout_actions (Philox4x32-10 on the device) == oracle/philox_oracle.c, every bit, for
    several shapes, seeds, stream ids and car offsets (including offsets beyond 2^32)."""
    for (n, a, seed, sid, off) in [(1, 1, 42, 0, 0), (4099, 5, 42, 0, 0), (513, 7, 2**40 + 3, 9, 2**33 + 5),
                                   (65536, 5, 42, 0, 0), (1000, 5, 0, 0xFFFFFFFF, 123456789)]:
        got = car.random_actions(n, a, seed=seed, stream_id=sid, car_offset=off).cpu().numpy()
        want = orc.rollout_actions(n, a, seed=seed, stream_id=sid, car_offset=off)
        assert got.shape == (n, a, 2)
        assert np.array_equal(got, want), (n, a, seed, sid, off)
    got = car.random_actions(64, 3, seed=7, speed_range=(1.0, 2.0), steer_range=(-0.1, 0.3)).cpu().numpy()
    assert np.array_equal(got, orc.rollout_actions(64, 3, seed=7, speed_range=(1.0, 2.0), steer_range=(-0.1, 0.3)))


def test_rollout_with_device_drawn_actions(orc, car, col):
    """actions=None: the schedule never exists on the host; the rollout equals the one fed with the
    oracle's copy of the same schedule, and a rank's slice (car_offset) equals the whole job's."""
    import torch
    n, steps = 64, 50
    start = maps.sample_free_poses(col["dist"], n, 5, col["res"], col["origin"], min_clear_px=6.0)
    s0 = np.zeros((n, 11))
    s0[:, :3] = start
    s0[:, 3] = 2.0
    st_a = torch.from_numpy(s0.copy()).cuda()
    out_a = car.rollout(col["rm"], st_a, None, steps, FOV, seed=42)
    acts = orc.rollout_actions(n, 5, seed=42)
    assert np.array_equal(out_a["actions"].cpu().numpy(), acts)
    st_b = torch.from_numpy(s0.copy()).cuda()
    out_b = car.rollout(col["rm"], st_b, torch.from_numpy(acts).cuda(), steps, FOV)
    for k in ("crash_index", "reward", "poses", "vsum"):
        assert torch.equal(out_a[k], out_b[k]), k
    assert torch.equal(st_a, st_b)
    # the node value MCTS.rollout returns: reward / abs(node.action), same bits as numpy's division
    na = np.random.default_rng(3).uniform(-0.41, 0.41, n)
    st_v = torch.from_numpy(s0.copy()).cuda()
    out_v = car.rollout(col["rm"], st_v, None, steps, FOV, seed=42, node_action=torch.from_numpy(na).cuda())
    assert torch.equal(out_v["reward"], out_a["reward"])
    assert np.array_equal(out_v["value"].cpu().numpy(), out_a["reward"].cpu().numpy() / np.abs(na))
    st_c = torch.from_numpy(s0[32:].copy()).cuda()
    out_c = car.rollout(col["rm"], st_c, None, steps, FOV, seed=42, car_offset=32)
    assert torch.equal(out_c["crash_index"], out_a["crash_index"][32:])
    assert torch.equal(out_c["poses"], out_a["poses"][:, 32:])


def test_rollout_is_cuda_graph_capturable(orc, car, col):
    """The device-pointer entry points only enqueue kernels on the caller's stream (no allocation, no
    synchronisation, the L2 access-policy window travels as a launch attribute), so an MCTS loop can
    capture action draw + rollout + scans + crash test into one CUDA graph and replay it."""
    import torch
    n, steps = 256, 30
    start = maps.sample_free_poses(col["dist"], n, 8, col["res"], col["origin"], min_clear_px=6.0)
    s0 = np.zeros((n, 11))
    s0[:, :3] = start
    s0[:, 3] = 2.0
    d_s0 = torch.from_numpy(s0).cuda()
    eager_states = d_s0.clone()
    eager = car.rollout(col["rm"], eager_states, None, steps, FOV, seed=7)
    states = d_s0.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm-up outside capture, as torch recommends
        car.rollout(col["rm"], states, None, steps, FOV, seed=7)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    states.copy_(d_s0)
    with torch.cuda.graph(g):
        out = car.rollout(col["rm"], states, None, steps, FOV, seed=7)
    for _ in range(3):                                  # replay: same inputs -> same outputs
        states.copy_(d_s0)
        out["crash_index"].fill_(12345)
        g.replay()
        torch.cuda.synchronize()
        for k in ("crash_index", "reward", "poses", "vsum", "actions"):
            assert torch.equal(out[k], eager[k]), k
        assert torch.equal(states, eager_states)


def test_argument_errors(car, col):
    import torch
    fresh = BatchedCar()
    with pytest.raises(ValueError):          # edge distances not set yet
        fresh.is_crashed_many(torch.zeros(NRAYS, dtype=torch.float32, device="cuda"), 1, 1)
    with pytest.raises(ValueError):
        car.step(torch.zeros((4, 10), dtype=torch.float64, device="cuda"), torch.zeros(4, dtype=torch.float64, device="cuda"),
                 torch.zeros(4, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        car.rollout(col["rm"], torch.zeros((4, 11), dtype=torch.float64, device="cuda"),
                    torch.zeros((4, 3, 2), dtype=torch.float64, device="cuda"), 50, FOV)
    with pytest.raises(ValueError):
        car.random_actions(4, 0)
    with pytest.raises(ValueError):
        car.random_actions(4, 5, car_offset=-1)
