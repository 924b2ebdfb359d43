"""BASELINE.json configs 3-5 as parity cases (config 1 and 2 live in test_gpu_scan_simulator.py and
test_gpu_march.py): full-size runs checked against the oracle where it finishes in seconds, and
through size-independent properties (determinism, permutation equivariance, subset == oracle)
where it does not."""
import numpy as np
import pytest

from pyracecarsimulator_b200 import maps, range_libc
from pyracecarsimulator_b200.racecar import BatchedCar
from gpu_util import assert_ranges_match, build_synth

pytestmark = pytest.mark.gpu
FOV = 4.71


def test_config3_particle_filter_full_size(orc):
    """1 M poses x 60 angles via calc_range_repeat_angles on the big-teach stand-in (2049^2)."""
    omap, y, occ, dist = build_synth(orc, 2049, 1234)
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    n, a = 1_000_000, 60
    poses = maps.sample_free_poses(dist, n, 303, y.resolution, y.origin)
    angles = np.linspace(-FOV / 2, FOV / 2, a, endpoint=False).astype(np.float32)
    out = np.zeros(n * a, np.float32)
    rm.calc_range_repeat_angles(poses, angles, out)
    want = orc.Marcher(dist, 300, y.resolution, y.origin).calc_range_repeat_angles(poses, angles, threads=0)
    assert_ranges_match(out, want, y.resolution)


def test_config4_fused_rollout_full_size(orc, colombia, colombia_scan):
    """65 536 cars x 50 steps x 1080 beams on maps/colombia: 3.54 G rays, nothing materialised.
    Checked on a random subset of cars against the oracle, plus determinism."""
    import torch
    binar = np.where(colombia_scan["grid"] > 0, 255, 0).ravel()
    omap = range_libc.PyOMap(maps.OccupancyGrid.make(binar, 435, 350, colombia["resolution"], colombia["origin"]))
    dist = orc.edt_float(colombia_scan["occ"])
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    car = BatchedCar()
    car.setCarEdgeDistances(1080, -FOV / 2.0, FOV / 1080, 0.275)
    n, steps = 65536, 50
    rng = np.random.default_rng(42)
    start = maps.sample_free_poses(dist, n, 404, colombia["resolution"], colombia["origin"], min_clear_px=6.0)
    s0 = np.zeros((n, 11))
    s0[:, :3] = start
    s0[:, 3] = 2.0
    # action schedule as SURVEY.md 8d fixes it: scripts/mcts.py:216-222 from Philox, seed 42, drawn on
    # the device; the oracle's copy of the same schedule drives the CPU check below
    actions = orc.rollout_actions(n, 5, seed=42)
    st = torch.from_numpy(s0.copy()).cuda()
    out = car.rollout(rm, st, None, steps, FOV, seed=42)
    d_actions = out["actions"]
    assert np.array_equal(d_actions.cpu().numpy(), actions)
    idx = out["crash_index"].cpu().numpy()
    assert np.all((idx == -(steps + 1)) | ((idx >= 0) & (idx < steps)))
    st2 = torch.from_numpy(s0.copy()).cuda()
    out2 = car.rollout(rm, st2, d_actions, steps, FOV)
    assert torch.equal(out["crash_index"], out2["crash_index"]) and torch.equal(st, st2)
    # subset against the oracle
    p = orc.car_params()
    edge = car.edge_distances()
    m = orc.Marcher(dist, 300, colombia["resolution"], colombia["origin"])
    sub = rng.choice(n, 48, replace=False)
    agree = 0
    for c in sub:
        s = s0[c].copy()
        poses = np.zeros((steps, 3), np.float32)
        for i in range(steps):
            orc.car_step(p, s, actions[c, i // 10, 0], actions[c, i // 10, 1], 0.01)
            poses[i] = s[:3]
        want = orc.car_is_crashed(m.calc_range_fan(poses, 1080, FOV), edge, 1080, steps, 0.001)
        agree += int(want == idx[c])
    assert agree == 48   # the f32 poses of all 65 536 cars equal the reference Car's (next test), so every index must agree
    print(f"config 4: {np.mean(idx >= 0) * 100:.1f}% of cars crash within {steps} steps")


def test_config4_crash_indices_vs_reference_car_all_cars(orc, colombia, colombia_scan):
    """the device's fp64 cos/sin/tan are CUDA's, not the host libm's, so a car state can
    differ from the reference Car in the last ulp.  Count what that does to the RESULT over all 65 536 cars of
    config 4: run the vehicle half on the CPU with the oracle (bit-identical to the unmodified reference Car,
    tests/test_car_oracle.py), scan those poses with the same marcher, compare every crash index."""
    import torch
    binar = np.where(colombia_scan["grid"] > 0, 255, 0).ravel()
    omap = range_libc.PyOMap(maps.OccupancyGrid.make(binar, 435, 350, colombia["resolution"], colombia["origin"]))
    dist = orc.edt_float(colombia_scan["occ"])
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    car = BatchedCar()
    car.setCarEdgeDistances(1080, -FOV / 2.0, FOV / 1080, 0.275)
    n, steps = 65536, 50
    start = maps.sample_free_poses(dist, n, 404, colombia["resolution"], colombia["origin"], min_clear_px=6.0)
    s0 = np.zeros((n, 11))
    s0[:, :3] = start
    s0[:, 3] = 2.0
    out = car.rollout(rm, torch.from_numpy(s0.copy()).cuda(), None, steps, FOV, seed=42)
    got_idx = out["crash_index"].cpu().numpy()
    got_poses = out["poses"].cpu().numpy()
    actions = orc.rollout_actions(n, 5, seed=42)
    ref_poses, ref_states = orc.car_rollout_poses(orc.car_params(), s0.copy(), actions, steps)
    pose_diff = int((got_poses.view(np.uint32) != ref_poses.view(np.uint32)).any(axis=2).sum())
    # crash indices of the CPU-produced poses through the same scan + crash kernel (group-major: car, step)
    gm = torch.from_numpy(np.ascontiguousarray(ref_poses.transpose(1, 0, 2)).reshape(n * steps, 3)).cuda()
    ref_idx, _ = car.scan_crash(rm, gm, n, steps, FOV)
    ref_idx = ref_idx.cpu().numpy()
    idx_diff = int((ref_idx != got_idx).sum())
    state_rel = float(np.max(np.abs(out["poses"].cpu().numpy().astype(np.float64) - ref_poses) / (np.abs(ref_poses) + 1e-12)))
    print(f"config 4 vs reference Car, all {n} cars: {pose_diff} of {n * steps} fp32 poses differ in any bit "
          f"(max relative difference {state_rel:.2e}); {idx_diff} crash indices differ")
    # measured on B200 (CUDA 12.9 libdevice): 0 of 3 276 800 poses differ in any bit, 0 crash indices differ
    assert idx_diff == 0, idx_diff
    assert pose_diff <= n * steps // 1000, pose_diff


def test_config5_large_map_reduced_pose_count(orc):
    """8192^2 synthetic map (256 MiB fp32 field, not L2-resident); 270 beams."""
    n = 8192
    img = maps.synth_map(n, 5678)
    grid = orc.mapserver_occupancy(img)
    occ = orc.omap_from_grid(grid, True)
    y = maps.synth_yaml(n)
    omap = range_libc.PyOMap(maps.OccupancyGrid.make(grid.ravel(), n, n, y.resolution, y.origin))
    d2 = orc.edt_exact(occ)
    assert np.array_equal(omap.dist2(), d2)                      # exact integer EDT is the definition here
    dist = orc.sqrt_dist2(d2)
    assert np.array_equal(omap.dist(), dist)
    differs = int((orc.edt_float(occ) != dist).sum())
    print(f"8192^2: float-Felzenszwalb restatement differs from the exact EDT on {differs} cells")
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    poses = maps.sample_free_poses(dist, 100_000, 505, y.resolution, y.origin)
    out = np.zeros(100_000 * 270, np.float32)
    rm.calc_range_fan(poses, out, FOV, 270)
    want = orc.Marcher(dist, 300, y.resolution, y.origin).calc_range_fan(poses, 270, FOV, threads=0)
    assert_ranges_match(out, want, y.resolution)


def test_more_than_2_31_rays_in_one_call(orc):
    """One call with > 2^31 rays (config 5 is 4.3 G rays over 8 GPUs, > 2^31 per call on fewer GPUs):
    exercises the 64-bit ray-index path.  Checked on the device against smaller calls over slices
    (which take the 32-bit path, itself checked against the oracle above)."""
    import torch
    omap, y, occ, dist = build_synth(orc, 1025, 11)
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    base = maps.sample_free_poses(dist, 65536, 606, y.resolution, y.origin)
    beams = 270
    n = 8_060_928                      # 123 * 65536 poses -> 2.176e9 rays
    assert n * beams > 2 ** 31
    poses = torch.from_numpy(np.tile(base, (n // 65536, 1))).cuda()
    out = torch.empty(n * beams, dtype=torch.float32, device="cuda")
    rm.calc_range_fan(poses, out, FOV, beams)
    ref = torch.empty(65536 * beams, dtype=torch.float32, device="cuda")
    rm.calc_range_fan(poses[:65536].contiguous(), ref, FOV, beams)
    want = orc.Marcher(dist, 300, y.resolution, y.origin).calc_range_fan(base[:512], beams, FOV, threads=0)
    assert np.array_equal(ref[:512 * beams].cpu().numpy(), want)
    for blk in (0, 61, 122):           # first, middle (beyond 2^31 / 2), last tile
        lo = blk * 65536 * beams
        assert torch.equal(out[lo:lo + 65536 * beams], ref), blk
    del out, poses
    torch.cuda.empty_cache()
