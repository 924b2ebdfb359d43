"""Regenerates the committed fixtures under tests/golden/ (run from the repo root, in the build
container where /root/reference exists):

    python tests/golden/make_golden.py

colombia_map.npz   the only map image present in the reference checkout
                   (maps/colombia/map.pgm + map.yaml), stored as a compressed uint8 array so the
                   GPU box -- which has no /root/reference -- can run the same tests.
colombia_scan.npz  outputs of the CPU oracle (oracle/rangelib_oracle.c) on that map: occupancy,
                   exact d^2, and ranges + step counts for seeded poses through all three entry
                   points.  The scan path's arithmetic lives in the external, un-vendored
                   range_libc, so these pin the ORACLE (regressions, platform drift), not the
                   reference: parity at that boundary is unpinned.
car_golden.npz     outputs of the UNMODIFIED reference Car class (racecar/src/racecar.cpp via
                   oracle/_ref/libracecar_ref.so): state trajectories under the MCTS action
                   schedule, lidar poses, edge distances probed through isCrashed.  These DO pin
                   the vehicle-model half of the fused rollout to the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pyracecarsimulator_b200 import maps  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def colombia():
    y = maps.load_map_yaml(os.path.join(REF, "maps/colombia/map.yaml"))
    img = maps.read_pgm(y.image)
    np.savez_compressed(os.path.join(OUT, "colombia_map.npz"), img=img, resolution=y.resolution,
                        origin=np.array(y.origin), negate=y.negate,
                        occupied_thresh=y.occupied_thresh, free_thresh=y.free_thresh)
    grid = oracle.mapserver_occupancy(img, y.negate, y.occupied_thresh, y.free_thresh)
    occ = oracle.omap_from_grid(grid, True)
    d2 = oracle.edt_exact(occ)
    dist = oracle.edt_float(occ)
    assert np.array_equal(oracle.sqrt_dist2(d2), dist)
    m = oracle.Marcher(dist, 300, y.resolution, y.origin)
    poses = maps.sample_free_poses(dist, 48, 2024, y.resolution, y.origin)
    poses[0] = (0.275, 0.0, 0.0)  # lidar pose of the zero-state car
    fan, fan_steps = m.calc_range_fan(poses, 1080, 4.71, steps=True)
    rng = np.random.default_rng(99)
    rays = maps.sample_free_poses(dist, 4096, 2025, y.resolution, y.origin)
    rays[:64, 0] += rng.uniform(-30, 30, 64).astype(np.float32)  # some start outside the map
    many, many_steps = m.calc_range_many(rays, steps=True)
    angles = np.linspace(-4.71 / 2, 4.71 / 2, 60, endpoint=False).astype(np.float32)
    rep, rep_steps = m.calc_range_repeat_angles(poses, angles, steps=True)
    np.savez_compressed(os.path.join(OUT, "colombia_scan.npz"), grid=grid, occ=occ, d2=d2,
                        poses=poses, fan=fan, fan_steps=fan_steps, rays=rays, many=many,
                        many_steps=many_steps, angles=angles, rep=rep, rep_steps=rep_steps)


def car():
    p = oracle.car_params()
    rng = np.random.default_rng(42)
    n_traj, n_steps = 16, 120
    init = np.zeros((n_traj, 11))
    init[:, 0:2] = rng.uniform(-5, 5, (n_traj, 2))
    init[:, 2] = rng.uniform(-np.pi, np.pi, n_traj)
    init[:, 3] = rng.uniform(0, 7, n_traj)
    init[0] = 0.0  # the survey's known-answer start
    actions = np.zeros((n_traj, n_steps // 10, 2))
    actions[..., 0] = rng.uniform(0, 7.0, actions.shape[:2])          # speed, scripts/mcts.py:220-221
    actions[..., 1] = rng.uniform(-0.4189, 0.4189, actions.shape[:2])  # steer, scripts/mcts.py:217-219
    actions[0, :, 0], actions[0, :, 1] = 2.0, 0.2
    states = np.zeros((n_traj, n_steps, 11))
    scan_poses = np.zeros((n_traj, n_steps, 3))
    car_ = oracle.RefCar(p)
    for t in range(n_traj):
        car_.set_state(init[t])
        for i in range(n_steps):
            if i % 10 == 0:
                car_.control(*actions[t, i // 10])
            car_.update(0.01)
            states[t, i] = car_.get_state()
            scan_poses[t, i] = car_.scan_pose(0.275)
    # edge distances are private in the reference; recover each one exactly by bisection on
    # isCrashed (crash <=> (double)ray - edge[j] < 0.001) over float32 ray values
    num_rays, fov = 1080, 4.71
    car_.set_edges(num_rays, -fov / 2.0, fov / num_rays, 0.275)
    first_crash_ray = np.zeros(num_rays, dtype=np.float32)  # smallest f32 range that does NOT crash
    for j in range(num_rays):
        lo, hi = np.float32(-1.0), np.float32(2.0)
        while True:
            mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
            if mid == lo or mid == hi:
                break
            r = np.full(num_rays, 100.0, dtype=np.float32)
            r[j] = mid
            if car_.is_crashed(r, num_rays, 1) == 0:
                lo = mid
            else:
                hi = mid
        first_crash_ray[j] = hi
    np.savez_compressed(os.path.join(OUT, "car_golden.npz"), params=p.as_array(), init=init,
                        actions=actions, states=states, scan_poses=scan_poses,
                        first_safe_ray=first_crash_ray)


def philox():
    """Frozen head of the default rollout action schedule (seed 42, stream 0): 8 cars x 5 actions."""
    np.save(os.path.join(OUT, "philox_seed42.npy"), oracle.rollout_actions(8, 5, seed=42))


if __name__ == "__main__":
    colombia()
    car()
    philox()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
