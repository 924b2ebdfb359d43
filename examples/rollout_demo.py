"""The reference's MCTS inner loop on the GPU path, end to end on maps/colombia (from the committed fixture):

  expansion  scan at the car's lidar pose -> follow-the-gap steering proposal      (scripts/mcts.py:187-200, :262-267)
  rollout    N perturbed copies of the proposal, 200 bicycle steps each with a scan + crash test after
             every step, nothing leaving the GPU                                  (scripts/mcts.py:202-245)
  choice     reward = sum of velocities before the crash / |action|, best action wins

    python examples/rollout_demo.py [n_candidates]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from pyracecarsimulator_b200.followgap import PyFollowGap  # noqa: E402
from pyracecarsimulator_b200.racecar import DEFAULT_CAR_CONFIG, BatchedCar  # noqa: E402
from pyracecarsimulator_b200.racecar_simulator import RacecarSimulator  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    z = np.load(os.path.join(ROOT, "tests", "golden", "colombia_map.npz"))
    path = "/tmp/_demo_colombia.pgm"
    maps.write_pgm(path, z["img"])
    yaml = maps.MapYaml(path, float(z["resolution"]), tuple(float(v) for v in z["origin"]))
    omap = range_libc.PyOMap(yaml)                                     # GPU ingest: occupancy + exact EDT

    cfg = dict(DEFAULT_CAR_CONFIG, scan_dist_to_base=0.275, batch_size=200, scan_beams=1080, scan_fov=4.71,
               scan_std=0.01, scan_max_range=15.0, free_thresh=0.8)   # params.yaml
    sim = RacecarSimulator(cfg)
    sim.setMap(omap, yaml.resolution, yaml.origin)
    sim.setRaytracingMethod("RMGPU")
    sim.drive(2.0, 0.0)
    for _ in range(30):
        sim.updatePose()
    sim.runScan()                                                      # config 1: one 1080-beam scan
    fg = PyFollowGap(10, 15.0, cfg["max_steer_ang"], cfg["scan_fov"] / cfg["scan_beams"])
    proposal = fg.eval(sim.getScan(), cfg["scan_beams"])
    print(f"scan min/max {sim.getScan().min():.2f}/{sim.getScan().max():.2f} m, follow-the-gap proposes {proposal:+.3f} rad")

    # N candidate steering actions around the proposal, each rolled out 200 steps with random speed changes
    rng = np.random.default_rng(0)
    cand = np.clip(proposal + rng.uniform(-0.2, 0.2, n), -cfg["max_steer_ang"], cfg["max_steer_ang"])
    cand[np.abs(cand) < 1e-3] = 1e-3
    steps = cfg["batch_size"]
    actions = np.zeros((n, steps // 10, 2))
    actions[:, :, 0] = rng.uniform(0, cfg["max_speed"], (n, steps // 10))
    actions[:, :, 1] = rng.uniform(-cfg["max_steer_ang"], cfg["max_steer_ang"], (n, steps // 10))
    actions[:, 0, 1] = cand                                            # first action = the candidate
    states = torch.from_numpy(np.tile(sim.getState(), (n, 1))).cuda()
    car = BatchedCar(cfg)
    car.setCarEdgeDistances(cfg["scan_beams"], -cfg["scan_fov"] / 2.0, cfg["scan_fov"] / cfg["scan_beams"],
                            cfg["scan_dist_to_base"])
    d_actions = torch.from_numpy(actions).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = car.rollout(sim.scan_simulator.scan_method, states, d_actions, steps, cfg["scan_fov"])
    reward = (out["reward"] / torch.from_numpy(np.abs(cand)).cuda()).cpu().numpy()
    dt = time.perf_counter() - t0
    crash = out["crash_index"].cpu().numpy()
    best = int(np.argmax(reward))
    print(f"{n} rollouts x {steps} steps x 1080 beams = {n * steps * 1080 / 1e9:.2f} G nominal rays in {dt * 1e3:.1f} ms; "
          f"{(crash >= 0).mean() * 100:.0f}% crash; best action {cand[best]:+.3f} rad (reward {reward[best]:.1f})")


if __name__ == "__main__":
    main()
