"""Throughput of the BASELINE.json configs that bench.py does not report (configs 3, 4, 5), for
DESIGN.md.  Device time with CUDA events, inputs resident in HBM, 3 warm-ups, best-of/mean of N.
    python tools/bench_configs.py [--reps 10]
Prints one JSON line per config."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from pyracecarsimulator_b200.racecar import BatchedCar  # noqa: E402

FOV = 4.71


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def synth(n, seed):
    img = maps.synth_map(n, seed)
    y = maps.synth_yaml(n)
    path = f"/tmp/_rl_cfg_{n}.pgm"
    maps.write_pgm(path, img)
    y.image = path
    omap = range_libc.PyOMap(y)
    return omap, y, omap.dist()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()

    # config 1: single-pose 1080-beam scan through ScanSimulator2D.scan (latency, host in / host out)
    import time
    from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D
    golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "colombia_map.npz")
    z = np.load(golden)
    path = "/tmp/_rl_cfg_colombia.pgm"
    maps.write_pgm(path, z["img"])
    yc = maps.MapYaml(path, float(z["resolution"]), tuple(float(v) for v in z["origin"]))
    omap1 = range_libc.PyOMap(yc)
    sim = ScanSimulator2D(1080, FOV, 0.01, batch_size=200)
    sim.setMap(omap1, 300, yc.resolution, yc.origin)
    sim.setRaytracingMethod("RMGPU")
    for _ in range(20):
        sim.scan(0.275, 0.0, 0.0)
    t0 = time.perf_counter()
    for _ in range(500):
        sim.scan(0.275, 0.0, 0.0)
    us = (time.perf_counter() - t0) / 500 * 1e6
    poses200 = maps.sample_free_poses(omap1.dist(), 200, 7, yc.resolution, yc.origin)
    for _ in range(20):
        sim.scanMany(poses200)
    t0 = time.perf_counter()
    for _ in range(300):
        sim.scanMany(poses200)
    us200 = (time.perf_counter() - t0) / 300 * 1e6
    print(json.dumps({"config": 1, "workload": "ScanSimulator2D.scan: 1 pose x 1080 beams on maps/colombia, host floats in, host ranges out",
                      "us_per_scan": us, "rays_per_s": 1080 / (us * 1e-6),
                      "scanMany_200_poses_us": us200, "scanMany_200_rays_per_s": 200 * 1080 / (us200 * 1e-6),
                      "note": "reference budget: 20 Hz real-time = 50 000 us per scan; MCTS rollout batch = 200 poses"}))

    # config 3: particle filter, 1M poses x 60 angles on the 2049^2 stand-in
    omap, y, dist = synth(2049, 1234)
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    n, a = 1_000_000, 60
    poses = torch.from_numpy(maps.sample_free_poses(dist, n, 303, y.resolution, y.origin)).cuda()
    angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, a, endpoint=False).astype(np.float32)).cuda()
    out = torch.empty(n * a, dtype=torch.float32, device="cuda")
    mean, best = timeit(lambda: rm.calc_range_repeat_angles(poses, angles, out), args.reps)
    print(json.dumps({"config": 3, "workload": "1M poses x 60 angles (calc_range_repeat_angles), 2049^2 map",
                      "rays": n * a, "ms_mean": mean, "ms_best": best, "grays_per_s": n * a / (mean * 1e-3) / 1e9,
                      "ingest_ms": omap.ingest_ms}))
    del poses, out

    # config 5 (single GPU share): 8192^2 map, 2M poses x 270 beams
    omap5, y5, dist5 = synth(8192, 5678)
    rm5 = range_libc.PyRayMarchingGPU(omap5, 300)
    n5, b5 = 2_000_000, 270
    poses5 = torch.from_numpy(maps.sample_free_poses(dist5, n5, 505, y5.resolution, y5.origin)).cuda()
    out5 = torch.empty(n5 * b5, dtype=torch.float32, device="cuda")
    mean, best = timeit(lambda: rm5.calc_range_fan(poses5, out5, FOV, b5), args.reps)
    print(json.dumps({"config": 5, "workload": "2M poses x 270 beams (one GPU's share of 16M), 8192^2 map (256 MiB field)",
                      "rays": n5 * b5, "ms_mean": mean, "ms_best": best, "grays_per_s": n5 * b5 / (mean * 1e-3) / 1e9,
                      "ingest_ms": omap5.ingest_ms}))
    del poses5, out5, rm5, omap5

    # config 4: fused rollout, 65536 cars x 50 steps x 1080 beams on a colombia-sized map region
    golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "colombia_map.npz")
    z = np.load(golden)
    path = "/tmp/_rl_cfg_colombia.pgm"
    maps.write_pgm(path, z["img"])
    yc = maps.MapYaml(path, float(z["resolution"]), tuple(float(v) for v in z["origin"]))
    omapc = range_libc.PyOMap(yc)
    distc = omapc.dist()
    rmc = range_libc.PyRayMarchingGPU(omapc, 300)
    car = BatchedCar()
    car.setCarEdgeDistances(1080, -FOV / 2.0, FOV / 1080, 0.275)
    ncars, steps = 65536, 50
    start = maps.sample_free_poses(distc, ncars, 404, yc.resolution, yc.origin, min_clear_px=6.0)
    s0 = np.zeros((ncars, 11))
    s0[:, :3] = start
    s0[:, 3] = 2.0
    s0d = torch.from_numpy(s0).cuda()
    st = s0d.clone()
    res = {}

    def roll():
        st.copy_(s0d)
        res["o"] = car.rollout(rmc, st, None, steps, FOV, seed=42)   # actions drawn on the device (Philox, seed 42)

    mean, best = timeit(roll, max(3, args.reps // 2))
    crash = res["o"]["crash_index"]
    scanned = torch.where(crash >= 0, crash + 1, torch.full_like(crash, steps)).sum().item()   # poses that had to be scanned
    print(json.dumps({"config": 4, "workload": "fused rollout: 65536 cars x 50 steps x 1080 beams, maps/colombia",
                      "nominal_rays": ncars * steps * 1080, "ms_mean": mean, "ms_best": best,
                      "nominal_grays_per_s": ncars * steps * 1080 / (mean * 1e-3) / 1e9,
                      "crashed_frac": float((crash >= 0).float().mean().item()),
                      "poses_needed": int(scanned), "poses_total": ncars * steps}))


if __name__ == "__main__":
    main()
