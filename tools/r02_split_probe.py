"""End-to-end scanMany (host poses in, host ranges out) for one RL_HOST_SPLIT value (read once per process):
    for s in 1 0.1 0.2 0.3; do RL_HOST_SPLIT=$s python tools/r02_split_probe.py; done"""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D  # noqa: E402

img = maps.synth_map(2049, 1234)
y = maps.synth_yaml(2049)
path = f"/tmp/_rl_split_{os.getpid()}.pgm"
maps.write_pgm(path, img)
y.image = path
omap = range_libc.PyOMap(y)
os.unlink(path)
B = int(os.environ.get("POSES", "4096"))
sim = ScanSimulator2D(1080, 4.71, 0.01, batch_size=B)
sim.setMap(omap, 300, y.resolution, y.origin)
sim.setRaytracingMethod("RMGPU")
poses = [maps.sample_free_poses(omap.dist(), B, 1000 + i, y.resolution, y.origin) for i in range(4)]
for i in range(8):
    out = sim.scanMany(poses[i & 3])
digest = hashlib.sha1(np.ascontiguousarray(sim.scanMany(poses[0])).tobytes()).hexdigest()[:12]
ts = []
for rep in range(5):
    t0 = time.perf_counter()
    for i in range(40):
        sim.scanMany(poses[i & 3])
    ts.append((time.perf_counter() - t0) / 40 * 1e3)
print(json.dumps({"probe": "host_split", "split": os.environ.get("RL_HOST_SPLIT", "default"), "poses": B,
                  "ms_per_call": float(np.median(ts)), "grays_per_s": B * 1080 / np.median(ts) / 1e6, "sha1_batch0": digest}), flush=True)
