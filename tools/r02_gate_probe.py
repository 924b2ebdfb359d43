"""Where map order stops paying on an L2-resident field: N poses x 60 angles, caller's order vs territories (forced)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from r02_probe import FOV, config2, timeit, with_env  # noqa: E402

omap, y, dist = config2()
angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, 60, endpoint=False).astype(np.float32)).cuda()
plain = with_env({"RL_SORT_POSES": "0"}, lambda: range_libc.PyRayMarchingGPU(omap, 300))
terr = with_env({"RL_SORT_POSES": "1", "RL_SORT_MIN_POSES": "1"}, lambda: range_libc.PyRayMarchingGPU(omap, 300))
for n in (62_500, 125_000, 250_000, 500_000, 1_000_000, 2_000_000):
    poses = torch.from_numpy(maps.sample_free_poses(dist, n, 303 + n, y.resolution, y.origin)).cuda()
    out = torch.empty(n * 60, dtype=torch.float32, device="cuda")
    a = timeit(lambda: plain.calc_range_repeat_angles(poses, angles, out), reps=9)
    b = timeit(lambda: terr.calc_range_repeat_angles(poses, angles, out), reps=9)
    print(json.dumps({"probe": "gate", "poses": n, "rays": n * 60, "poses_per_cell": n / (2049 * 2049), "caller_order_ms": a,
                      "territories_ms": b, "gain": a / b - 1.0}), flush=True)
