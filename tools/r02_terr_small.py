"""Sort + territory kernels on a small and a large batch (for an ncu duration pass)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from r02_probe import FOV, config2, with_env  # noqa: E402

omap, y, dist = config2()
angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, 60, endpoint=False).astype(np.float32)).cuda()
rm = with_env({"RL_SORT_POSES": "1", "RL_SORT_MIN_POSES": "1"}, lambda: range_libc.PyRayMarchingGPU(omap, 300))
for n in (2048, 65536, 1_000_000):
    poses = torch.from_numpy(maps.sample_free_poses(dist, n, 303, y.resolution, y.origin)).cuda()
    out = torch.empty(n * 60, dtype=torch.float32, device="cuda")
    for _ in range(2):
        rm.calc_range_repeat_angles(poses, angles, out)
    torch.cuda.synchronize()
    print("done", n, flush=True)
