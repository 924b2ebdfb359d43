"""Sort + territory kernels on uniform poses and on a particle cloud, per cell size (for an ncu duration pass)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from r02_probe import FOV, config2, with_env  # noqa: E402

omap, y, dist = config2()
angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, 60, endpoint=False).astype(np.float32)).cuda()
n = 1_000_000
uniform = maps.sample_free_poses(dist, n, 303, y.resolution, y.origin)
rng = np.random.default_rng(9)
c = uniform[12345]
cloud = np.empty((n, 3), np.float32)
cloud[:, 0] = c[0] + rng.normal(0, 0.5, n)
cloud[:, 1] = c[1] + rng.normal(0, 0.5, n)
cloud[:, 2] = c[2] + rng.normal(0, 0.3, n)
out = torch.empty(n * 60, dtype=torch.float32, device="cuda")
for shift in ("2", "4", "6"):
    rm = with_env({"RL_SORT_POSES": "1", "RL_SORT_SHIFT": shift}, lambda: range_libc.PyRayMarchingGPU(omap, 300))
    for label, ps in (("uniform", uniform), ("cloud", cloud)):
        poses = torch.from_numpy(ps).cuda()
        rm.calc_range_repeat_angles(poses, angles, out)
        torch.cuda.synchronize()
        print("done", shift, label, flush=True)
    del rm
