"""One launch per variant of config 3 (and a 500k-pose slice of config 5) for an ncu metrics pass:
    ncu --metrics ... python tools/r02_terr_ncu.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from r02_probe import FOV, VARIANTS, config2, with_env  # noqa: E402

which = VARIANTS[:2]
omap, y, dist = config2()
n, a = 1_000_000, 60
poses = torch.from_numpy(maps.sample_free_poses(dist, n, 303, y.resolution, y.origin)).cuda()
angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, a, endpoint=False).astype(np.float32)).cuda()
out = torch.empty(n * a, dtype=torch.float32, device="cuda")
rng = np.random.default_rng(9)
c = poses[12345].cpu().numpy()
cloud = np.empty((n, 3), np.float32)
cloud[:, 0] = c[0] + rng.normal(0, 0.5, n)
cloud[:, 1] = c[1] + rng.normal(0, 0.5, n)
cloud[:, 2] = c[2] + rng.normal(0, 0.3, n)
cloud = torch.from_numpy(cloud).cuda()
for name, env in which:
    rm = with_env(env, lambda: range_libc.PyRayMarchingGPU(omap, 300))
    for label, ps in (("cfg3", poses), ("cloud", cloud)):
        rm.calc_range_repeat_angles(ps, angles, out)
        torch.cuda.synchronize()
        print(label, name, flush=True)
    del rm
