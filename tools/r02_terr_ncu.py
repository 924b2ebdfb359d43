"""One launch per order (caller's / map order by SM territories) of config 3 and of one GPU's share of config 5, for an
ncu metrics pass:   cd tools && ncu -k regex:"march|sort" --metrics ... python r02_terr_ncu.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from r02_probe import FOV, config2, with_env  # noqa: E402

ORDERS = (("caller order", {"RL_SORT_POSES": "0"}), ("territories (product default)", {}))
omap, y, dist = config2()
n, a = 1_000_000, 60
poses = torch.from_numpy(maps.sample_free_poses(dist, n, 303, y.resolution, y.origin)).cuda()
angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, a, endpoint=False).astype(np.float32)).cuda()
out = torch.empty(n * a, dtype=torch.float32, device="cuda")
for name, env in ORDERS:
    rm = with_env(env, lambda: range_libc.PyRayMarchingGPU(omap, 300))
    rm.calc_range_repeat_angles(poses, angles, out)
    torch.cuda.synchronize()
    print("cfg3", name, flush=True)
    del rm
del omap, poses, out
img = maps.synth_map(8192, 5678)
y5 = maps.synth_yaml(8192)
path = f"/tmp/_rl_ncu5_{os.getpid()}.pgm"
maps.write_pgm(path, img)
y5.image = path
omap5 = range_libc.PyOMap(y5)
os.unlink(path)
n5, b5 = 2_000_000, 270
poses5 = torch.from_numpy(maps.sample_free_poses(omap5.dist(), n5, 505, y5.resolution, y5.origin)).cuda()
out5 = torch.empty(n5 * b5, dtype=torch.float32, device="cuda")
for name, env in ORDERS:
    rm5 = with_env(env, lambda: range_libc.PyRayMarchingGPU(omap5, 300))
    rm5.calc_range_fan(poses5, out5, FOV, b5)
    torch.cuda.synchronize()
    print("cfg5 share", name, flush=True)
    del rm5
