"""One fused rollout of BASELINE config 4 (65 536 cars x 50 steps x 1080 beams, maps/colombia) for an ncu metrics pass."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from pyracecarsimulator_b200.racecar import BatchedCar  # noqa: E402

FOV = 4.71
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "colombia_map.npz"))
path = "/tmp/_rl_cfg4_ncu.pgm"
maps.write_pgm(path, z["img"])
yc = maps.MapYaml(path, float(z["resolution"]), tuple(float(v) for v in z["origin"]))
omapc = range_libc.PyOMap(yc)
rmc = range_libc.PyRayMarchingGPU(omapc, 300)
car = BatchedCar()
car.setCarEdgeDistances(1080, -FOV / 2.0, FOV / 1080, 0.275)
ncars, steps = 65536, 50
s0 = np.zeros((ncars, 11))
s0[:, :3] = maps.sample_free_poses(omapc.dist(), ncars, 404, yc.resolution, yc.origin, min_clear_px=6.0)
s0[:, 3] = 2.0
st = torch.from_numpy(s0).cuda()
car.rollout(rmc, st, None, steps, FOV, seed=42)
torch.cuda.synchronize()
print("done")
