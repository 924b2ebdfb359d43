"""Config 2 through the plain kernel and through the territory kernel in the caller's order (no sort)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from r02_probe import FOV, config2, timeit, with_env  # noqa: E402

omap, y, dist = config2()
p2 = [torch.from_numpy(maps.sample_free_poses(dist, 4096, 1000 + i, y.resolution, y.origin)).cuda() for i in range(4)]
out = [torch.empty(4096 * 1080, dtype=torch.float32, device="cuda") for _ in range(4)]
for name, env in (("plain", {"RL_SORT_POSES": "0"}),
                  ("territories, caller order", {"RL_SORT_POSES": "1", "RL_SORT_MIN_POSES": "1", "RL_TERRITORY_IDENTITY": "1"})):
    rm = with_env(env, lambda: range_libc.PyRayMarchingGPU(omap, 300))
    single = timeit(lambda: rm.calc_range_fan(p2[0], out[0], FOV, 1080), reps=20)
    many = timeit(lambda: [rm.calc_range_fan(p2[i & 3], out[i & 3], FOV, 1080) for i in range(50)], reps=5) / 50
    rm.set_pipelined(1)
    piped = timeit(lambda: ([rm.calc_range_fan(p2[i & 3], out[i & 3], FOV, 1080) for i in range(50)], rm.join()), reps=5) / 50
    rm.set_pipelined(0)
    print(name, "single %.4f ms, back to back %.4f ms, pipelined %.4f ms" % (single, many, piped), flush=True)
    del rm
