"""BASELINE config 5 on ONE GPU in ONE call: 16 M poses x 270 beams = 4.32 G rays (more than 2^31: the 64-bit index
variants of the kernels), by territories and in the caller's order; the two 17.3 GB results must be identical."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import _native, maps, range_libc  # noqa: E402
from r02_probe import FOV, timeit  # noqa: E402

img = maps.synth_map(8192, 5678)
y = maps.synth_yaml(8192)
path = f"/tmp/_rl_cfg5_{os.getpid()}.pgm"
maps.write_pgm(path, img)
y.image = path
omap = range_libc.PyOMap(y)
os.unlink(path)
n, b = 16_000_000, 270
poses = torch.from_numpy(maps.sample_free_poses(omap.dist(), n, 505, y.resolution, y.origin)).cuda()
terr = range_libc.PyRayMarchingGPU(omap, 300)
plain = range_libc.PyRayMarchingGPU(omap, 300, flags=_native.RL_FLAG_NO_POSE_SORT)
a = torch.empty(n * b, dtype=torch.float32, device="cuda")
c = torch.empty(n * b, dtype=torch.float32, device="cuda")
ms_t = timeit(lambda: terr.calc_range_fan(poses, a, FOV, b), reps=3)
ms_p = timeit(lambda: plain.calc_range_fan(poses, c, FOV, b), reps=3)
same = bool(torch.equal(a, c))
# and against the same batch marched in pieces of 2 M poses (32-bit index kernels)
c.fill_(-1.0)
for lo in range(0, n, 2_000_000):
    plain.calc_range_fan(poses[lo:lo + 2_000_000], c[lo * b:(lo + 2_000_000) * b], FOV, b)
torch.cuda.synchronize()
print(json.dumps({"probe": "cfg5_one_call", "poses": n, "beams": b, "rays": n * b, "territories_ms": ms_t, "caller_order_ms": ms_p,
                  "territories_grays_per_s": n * b / ms_t / 1e6, "caller_order_grays_per_s": n * b / ms_p / 1e6,
                  "identical": same, "identical_to_pieces": bool(torch.equal(a, c))}), flush=True)
