"""Writes the inputs of tools/tune_march (distance field from the product's GPU ingest, seeded
poses, marcher constants) to a directory.  Development helper, not on the product path."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/tune"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2049
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1234
num_poses = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
num_beams = int(sys.argv[5]) if len(sys.argv) > 5 else 1080
os.makedirs(out, exist_ok=True)
img = maps.synth_map(n, seed)
y = maps.synth_yaml(n)
path = os.path.join(out, "map.pgm")
maps.write_pgm(path, img)
y.image = path
omap = range_libc.PyOMap(y)
dist = omap.dist()
poses = maps.sample_free_poses(dist, num_poses, 1000, y.resolution, y.origin)
dist.tofile(os.path.join(out, "dist.bin"))
poses.tofile(os.path.join(out, "poses.bin"))
scale = np.float32(y.resolution)
angle = -0.0
world = [scale, np.float32(angle), np.float32(y.origin[0]), np.float32(y.origin[1]), np.float32(np.sin(angle)),
         np.float32(np.cos(angle)), np.float32(1.0 / float(scale)), np.float32(-1.0 * float(np.float32(angle)) - 3.0 * np.pi / 2.0)]
with open(os.path.join(out, "meta.bin"), "wb") as f:
    f.write(struct.pack("<4i", n, n, num_poses, num_beams))
    f.write(struct.pack("<2f", 300.0, 4.71))
    f.write(struct.pack("<8f", *[float(v) for v in world]))
print("wrote", out, dist.shape, poses.shape)
