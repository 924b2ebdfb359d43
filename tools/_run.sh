(timeout 900 python -m pytest tests/test_gpu_scan_simulator.py tests/test_gpu_march.py -x -q 2>&1 | tail -3)
python - <<'PY'
import os, time, numpy as np
from pyracecarsimulator_b200 import maps, range_libc
from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D
z = np.load("tests/golden/colombia_map.npz")
maps.write_pgm("/tmp/_c.pgm", z["img"])
yc = maps.MapYaml("/tmp/_c.pgm", float(z["resolution"]), tuple(float(v) for v in z["origin"]))
omap = range_libc.PyOMap(yc)
sim = ScanSimulator2D(1080, 4.71, 0.01, batch_size=200)
sim.setMap(omap, 300, yc.resolution, yc.origin); sim.setRaytracingMethod("RMGPU")
for rep in range(3):
    for _ in range(50): sim.scan(0.275, 0.0, 0.0)
    t0 = time.perf_counter()
    for _ in range(1000): sim.scan(0.275, 0.0, 0.0)
    print("us per scan", (time.perf_counter() - t0) / 1000 * 1e6)
PY
