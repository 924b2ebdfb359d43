set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2b_tests.log 2>&1
timeout 600 python tools/r02_probe.py edt > gpurun_out/r2b_probe.jsonl 2> gpurun_out/r2b_probe.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
cat > /tmp/edt_prof.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from pyracecarsimulator_b200 import maps, range_libc
import oracle
g = oracle.mapserver_occupancy(maps.synth_map(8192, 5678))
msg = maps.OccupancyGrid.make(g.ravel(), 8192, 8192, 0.05, (0.0, 0.0, 0.0))
lone = np.zeros((8192, 8192), bool); lone[2730, 5461] = True
for which in ("scan", "dc"):
    os.environ["RL_EDT_ROWS"] = which
    range_libc.PyOMap(msg)
os.environ["RL_EDT_ROWS"] = "dc"
range_libc.PyOMap(lone)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_ -c 16 -o gpurun_out/r2b_edt python /tmp/edt_prof.py > gpurun_out/r2b_edt_ncu.log 2>&1
tail -5 gpurun_out/r2b_tests.log
cat gpurun_out/r2b_probe.jsonl
cat gpurun_out/r2b_bench.json
tail -5 gpurun_out/r2b_bench.err
cat gpurun_out/r2b_bench_ref.json
