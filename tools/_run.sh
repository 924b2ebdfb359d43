mkdir -p gpurun_out
cd tools && timeout 300 ncu -k regex:edt --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file ../gpurun_out/r2n_edt_ncu.csv python -c "
import sys, os
sys.path.insert(0, os.path.dirname(os.getcwd()))
import torch
from pyracecarsimulator_b200 import maps, range_libc
img = maps.synth_map(8192, 5678); y = maps.synth_yaml(8192); p='/tmp/_x.pgm'; maps.write_pgm(p, img); y.image = p
om = range_libc.PyOMap(y); torch.cuda.synchronize(); print(om.ingest_ms)
" 2>&1 | tail -2
