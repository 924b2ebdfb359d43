mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_fuzz.py tests/test_gpu_round2.py -x -q 2>&1 | tail -4)
timeout 900 python tools/r02_probe.py edt 2> gpurun_out/r2m_edt.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['map'], {k:round(v,3) for k,v in d.items() if k.endswith('_ms') and 'budget' not in k})
"
tail -2 gpurun_out/r2m_edt.err
cd tools && timeout 300 ncu -k regex:edt --metrics gpu__time_duration.sum --clock-control none --csv --log-file ../gpurun_out/r2m_edt_ncu.csv python -c "
import sys, os
sys.path.insert(0, os.path.dirname(os.getcwd()))
from r02_probe import config2
import torch
config2(); torch.cuda.synchronize()
" > /dev/null 2>&1; grep -v "^==" ../gpurun_out/r2m_edt_ncu.csv | cut -d, -f5,15 | cut -c1-50,60- | tail -6
