mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_fuzz.py tests/test_gpu_round2.py tests/test_gpu_configs.py tests/test_gpu_scan_simulator.py -x -q 2>&1 | tail -4)
timeout 900 python tools/r02_probe.py edt > gpurun_out/r2q_probe_edt.jsonl 2> gpurun_out/r2q_edt.err
python -c "
import json
for l in open('gpurun_out/r2q_probe_edt.jsonl'):
    d=json.loads(l); print(d['map'], {k:round(v,3) for k,v in d.items() if k.endswith('_ms') and 'budget' not in k})
"
