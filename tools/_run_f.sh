set -x
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2f_tests.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_configs.py -q -s -k "crash_indices" 2>&1 | grep -i "config 4\|passed\|failed") > gpurun_out/r2f_crash_count.log 2>&1
timeout 600 python tools/r02_probe.py steady cfg > gpurun_out/r2f_probe.jsonl 2> gpurun_out/r2f_probe.err
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -6 gpurun_out/r2f_tests.log
cat gpurun_out/r2f_crash_count.log
cat gpurun_out/r2f_probe.jsonl | cut -c1-400
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2f_bench_n1.json').read().strip().splitlines()[-1])
r=j['roofline']
print('value', j['value']/1e9, j['ms_per_step'], 'kernel cold/warm/pinned', r['kernel_ms'], r['kernel_ms_warm_l2'], r['kernel_ms_flushed_field_pinned'], 'steady', j['steady_state']['value']/1e9, j['steady_state']['stream_order']['value']/1e9, 'e2e', j['e2e']['value']/1e9, j['e2e']['pinned_d2h_gbs'], 'pageable', j['e2e']['numpy_pageable_buffers']['ms_per_call'], 'fused crash', j['e2e_fused_crash']['ms_per_call'])
for k,v in j['configs'].items():
    if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('kernel_ms','rays_per_s','nominal_rays_per_s','us_per_scan','ingest_ms')})
PY
