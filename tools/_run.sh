mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2C_tests.log 2>&1
tail -5 gpurun_out/r2C_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2C_bench_n1.json 2> gpurun_out/r2C_bench_n1.err
tail -c 300 gpurun_out/r2C_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2C_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value']/1e9,'ms',d['ms_per_step'],'kernel',r['kernel_ms'],'warm',r['kernel_ms_warm_l2'],'pinned',r['kernel_ms_flushed_field_pinned'],'steady',d['steady_state']['ms_per_launch'],'so',d['steady_state']['stream_order']['ms_per_launch'],'e2e',d['e2e']['value']/1e9, 'crash', d['e2e_fused_crash']['ms_per_call'])
for k in ('config1','config3','config4','config5'):
    c=d['configs'][k]; print(k,{x:c[x] for x in c if x in('kernel_ms','rays_per_s','us_per_scan','nominal_rays_per_s')})
PY
