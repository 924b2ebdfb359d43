mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r2r_tests.log 2>&1
tail -6 gpurun_out/r2r_tests.log
timeout 900 python bench.py > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err
tail -c 600 gpurun_out/r2r_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k in ('config3','config5','config4','config1'):
    c=d['configs'][k]; print(k,{x:c[x] for x in c if x in('kernel_ms','rays_per_s','us_per_scan')})
PY
