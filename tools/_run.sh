mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err
tail -c 200 gpurun_out/r2v_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2v_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['cpu_baseline']['value']/1e9, d['ingest_ms'], 'counters' in d['configs']['config3'])
"
