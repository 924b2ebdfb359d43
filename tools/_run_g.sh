set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30) > gpurun_out/r2g_tests_multi.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/r2g_bench_n4.json 2> gpurun_out/r2g_bench_n4.err
tail -4 gpurun_out/r2g_tests_multi.log
for f in n8 n4; do python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2g_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', 'value', j['value']/1e9, j['ms_per_step'], 'sharded', j['sharded']['ms_per_step'], 'nccl', j['gather_nccl']['ms_per_step'], 'e2e', j['e2e']['value']/1e9, j.get('gather_check'))
    print('   nvlink', j['roofline']['nvlink'])
    print('   steady', j['steady_state']['value']/1e9)
    for k,v in j['configs'].items():
        if isinstance(v,dict): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('kernel_ms','rays_per_s','nominal_rays_per_s','us_per_scan','with_gather')})
except Exception as e:
    print('$f', 'ERR', e)
PY
done
tail -3 gpurun_out/r2g_bench_n8.err
