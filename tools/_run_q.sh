set -x
mkdir -p gpurun_out
N=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2q_bench_n$N.json 2> gpurun_out/r2q_bench_n$N.err; tail -c 300 gpurun_out/r2q_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_bench_n$N.json').read().strip().splitlines()[-1])
print('value',d['value']/1e9,'ms',d['ms_per_step'], 'stores_only', d['roofline']['nvlink']['stores_only_ms'], 'e2e', d['e2e']['value']/1e9, 'steady', d['steady_state']['value']/1e9, 'sharded', d['sharded']['value']/1e9, d['sharded']['ms_per_step'], 'nccl', d['gather_nccl']['value']/1e9, d['gather_nccl']['ms_per_step'], d.get('gather_check'))
for k in ('config3','config5','config4'):
    c=d['configs'][k]; print('  ',k,'sharded',c.get('rays_per_s',c.get('nominal_rays_per_s',0))/1e9, 'with_gather',c.get('with_gather',{}).get('rays_per_s',0)/1e9, c.get('with_gather',{}).get('own_slot_check'), c.get('with_gather',{}).get('check'))
PY
