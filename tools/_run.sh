mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/$1.json 2> gpurun_out/$1.err; tail -c 200 gpurun_out/$1.err; }
run r2z_bench_n4
RL_GATHER_MODE=uc run r2z_bench_n4_uc
python - <<'PY'
import json
for f in ('r2z_bench_n4','r2z_bench_n4_uc'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f,'value',d['value']/1e9,'ms',d['ms_per_step'], 'stores_only', d['roofline']['nvlink']['stores_only_ms'], 'e2e', d['e2e']['value']/1e9, 'steady', d['steady_state']['value']/1e9, 'sharded', d['sharded']['value']/1e9, 'nccl', d['gather_nccl']['value']/1e9)
    for k in ('config3','config5','config4'):
        c=d['configs'][k]; print('  ',k,'sharded',c.get('rays_per_s',c.get('nominal_rays_per_s',0))/1e9, 'with_gather',c.get('with_gather',{}).get('rays_per_s',0)/1e9, c.get('with_gather',{}).get('own_slot_check'), c.get('with_gather',{}).get('check'))
PY
