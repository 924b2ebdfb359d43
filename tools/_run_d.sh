set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1
lscpu | grep -i "numa\|model name\|^CPU(s)\|socket" >> gpurun_out/r2d_topo.txt 2>&1
(timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30) > gpurun_out/r2d_tests_multi.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2d_bench_n8.json 2> gpurun_out/r2d_bench_n8.err
RL_GATHER_MODE=mc_weak timeout 300 $TR --master-port 29552 bench.py --gpus 8 --steps 100 --warmup 10 --no-configs > gpurun_out/r2d_bench_n8_weak.json 2> gpurun_out/r2d_bench_n8_weak.err
RL_GATHER_MODE=uc timeout 300 $TR --master-port 29553 bench.py --gpus 8 --steps 100 --warmup 10 --no-configs > gpurun_out/r2d_bench_n8_uc.json 2> gpurun_out/r2d_bench_n8_uc.err
RL_GATHER_VEC=0 timeout 300 $TR --master-port 29554 bench.py --gpus 8 --steps 100 --warmup 10 --no-configs > gpurun_out/r2d_bench_n8_scalar.json 2> gpurun_out/r2d_bench_n8_scalar.err
RL_BIND=1 timeout 300 $TR --master-port 29555 tools/d2h_probe.py > gpurun_out/r2d_d2h_bind1.jsonl 2> gpurun_out/r2d_d2h_bind1.err
RL_BIND=0 timeout 300 $TR --master-port 29556 tools/d2h_probe.py > gpurun_out/r2d_d2h_bind0.jsonl 2> gpurun_out/r2d_d2h_bind0.err
tail -4 gpurun_out/r2d_tests_multi.log
for f in n8 n8_weak n8_uc n8_scalar; do python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2d_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', 'value', j['value']/1e9, j['ms_per_step'], 'sharded', j['sharded']['ms_per_step'], 'nccl', j['gather_nccl']['ms_per_step'], 'e2e', j['e2e']['value']/1e9, j.get('gather_check'))
except Exception as e:
    print('$f', 'ERR', e)
PY
done
cat gpurun_out/r2d_d2h_bind1.jsonl gpurun_out/r2d_d2h_bind0.jsonl | cut -c1-600
tail -3 gpurun_out/r2d_bench_n8.err
