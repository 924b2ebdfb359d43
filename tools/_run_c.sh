set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m > gpurun_out/r2c_topo.txt 2>&1
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2c_tests.log 2>&1
timeout 600 python tools/r02_probe.py edt steady > gpurun_out/r2c_probe.jsonl 2> gpurun_out/r2c_probe.err
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
tail -5 gpurun_out/r2c_tests.log
cat gpurun_out/r2c_probe.jsonl
cat gpurun_out/r2c_bench_n1.json
tail -3 gpurun_out/r2c_bench_n1.err
cat gpurun_out/r2c_bench_n2.json
tail -20 gpurun_out/r2c_bench_n2.err
