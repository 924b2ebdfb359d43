mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "map_order or territories" 2>&1 | tail -4)
timeout 900 python tools/r02_probe.py terr 2> gpurun_out/r2t_terr.err | grep -v terr_threshold > gpurun_out/r2t_terr.jsonl
cat gpurun_out/r2t_terr.jsonl
cd tools
timeout 600 ncu -k regex:"sort|territory" --metrics gpu__time_duration.sum --clock-control none --csv --log-file ../gpurun_out/r2t_ncu.csv python r02_terr_small.py > ../gpurun_out/r2t.log 2>&1
