mkdir -p gpurun_out
timeout 900 python tools/r02_cfg5_one_call.py 2> gpurun_out/r2s_cfg5.err | tee gpurun_out/r2s_cfg5_one_call.jsonl
tail -3 gpurun_out/r2s_cfg5.err
