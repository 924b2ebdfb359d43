mkdir -p gpurun_out
timeout 600 python tools/r02_gate_probe.py 2>/dev/null | grep gate | tee gpurun_out/r2t_gate.jsonl
