mkdir -p gpurun_out
for s in 1 0.1 0.15 0.2 0.25 0.3 0.4 0.6; do RL_HOST_SPLIT=$s timeout 300 python tools/r02_split_probe.py 2>/dev/null | grep host_split; done | tee gpurun_out/r2r_split.jsonl
for s in 1 0.2; do POSES=1024 RL_HOST_SPLIT=$s timeout 300 python tools/r02_split_probe.py 2>/dev/null | grep host_split; done | tee -a gpurun_out/r2r_split.jsonl
for s in 1 0.2; do POSES=16384 RL_HOST_SPLIT=$s timeout 300 python tools/r02_split_probe.py 2>/dev/null | grep host_split; done | tee -a gpurun_out/r2r_split.jsonl
