mkdir -p gpurun_out
K='map_order or territories or fused_allgather'
(timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -k "$K" 2>&1 | tail -12) > gpurun_out/r2B_memcheck.log 2>&1
tail -6 gpurun_out/r2B_memcheck.log
(timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -k "territories_on_small" 2>&1 | tail -12) > gpurun_out/r2B_racecheck.log 2>&1
tail -6 gpurun_out/r2B_racecheck.log
(timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -k "territories_on_small" 2>&1 | tail -12) > gpurun_out/r2B_initcheck.log 2>&1
tail -6 gpurun_out/r2B_initcheck.log
