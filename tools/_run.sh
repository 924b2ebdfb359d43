mkdir -p gpurun_out
(timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_fuzz.py -x -q -k "not large_map and not empty_large" 2>&1 | tail -8) > gpurun_out/r2w_memcheck_ingest.log 2>&1
tail -5 gpurun_out/r2w_memcheck_ingest.log
(timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ingest.py -x -q -k "edge_case or synthetic_maps" 2>&1 | tail -8) > gpurun_out/r2w_racecheck_ingest.log 2>&1
tail -4 gpurun_out/r2w_racecheck_ingest.log
