set -x
mkdir -p gpurun_out
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_march.py tests/test_gpu_round2.py tests/test_gpu_car.py -x -q -k "not full_size and not sparse_large and not l2_carve" 2>&1 | tail -12) > gpurun_out/r2p_memcheck.log 2>&1
(timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_march.py -x -q -k "not full_size and not concurrent" 2>&1 | tail -12) > gpurun_out/r2p_racecheck.log 2>&1
tail -6 gpurun_out/r2p_memcheck.log; tail -6 gpurun_out/r2p_racecheck.log
