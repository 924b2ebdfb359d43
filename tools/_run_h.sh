set -x
mkdir -p gpurun_out
timeout 900 python tools/e2e_probe.py > gpurun_out/r2h_e2e_probe.txt 2>&1
timeout 300 python tools/e2e_probe.py pageable >> gpurun_out/r2h_e2e_probe.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q -k "scan_simulator or march or registry or growing or host" 2>&1 | tail -8) > gpurun_out/r2h_tests.log 2>&1
cat gpurun_out/r2h_e2e_probe.txt
tail -4 gpurun_out/r2h_tests.log
