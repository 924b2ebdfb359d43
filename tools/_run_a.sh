set -x
mkdir -p gpurun_out
nvidia-smi -L
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2a_tests.log 2>&1
timeout 600 python tools/r02_probe.py edt steady calib cfg > gpurun_out/r2a_probe.jsonl 2> gpurun_out/r2a_probe.err
timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:march_pose_kernel -c 12 --csv --log-file gpurun_out/r2a_l2_metrics.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
tail -5 gpurun_out/r2a_tests.log
cat gpurun_out/r2a_probe.jsonl
