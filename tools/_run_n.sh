set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:march_pose_kernel -c 14 --csv --log-file gpurun_out/r2n_l2_metrics.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2n_ncu1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2n_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:march_pose_kernel -s 6 -c 2 -o gpurun_out/r2n_march python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2n_ncu3.log 2>&1
ls -la gpurun_out/ | tail -5
