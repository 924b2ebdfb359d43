mkdir -p gpurun_out
cd tools
timeout 600 python r02_cfg2_terr_ncu.py 2>&1 | grep "single"
M=gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__grid_size
timeout 600 ncu -k regex:"territory" -c 2 --metrics $M --clock-control none --csv --log-file ../gpurun_out/r2p_cfg2_terr_ncu.csv python r02_cfg2_terr_ncu.py > /dev/null 2>&1
