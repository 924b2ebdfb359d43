mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_march.py -x -q 2>&1 | tail -6) > gpurun_out/r2o_tests.log 2>&1
timeout 900 python tools/r02_probe.py terr > gpurun_out/r2o_terr.jsonl 2> gpurun_out/r2o_terr.err
tail -6 gpurun_out/r2o_tests.log
cat gpurun_out/r2o_terr.jsonl
tail -3 gpurun_out/r2o_terr.err
M=gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,launch__grid_size,launch__registers_per_thread
cd tools
timeout 900 ncu -k regex:"march|sort" --metrics $M --clock-control none --csv --log-file ../gpurun_out/r2o_terr_ncu.csv python r02_terr_ncu.py > ../gpurun_out/r2o_terr_ncu.log 2>&1
tail -3 ../gpurun_out/r2o_terr_ncu.log
