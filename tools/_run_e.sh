set -x
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2e_tests.log 2>&1
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
RL_FIELD_PREFETCH=0 timeout 900 python bench.py --steps 100 --warmup 10 --no-configs --no-cpu-baseline > gpurun_out/r2e_bench_n1_noprefetch.json 2> gpurun_out/r2e_bench_n1_noprefetch.err
M=gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:march_pose_kernel -c 14 --csv --log-file gpurun_out/r2e_l2_metrics.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2e_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2e_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_pose_kernel -s 6 -c 2 -o gpurun_out/r2e_march python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2e_ncu3.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:march_pose_kernel -c 12 --csv --log-file gpurun_out/r2e_cfg35_metrics.csv python tools/r02_probe.py cfg > gpurun_out/r2e_ncu4.log 2>&1
tail -6 gpurun_out/r2e_tests.log
python - <<'PY'
import json
for f in ('r2e_bench_n1','r2e_bench_n1_noprefetch'):
    j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    r=j['roofline']
    print(f, j['value']/1e9, j['ms_per_step'], 'kernel cold/warm/pinned', r['kernel_ms'], r['kernel_ms_warm_l2'], r['kernel_ms_flushed_field_pinned'], 'steady', j['steady_state']['value']/1e9, 'e2e', j['e2e']['value']/1e9)
PY
