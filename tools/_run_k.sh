set -x
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r2k_tests.log 2>&1
M=gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:march_pose_kernel -c 14 --csv --log-file gpurun_out/r2k_l2_metrics.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2k_ncu1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"sector_kernel|gather_kernel" -c 4 --csv --log-file gpurun_out/r2k_calibration_metrics.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2k_ncu1b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2k_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_pose_kernel -s 6 -c 2 -o gpurun_out/r2k_march python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2k_ncu3.log 2>&1
(cd tools && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"territory|pose_sort" -c 2 -o ../gpurun_out/r2k_territory python r02_terr_ncu.py > ../gpurun_out/r2k_ncu4.log 2>&1)
timeout 900 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err
timeout 600 python tools/r02_probe.py edt > gpurun_out/r2k_probe_edt.jsonl 2> gpurun_out/r2k_probe_edt.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1
(timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_march.py tests/test_gpu_car.py -x -q -k "not full_size and not sparse_large and not l2_carve" 2>&1 | tail -15) > gpurun_out/r2k_memcheck.log 2>&1
tail -5 gpurun_out/r2k_tests.log
cat gpurun_out/r2k_smoke.log
tail -8 gpurun_out/r2k_memcheck.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2k_bench_n1.json').read().strip().splitlines()[-1])
r=j['roofline']
print('value', j['value']/1e9, j['ms_per_step'], 'kernel cold/warm/pinned', r['kernel_ms'], r['kernel_ms_warm_l2'], r['kernel_ms_flushed_field_pinned'], 'steady', j['steady_state']['value']/1e9, 'e2e', j['e2e']['value']/1e9)
print(r['at_steady_state'])
for k in ('config1','config3','config4','config5'):
    c=j['configs'][k]; print(k,{x:c[x] for x in c if x in('kernel_ms','rays_per_s','us_per_scan','nominal_rays_per_s')})
jr=json.loads(open('gpurun_out/r2k_bench_ref.json').read().strip().splitlines()[-1])
print('ref', jr['value']/1e9, jr['ms_per_step'], jr['cpu_baseline'])
PY
