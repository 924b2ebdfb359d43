set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2m_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
tail -4 gpurun_out/r2m_tests.log; cat gpurun_out/r2m_smoke.log; tail -c 300 gpurun_out/r2m_bench_n1.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2m_bench_n1.json').read().strip().splitlines()[-1])
r=j['roofline']
print("value", j["value"]/1e9, j["ms_per_step"], "steady", j["steady_state"]["value"]/1e9, "e2e", j["e2e"]["value"]/1e9)
for k in ("config1","config3","config4","config5"):
    c=j["configs"][k]; print(k,{x:c[x] for x in c if x in("kernel_ms","rays_per_s","us_per_scan","nominal_rays_per_s","ms")})
PY
