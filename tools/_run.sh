cd tools; timeout 600 python r02_cfg2_terr_ncu.py 2>&1 | grep "^plain"
cd ..; timeout 900 python tools/r02_probe.py terr 2>/dev/null | grep -v terr_threshold | grep "default\|caller order" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], '|', d['variant'], '|', round(d['ms'],4))
"
timeout 600 python tools/bench_configs.py --reps 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('config', d['config'], round(d['ms_mean'],3))
"
