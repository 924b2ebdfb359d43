timeout 900 python tools/r02_probe.py terr 2>/dev/null | grep -v terr_threshold | grep "default" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], '|', d['variant'], '|', round(d['ms'],4))
"
