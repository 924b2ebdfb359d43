"""profiles/traffic_configs.json from profiles/r02_territories_ncu.csv (tools/r02_terr_ncu.py under ncu): per-launch
counters of configs 3 and 5 (one GPU's share: 2 M poses) in the caller's order and by territories.  bench.py reads it
to put the issue / SM->L2 request / DRAM views next to the measured times of those configs."""
import collections
import csv
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_territories_ncu.csv"
rows = list(csv.reader(open(src)))
hdr = next(r for r in rows if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows:
    if not r or not r[0].isdigit():
        continue
    k = int(r[ix["ID"]])
    per.setdefault(k, {"kernel": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
launches = list(per.values())
# order of tools/r02_terr_ncu.py: cfg3 plain, cfg3 sort, cfg3 territory, cfg5 plain, cfg5 sort, cfg5 territory
names = ["config3.caller_order", "config3.sort", "config3.territories", "config5_share.caller_order", "config5_share.sort",
         "config5_share.territories"]
assert len(launches) == 6, len(launches)
out = {"source": src + " (ncu --metrics ... -k regex:'march|sort' python tools/r02_terr_ncu.py, --clock-control none)"}
for name, m in zip(names, launches):
    cfg, what = name.split(".")
    out.setdefault(cfg, {})[what] = {
        "kernel": m["kernel"].split("(")[0].replace("<unnamed>::", "").replace("void ", ""),
        "ncu_time_ms": m["gpu__time_duration.sum"] / 1e6,
        "warp_insts": m["smsp__inst_executed.sum"],
        "l1_sectors_requested": m["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"],
        "l1_sector_hits": m["l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum"],
        "l2_read_sectors_from_sm": m["lts__t_sectors_srcunit_tex_op_read.sum"],
        "l2_read_sector_hits": m["lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"],
        "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_written": m["dram__bytes_write.sum"],
        "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "sm_to_l2_request_path_busy_pct": m["l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed"],
    }
json.dump(out, open("profiles/traffic_configs.json", "w"), indent=1)
print(json.dumps(out, indent=1)[:1200])
