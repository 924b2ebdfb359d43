"""profiles/traffic.json from an ncu metrics CSV of the shipped march kernel:
    ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_read.sum,\
lts__t_sectors_srcunit_tex_op_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,\
l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,\
dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:march_pose_kernel \
-c 12 --csv --log-file gpurun_out/X.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs
    python tools/make_traffic.py profiles/r02_l2_metrics.csv
Only launches of the plain kernel (march_pose_kernel<1, 0, 1, 0, 1>: FAN, no step counter, 32-bit index, local output, padded field)
over the full 4096 x 1080 batch are averaged.  bench.py reads the result for roofline.traffic / .l2 / .issue."""
import collections
import csv
import json
import sys

src = sys.argv[1]
rows = list(csv.reader(open(src)))
hdr = next(r for r in rows if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(dict)
meta = {}
for r in rows:
    if not r or not r[0].isdigit():
        continue
    k = int(r[ix["ID"]])
    per[k][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    meta[k] = (r[ix["Kernel Name"]], r[ix["Grid Size"]])
keep = [k for k in per if "march_pose_kernel<1, 0, 1, 0" in meta[k][0] and meta[k][1].replace(" ", "").startswith("(34560,")]
if not keep:
    keep = [k for k in per if "march_pose_kernel<1, 0, 1, 0" in meta[k][0]]
# the end-to-end scanMany launches store their ranges into the caller's pinned HOST buffer (no L2 write sectors,
# PCIe-bound duration): only the device-output launches describe the kernel
dev = [k for k in keep if per[k].get("lts__t_sectors_srcunit_tex_op_write.sum", 0) > 0.9 * 4096 * 1080 * 4 / 32]
keep = dev or keep


def avg(name):
    v = [per[k][name] for k in keep if name in per[k]]
    return sum(v) / len(v) if v else None


out = {
    "kernel": "march_pose_kernel<FAN=1, COUNT=0, SMALL=1, OUT_LOCAL, PADDED=1> (4096 poses x 1080 beams, 2049^2 field)",
    "launches": len(keep),
    "dram_bytes_per_launch": (avg("dram__bytes_read.sum") or 0) + (avg("dram__bytes_write.sum") or 0),
    "warp_insts_per_launch": avg("smsp__inst_executed.sum"),
    "lts_bytes_per_launch": avg("lts__t_bytes.sum"),
    "lts_sectors_per_launch": avg("lts__t_sectors.sum"),
    "lts_tex_read_sectors_per_launch": avg("lts__t_sectors_srcunit_tex_op_read.sum"),
    "lts_tex_write_sectors_per_launch": avg("lts__t_sectors_srcunit_tex_op_write.sum"),
    "l1_global_load_sectors_per_launch": avg("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
    "l1_global_load_sector_hits_per_launch": avg("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum"),
    "l1_global_load_requests_per_launch": avg("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
    "ncu_time_ns_per_launch": avg("gpu__time_duration.sum"),
    "source": f"{src}: ncu --metrics ... --clock-control none -k regex:march_pose_kernel on bench.py (command in tools/make_traffic.py); "
              "per-launch mean over the plain full-batch launches",
}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
