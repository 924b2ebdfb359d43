"""Turn the ncu artefacts of a round into the tracked summaries under profiles/.
    python tools/make_profiles.py gpurun_out/r01_prof.ncu-rep gpurun_out/r01_launches.csv r01
Writes profiles/<tag>_ncu_full_summary.md and profiles/<tag>_launches.csv + _launches_summary.md.
(profiles/traffic.json, which bench.py reads, comes from tools/make_traffic.py.)"""
import collections
import csv
import json
import shutil
import subprocess
import sys

rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
cols = ['kernel'] + [w for w in want if w in idx]
out = [f"# {tag} - ncu `--set full --clock-control none` summary (one row per captured launch)", "",
       "Command: `ncu --set full --clock-control none --import-source on -k regex:march_pose_kernel -s 6 -c 2 python bench.py "
       "--steps 4 --warmup 3 --no-cpu-baseline --no-configs` (B200, one GPU; times under ncu are cold-cache and serialised: compare "
       "shares, not absolutes).", "", '| ' + ' | '.join(cols) + ' |', '|' + '---|' * len(cols)]


def num(r, k):
    v = float(r[idx[k]].replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[idx[k]], 1)


march = []
for r in data:
    name = r[idx['Kernel Name']].split('(')[0].replace('<unnamed>::', '').replace('void ', '')
    out.append('| ' + ' | '.join([name] + [r[idx[w]] + ' ' + units[idx[w]] for w in want if w in idx]) + ' |')
    if 'march_pose' in name:
        march.append((num(r, 'dram__bytes_read.sum') + num(r, 'dram__bytes_write.sum'), num(r, 'smsp__inst_executed.sum')))
open(f'profiles/{tag}_ncu_full_summary.md', 'w').write('\n'.join(out) + '\n')
shutil.copy(launches, f'profiles/{tag}_launches.csv')
lr = [r for r in csv.reader(open(launches)) if r and r[0].isdigit()]
agg = collections.OrderedDict()
names = [r[4].split('(')[0].replace('<unnamed>::', '').replace('void ', '')[:70] for r in lr]
plain = sorted(float(r[-1]) for r, n in zip(lr, names) if n.startswith('march_pose_kernel<1, 0'))
med = plain[len(plain) // 2] if plain else 0.0
for r, n in zip(lr, names):
    # the end-to-end calls store ranges straight into the caller's page-locked HOST buffer (zero-copy): those
    # launches last as long as 17.7 MB take to cross PCIe, not as long as the march
    if n.startswith('march_pose_kernel<1, 0') and float(r[-1]) > 2.5 * med:
        n += ' [e2e: zero-copy stores into pinned host memory]'
    agg.setdefault(n, []).append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
lines = [f"# {tag} - launch list of `python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs` (ncu gpu__time_duration.sum, "
         "--clock-control none)", "", f"Raw: profiles/{tag}_launches.csv. Times are cold-cache and serialised under ncu: "
         "compare shares.  The fill kernel is bench.py's L2 flush, gather_kernel the roofline calibration (both outside "
         "the timed step; sector_kernel is the L2 full-sector calibration); march_pose_kernel<1, 1, ...> with COUNT is the step-counting "
         "pass; most plain march launches belong to the steady-state section (3 modes x 3 repetitions x K launches, warm L2); the march launches marked "
         "[e2e] are the end-to-end scanMany calls, whose ranges are stored by the kernel straight into the caller's "
         "page-locked host buffer (PCIe-bound); march_crash_kernel is the fused checkCollisionMany figure.", "", "| kernel | launches | mean us | total us | share |",
         "|---|---|---|---|---|"]
for k, v in agg.items():
    lines.append(f"| {k} | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/1e3:.1f} | {sum(v)/tot*100:.1f}% |")
open(f'profiles/{tag}_launches_summary.md', 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
