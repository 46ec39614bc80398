#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + per-line tables).  usage: ncu_summary.py rep events_per_launch"""
import collections, csv, subprocess, sys
rep, ev = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f'{k:82s} {r[i]:>20s} {units[i]}')
for i, k in enumerate(hdr):
    if 'sm__inst_executed_pipe' in k and 'pct_of_peak_sustained_elapsed' in k:
        try:
            if float(r[i]) > 5: print(f'{k:82s} {r[i]:>20s} {units[i]}')
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, ""])
cur = hdr = None
def num(x):
    try: return int(float(x))
    except Exception: return 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit() or len(r) < 10: continue
    a = agg[(cur, int(r[0]))]
    a[0] += num(r[hdr.index("# Samples")]); a[1] += num(r[hdr.index("Instructions Executed")])
    a[2] += num(r[hdr.index("L1 Tag Requests Global")]); a[3] += num(r[hdr.index("L1 Wavefronts Shared")])
    a[4] += num(r[hdr.index("Thread Instructions Executed")]); a[5] = r[1]
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
print(f"per event: warp inst {ti/ev:.1f}  L1 global tag requests {sum(a[2] for a in agg.values())/ev:.2f}  shared wavefronts {sum(a[3] for a in agg.values())/ev:.2f}  (inlined lines are attributed to caller and callee)")
for title, idx in (("stall samples", 0), ("warp instructions", 1), ("L1 global tag requests", 2), ("shared wavefronts", 3)):
    print("== top lines by", title)
    tot = sum(a[idx] for a in agg.values()) or 1
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:14]:
        print(f"{f}:{ln:4d} {100*a[idx]/tot:5.1f}%  lanes {a[4]/max(a[1],1):4.1f} | {a[5].strip()[:100]}")
