#!/usr/bin/env python
"""Table of the per-launch metrics tools/gpu_launch_metrics.sh collects: the launches of the last full-size frame in the log (a warm one)."""
import csv, sys, collections

rows = collections.OrderedDict()
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    key = int(r["ID"])
    e = rows.setdefault(key, {"kernel": r["Kernel Name"].split("(")[0]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    n = r["Metric Name"]
    if n == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    if n.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    e[n] = v
ids = sorted(rows)
# frames start at a k_primary launch; bench.py also renders small frames (checks), so take the LAST of the full-size frames (a warm one)
frames = []
for i in ids:
    if rows[i]["kernel"].startswith("k_primary"):
        frames.append([])
    elif frames:
        frames[-1].append(i)
total = [sum(rows[i].get("smsp__inst_executed.sum", 0.0) for i in fr) for fr in frames]
ids = [fr for fr, t in zip(frames, total) if t > 0.9 * max(total)][-1]
print("launch kernel                           ms    warp inst lanes/inst occupancy %  DRAM rd+wr MB  DRAM % stall no_inst  long_sb  barrier")
for i in ids:
    e = rows[i]
    g = lambda k: e.get(k, float("nan"))
    print("%6d %-28s %7.3f %12d %10.2f %11.1f %14.0f %7.1f %13.2f %8.2f %8.2f" % (
        i, e["kernel"], g("gpu__time_duration.sum"), g("smsp__inst_executed.sum"), g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        g("sm__warps_active.avg.pct_of_peak_sustained_active"), (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / 1e6,
        g("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        g("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        g("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        g("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")))
