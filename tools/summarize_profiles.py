#!/usr/bin/env python
"""Turns the ncu outputs of tools/profile_round.sh (in gpurun_out/) into the tracked summaries under
profiles/: a per-kernel aggregate of the launch list and a table of the --set full metrics."""
import collections
import csv
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rows = [l for l in open("gpurun_out/%s_launches_raw.csv" % tag) if not l.startswith("==")]
agg = collections.OrderedDict()
order = []
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "usecond": v, "nsecond": v / 1e3, "msecond": v * 1e3}[r["Metric Unit"]]
    name = r["Kernel Name"].split("(")[0]
    agg.setdefault(name, []).append(v)
    order.append((name, v))
step_kernels = ("k_step_scan", "k_step_player", "k_step_monsters", "k_step_gen", "k_step_end", "k_prefetch")
tot = sum(sum(v) for k, v in agg.items() if k in step_kernels)
with open("profiles/%s_launches.csv" % tag, "w") as f:
    f.write("kernel,launches,avg_us,min_us,max_us,total_us,share_of_step_path_pct\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        share = 100.0 * sum(v) / tot if k in step_kernels else float("nan")
        f.write("%s,%d,%.2f,%.2f,%.2f,%.1f,%.2f\n" % (k, len(v), sum(v) / len(v), min(v), max(v), sum(v), share))
# steady-state part only (second half of the launches) for the per-step shares
half = order[len(order) // 2:]
agg2 = collections.defaultdict(list)
for k, v in half:
    agg2[k].append(v)
tot2 = sum(sum(v) for k, v in agg2.items() if k in step_kernels)
with open("profiles/%s_launches_steady.csv" % tag, "w") as f:
    f.write("kernel,launches,avg_us,total_us,share_of_step_path_pct\n")
    for k, v in sorted(agg2.items(), key=lambda kv: -sum(kv[1])):
        if k in step_kernels:
            f.write("%s,%d,%.2f,%.1f,%.2f\n" % (k, len(v), sum(v) / len(v), sum(v), 100.0 * sum(v) / tot2))

r = list(csv.reader(open("gpurun_out/%s_step_full_raw.csv" % tag)))
h = r[0]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
idx = [h.index(c) for c in want if c in h]
with open("profiles/%s_step_full_metrics.csv" % tag, "w") as f:
    w = csv.writer(f)
    w.writerow([h[i] for i in idx])
    w.writerow([r[1][i] for i in idx])
    for row in r[2:]:
        w.writerow([row[i].split("(")[0] if h[i] == "Kernel Name" else row[i] for i in idx])
print(open("profiles/%s_launches_steady.csv" % tag).read())
print(open("profiles/%s_step_full_metrics.csv" % tag).read())
