#!/usr/bin/env python
"""Turns the ncu outputs of tools/profile_r2.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/."""
import collections
import csv
import shutil
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
# ---- launch list
rows = [l for l in open("gpurun_out/%s_launches_raw.csv" % tag) if not l.startswith("==")]
agg = collections.OrderedDict()
seq = []
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "usecond": v, "nsecond": v / 1e3, "msecond": v * 1e3}[r["Metric Unit"]]
    name = r["Kernel Name"].split("(")[0].replace("rg::", "")
    seq.append((name, v))
half = seq[len(seq) // 2:]  # the second half: past the start-up
for name, v in half:
    agg.setdefault(name, []).append(v)
step = ("k_step_scan", "k_step_fast", "k_step_player", "k_step_monsters", "k_step_gen", "k_prefetch", "k_spec_build")
tot = sum(sum(v) for k, v in agg.items() if k in step)
with open("profiles/%s_launches.csv" % tag, "w") as f:
    f.write("kernel,launches,avg_us,min_us,max_us,total_us,share_of_step_path_pct\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        share = 100.0 * sum(v) / tot if k in step else float("nan")
        f.write("%s,%d,%.2f,%.2f,%.2f,%.1f,%.2f\n" % (k, len(v), sum(v) / len(v), min(v), max(v), sum(v), share))
# ---- --set full extracts
COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for part in ("light", "heavy", "other"):
    rows = list(csv.reader(open("gpurun_out/%s_%s_raw.csv" % (tag, part))))
    hdr = rows[0]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open("profiles/%s_%s_metrics.csv" % (tag, part), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            r = list(r)
            r[hdr.index("Kernel Name")] = r[hdr.index("Kernel Name")].split("(")[0].replace("rg::", "")
            w.writerow([r[i] for i in idx])
shutil.copyfile("gpurun_out/%s_light_dram_appreplay.csv" % tag, "profiles/%s_light_dram_appreplay.csv" % tag)
print(open("profiles/%s_launches.csv" % tag).read())
for part in ("light", "heavy", "other"):
    print("==", part)
    print(open("profiles/%s_%s_metrics.csv" % (tag, part)).read())
