#!/usr/bin/env python
"""Summarise gpurun_out/<tag>_launches.csv (if present) and <tag>_prof.ncu-rep into profiles/<tag>_*.txt.
With a second argument, also record dram__bytes_read.sum + dram__bytes_write.sum per launch of the captured kernel under
that key in profiles/traffic.json (what bench.py reports as roofline.traffic):  summarize_ncu.py <tag> [traffic-key]"""
import json
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
src = os.path.join(ROOT, "gpurun_out")

if not os.path.exists(f"{src}/{tag}_launches.csv"):
    rows = []
else:
    rows = [r for r in csv.reader(open(f"{src}/{tag}_launches.csv")) if len(r) > 10]
hdr = rows[0] if rows else ["Kernel Name", "Metric Value"]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
if rows:
  with open(f"{out_dir}/{tag}_launches.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write(f"# source: gpurun_out/{tag}_launches.csv ({sum(len(v) for v in agg.values())} launches)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{sum(v) / tot:6.1%}  n={len(v):4d}  avg={sum(v) / len(v) / 1e3:9.2f} us  {k[:110]}\n")
  print(open(f"{out_dir}/{tag}_launches.txt").read())

if os.path.exists(f"{src}/{tag}_raw.csv"):       # converted on the GPU box (tools/ncu_all.sh): the reports exceed gpurun's 64 MiB
    raw = open(f"{src}/{tag}_raw.csv").read()
else:
    raw = subprocess.run(["ncu", "-i", f"{src}/{tag}_prof.ncu-rep", "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
with open(f"{out_dir}/{tag}_full.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none, 3 launches; source gpurun_out/{tag}_prof.ncu-rep (raw metric page)\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"{w} [{units[i]}]: " + " | ".join(r[i][:90] for r in data) + "\n")
print(open(f"{out_dir}/{tag}_full.txt").read())

if len(sys.argv) > 2 and "dram__bytes_read.sum" in hdr:
    def col(name):
        i = hdr.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return [float(r[i].replace(",", "")) * scale for r in data]
    rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    per_launch = sum(a + b for a, b in zip(rd, wr)) / max(len(rd), 1)
    tj = os.path.join(out_dir, "traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    name = data[0][hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
    t[sys.argv[2]] = {"kernel": name[:80], "bytes_per_launch": round(per_launch),
                      "source": f"profiles/{tag}_full.txt (read {sum(rd) / len(rd) / 1e6:.1f} MB + write {sum(wr) / len(wr) / 1e6:.1f} MB per launch, {len(rd)} launches)"}
    json.dump(t, open(tj, "w"), indent=1)
    print("traffic", sys.argv[2], round(per_launch))
