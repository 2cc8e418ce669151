#!/usr/bin/env python
"""Top stall sites of a kernel from an `ncu --page source --csv --print-source sass` export (tools/ncu_source.sh).
Usage: ncu_hot.py gpurun_out/<tag>_source_sass.csv [top N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
S = ci["# Samples"]; E = ci["Instructions Executed"]
tot = sum(int(r[S]) for r in data); tote = sum(int(r[E]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_")]
print(f"kernel: {rows[0][1][:100]}\nsamples {tot}, warp instructions executed {tote}")
agg = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
print("stall totals:", ", ".join(f"{h[6:]} {v / tot:.1%}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
print(f"{'idx':>5} {'samples':>8} {'share':>6} {'exec':>9}  instruction | top stall reasons")
order = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ci[h]]), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{i:5d} {int(r[S]):8d} {int(r[S]) / tot:6.1%} {int(r[E]):9d}  {r[ci['Source']].strip()[:70]:70s} | " +
          ", ".join(f"{n}:{c}" for c, n in st if c))
