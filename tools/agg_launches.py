#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
into per-kernel totals for one control step (from stage_inputs_kernel to clamp_copy_kernel).
  python tools/agg_launches.py launches.csv [out.txt] [title]"""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
title = sys.argv[3] if len(sys.argv) > 3 else "one bs=1 control step"
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
# one record per launch ID with all of its metrics
launches = collections.OrderedDict()
for r in rows:
    d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time_duration"):
        d["us"] = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    else:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r["Metric Name"]] = val * scale
recs = list(launches.values())
names = [r["name"] for r in recs]
starts = [i for i, n in enumerate(names) if n.startswith("stage_inputs")]
ends = [i for i, n in enumerate(names) if "clamp_copy" in n]
s = starts[0] if starts else 0
e = ([x for x in ends if x > s] or [len(recs) - 1])[0]
step = recs[s:e + 1]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
tot = 0.0
for r in step:
    name = re.sub(r"\(.*", "", r["name"]).replace("void ", "")
    if "gemm_tc" in name or "attn" in name:
        name += " grid=" + r["grid"]
    a = agg[name]
    a[0] += 1; a[1] += r.get("us", 0.0); a[2] += r.get("dram__bytes_read.sum", 0.0); a[3] += r.get("dram__bytes_write.sum", 0.0)
    tot += r.get("us", 0.0)
out = [f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; {title} (cold-cache, serialised: compare shares)",
       f"launches {len(step)}  total kernel time {tot:.1f} us",
       f"{'total us':>10s} {'share':>6s} {'count':>5s} {'avg us':>8s} {'DRAM rd MB/launch':>18s} {'wr MB/launch':>13s} {'DRAM GB/s':>10s}  kernel"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rd, wr = v[2] / v[0] / 1e6, v[3] / v[0] / 1e6
    gbs = (v[2] + v[3]) / (v[1] * 1e-6) / 1e9 if v[1] > 0 else 0.0
    out.append(f"{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d} {v[1] / v[0]:8.2f} {rd:18.2f} {wr:13.2f} {gbs:10.0f}  {k[:100]}")
text = "\n".join(out) + "\n"
print(text)
if dst:
    open(dst, "w").write(text)
