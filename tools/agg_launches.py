#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for
one control step (from stage_inputs_kernel to clamp_copy_kernel)."""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
starts = [i for i, n in enumerate(names) if n.startswith("stage_inputs")]
ends = [i for i, n in enumerate(names) if "clamp_copy" in n]
s = starts[0]
e = [x for x in ends if x > s][0]
step = rows[s:e + 1]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in step:
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    if "gemm_tc" in name or "attn" in name:
        name += " grid=" + row["Grid Size"]
    t = float(row["Metric Value"].replace(",", "")) / 1000.0
    agg[name][0] += 1
    agg[name][1] += t
    tot += t
out = [f"ncu --metrics gpu__time_duration.sum --clock-control none; one bs=1 control step (cold-cache, serialised)",
       f"launches {len(step)}  total kernel time {tot:.1f} us"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}% {v[0]:4d}x avg {v[1] / v[0]:7.2f}  {k[:100]}")
text = "\n".join(out) + "\n"
print(text)
if dst:
    open(dst, "w").write(text)
