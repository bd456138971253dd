#!/usr/bin/env python3
"""Print an ncu launch list (gpu__time_duration + dram bytes, --csv) of tools/openvla_step.py, starting at the first
kernel of the generate call:  python tools/llm_launches.py launches.csv [first] [last]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
L = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = L.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time"):
        d["us"] = v / 1000 if u in ("ns", "nsecond") else v
    else:
        d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
recs = list(L.values())
start = [i for i, r in enumerate(recs) if "consumer_kernel" in r["name"]][0]
recs = recs[start:]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(recs)
tot = 0.0
for i, r in enumerate(recs[lo:hi], lo):
    nm = re.sub(r"\(.*", "", r["name"]).replace("void ", "").replace("blurr::", "")
    tot += r["us"]
    print(f"{i:3d} {nm[:40]:40s} grid {r['grid']:>14s} {r['us']:8.1f} us  rd {r.get('dram__bytes_read.sum', 0) / 1e6:8.1f} MB "
          f"wr {r.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB")
print(f"total {tot:.1f} us over {hi - lo} launches")
