#!/usr/bin/env python3
"""Hottest SASS instructions of an `ncu --set full --import-source on` report:
  ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_hot.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for b in blocks[:2]:
    hdr, data = b["hdr"], b["data"]
    idx = {h: i for i, h in enumerate(hdr)}
    samp = lambda r: int(r[idx["# Samples"]]) if r[idx["# Samples"]].isdigit() else 0
    tot = sum(samp(r) for r in data) or 1
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"== {b['name'][:100]}: {tot} samples, {len(data)} instructions")
    agg = {h: sum(int(r[idx[h]]) for r in data if r[idx[h]].isdigit()) for h in stalls}
    print("   stall mix: " + ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = {id(r): i for i, r in enumerate(data)}
    for r in sorted(data, key=lambda r: -samp(r))[:top_n]:
        s = {h: int(r[idx[h]]) for h in stalls if r[idx[h]].isdigit() and int(r[idx[h]]) > 0}
        main = ", ".join(f"{k[6:]} {v}" for k, v in sorted(s.items(), key=lambda kv: -kv[1])[:2])
        print(f"  #{order[id(r)]:5d} {samp(r):6d} {100 * samp(r) / tot:5.1f}%  {r[idx['Source']].strip()[:64]:64s} {main}")
