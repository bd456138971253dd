#!/usr/bin/env python3
"""Per-kernel SASS census of lib/libblurr_pi0.so: the mnemonics that prove (or disprove) a Blackwell-native kernel.
  python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "blurr-a-boosted-low-resource-inference-for-vision-language-action-model_b200", "lib", "libblurr_pi0.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "LDGSTS", "LDSM", "BRA.U.ANY", "SYNCS"]
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for mn in MN:
        if re.search(r"\b" + re.escape(mn) + r"\b", line) or (mn.endswith("ANY") and mn in line):
            counts[cur][mn] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("SASS census of libblurr_pi0.so (cuobjdump -sass; instruction counts per kernel)")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA load, HMMA = legacy mma.sync, LDGSTS = cp.async,")
print("BRA.U.ANY = ptxas' per-lane uniformisation loop around a single-lane tcgen05 / TMA issue (must be 0)")
print(f"{'kernel':90s} " + " ".join(f"{m:>9s}" for m in MN))
tot = collections.Counter()
for (name, c), d in zip(counts.items(), dem):
    cut = d.rfind(">(")
    short = (d[:cut + 1] if cut > 0 else d.split("(")[0]).replace("void ", "").replace("blurr::", "").replace("(int)", "")
    print(f"{short[:90]:90s} " + " ".join(f"{c[m]:9d}" for m in MN))
    tot.update(c)
print(f"{'TOTAL':90s} " + " ".join(f"{tot[m]:9d}" for m in MN))
