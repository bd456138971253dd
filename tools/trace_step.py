#!/usr/bin/env python3
"""In-graph timeline of one control step: every GEMM / consumer / RoPE / attention kernel stamps
%globaltimer while the step runs exactly as in production (CUDA graph + PDL + three streams).
Prints, per stream, each kernel's start, duration, time its first CTA sat in the programmatic-
dependency wait, and the gap since the previous kernel on that stream ended; then totals by label.
  python tools/trace_step.py [batch] [out.txt]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dst = sys.argv[2] if len(sys.argv) > 2 else None
cfg = bridge_config(1)
model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
inp = synth.synthetic_inputs(cfg, B, dtype=torch.bfloat16, device=dev, vary_text=B > 1)
args = synth.call_args(inp)
with torch.inference_mode():
    model(**args, noise=inp["noise"])       # builds the engine
    rows = model._engine.trace(lambda: model(**args, noise=inp["noise"]))
out = [f"in-graph timeline, batch {B}: {len(rows)} traced kernels (globaltimer, us)"]
names = {0: "main (SigLIP + VLM)", 1: "proprio expert", 2: "action expert"}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for stream in (0, 1, 2):
    rs = sorted([r for r in rows if r[1] == stream], key=lambda r: r[2])
    if not rs:
        continue
    busy = sum(r[4] - r[2] for r in rs)
    out.append(f"--- stream {stream}: {names[stream]}: {len(rs)} kernels, first start {rs[0][2]:.1f}, last end {rs[-1][4]:.1f}, "
               f"sum of durations {busy:.1f}")
    prev_end = None
    for idx, _, s, w, e, label in rs:
        gap = (s - prev_end) if prev_end is not None else 0.0
        waited = (w - s) if w >= 0 else 0.0
        out.append(f"{s:9.2f} dur {e - s:7.2f} wait {waited:6.2f} gap {gap:7.2f}  {label}")
        a = agg[(stream, label)]
        a[0] += 1; a[1] += e - s; a[2] += max(waited, 0.0); a[3] += (e - max(prev_end, s)) if prev_end is not None else e - s
        prev_end = max(e, prev_end) if prev_end is not None else e
out.append("--- totals by label: count, sum of durations, sum of dependency waits, exclusive time (end - max(start, previous end))")
for (stream, label), a in sorted(agg.items(), key=lambda kv: -kv[1][3]):
    out.append(f"s{stream} {a[0]:4d}x dur {a[1]:8.1f} (avg {a[1] / a[0]:6.2f}) wait {a[2]:8.1f} excl {a[3]:8.1f} (avg {a[3] / a[0]:6.2f})  {label}")
text = "\n".join(out) + "\n"
if dst:
    open(dst, "w").write(text)
    print("\n".join(out[-45:]))
else:
    print(text)
