#!/usr/bin/env python3
"""In-graph timeline of one OpenVLA-7B-shaped generate call (option "trace"): per kernel start / dependency-wait /
end from %globaltimer, plus the gaps between consecutive kernels.  python tools/llm_trace.py [batch] [layers] [n_new]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import openvla

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
cfg = openvla.LlamaShapedConfig(num_layers=L)
dec = openvla.LlamaDecoder.from_state_dict(cfg, openvla.synthetic_llama_state_dict(cfg, dev, 0), dev, max_batch=B)
dec.set_option("trace", 1)
x = (torch.randn((B, 281, cfg.hidden), device=dev) * 0.5).to(torch.bfloat16)
for _ in range(4):
    dec.generate(x, N)
dec.check()
rows = []
for line in dec.trace_report().splitlines():
    if line.startswith("#"):
        continue
    i, s, w, e, label = line.split(" ", 4)
    rows.append((float(s), float(w), float(e), label))
prev_end = None
print(f"batch {B}, {L} layers, {N} new tokens: {len(rows)} traced kernels")
for s, w, e, label in rows:
    gap = (s - prev_end) if prev_end is not None else 0.0
    print(f"{s:10.2f} dur {e - s:8.2f} wait {max(w - s, 0):7.2f} gap {gap:8.2f}  {label}")
    prev_end = e
