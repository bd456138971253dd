#!/usr/bin/env python3
"""Per-kernel warm-cache timing of the control step (engine option "profile": eager launches,
each bracketed by CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = bridge_config(1)
model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
inp = synth.synthetic_inputs(cfg, B, dtype=torch.bfloat16, device=dev, vary_text=B > 1)
with torch.inference_mode():
    fn = lambda: model(**synth.call_args(inp), noise=inp["noise"])
    for _ in range(3):
        fn()
    iters = 5 if B == 1 else 2
    rep = model._engine.profile(fn, iters)
print(f"B={B}: per-kernel totals over {iters} steps (divide by {iters} for one step)")
print(rep)
