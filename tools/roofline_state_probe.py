#!/usr/bin/env python3
"""Which earlier activity flips bench.py's gate/up measurement between its two modes (38 vs 47 us per launch)?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference

dev = torch.device("cuda:0")
peaks = bench.measured_peaks()


def probe(tag):
    r = bench.dominant_kernel_roofline(dev, peaks)
    print(f"{tag:50s}: {r['ms_per_launch'] * 1e3:6.2f} us back to back, {r['ms_isolated_launch_median'] * 1e3:6.2f} isolated", flush=True)


probe("fresh process")
cfg = bridge_config(1)
model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
probe("after building the model (weights uploaded)")
inp = synth.synthetic_inputs(cfg, 1, dtype=torch.bfloat16, device=dev, vary_text=False)
args = synth.call_args(inp)
with torch.inference_mode():
    for _ in range(20):
        model(**args, noise=inp["noise"])
torch.cuda.synchronize()
probe("after 20 bs=1 control steps")
model.set_engine_options(reserve_batch=64)
inp64 = synth.synthetic_inputs(cfg, 64, dtype=torch.bfloat16, device=dev, vary_text=True)
args64 = synth.call_args(inp64)
with torch.inference_mode():
    for _ in range(5):
        model(**args64, noise=inp64["noise"])
torch.cuda.synchronize()
probe("after 5 bs=64 steps")
model.release_engine()
del model
torch.cuda.empty_cache()
probe("after releasing the engine")
