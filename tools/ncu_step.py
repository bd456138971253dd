#!/usr/bin/env python3
"""One control step (or a subset of its stages) between cudaProfilerStart/Stop, for ncu:
  python tools/ncu_step.py [batch] [stage_mask: 1 vision, 2 prefill, 4 action, 7 all] [streams 0/1]
With `--profile-from-start off` ncu sees exactly the kernels of that step, in issue order."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mask = int(sys.argv[2]) if len(sys.argv) > 2 else 7
streams = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = bridge_config(1)
model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
inp = synth.synthetic_inputs(cfg, B, dtype=torch.bfloat16, device=dev, vary_text=B > 1)
args = synth.call_args(inp)
with torch.inference_mode():
    model(**args, noise=inp["noise"])
    model._engine.set_option("use_streams", streams)
    model._engine.set_option("stage_mask", mask)
    for _ in range(3):
        model(**args, noise=inp["noise"])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = model(**args, noise=inp["noise"])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("launches", model.last_launch_count)
