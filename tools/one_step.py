#!/usr/bin/env python3
"""One full-size Bridge control step between cudaProfilerStart/Stop, for ncu launch lists:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/one_step.py [batch]
Without ncu it just runs the step and prints the launch count."""
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
args = synth.call_args(inp)
with torch.inference_mode():
    for _ in range(3):
        model(**args, noise=inp["noise"])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = model(**args, noise=inp["noise"])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
model._engine.check()
print("launches", model.last_launch_count, "actions", out.float().flatten()[:4].tolist())
