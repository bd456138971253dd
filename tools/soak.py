#!/usr/bin/env python3
"""Soak: thousands of control steps / generate calls back to back; device-side error flags and result stability checked."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import openvla, synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cfg = bridge_config(1)
model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
inp = synth.synthetic_inputs(cfg, 1, dtype=torch.bfloat16, device=dev, vary_text=False)
args = synth.call_args(inp)
t0 = time.time()
with torch.inference_mode():
    first = model(**args, noise=inp["noise"]).clone()
    for i in range(N):
        out = model(**args, noise=inp["noise"])
        if (i + 1) % 1000 == 0:
            model.check()
            assert torch.equal(out, first), f"step {i}: result changed"
torch.cuda.synchronize()
print(f"pi0: {N} control steps, identical results, no device flags, {time.time() - t0:.1f} s", flush=True)
lcfg = openvla.LlamaShapedConfig(num_layers=4)
dec = openvla.LlamaDecoder.from_state_dict(lcfg, openvla.synthetic_llama_state_dict(lcfg, dev, 0), dev, max_batch=4)
x = (torch.randn((4, 281, lcfg.hidden), device=dev) * 0.5).to(torch.bfloat16)
ids0 = dec.generate(x, 7).clone()
for i in range(300):
    ids = dec.generate(x, 7)
dec.check()
assert torch.equal(ids, ids0)
print("llm: 300 generate calls, identical tokens, no device flags", flush=True)
