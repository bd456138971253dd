#!/usr/bin/env python3
"""Batched-episode step time for same-box A/B runs: sustained (K steps back to back, the bench's regime,
power-capped on B200) and burst (a synchronize between steps).  B=64 by default; OPTS / BLURR_PI0_LIB as ab_step.py."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "64"))
K = int(os.environ.get("K", "20"))
if os.environ.get("OPTS"):
    from blurr_b200 import capi
    for kv in os.environ["OPTS"].split(","):
        k, v = kv.split("=")
        capi.check(capi.load_library().blurr_set_global_option(k.encode(), int(v)))
cfg = bridge_config(1)
model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
inp = synth.synthetic_inputs(cfg, B, dtype=torch.bfloat16, device=dev, vary_text=True)
args = synth.call_args(inp)
with torch.inference_mode():
    for _ in range(3):
        model(**args, noise=inp["noise"])
    torch.cuda.synchronize()
    res = []
    for rep in range(2):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(K):
            model(**args, noise=inp["noise"])
        e.record(); torch.cuda.synchronize()
        res.append(s.elapsed_time(e) / K)
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); model(**args, noise=inp["noise"]); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
print(f"{os.environ.get('TAG', '')}: B={B} sustained {res[0]:.2f} / {res[1]:.2f} ms per step ({K} steps), burst median {statistics.median(ts):.2f} ms", flush=True)
