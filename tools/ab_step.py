#!/usr/bin/env python3
"""bs=1 control-step latency (median of 60, CUDA events) for A/B runs under env toggles."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference
dev = torch.device("cuda:0")
if os.environ.get("OPTS"):          # e.g. OPTS="gemm_persistent=0,gemm_max_stages=4": process-wide tuning knobs
    from blurr_b200 import capi
    for kv in os.environ["OPTS"].split(","):
        k, v = kv.split("=")
        capi.check(capi.load_library().blurr_set_global_option(k.encode(), int(v)))
cfg = bridge_config(1)
model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
inp = synth.synthetic_inputs(cfg, 1, dtype=torch.bfloat16, device=dev)
args = synth.call_args(inp)
with torch.inference_mode():
    for _ in range(10):
        model(**args, noise=inp["noise"])
    torch.cuda.synchronize()
    ts = []
    for _ in range(60):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); model(**args, noise=inp["noise"]); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
print(f"{os.environ.get('TAG', '')}: median {statistics.median(ts):.3f} ms  min {min(ts):.3f}")
