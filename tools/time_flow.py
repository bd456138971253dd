#!/usr/bin/env python3
"""Flow-loop cost: one action-expert pass (stage_mask 4) with per-op kernels vs the persistent step kernel,
and the full 10-step Bridge control step."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference

dev = torch.device("cuda:0")
if os.environ.get("ATTN_TC"):
    from blurr_b200 import capi
    capi.check(capi.load_library().blurr_set_global_option(b"attn_tc", int(os.environ["ATTN_TC"])))
def timed(model, args, noise, n=30):
    with torch.inference_mode():
        for _ in range(5):
            model(**args, noise=noise)
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); model(**args, noise=noise); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
    return statistics.median(ts)

for steps in (1, 10):
    cfg = bridge_config(steps)
    model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
    inp = synth.synthetic_inputs(cfg, 1, dtype=torch.bfloat16, device=dev)
    args = synth.call_args(inp)
    with torch.inference_mode():
        model(**args, noise=inp["noise"])
    print(f"steps={steps} all (graph, streams): {timed(model, args, inp['noise']):.3f} ms", flush=True)
    model._engine.set_option("stage_mask", 4)
    print(f"steps={steps} action only: {timed(model, args, inp['noise']):.3f} ms", flush=True)
    model._engine.set_option("stage_mask", 7)
    model.release_engine()
    del model
