#!/usr/bin/env python3
"""Per-stage timing of the bs=1 control step under the real launch regime (CUDA graph + PDL):
runs the step with subsets of the stages enabled (results are meaningless, timings are not)."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference

from blurr_b200 import capi

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for kv in filter(None, os.environ.get("OPTS", "").split(",")):      # e.g. OPTS="attn_tc=1,gemm_pair_small=2"
    k, v = kv.split("=")
    capi.check(capi.load_library().blurr_set_global_option(k.encode(), int(v)))
cfg = bridge_config(1)
model = PiZeroInference.from_state_dict(cfg, synth.random_state_dict_on_device(cfg, dev), device=dev)
inp = synth.synthetic_inputs(cfg, B, dtype=torch.bfloat16, device=dev, vary_text=B > 1)
args = synth.call_args(inp)


def timed(n=30):
    with torch.inference_mode():
        for _ in range(5):
            model(**args, noise=inp["noise"])
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); model(**args, noise=inp["noise"]); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
    return statistics.median(ts), model.last_launch_count


with torch.inference_mode():
    model(**args, noise=inp["noise"])
for kv in filter(None, os.environ.get("ENGINE_OPTS", "").split(",")):      # e.g. ENGINE_OPTS="chunked_splitk=0"
    k, v = kv.split("=")
    model._engine.set_option(k, int(v))
for name, mask in [("all", 7), ("vision only", 1), ("prefill only", 2), ("action only", 4), ("staging only", 0)]:
    model._engine.set_option("stage_mask", mask)
    ms, launches = timed(30 if B == 1 else 5)
    print(f"B={B} {name:14s}: {ms:8.3f} ms  launches={launches}", flush=True)
model._engine.set_option("stage_mask", 7)
