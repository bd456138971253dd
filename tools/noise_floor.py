#!/usr/bin/env python3
"""How far apart are two bf16 runs of the SAME full-size model that differ only in summation order?
Compares, on 16 episodes (clamped actions, max-abs per run pair):
  ours(bs=16) vs oracle(bs=8 blocks) | ours(bs=1 each) vs oracle(bs=1 each) | oracle(bs=8 blocks) vs oracle(bs=1 each)
  | ours(bs=16) vs ours(bs=1 each) | each vs the oracle's fp32 run
The third number is the reference op sequence's own reproducibility floor on this GPU (cuBLAS picks other kernels /
reduction orders at another batch size)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import dist as bdist
from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference
from oracle import pi0_oracle as O

dev = "cuda"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = bridge_config(steps)
cfg.final_action_clip_value = None
sd = synth.synthetic_state_dict(cfg, 0, torch.bfloat16)
model = PiZeroInference.from_state_dict(cfg, sd, device=dev)
model.set_engine_options(reserve_batch=n)
sd_gpu = {k: v.to(dev) for k, v in sd.items()}
sd32 = {k: v.float() for k, v in sd_gpu.items()}
inp = synth.synthetic_inputs(cfg, n, seed=4242, dtype=torch.bfloat16, vary_text=True, device=dev)


def sub(lo, hi):
    return {k: (v[lo:hi] if k in bdist.BATCH_KEYS else v) for k, v in inp.items()}


def oracle(blk, f32=False):
    f = (lambda t: t.float() if (f32 and t.is_floating_point()) else t)
    with torch.inference_mode():
        return O.infer_action(sd32 if f32 else sd_gpu, cfg, blk["input_ids"], f(blk["pixel_values"]).clone(),
                              f(blk["image_text_proprio_mask"]), f(blk["action_mask"]), blk["vlm_position_ids"],
                              blk["proprio_position_ids"], blk["action_position_ids"], f(blk["proprios"]),
                              noise=blk["noise"], rope_dtype=torch.bfloat16).float()


def ours(blk):
    with torch.inference_mode():
        return model(**synth.call_args(blk), noise=blk["noise"]).float().clone()


o_all = ours(inp)
o_one = torch.cat([ours(sub(i, i + 1)) for i in range(n)])
r_blk = torch.cat([oracle(sub(i, i + 8)) for i in range(0, n, 8)])
r_one = torch.cat([oracle(sub(i, i + 1)) for i in range(n)])
r_32 = torch.cat([oracle(sub(i, i + 1), True) for i in range(n)])
model._engine.check()
c = lambda a, b: (a.clamp(-1, 1) - b.clamp(-1, 1)).abs().max().item()
u = lambda a, b: (a - b).abs().max().item()
print(f"flow steps {steps}, {n} episodes, clamped (un-clamped) max-abs action difference:")
print(f"  ours(bs={n}) vs oracle(bs=8)      {c(o_all, r_blk):.3e} ({u(o_all, r_blk):.3e})")
print(f"  ours(bs=1)  vs oracle(bs=1)      {c(o_one, r_one):.3e} ({u(o_one, r_one):.3e})")
print(f"  oracle(bs=8) vs oracle(bs=1)     {c(r_blk, r_one):.3e} ({u(r_blk, r_one):.3e})   <- the reference's own bf16 reproducibility")
print(f"  ours(bs={n}) vs ours(bs=1)        {c(o_all, o_one):.3e} ({u(o_all, o_one):.3e})")
print(f"  vs fp32 oracle: ours(bs={n}) {c(o_all, r_32):.3e} | ours(bs=1) {c(o_one, r_32):.3e} | oracle(bs=8) {c(r_blk, r_32):.3e} | oracle(bs=1) {c(r_one, r_32):.3e}")
m = lambda a, b: (a.clamp(-1, 1) - b.clamp(-1, 1)).abs().mean().item()
print(f"  mean abs vs fp32 oracle: ours(bs={n}) {m(o_all, r_32):.3e} | ours(bs=1) {m(o_one, r_32):.3e} | oracle(bs=8) {m(r_blk, r_32):.3e} | oracle(bs=1) {m(r_one, r_32):.3e}")
