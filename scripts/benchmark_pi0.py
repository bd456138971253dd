#!/usr/bin/env python3
"""Drop-in for the reference's `scripts/benchmark_pi0.py` on the B200 path: same flags, same
warm-up / timed loop / report (`benchmark_pi0.py:255-300`), `PiZeroInference` supplied by
`blurr_b200.pizero`.

Differences forced by the offline box: `--checkpoint random` (default) draws random-init weights
instead of loading a `.pt`; when the PaliGemma tokenizer files are missing the prompt is replaced
by synthetic token ids in the `VLAProcessor` layout; `--preset` applies the reference's presets.
"""

from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

from blurr_b200 import synth
from blurr_b200.config import bridge_config, load_yaml_config
from blurr_b200.pizero import PiZeroInference
from blurr_b200.presets import apply_preset


def parse_args():
    p = argparse.ArgumentParser(description="Benchmark latency / VRAM for a PiZero model on the B200 path.")
    p.add_argument("--config", type=str, default="", help="open-pi-zero eval yaml (default: built-in bridge.yaml values)")
    p.add_argument("--checkpoint", type=str, default="random", help="path to a .pt checkpoint, or 'random'")
    p.add_argument("--preset", type=str, default="blurr")
    p.add_argument("--use-bf16", action="store_true")
    p.add_argument("--use-fp16", action="store_true")
    p.add_argument("--use-torch-compile", action="store_true", help="accepted for CLI compatibility (no-op)")
    p.add_argument("--no-prefix-kv-cache", action="store_true")
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--iters", type=int, default=50)
    p.add_argument("--skip-flops", action="store_true", help="accepted for CLI compatibility")
    p.add_argument("--batch", type=int, default=1)
    p.add_argument("--proprio-mode", type=str, default="zeros", choices=["zeros", "random"])
    p.add_argument("--device", type=str, default="cuda:0")
    return p.parse_args()


def main():
    args = parse_args()
    if args.use_fp16:
        raise SystemExit("the B200 path computes in bf16 only")
    cfg = load_yaml_config(args.config) if args.config else bridge_config()
    apply_preset(cfg, args.preset)
    if args.no_prefix_kv_cache:
        cfg["use_prefix_kv_cache"] = False
    device = torch.device(args.device)
    if args.checkpoint == "random":
        sd = synth.random_state_dict_on_device(cfg, device, seed=0)
        model = PiZeroInference.from_state_dict(cfg, sd, device=device)
    else:
        model = PiZeroInference(cfg, use_ddp=False)
        data = torch.load(os.path.expanduser(args.checkpoint), map_location="cpu")
        data["model"] = {k.replace("_orig_mod.", ""): v for k, v in data["model"].items()}
        model.load_state_dict(data["model"], strict=True)
        model.freeze_all_weights()
        model.to(torch.bfloat16)
        model.to(device)
    model.eval()
    inp = synth.synthetic_inputs(cfg, args.batch, dtype=torch.bfloat16, device=device, vary_text=args.batch > 1)
    if args.proprio_mode == "zeros":
        inp["proprios"] = torch.zeros_like(inp["proprios"])
    if cfg["use_prefix_kv_cache"]:
        call = lambda: model(**synth.call_args(inp))
    else:
        call = lambda: model.infer_action_naive(inp["input_ids"], inp["pixel_values"], inp["causal_mask"],
                                                inp["vlm_position_ids"], inp["proprio_position_ids"],
                                                inp["action_position_ids"], inp["proprios"])
    with torch.inference_mode():
        call()
        torch.cuda.synchronize(device)
        torch.cuda.reset_peak_memory_stats(device=device)
        for _ in range(args.warmup):
            call()
        torch.cuda.synchronize(device)
        start = time.time()
        for _ in range(args.iters):
            call()
        torch.cuda.synchronize(device)
    avg = (time.time() - start) / max(args.iters, 1)
    peak = torch.cuda.max_memory_reserved(device=device) / 1024 ** 3
    print(f"[{args.preset}] latency {avg * 1e3:.3f} ms | peak reserved (torch) {peak:.2f} GB | "
          f"engine weights {model._engine.weight_bytes() / 1e9:.2f} GB | flow steps {cfg['num_inference_steps']} | "
          f"kernel launches/step {model.last_launch_count}")


if __name__ == "__main__":
    main()
