#!/usr/bin/env python3
"""Generate the golden vectors of tests/golden/ with the UNMODIFIED reference
(`/root/reference/third_party/open_pi_zero`, imported through oracle/ref_harness.py).

Run in the build container (the only place the reference exists):
    python tests/golden/make_golden.py [--full]

Weights come from `blurr_b200.synth.synthetic_state_dict(cfg, seed=0)` (CPU-seeded, reproducible on
any machine), inputs from `synth.synthetic_inputs(cfg, batch, seed=1234)`; the flow noise is the
injected bf16-representable tensor of the inputs dict.  Stored: the reference's actions (and a few
activation slices) on CPU in fp32 and bf16, un-clamped (`final_action_clip_value=None`) and clamped.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from blurr_b200 import synth
from blurr_b200.config import bridge_config, fractal_config, shrink_config
from oracle import ref_harness

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_model(cfg, sd, dtype):
    pz = ref_harness.import_reference()
    with torch.device("meta"):
        model = pz.PiZeroInference(cfg, use_ddp=False)
    model.load_state_dict(sd, strict=True, assign=True)
    for m in model.modules():
        if type(m).__name__ == "GemmaRotaryEmbedding":
            m.inv_freq = 1.0 / (m.base ** (torch.arange(0, m.dim, 2, dtype=torch.int64).float() / m.dim))
        if type(m).__name__ == "SiglipVisionEmbeddings":
            m.position_ids = torch.arange(m.num_positions).expand((1, -1))
    model.freeze_all_weights()
    model.to(dtype)
    model.eval()
    return model


def run(cfg, sd32, batch, dtype, vary_text, stress=False):
    inp = synth.synthetic_inputs(cfg, batch, dtype=dtype, vary_text=vary_text)
    model = ref_model(cfg, {k: v.clone() for k, v in sd32.items()}, dtype)
    taps = {}
    handles = ref_harness.install_taps(model, lambda n, t: taps.__setitem__(n, t.detach().float()[..., :4, :8].clone()))
    with torch.inference_mode(), ref_harness.patched_randn(inp["noise"]):
        t0 = time.time()
        out = model(**{k: (v.clone() if k == "pixel_values" else v) for k, v in synth.call_args(inp).items()})
        dt = time.time() - t0
    for h in handles:
        h.remove()
    return out.float(), taps, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also the full-size Bridge model (minutes, ~30 GB RAM)")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    out = {"torch": torch.__version__, "recipe": "synth.synthetic_state_dict(seed=0) / synthetic_inputs(seed=1234)"}
    cases = [
        ("shrunk_bridge_s1", shrink_config(bridge_config(1), 2, 3), 2, True, False),
        ("shrunk_bridge_s1_stress", shrink_config(bridge_config(1), 2, 3), 2, True, True),
        ("shrunk_fractal_s10", shrink_config(fractal_config(10), 2, 3), 1, False, False),
    ]
    if args.full:
        cases.append(("full_bridge_s1", bridge_config(1), 1, False, False))
        cases.append(("full_fractal_s10", fractal_config(10), 1, False, False))
    for name, cfg, batch, vary, stress in cases:
        cfg.final_action_clip_value = None
        sd32 = synth.synthetic_state_dict(cfg, 0, torch.float32, stress=stress)
        entry = {"batch": batch, "vary_text": vary, "stress": stress}
        for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            actions, taps, dt = run(cfg, sd32, batch, dtype, vary)
            entry[f"actions_{tag}"] = actions
            entry[f"taps_{tag}"] = taps
            entry[f"seconds_{tag}"] = dt
            print(f"{name} {tag}: {dt:.1f}s actions[0,0]={actions[0, 0].tolist()}", flush=True)
        d = (entry["actions_fp32"] - entry["actions_bf16"]).abs().max().item()
        print(f"{name}: reference bf16 vs fp32 max_abs (un-clamped) = {d:.4e}", flush=True)
        out[name] = entry
        del sd32
    path = os.path.join(HERE, "pi0_reference_golden_full.pt" if args.full else "pi0_reference_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
