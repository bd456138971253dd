"""The reference's named presets (`scripts/eval_pi0_simpler.py:21-41`,
`scripts/paper/pi0_microbench.py:331-349`) restated for the B200 path.

`blurr` (= `blurr_step1`, `step1`): prefix KV cache, bf16, compiled, 1 flow step — the path this
repo accelerates.  `prefix_cache`/`cached` and `baseline`/`vanilla` select 10 flow steps; they run on
the same bf16 engine here (the reference's fp32 eager arithmetic is not rebuilt), and the naive
no-cache mode is served by the cached schedule because its result is identical
(`src/agent/eval.py:213-214`).
"""

from __future__ import annotations

PRESETS = {
    "baseline": dict(use_prefix_kv_cache=False, use_bf16=False, use_torch_compile=False, num_inference_steps=10),
    "vanilla": dict(use_prefix_kv_cache=False, use_bf16=False, use_torch_compile=False, num_inference_steps=10),
    "prefix_cache": dict(use_prefix_kv_cache=True, use_bf16=False, use_torch_compile=False, num_inference_steps=10),
    "cached": dict(use_prefix_kv_cache=True, use_bf16=False, use_torch_compile=False, num_inference_steps=10),
    "blurr": dict(use_prefix_kv_cache=True, use_bf16=True, use_torch_compile=True, num_inference_steps=1),
    "blurr_step1": dict(use_prefix_kv_cache=True, use_bf16=True, use_torch_compile=True, num_inference_steps=1),
    "step1": dict(use_prefix_kv_cache=True, use_bf16=True, use_torch_compile=True, num_inference_steps=1),
}


def apply_preset(cfg, preset: str) -> None:
    """Mutates `cfg` exactly like the reference's `_apply_preset`."""
    key = preset.lower().strip()
    if key not in PRESETS:
        raise ValueError(f"Unknown preset: {preset}")
    cfg["use_prefix_kv_cache"] = cfg.get("use_prefix_kv_cache", True)
    for k, v in PRESETS[key].items():
        cfg[k] = v
