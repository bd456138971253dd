"""Synthetic weights and inputs for the Pi-0 control step (no checkpoint, tokenizer or simulator
is available offline).

* `synthetic_state_dict`: a full reference-layout `state_dict` drawn from
  a per-tensor seeded CPU generator, so the *same* weights can be produced in the build
  container (where golden vectors are made with the real reference) and on the GPU box.
  Distributions follow PyTorch's defaults for the reference's modules (uniform ±1/sqrt(fan_in)
  for Linear/Conv weights and biases, N(0,1) embeddings) except that norm weights/biases get a
  small perturbation so the `(1 + w)` / affine terms are exercised.
* `synthetic_inputs`: the 8 call tensors in the format `VLAProcessor` + the mask builders
  produce (reference `src/model/vla/processing.py:9-22,48-60,96-136`), plus the
  injected flow noise.
"""

from __future__ import annotations

import math
import zlib
from typing import Dict, List, Optional, Tuple

import torch

from . import masks


def state_dict_spec(cfg) -> List[Tuple[str, Tuple[int, ...], str, int]]:
    """(key, shape, kind, fan_in) for every entry of the reference state_dict."""
    vc, jc = cfg["vision"]["config"], cfg["joint"]["config"]
    vh, vi = vc["hidden_size"], vc["intermediate_size"]
    ps, nch = vc["patch_size"], vc["num_channels"]
    npos = (vc["image_size"] // ps) ** 2
    spec: List[Tuple[str, Tuple[int, ...], str, int]] = []
    hid = jc["mixture"]["vlm"]["hidden_size"]
    spec.append(("embed_tokens.weight", (cfg["vocab_size"], hid), "embedding_pad0", 0))
    p = "vision_tower.vision_model."
    spec.append((p + "embeddings.patch_embedding.weight", (vh, nch, ps, ps), "uniform", nch * ps * ps))
    spec.append((p + "embeddings.patch_embedding.bias", (vh,), "uniform", nch * ps * ps))
    spec.append((p + "embeddings.position_embedding.weight", (npos, vh), "embedding", 0))
    for l in range(vc["num_hidden_layers"]):
        lp = f"{p}encoder.layers.{l}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            spec.append((lp + f"self_attn.{n}.weight", (vh, vh), "uniform", vh))
            spec.append((lp + f"self_attn.{n}.bias", (vh,), "uniform", vh))
        spec.append((lp + "layer_norm1.weight", (vh,), "ln_weight", 0))
        spec.append((lp + "layer_norm1.bias", (vh,), "ln_bias", 0))
        spec.append((lp + "mlp.fc1.weight", (vi, vh), "uniform", vh))
        spec.append((lp + "mlp.fc1.bias", (vi,), "uniform", vh))
        spec.append((lp + "mlp.fc2.weight", (vh, vi), "uniform", vi))
        spec.append((lp + "mlp.fc2.bias", (vh,), "uniform", vi))
        spec.append((lp + "layer_norm2.weight", (vh,), "ln_weight", 0))
        spec.append((lp + "layer_norm2.bias", (vh,), "ln_bias", 0))
    spec.append((p + "post_layernorm.weight", (vh,), "ln_weight", 0))
    spec.append((p + "post_layernorm.bias", (vh,), "ln_bias", 0))
    pc = cfg["vision_projector"]["config"]["vision_config"]
    spec.append(("multi_modal_projector.linear.weight", (pc["projection_dim"], pc["hidden_size"]), "uniform",
                 pc["hidden_size"]))
    spec.append(("multi_modal_projector.linear.bias", (pc["projection_dim"],), "uniform", pc["hidden_size"]))
    nh, nkv, hd = jc["num_attention_heads"], jc["num_key_value_heads"], jc["head_dim"]
    for name, mc in jc["mixture"].items():
        h, i = mc["hidden_size"], mc["intermediate_size"]
        for l in range(jc["num_hidden_layers"]):
            lp = f"joint_model.mixtures.{name}.layers.{l}."
            spec.append((lp + "self_attn.q_proj.weight", (nh * hd, h), "uniform", h))
            spec.append((lp + "self_attn.k_proj.weight", (nkv * hd, h), "uniform", h))
            spec.append((lp + "self_attn.v_proj.weight", (nkv * hd, h), "uniform", h))
            spec.append((lp + "self_attn.o_proj.weight", (h, nh * hd), "uniform", nh * hd))
            spec.append((lp + "mlp.gate_proj.weight", (i, h), "uniform", h))
            spec.append((lp + "mlp.up_proj.weight", (i, h), "uniform", h))
            spec.append((lp + "mlp.down_proj.weight", (h, i), "uniform", i))
            spec.append((lp + "input_layernorm.weight", (h,), "rms_weight", 0))
            spec.append((lp + "post_attention_layernorm.weight", (h,), "rms_weight", 0))
        if mc["use_final_norm"]:
            spec.append((f"joint_model.mixtures.{name}.norm.weight", (h,), "rms_weight", 0))
    w = jc["mixture"]["action"]["hidden_size"]
    ad, pd = cfg["action_dim"], cfg["proprio_dim"]
    spec.append(("action_encoder.linear_1.weight", (w, ad), "uniform", ad))
    spec.append(("action_encoder.linear_1.bias", (w,), "uniform", ad))
    spec.append(("action_encoder.linear_2.weight", (w, 2 * w), "uniform", 2 * w))
    spec.append(("action_encoder.linear_2.bias", (w,), "uniform", 2 * w))
    spec.append(("action_encoder.linear_3.weight", (w, w), "uniform", w))
    spec.append(("action_encoder.linear_3.bias", (w,), "uniform", w))
    spec.append(("proprio_encoder.weight", (jc["mixture"]["proprio"]["hidden_size"], pd), "uniform", pd))
    spec.append(("proprio_encoder.bias", (jc["mixture"]["proprio"]["hidden_size"],), "uniform", pd))
    spec.append(("action_decoder.weight", (ad, w), "uniform", w))
    spec.append(("action_decoder.bias", (ad,), "uniform", w))
    return spec


def _draw(key: str, shape, kind: str, fan_in: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFFFFFFFFFF)
    if kind == "uniform":
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
    if kind == "embedding":
        return torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "embedding_pad0":
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        t[0].zero_()
        return t
    if kind == "ln_weight":
        return 1.0 + 0.1 * (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0)
    if kind == "ln_bias":
        return 0.1 * (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0)
    if kind == "rms_weight":
        return 0.1 * (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0)
    raise ValueError(kind)


def synthetic_state_dict(cfg, seed: int = 0, dtype: torch.dtype = torch.bfloat16, device="cpu",
                         stress: bool = False, tie_experts: bool = False) -> Dict[str, torch.Tensor]:
    """Deterministic (CPU generator) weights in the reference state_dict layout.

    `stress=True` multiplies every mixture `q_proj`/`k_proj` weight by 8 (exact in bf16) so attention
    logits grow 64x: soft-clamp, masking and RoPE then matter.
    `tie_experts=True` gives the proprio expert the action expert's weights (what a trained
    checkpoint holds, pizero.py:270-272)."""
    out: Dict[str, torch.Tensor] = {}
    for key, shape, kind, fan_in in state_dict_spec(cfg):
        src_key = key
        if tie_experts and key.startswith("joint_model.mixtures.proprio."):
            src_key = key.replace(".proprio.", ".action.", 1)
        t = _draw(src_key, shape, kind, fan_in, seed)
        if stress and key.startswith("joint_model.mixtures.") and (
                key.endswith("q_proj.weight") or key.endswith("k_proj.weight")):
            t = t * 8.0
        out[key] = t.to(dtype).to(device)
    return out


def process_images(images_u8: torch.Tensor) -> torch.Tensor:
    """`process_images(img, 1/255, mean=.5, std=.5)` (processing.py:48-60) in fp32."""
    x = images_u8 * (1 / 255.0)
    mean = torch.tensor([0.5, 0.5, 0.5])[None, :, None, None]
    std = torch.tensor([0.5, 0.5, 0.5])[None, :, None, None]
    return (x - mean) / std


def synthetic_inputs(cfg, batch: int = 1, seed: int = 1234, dtype: torch.dtype = torch.bfloat16,
                     vary_text: bool = False, device="cpu") -> Dict[str, torch.Tensor]:
    """The 8 call tensors of `PiZeroInference.forward` + `noise` + `causal_mask`/`attention_mask`.

    `input_ids` = `<image>`x256, BOS, text ids, "\\n" (108), pads — the `VLAProcessor` layout
    (processing.py:9-22,128-134).  With `vary_text` the text length differs per sample (6..18 ids)
    so every row has its own valid-token count."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_img = cfg["vision"]["config"]["num_image_tokens"]
    n_it = cfg["max_image_text_tokens"]
    img_tok, pad = cfg["image_token_index"], cfg["pad_token_id"]
    max_text = n_it - n_img - 2
    ids = torch.full((batch, n_it), pad, dtype=torch.int64)
    ids[:, :n_img] = img_tok
    ids[:, n_img] = 2
    for b in range(batch):
        n_text = 10 if not vary_text else int(torch.randint(6, max_text + 1, (1,), generator=g))
        n_text = min(n_text, max_text)
        ids[b, n_img + 1:n_img + 1 + n_text] = torch.randint(3, 257000, (n_text,), generator=g)
        ids[b, n_img + 1 + n_text] = 108
    size = cfg["vision"]["config"]["image_size"]
    img = torch.randint(0, 256, (batch, 3, size, size), dtype=torch.uint8, generator=g)
    px = process_images(img)
    proprios = torch.rand(batch, cfg["cond_steps"], cfg["proprio_dim"], generator=g) * 2 - 1
    noise = torch.randn(batch, cfg["horizon_steps"], cfg["action_dim"], generator=g).bfloat16().float()
    attention_mask = (ids != pad).long()
    causal_mask, vp, pp, ap = masks.build_causal_mask_and_position_ids(
        attention_mask, dtype, n_it, cfg["cond_steps"], cfg["horizon_steps"])
    m1, m2 = masks.split_full_mask_into_submasks(causal_mask, n_it, cfg["cond_steps"], cfg["horizon_steps"])
    out = {
        "input_ids": ids,
        "pixel_values": px.to(dtype),
        "image_text_proprio_mask": m1,
        "action_mask": m2,
        "vlm_position_ids": vp,
        "proprio_position_ids": pp,
        "action_position_ids": ap,
        "proprios": proprios.to(dtype),
        "noise": noise.to(dtype),
        "causal_mask": causal_mask,
        "attention_mask": attention_mask,
    }
    if device != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out


CALL_KEYS = ("input_ids", "pixel_values", "image_text_proprio_mask", "action_mask", "vlm_position_ids",
             "proprio_position_ids", "action_position_ids", "proprios")


def call_args(inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k: inputs[k] for k in CALL_KEYS}


def random_state_dict_on_device(cfg, device, seed: int = 0, dtype: torch.dtype = torch.bfloat16
                                ) -> Dict[str, torch.Tensor]:
    """Same layout and distributions as `synthetic_state_dict`, drawn directly on `device` with
    its own generator (seconds instead of a minute for 3.5 B parameters).  Used by `bench.py`,
    where only the shapes and value ranges matter; parity tests use the CPU-seeded recipe."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for key, shape, kind, fan_in in state_dict_spec(cfg):
        if kind == "uniform":
            bound = 1.0 / math.sqrt(fan_in)
            t = torch.empty(shape, device=device, dtype=torch.float32).uniform_(-bound, bound, generator=g)
        elif kind in ("embedding", "embedding_pad0"):
            t = torch.empty(shape, device=device, dtype=dtype).normal_(generator=g)
            if kind == "embedding_pad0":
                t[0].zero_()
        elif kind == "ln_weight":
            t = 1.0 + torch.empty(shape, device=device, dtype=torch.float32).uniform_(-0.1, 0.1, generator=g)
        else:
            t = torch.empty(shape, device=device, dtype=torch.float32).uniform_(-0.1, 0.1, generator=g)
        out[key] = t.to(dtype)
    return out
