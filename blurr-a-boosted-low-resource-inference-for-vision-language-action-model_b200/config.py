"""Pi-0 config handling for the B200 path.

The reference drives ``PiZeroInference(cfg)`` with an OmegaConf ``DictConfig``
loaded from ``third_party/open_pi_zero/config/eval/*.yaml`` (attribute access,
``.get()``, ``${...}`` interpolation; reference ``scripts/benchmark_pi0.py:127``,
``src/model/vla/pizero.py:35-120``).  OmegaConf/hydra are not required here:
``AttrDict`` gives the same access pattern, ``load_yaml_config`` resolves the
interpolations the Pi-0 YAMLs use, and ``bridge_config()/fractal_config()``
restate the shipped values (``config/eval/bridge.yaml``, ``fractal_*.yaml``) so
the path works without the YAML files (they do not exist on the GPU box).

A real OmegaConf ``DictConfig`` is accepted everywhere an ``AttrDict`` is.
"""

from __future__ import annotations

import copy
import os
import re
import time
from typing import Any, Dict, Iterator, Mapping


class AttrDict(dict):
    """dict with attribute access and OmegaConf-like ``get``; nested dicts wrap lazily."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in list(self.items()):
            if isinstance(v, Mapping) and not isinstance(v, AttrDict):
                super().__setitem__(k, AttrDict(v))

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as exc:  # same error class OmegaConf raises in struct mode
            raise AttributeError(name) from exc

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __setitem__(self, key, value):
        if isinstance(value, Mapping) and not isinstance(value, AttrDict):
            value = AttrDict(value)
        super().__setitem__(key, value)

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def keys_list(self):
        return list(self.keys())


def merge(base: Mapping, override: Mapping) -> AttrDict:
    """Recursive merge with ``override`` winning: the behaviour of ``OmegaConf.merge``
    that ``JointModel.__init__`` relies on (reference ``joint_model.py:328-330``)."""
    out = AttrDict(copy.deepcopy(dict(base)))
    for k, v in override.items():
        if isinstance(v, Mapping) and isinstance(out.get(k), Mapping):
            out[k] = merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


_INTERP = re.compile(r"\$\{([^${}]+)\}")


def _lookup(root: Mapping, dotted: str) -> Any:
    node: Any = root
    for part in dotted.split("."):
        node = node[part]
    return node


def _resolve_str(root: Mapping, s: str, depth: int = 0) -> Any:
    if depth > 16:
        raise ValueError(f"interpolation too deep: {s}")
    m = _INTERP.fullmatch(s.strip())
    if m:  # whole-value interpolation keeps the node type
        val = _resolve_expr(root, m.group(1), depth)
        return val
    def sub(match):
        return str(_resolve_expr(root, match.group(1), depth))
    prev = None
    while prev != s and _INTERP.search(s):
        prev = s
        s = _INTERP.sub(sub, s)
    return s


def _resolve_expr(root: Mapping, expr: str, depth: int) -> Any:
    expr = expr.strip()
    if expr.startswith("oc.env:"):
        name = expr[len("oc.env:"):].split(",")[0]
        return os.environ.get(name, f"<env:{name}>")
    if expr.startswith("now:"):
        return time.strftime(expr[len("now:"):])
    if expr.startswith("eval:"):
        src = expr[len("eval:"):].strip().strip("'\"")
        return eval(src, {"__builtins__": {}}, {})  # arithmetic only, as in fractal_*.yaml
    val = _lookup(root, expr)
    if isinstance(val, str) and _INTERP.search(val):
        val = _resolve_str(root, val, depth + 1)
    elif isinstance(val, Mapping):
        val = _resolve_tree(root, copy.deepcopy(val), depth + 1)
    return val


_FLOAT_LIKE = re.compile(r"[-+]?(\d+\.?\d*|\.\d+)[eE][-+]?\d+")


def _resolve_tree(root: Mapping, node: Any, depth: int = 0) -> Any:
    if isinstance(node, str) and _FLOAT_LIKE.fullmatch(node.strip()):
        return float(node)          # PyYAML (YAML 1.1) reads `1e-6` as a string; OmegaConf as a float
    if isinstance(node, Mapping):
        return AttrDict({k: _resolve_tree(root, v, depth) for k, v in node.items()})
    if isinstance(node, list):
        return [_resolve_tree(root, v, depth) for v in node]
    if isinstance(node, str) and _INTERP.search(node):
        try:
            return _resolve_str(root, node, depth)
        except (KeyError, TypeError):
            return node
    return node


def load_yaml_config(path: str) -> AttrDict:
    """Load one of the open-pi-zero eval YAMLs and resolve its ``${...}`` references."""
    import yaml

    with open(path, "r", encoding="utf-8") as f:
        raw = yaml.safe_load(f)
    if raw is None:
        raise ValueError(f"{path} is empty (e.g. bridge_pool64_steps1.yaml ships empty)")
    if "mixture" not in raw:
        raise ValueError(
            f"{path} needs hydra `defaults` composition (no script in the reference performs it)"
        )
    return _resolve_tree(raw, raw)


# ---------------------------------------------------------------------------
# Shipped configs restated (values: config/eval/bridge.yaml:35-136)
# ---------------------------------------------------------------------------

def _mixture(hidden, inter, final_norm, cache) -> Dict[str, Any]:
    return dict(hidden_size=hidden, intermediate_size=inter, use_final_norm=final_norm,
                cache=cache, use_quantize=False, use_lora=False, adaptive_mode=None,
                rope_theta=10000.0)


def bridge_config(num_inference_steps: int = 10, **overrides) -> AttrDict:
    """``config/eval/bridge.yaml`` (``bridge_step1.yaml`` is ``num_inference_steps=1``)."""
    mixture = dict(
        vlm=_mixture(2048, 16384, False, True),       # bridge.yaml:55-64
        proprio=_mixture(1024, 4096, True, True),     # bridge.yaml:65-73
        action=_mixture(1024, 4096, True, False),     # bridge.yaml:74-82
    )
    cfg = dict(
        max_seq_len=276, num_inference_steps=num_inference_steps,
        final_action_clip_value=1.0, use_torch_compile=True, use_bf16=False,
        action_quantization=dict(mode=None, activation_clip=1.0, cache_fp_weight=True),
        kv_quantization=dict(mode=None, activation_clip=1.0, dtype="bfloat16"),
        use_flash_attn=False,
        cond_steps=1, horizon_steps=4, act_steps=4, action_dim=7, proprio_dim=7,
        mixture=mixture,
        action_expert_adaptive_mode=None, time_hidden_size=256, time_max_period=10000.0,
        action_expert_rope_theta=10000.0, quantize=False, lora=False,
        max_image_text_tokens=276,
        image_token_index=257152, vocab_size=257216, pad_token_id=0,
        vision=dict(
            _target_="src.model.paligemma.siglip.SiglipVisionModel",
            config=dict(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27,
                        num_attention_heads=16, num_channels=3, image_size=224, patch_size=14,
                        layer_norm_eps=1e-6, attention_dropout=0.0, num_image_tokens=256),
        ),
        vision_projector=dict(
            _target_="src.model.paligemma.siglip.PaliGemmaMultiModalProjector",
            config=dict(vision_config=dict(hidden_size=1152, projection_dim=2048)),
        ),
        joint=dict(
            _target_="src.model.vla.joint_model.JointModel",
            config=dict(
                action_expert_adaptive_mode=None, time_hidden_size=256,
                mixture=copy.deepcopy(mixture), lora=dict(r=32, dropout=0.0),
                num_hidden_layers=18, num_attention_heads=8, num_key_value_heads=1,
                head_dim=256, rms_norm_eps=1e-6, attention_bias=False,
                attention_dropout=0.0, pad_token_id=0,
            ),
        ),
    )
    out = AttrDict(cfg)
    for k, v in overrides.items():
        out[k] = v
    return out


def fractal_config(num_inference_steps: int = 10, **overrides) -> AttrDict:
    """``config/eval/fractal_*.yaml``: Bridge with ``proprio_dim=8`` and ``act_steps=2``."""
    cfg = bridge_config(num_inference_steps)
    cfg.proprio_dim = 8
    cfg.act_steps = 2
    for k, v in overrides.items():
        cfg[k] = v
    return cfg


def shrink_config(cfg: Mapping, vision_layers: int, joint_layers: int) -> AttrDict:
    """Same widths, fewer layers: the reduced-depth configs the fast parity tests use."""
    out = AttrDict(copy.deepcopy(dict(cfg)))
    out.vision.config.num_hidden_layers = vision_layers
    out.joint.config.num_hidden_layers = joint_layers
    return out


def iter_mixture_names(cfg: Mapping) -> Iterator[str]:
    return iter(cfg["joint"]["config"]["mixture"].keys())
