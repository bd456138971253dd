"""Host-side mirror of the reference's `PiZero` / `PiZeroInference`
(`third_party/open_pi_zero/src/model/vla/pizero.py:33-120, 328-393, 473-614, 721-742`).

Same constructor, same `state_dict` keys and shapes, same mask / position
builders and the same 8-tensor `forward` = `infer_action` call, so it drops into
`scripts/benchmark_pi0.py:127-146,238-242` and `src/agent/eval.py:68-85,215-218`.  The modules
below only *hold* parameters (so `load_state_dict`, `.to()`, `.parameters()`, `freeze_all_weights`
behave like the reference); all arithmetic of the control step runs in the CUDA library behind
`include/blurr_pi0.h`.  There is no PyTorch fallback: without the library or a B200 the call
raises.

Module construction order follows the reference so that `torch.manual_seed(s); PiZeroInference(cfg)`
draws the same default-init weights as the reference does.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import capi, masks


def _cfg_get(node, key, default=None):
    if hasattr(node, "get"):
        try:
            return node.get(key, default)
        except TypeError:
            pass
    return getattr(node, key, default)


# ---------------------------------------------------------------------------
# parameter containers (names = the reference's attribute names)
# ---------------------------------------------------------------------------
class _RMSNormParams(nn.Module):
    """`GemmaRMSNorm` weight (paligemma/modules.py:7-11): zeros, applied as (1 + w)."""

    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(dim))


class _RotaryParams(nn.Module):
    """`GemmaRotaryEmbedding.inv_freq` (paligemma/modules.py:41-45): a non-persistent buffer, so
    `model.to(bfloat16)` rounds it to bf16 exactly like the reference."""

    def __init__(self, dim: int, base: float):
        super().__init__()
        self.dim, self.base = dim, base
        self.register_buffer("inv_freq", tensor=self.fresh_inv_freq(), persistent=False)

    def fresh_inv_freq(self) -> torch.Tensor:
        return 1.0 / (self.base ** (torch.arange(0, self.dim, 2, dtype=torch.int64, device="cpu").float()
                                    / self.dim))


class _SiglipAttentionParams(nn.Module):
    def __init__(self, hidden: int):
        super().__init__()
        # registration order as in siglip.py:115-119 (k, v, q, out): it fixes the RNG draw order
        self.k_proj = nn.Linear(hidden, hidden)
        self.v_proj = nn.Linear(hidden, hidden)
        self.q_proj = nn.Linear(hidden, hidden)
        self.out_proj = nn.Linear(hidden, hidden)


class _SiglipMLPParams(nn.Module):
    def __init__(self, hidden: int, inter: int):
        super().__init__()
        self.fc1 = nn.Linear(hidden, inter)
        self.fc2 = nn.Linear(inter, hidden)


class _SiglipLayerParams(nn.Module):
    def __init__(self, hidden: int, inter: int, eps: float):
        super().__init__()
        self.self_attn = _SiglipAttentionParams(hidden)
        self.layer_norm1 = nn.LayerNorm(hidden, eps=eps)
        self.mlp = _SiglipMLPParams(hidden, inter)
        self.layer_norm2 = nn.LayerNorm(hidden, eps=eps)


class _SiglipEmbeddingsParams(nn.Module):
    def __init__(self, vc):
        super().__init__()
        self.patch_embedding = nn.Conv2d(vc["num_channels"], vc["hidden_size"],
                                         kernel_size=vc["patch_size"], stride=vc["patch_size"],
                                         padding="valid")
        n = (vc["image_size"] // vc["patch_size"]) ** 2
        self.position_embedding = nn.Embedding(n, vc["hidden_size"])
        self.register_buffer("position_ids", torch.arange(n).expand((1, -1)), persistent=False)


class _SiglipEncoderParams(nn.Module):
    def __init__(self, vc):
        super().__init__()
        self.layers = nn.ModuleList([
            _SiglipLayerParams(vc["hidden_size"], vc["intermediate_size"], vc["layer_norm_eps"])
            for _ in range(vc["num_hidden_layers"])
        ])


class _SiglipTransformerParams(nn.Module):
    def __init__(self, vc):
        super().__init__()
        self.embeddings = _SiglipEmbeddingsParams(vc)
        self.encoder = _SiglipEncoderParams(vc)
        self.post_layernorm = nn.LayerNorm(vc["hidden_size"], eps=vc["layer_norm_eps"])


class _SiglipVisionModelParams(nn.Module):
    def __init__(self, vc):
        super().__init__()
        self.vision_model = _SiglipTransformerParams(vc)


class _ProjectorParams(nn.Module):
    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.linear = nn.Linear(in_dim, out_dim, bias=True)


class _MixtureAttentionParams(nn.Module):
    def __init__(self, hidden: int, n_heads: int, n_kv: int, head_dim: int, rope_theta: float):
        super().__init__()
        self.q_proj = nn.Linear(hidden, n_heads * head_dim, bias=False)
        self.k_proj = nn.Linear(hidden, n_kv * head_dim, bias=False)
        self.v_proj = nn.Linear(hidden, n_kv * head_dim, bias=False)
        self.o_proj = nn.Linear(n_heads * head_dim, hidden, bias=False)
        self.rotary_emb = _RotaryParams(head_dim, rope_theta)


class _GemmaMLPParams(nn.Module):
    def __init__(self, hidden: int, inter: int):
        super().__init__()
        self.gate_proj = nn.Linear(hidden, inter, bias=False)
        self.up_proj = nn.Linear(hidden, inter, bias=False)
        self.down_proj = nn.Linear(inter, hidden, bias=False)


class _MixtureLayerParams(nn.Module):
    def __init__(self, hidden, inter, n_heads, n_kv, head_dim, rope_theta):
        super().__init__()
        self.self_attn = _MixtureAttentionParams(hidden, n_heads, n_kv, head_dim, rope_theta)
        self.mlp = _GemmaMLPParams(hidden, inter)
        self.input_layernorm = _RMSNormParams(hidden)
        self.post_attention_layernorm = _RMSNormParams(hidden)


class _MixtureParams(nn.Module):
    def __init__(self, jc, mc):
        super().__init__()
        self.layers = nn.ModuleList([
            _MixtureLayerParams(mc["hidden_size"], mc["intermediate_size"], jc["num_attention_heads"],
                                jc["num_key_value_heads"], jc["head_dim"], mc["rope_theta"])
            for _ in range(jc["num_hidden_layers"])
        ])
        if mc["use_final_norm"]:
            self.norm = _RMSNormParams(mc["hidden_size"])


class _JointModelParams(nn.Module):
    def __init__(self, jc):
        super().__init__()
        self.num_hidden_layers = jc["num_hidden_layers"]
        self.cache_names = [n for n in jc["mixture"] if jc["mixture"][n]["cache"]]
        self.mixtures = nn.ModuleDict()
        for name, mc in jc["mixture"].items():
            self.mixtures[name] = _MixtureParams(jc, mc)
        self.mixture_names = list(jc["mixture"].keys())


class _ActionEncoderParams(nn.Module):
    def __init__(self, action_dim: int, width: int):
        super().__init__()
        self.linear_1 = nn.Linear(action_dim, width)
        self.linear_2 = nn.Linear(2 * width, width)
        self.linear_3 = nn.Linear(width, width)


def sinusoidal_time_table(num_steps: int, width: int, max_period: float, device, dtype) -> torch.Tensor:
    """`time_cond` for each flow step, with the reference's own ops and dtype
    (`SinusoidalPosEmb.forward`, vla/modules.py:15-22; `t += delta_t` accumulates in the model
    dtype, pizero.py:516-538).  Returns [num_steps, width]."""
    half = width // 2
    t = torch.zeros(1, device=device, dtype=dtype)
    delta_t = 1.0 / num_steps
    rows = []
    for _ in range(num_steps):
        emb = math.log(max_period) / (half - 1)
        emb = torch.exp(torch.arange(half, device=t.device, dtype=t.dtype) * -emb)
        emb = t[:, None] * emb[None, :]
        rows.append(torch.cat((emb.sin(), emb.cos()), dim=-1)[0])
        t += delta_t
    return torch.stack(rows).contiguous()


# ---------------------------------------------------------------------------
# the model
# ---------------------------------------------------------------------------
class PiZero(nn.Module):
    """Drop-in for the reference `PiZero` on the inference path."""

    def __init__(self, cfg, use_ddp: bool = False):
        super().__init__()
        self.cfg = cfg
        self.use_ddp = use_ddp
        self.vocab_size = cfg.vocab_size
        self.pad_token_id = cfg.pad_token_id
        self.image_token_index = cfg.image_token_index
        if _cfg_get(cfg, "use_lm_head", False):
            raise NotImplementedError("use_lm_head (text generation) is outside the control-step path")
        if _cfg_get(cfg, "action_expert_adaptive_mode", None):
            raise NotImplementedError(
                "action_expert_adaptive_mode (adaLN) is null in every shipped eval config and is not built")

        self.max_image_text_tokens = cfg.max_image_text_tokens
        self.num_proprio_tokens = cfg.cond_steps
        self.num_action_tokens = cfg.horizon_steps
        self.total_num_tokens = (self.max_image_text_tokens + self.num_proprio_tokens
                                 + self.num_action_tokens)
        self.image_text_hidden_size = cfg.mixture.vlm.hidden_size
        self.proprio_hidden_size = cfg.mixture.proprio.hidden_size
        self.action_hidden_size = cfg.mixture.action.hidden_size

        self.num_inference_steps = cfg.num_inference_steps
        self.horizon_steps = cfg.horizon_steps
        self.action_dim = cfg.action_dim
        self.proprio_dim = cfg.proprio_dim
        self.final_action_clip_value = cfg.final_action_clip_value
        self.flow_sig_min = _cfg_get(cfg, "flow_sig_min", 0.001)
        self.time_max_period = cfg.time_max_period

        vc = cfg.vision.config
        jc = cfg.joint.config
        for name in jc["mixture"]:
            mc = jc["mixture"][name]
            if mc["use_quantize"] or mc["use_lora"]:
                raise RuntimeError("LoRA / 4-bit layers are not shipped (as in the reference's lora.py:17-30)")

        # construction order = reference order (pizero.py:66-111)
        self.embed_tokens = nn.Embedding(cfg.vocab_size, self.image_text_hidden_size, self.pad_token_id)
        self.vision_tower = _SiglipVisionModelParams(vc)
        pc = cfg.vision_projector.config.vision_config
        self.multi_modal_projector = _ProjectorParams(pc["hidden_size"], pc["projection_dim"])
        self.joint_model = _JointModelParams(jc)
        self.action_expert_adaptive_mode = None
        self.action_encoder = _ActionEncoderParams(self.action_dim, self.action_hidden_size)
        self.proprio_encoder = nn.Linear(self.proprio_dim, self.proprio_hidden_size)
        self.action_quant_config = _cfg_get(cfg, "action_quantization")
        self._action_quant_enabled = False
        self.action_decoder = nn.Linear(self.action_hidden_size, self.action_dim)

        # engine state (not part of the module tree)
        object.__setattr__(self, "_engine", None)
        object.__setattr__(self, "_engine_key", None)
        object.__setattr__(self, "_weights_version", 0)
        object.__setattr__(self, "_debug_taps", False)
        object.__setattr__(self, "_use_cuda_graph", True)
        object.__setattr__(self, "_reserve_batch", 1)

    @classmethod
    def from_state_dict(cls, cfg, state_dict: Dict[str, torch.Tensor], device=None, dtype=None):
        """Build without the ~45 s default initialisation of 3.5 B parameters: construct on the
        meta device, adopt the given tensors (`load_state_dict(assign=True)`), rebuild the
        non-persistent buffers, then apply `.to(dtype)` / `.to(device)` like
        `scripts/benchmark_pi0.py:139-142`."""
        with torch.device("meta"):
            model = cls(cfg)
        model.load_state_dict(state_dict, strict=True, assign=True)
        for m in model.modules():
            if isinstance(m, _RotaryParams):
                m.inv_freq = m.fresh_inv_freq()
            elif isinstance(m, _SiglipEmbeddingsParams):
                m.position_ids = torch.arange(m.position_embedding.num_embeddings, device="cpu").expand((1, -1))
        model.freeze_all_weights()
        if dtype is None:   # the reference flow always ends with `model.to(dtype)`, which also
            dtype = next(iter(state_dict.values())).dtype   # rounds the RoPE `inv_freq` buffers
        model.to(dtype)
        if device is not None:
            model.to(device)
        model.eval()
        return model

    # ----- reference API surface --------------------------------------------------------
    def freeze_all_weights(self):
        for _, param in self.named_parameters():
            param.requires_grad = False

    def tie_action_proprio_weights(self):
        self.joint_model.mixtures["proprio"] = self.joint_model.mixtures["action"]
        self._bump()

    def enable_action_quantization(self):
        """pizero.py:274-321.  No-op when `action_quantization.mode` is null (every shipped config).  Modes "int8" /
        "int8_cached" are the reference's *fake* quantisation (int8_linear.py:20-100): each Linear of the action
        mixture and the action encoder (the calls on the bare-Linear `action_decoder` / `proprio_encoder` swap nothing,
        `quantize_module_int8` only replaces children, :94-100) gets per-output-channel symmetric int8 weights
        that `forward` de-quantises back to the activation dtype before a plain `F.linear`, after clamping its input
        to +-activation_clip.  Here the de-quantised weights replace the bf16 ones at upload time (same bits as the
        reference's per-call de-quantisation) and the engine clamps the inputs of those GEMMs (option
        "activation_clip_bits").  "bnb_int8" needs bitsandbytes, which neither this image nor the reference's
        requirements carry."""
        if self._action_quant_enabled:
            return
        cfg = self.action_quant_config or {}
        mode = str(_cfg_get(cfg, "mode", "") or "").lower()
        if mode in {"", "none"}:
            return
        if mode == "bnb_int8":
            raise NotImplementedError("action_quantization.mode='bnb_int8' needs bitsandbytes (absent)")
        if mode not in {"int8", "int8_cached"}:
            return                                  # the reference ignores unknown modes too (:319-320)
        clip = _cfg_get(cfg, "activation_clip", None)
        cache_fp = bool(_cfg_get(cfg, "cache_fp_weight", False))
        fp_dtype = getattr(torch, str(_cfg_get(cfg, "fp_dtype", "bfloat16")), torch.bfloat16)
        mixtures = self.joint_model.mixtures
        roots = [mixtures["action"], self.action_encoder, self.action_decoder, self.proprio_encoder]
        seen = set()
        with torch.no_grad():
            for root in roots:
                for m in root.modules():
                    if m is root or not isinstance(m, nn.Linear) or id(m) in seen:     # children only, like the reference
                        continue
                    seen.add(id(m))
                    w32 = m.weight.detach().to(torch.float32)
                    scale = w32.abs().amax(dim=1, keepdim=True).clamp(min=1e-6) / 127.0
                    q = torch.clamp((w32 / scale).round(), -128, 127).to(torch.int8)
                    deq = q.to(torch.float32) * scale
                    if cache_fp:
                        deq = deq.to(fp_dtype)
                    m.weight.data = deq.to(m.weight.dtype)
        object.__setattr__(self, "_activation_clip", float(clip) if clip is not None else None)
        self._action_quant_enabled = True
        self._bump()

    def _activation_clip_options(self):
        """(float32 bits of the bf16-rounded clip, module mask) for the engine; (0, 0) when off."""
        clip = getattr(self, "_activation_clip", None)
        if not self._action_quant_enabled or clip is None:
            return 0, 0
        c = torch.tensor(clip, dtype=torch.bfloat16).float()        # torch.clamp casts the scalar to the tensor dtype
        bits = int(c.view(torch.int32).item()) & 0xFFFFFFFF
        tied = self.joint_model.mixtures["proprio"] is self.joint_model.mixtures["action"]
        return bits, (1 << 2) | (1 << 3) | ((1 << 1) if tied else 0)

    def build_text_cache(self):
        raise NotImplementedError("text generation is outside the control-step path")

    def build_causal_mask_and_position_ids(self, attention_mask: torch.Tensor, dtype: torch.dtype
                                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Block-attention additive mask and the three position-id tensors; bit-exact with
        pizero.py:328-381, built without the per-sample Python loop (:353-357)."""
        return masks.build_causal_mask_and_position_ids(
            attention_mask, dtype, self.max_image_text_tokens, self.num_proprio_tokens,
            self.num_action_tokens)

    def split_full_mask_into_submasks(self, causal_mask: torch.Tensor):
        """pizero.py:383-393 (views, no copies)."""
        return masks.split_full_mask_into_submasks(
            causal_mask, self.max_image_text_tokens, self.num_proprio_tokens, self.num_action_tokens)

    # ----- weight bookkeeping -----------------------------------------------------------
    def _bump(self):
        object.__setattr__(self, "_weights_version", self._weights_version + 1)

    def refresh_weights(self):
        """Call after editing parameters in place (`p.data.copy_(...)`, an optimiser step, ...): the engine holds a repacked
        copy of the weights that only `.to()`, `load_state_dict`, `tie_action_proprio_weights` and
        `enable_action_quantization` invalidate on their own; the next call re-uploads."""
        self._bump()

    def _apply(self, fn, *args, **kwargs):          # .to() / .cuda() / .bfloat16()
        out = super()._apply(fn, *args, **kwargs)
        self._bump()
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):
        out = super().load_state_dict(state_dict, strict=strict, **kwargs)
        self._bump()
        return out

    def set_engine_options(self, *, debug_taps: Optional[bool] = None, use_cuda_graph: Optional[bool] = None,
                           reserve_batch: Optional[int] = None):
        """`reserve_batch`: size the engine's workspace for at least this many episodes up front (a
        larger batch later re-creates the engine and re-uploads the weights)."""
        if reserve_batch is not None:
            object.__setattr__(self, "_reserve_batch", int(reserve_batch))
        if debug_taps is not None:
            object.__setattr__(self, "_debug_taps", bool(debug_taps))
        if use_cuda_graph is not None:
            object.__setattr__(self, "_use_cuda_graph", bool(use_cuda_graph))
        eng = self._engine
        if eng is not None:
            eng.set_option("debug_taps", int(self._debug_taps))
            eng.set_option("use_cuda_graph", int(self._use_cuda_graph))

    def release_engine(self):
        eng = self._engine
        if eng is not None:
            eng.close()
        object.__setattr__(self, "_engine", None)
        object.__setattr__(self, "_engine_key", None)

    def _config_c(self) -> capi.Pi0ConfigC:
        cfg = self.cfg
        vc, jc = cfg.vision.config, cfg.joint.config
        mx = jc["mixture"]
        if mx["proprio"]["hidden_size"] != mx["action"]["hidden_size"] or \
                mx["proprio"]["intermediate_size"] != mx["action"]["intermediate_size"]:
            raise NotImplementedError("proprio and action experts must share their sizes")
        clip = self.final_action_clip_value
        return capi.Pi0ConfigC(
            abi_version=capi.ABI_VERSION,
            vision_layers=vc["num_hidden_layers"], vision_hidden=vc["hidden_size"],
            vision_intermediate=vc["intermediate_size"], vision_heads=vc["num_attention_heads"],
            image_size=vc["image_size"], patch_size=vc["patch_size"],
            num_image_tokens=vc["num_image_tokens"], layer_norm_eps=vc["layer_norm_eps"],
            joint_layers=jc["num_hidden_layers"], num_heads=jc["num_attention_heads"],
            num_kv_heads=jc["num_key_value_heads"], head_dim=jc["head_dim"],
            vlm_hidden=mx["vlm"]["hidden_size"], vlm_intermediate=mx["vlm"]["intermediate_size"],
            expert_hidden=mx["action"]["hidden_size"], expert_intermediate=mx["action"]["intermediate_size"],
            rms_norm_eps=jc["rms_norm_eps"],
            max_image_text_tokens=self.max_image_text_tokens, num_proprio_tokens=self.num_proprio_tokens,
            num_action_tokens=self.num_action_tokens, action_dim=self.action_dim, proprio_dim=self.proprio_dim,
            vocab_size=cfg.vocab_size, image_token_index=cfg.image_token_index, pad_token_id=cfg.pad_token_id,
            num_inference_steps=self.num_inference_steps, has_clip=int(clip is not None),
            final_action_clip_value=float(clip) if clip is not None else 0.0,
        )

    def _ensure_engine(self, device: torch.device, batch: int) -> "_Engine":
        key = (device.index if device.index is not None else torch.cuda.current_device(),
               self._weights_version)
        eng = self._engine
        if eng is not None and self._engine_key == key and eng.max_batch >= batch:
            if eng.num_steps != self.num_inference_steps:     # cheap: new time table + schedule
                eng.set_steps(self, self.num_inference_steps)
            return eng
        if eng is not None:
            eng.close()
        eng = _Engine(self, device, max_batch=max(batch, self._reserve_batch, 1))
        object.__setattr__(self, "_engine", eng)
        object.__setattr__(self, "_engine_key", key)
        return eng

    # ----- the hot path --------------------------------------------------------------------
    @torch.compiler.disable
    def infer_action(self, input_ids, pixel_values, image_text_proprio_mask, action_mask,
                     vlm_position_ids, proprio_position_ids, action_position_ids, proprios,
                     noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One control step (pizero.py:473-547).  `noise` (optional, [B, horizon, action_dim]) replaces
        the internal `torch.randn` draw (:511-513); when omitted it is drawn with the identical call."""
        dtype, device = pixel_values.dtype, pixel_values.device
        if device.type != "cuda":
            raise RuntimeError("blurr_b200 runs on a CUDA (B200, sm_100a) device only; there is no CPU path")
        if dtype != torch.bfloat16:
            raise NotImplementedError(
                "the B200 path computes in bf16 (the --preset blurr path); cast the model and inputs with "
                f".to(torch.bfloat16) (got {dtype})")
        bsz = pixel_values.size(0)
        eng = self._ensure_engine(device, bsz)
        if noise is None:
            noise = torch.randn((bsz, self.horizon_steps, self.action_dim), device=device, dtype=dtype)
        else:
            noise = noise.to(device=device, dtype=dtype)
        return eng.infer_action(input_ids, pixel_values, image_text_proprio_mask, action_mask,
                                vlm_position_ids, proprio_position_ids, action_position_ids, proprios, noise)

    def check(self) -> None:
        """Synchronise the current stream and raise `BlurrError` if a control step since the last check tripped a
        device-side error (a bounded pipeline wait that expired inside a kernel, a token id outside the embedding
        table, more image tokens than the config allows).  Such a step already returned NaN actions — the last kernel
        of the step poisons them when any of the sticky flags is set — so a caller that never checks still cannot act
        on garbage; `check()` says why and clears the flags.  (The reference would have raised inside the call; this
        path is asynchronous, so the report comes with the first synchronising read or with this call.)"""
        eng = self._engine
        if eng is not None:
            eng.check()

    @torch.compiler.disable
    def infer_action_naive(self, input_ids, pixel_values, causal_mask, vlm_position_ids,
                           proprio_position_ids, action_position_ids, proprios,
                           noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pizero.py:549-614.  The reference's naive mode re-runs the VLM every flow step but attends
        over the *same* K/V with the same mask rows, so its result equals `infer_action`'s (the
        reference asserts this itself, agent/eval.py:213-214); it is served by the cached schedule."""
        itp_mask, action_mask = self.split_full_mask_into_submasks(causal_mask)
        return self.infer_action(input_ids, pixel_values, itp_mask, action_mask, vlm_position_ids,
                                 proprio_position_ids, action_position_ids, proprios, noise=noise)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("flow-matching training loss is outside the control-step path; "
                                  "use PiZeroInference for inference")

    # ----- parity taps -----------------------------------------------------------------------
    def debug_tap(self, name: str) -> torch.Tensor:
        """Named intermediate of the last call (see include/blurr_pi0.h), as a flat bf16 tensor."""
        if self._engine is None:
            raise RuntimeError("no engine yet: run a step first")
        return self._engine.debug_tap(name)

    @property
    def last_launch_count(self) -> int:
        return 0 if self._engine is None else self._engine.last_launch_count()


class PiZeroInference(PiZero):
    """pizero.py:721-742."""

    @torch.compiler.disable
    def forward(self, input_ids, pixel_values, image_text_proprio_mask, action_mask,
                vlm_position_ids, proprio_position_ids, action_position_ids, proprios,
                noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.infer_action(input_ids, pixel_values, image_text_proprio_mask, action_mask,
                                 vlm_position_ids, proprio_position_ids, action_position_ids, proprios,
                                 noise=noise)


# ---------------------------------------------------------------------------
# engine wrapper
# ---------------------------------------------------------------------------
class _Engine:
    """Owns one `blurr_pi0_t` handle: uploads the module's weights and issues control steps."""

    def __init__(self, model: PiZero, device: torch.device, max_batch: int):
        self.lib = capi.load_library()
        self.device = device
        self.max_batch = max_batch
        self.handle = C.c_void_p()
        self.model_dims = (model.num_action_tokens, model.action_dim)
        self.num_steps = model.num_inference_steps
        cfg_c = model._config_c()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        capi.check(self.lib.blurr_pi0_create(C.byref(cfg_c), dev_index, max_batch, C.byref(self.handle)))
        try:
            self._upload(model)
        except Exception:
            self.close()
            raise
        self.set_option("debug_taps", int(model._debug_taps))
        self.set_option("use_cuda_graph", int(model._use_cuda_graph))
        bits, mask = model._activation_clip_options()
        self.set_option("activation_clip_mask", mask)
        self.set_option("activation_clip_bits", bits)

    def _upload(self, model: PiZero):
        sd = model.state_dict()
        # The repack kernels and the table copies of the C side run on the legacy default stream; the tensors they read
        # were produced on torch's CURRENT stream (`.to()`, `.contiguous()`, the time table).  Drain it first so that an
        # engine built under `torch.cuda.stream(side)` (a non-blocking stream) cannot read weights that are still being
        # written, and keep every temporary alive until the C calls have synchronised (finalize does).
        torch.cuda.current_stream(self.device).synchronize()
        keep = []
        with torch.cuda.device(self.device):
            for key, tensor in sd.items():
                if tensor.device != self.device and tensor.device.type != "cuda":
                    raise RuntimeError(f"parameter {key} is on {tensor.device}; move the model to {self.device}")
                if tensor.dtype != torch.bfloat16:
                    raise NotImplementedError(
                        f"parameter {key} is {tensor.dtype}; the B200 path needs model.to(torch.bfloat16)")
                t = tensor.detach().contiguous()
                keep.append(t)
                if t.data_ptr() != tensor.data_ptr():       # .contiguous() made a copy on torch's current stream
                    torch.cuda.current_stream(self.device).synchronize()
                shape = (C.c_int64 * t.dim())(*t.shape)
                capi.check(self.lib.blurr_pi0_set_weight(self.handle, key.encode(), C.c_void_p(t.data_ptr()),
                                                         shape, t.dim(), capi.BLURR_BF16))
            for name in ("vlm", "proprio", "action"):
                inv = model.joint_model.mixtures[name].layers[0].self_attn.rotary_emb.inv_freq
                inv = inv.detach().float().cpu().contiguous()
                arr = (C.c_float * inv.numel())(*inv.tolist())
                capi.check(self.lib.blurr_pi0_set_rope_inv_freq(self.handle, name.encode(), arr, inv.numel()))
            table = sinusoidal_time_table(model.num_inference_steps, model.action_hidden_size,
                                          model.time_max_period, self.device, torch.bfloat16)
            torch.cuda.current_stream(self.device).synchronize()
            capi.check(self.lib.blurr_pi0_set_time_table(self.handle, C.c_void_p(table.data_ptr()),
                                                         model.num_inference_steps))
            capi.check(self.lib.blurr_pi0_finalize_weights(self.handle))      # cudaDeviceSynchronize inside
        del keep

    def set_steps(self, model: "PiZero", steps: int):
        """Change `num_inference_steps` without re-uploading weights."""
        self.set_option("num_inference_steps", steps)
        with torch.cuda.device(self.device):
            table = sinusoidal_time_table(steps, model.action_hidden_size, model.time_max_period, self.device,
                                          torch.bfloat16)
            torch.cuda.current_stream(self.device).synchronize()      # the C side copies on the legacy stream (blocking copy)
            capi.check(self.lib.blurr_pi0_set_time_table(self.handle, C.c_void_p(table.data_ptr()), steps))
        self.num_steps = steps

    def set_option(self, name: str, value: int):
        capi.check(self.lib.blurr_pi0_set_option(self.handle, name.encode(), int(value)))

    def infer_action(self, input_ids, pixel_values, itp_mask, action_mask, vlm_pos, proprio_pos, action_pos,
                     proprios, noise) -> torch.Tensor:
        dev = self.device
        bsz = pixel_values.size(0)

        def i64(x):
            return x.to(device=dev, dtype=torch.int64).contiguous()

        def lastdim_contig(x):
            x = x.to(device=dev, dtype=torch.bfloat16)
            return x if x.stride(-1) == 1 else x.contiguous()

        ids, vp, pp, ap = i64(input_ids), i64(vlm_pos), i64(proprio_pos), i64(action_pos)
        px = pixel_values.to(device=dev)
        m1, m2 = lastdim_contig(itp_mask), lastdim_contig(action_mask)
        pr = proprios.to(device=dev, dtype=torch.bfloat16).contiguous()
        nz = noise.contiguous()
        if m1.dim() != 4 or m2.dim() != 4:
            raise ValueError("masks must be [B, 1, Q, KV] as built by split_full_mask_into_submasks")
        out = torch.empty((bsz, self.model_dims[0], self.model_dims[1]), device=dev, dtype=torch.bfloat16)
        inp = capi.Pi0InputsC()
        inp.input_ids = ids.data_ptr()
        inp.pixel_values = px.data_ptr()
        inp.pixel_strides = (C.c_int64 * 4)(*px.stride())
        inp.image_text_proprio_mask = m1.data_ptr()
        inp.itp_mask_bstride, inp.itp_mask_rstride = m1.stride(0), m1.stride(2)
        inp.action_mask = m2.data_ptr()
        inp.action_mask_bstride, inp.action_mask_rstride = m2.stride(0), m2.stride(2)
        inp.vlm_position_ids = vp.data_ptr()
        inp.proprio_position_ids = pp.data_ptr()
        inp.action_position_ids = ap.data_ptr()
        inp.proprios = pr.data_ptr()
        inp.noise = nz.data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            capi.check(self.lib.blurr_pi0_infer_action(self.handle, C.c_void_p(stream), bsz, C.byref(inp),
                                                       C.c_void_p(out.data_ptr())))
        # keep the staged sources alive until the stream has consumed them
        for t in (ids, vp, pp, ap, px, m1, m2, pr, nz):
            t.record_stream(torch.cuda.current_stream(dev))
        return out

    def check(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self.lib.blurr_pi0_check(self.handle, C.c_void_p(stream)))

    def debug_tap(self, name: str) -> torch.Tensor:
        torch.cuda.synchronize(self.device)
        nbytes = C.c_size_t(0)
        capi.check(self.lib.blurr_pi0_debug_tap(self.handle, name.encode(), None, 0, C.byref(nbytes)))
        out = torch.empty(nbytes.value // 2, device=self.device, dtype=torch.bfloat16)
        capi.check(self.lib.blurr_pi0_debug_tap(self.handle, name.encode(), C.c_void_p(out.data_ptr()),
                                                nbytes.value, C.byref(nbytes)))
        return out

    def last_launch_count(self) -> int:
        return int(self.lib.blurr_pi0_last_launch_count(self.handle))

    def profile(self, fn, iters: int = 3) -> str:
        """Run `fn()` `iters` times with every kernel bracketed by CUDA events; returns the report."""
        self.set_option("profile", 2)
        for _ in range(iters):
            fn()
        torch.cuda.synchronize(self.device)
        self.set_option("profile", 0)
        buf = C.create_string_buffer(1 << 16)
        capi.check(self.lib.blurr_pi0_profile_report(self.handle, buf, len(buf)))
        return buf.value.decode()

    def trace(self, fn, warm: int = 3):
        """Run `fn()` in its normal launch regime (CUDA graph, PDL, streams) with the in-kernel
        %globaltimer stamps on; returns [(idx, stream, start_us, waited_us, end_us, label)] of the
        last call."""
        self.set_option("trace", 1)
        try:
            for _ in range(warm + 1):
                fn()
            torch.cuda.synchronize(self.device)
            buf = C.create_string_buffer(1 << 19)
            capi.check(self.lib.blurr_pi0_trace_report(self.handle, buf, len(buf)))
        finally:
            self.set_option("trace", 0)
        rows = []
        for line in buf.value.decode().splitlines():
            if line.startswith("#") or not line.strip():
                continue
            idx, stream, s, w, e, label = line.split(" ", 5)
            rows.append((int(idx), int(stream), float(s), float(w), float(e), label))
        return rows

    def weight_bytes(self) -> int:
        return int(self.lib.blurr_pi0_weight_bytes(self.handle))

    def close(self):
        if self.handle:
            self.lib.blurr_pi0_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
