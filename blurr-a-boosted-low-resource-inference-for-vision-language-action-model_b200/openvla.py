"""OpenVLA-7B-shaped path (BASELINE.json configs[4]): the language-model half.

The reference has no source for this model — `scripts/benchmark_hf_vla.py:100-109` loads `openvla/openvla-7b` with
`trust_remote_code=True` and times `model.predict_action(**inputs, unnorm_key=..., do_sample=False)` (:141-197).
That call (public Prismatic/OpenVLA architecture) runs the fused SigLIP+DINOv2 backbone and the MLP projector, builds
`[BOS] + 256 patch embeddings + prompt tokens`, and lets a `LlamaForCausalLM` generate `action_dim` = 7 tokens greedily
with a KV cache; each token is mapped to one of 256 uniform bins on [-1, 1] and de-normalised with q01/q99.

Built here: the Llama-2-7B-shaped decoder (`LlamaDecoder`, csrc/llm_engine.cu behind include/blurr_llm.h) — prefill of
the multimodal prompt embeddings + greedy decode — and the action de-tokeniser (`ActionDetokenizer`).  NOT built: the
fused vision backbone and projector (the caller supplies the projected patch embeddings; the benchmark uses random ones
of the right shape — at 7B the vision tower is ~5 % of the step's weight bytes).  PARITY: the decoder is pinned against
transformers' `LlamaForCausalLM` (eager attention) in tests/test_gpu_llm.py; against the reference's remote code it is
UNPINNED (not vendored, not downloadable)."""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import capi


@dataclass
class LlamaShapedConfig:
    num_layers: int = 32
    hidden: int = 4096
    num_heads: int = 32
    num_kv_heads: int = 32
    head_dim: int = 128
    intermediate: int = 11008
    vocab: int = 32064            # Llama-2's 32000 padded to a multiple of 64 (OpenVLA adds <PAD>)
    max_positions: int = 320      # 1 + 256 patches + prompt + 7 action tokens fits
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0


def openvla_7b_config() -> LlamaShapedConfig:
    return LlamaShapedConfig()


def rope_tables(inv_freq: torch.Tensor, n_pos: int, dtype: torch.dtype = torch.bfloat16) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos / sin of `LlamaRotaryEmbedding.forward` (modeling_llama.py) for positions 0..n_pos-1, first half of the head
    dim: fp32 angle `inv_freq.float() * position`, cos/sin cast to the activation dtype.  `inv_freq` is the module's own
    buffer (a model cast with `.to(bfloat16)` carries a bf16-rounded one)."""
    pos = torch.arange(n_pos, device=inv_freq.device, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq.float()[None, :]
    return freqs.cos().to(dtype).float().contiguous(), freqs.sin().to(dtype).float().contiguous()


def default_inv_freq(head_dim: int, theta: float, device) -> torch.Tensor:
    """`ROPE_INIT_FUNCTIONS['default']`: 1 / theta^(2i / d), computed in fp32 on int64 indices."""
    return 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64, device=device).float() / head_dim))


class LlamaDecoder:
    """Greedy generation of a Llama-shaped decoder on one B200 through `include/blurr_llm.h`."""

    def __init__(self, cfg: LlamaShapedConfig, device, max_batch: int = 1):
        self.lib = capi.load_library()
        self.cfg = cfg
        self.device = torch.device(device)
        self.max_batch = max_batch
        self.handle = C.c_void_p()
        c = capi.LlmConfigC(capi.LLM_ABI_VERSION, cfg.num_layers, cfg.hidden, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim,
                            cfg.intermediate, cfg.vocab, cfg.max_positions, cfg.rms_eps)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        capi.check(self.lib.blurr_llm_create(C.byref(c), index, max_batch, C.byref(self.handle)))

    @classmethod
    def from_state_dict(cls, cfg: LlamaShapedConfig, state_dict: Dict[str, torch.Tensor], device, max_batch: int = 1,
                        inv_freq: Optional[torch.Tensor] = None, prefix: str = "") -> "LlamaDecoder":
        """`state_dict`: keys of transformers' LlamaForCausalLM (optionally under `prefix`, e.g. "language_model.")."""
        self = cls(cfg, device, max_batch)
        try:
            torch.cuda.current_stream(self.device).synchronize()
            for key, t in state_dict.items():
                if prefix and not key.startswith(prefix):
                    continue
                name = key[len(prefix):]
                if name.endswith("rotary_emb.inv_freq"):
                    continue
                self.set_weight(name, t)
            if inv_freq is None:
                inv_freq = default_inv_freq(cfg.head_dim, cfg.rope_theta, self.device)
            cos, sin = rope_tables(inv_freq.to(self.device), cfg.max_positions)
            torch.cuda.current_stream(self.device).synchronize()
            capi.check(self.lib.blurr_llm_set_rope_table(self.handle, C.c_void_p(cos.data_ptr()), C.c_void_p(sin.data_ptr()),
                                                         cfg.max_positions))
            capi.check(self.lib.blurr_llm_finalize(self.handle))
        except Exception:
            self.close()
            raise
        return self

    def set_weight(self, name: str, tensor: torch.Tensor):
        t = tensor.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        torch.cuda.current_stream(self.device).synchronize()
        shape = (C.c_int64 * t.dim())(*t.shape)
        capi.check(self.lib.blurr_llm_set_weight(self.handle, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))

    def embed(self, ids: torch.Tensor) -> torch.Tensor:
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty((*ids.shape, self.cfg.hidden), device=self.device, dtype=torch.bfloat16)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self.lib.blurr_llm_embed(self.handle, C.c_void_p(stream), C.c_void_p(ids.data_ptr()), ids.numel(),
                                            C.c_void_p(out.data_ptr())))
        return out

    def generate(self, inputs_embeds: torch.Tensor, n_new: int, return_logits: bool = False):
        """inputs_embeds: bf16 [B, T, hidden] on the device -> int64 [B, n_new] (and bf16 logits [B, n_new, vocab])."""
        if inputs_embeds.dim() != 3 or inputs_embeds.shape[2] != self.cfg.hidden:
            raise ValueError(f"inputs_embeds must be [B, T, {self.cfg.hidden}]")
        if inputs_embeds.dtype != torch.bfloat16 or inputs_embeds.device.type != "cuda":
            raise ValueError("inputs_embeds must be a bf16 CUDA tensor")
        x = inputs_embeds.contiguous()
        B, T = x.shape[0], x.shape[1]
        ids = torch.empty((B, n_new), device=self.device, dtype=torch.int64)
        logits = torch.empty((B, n_new, self.cfg.vocab), device=self.device, dtype=torch.bfloat16) if return_logits else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self.lib.blurr_llm_generate(self.handle, C.c_void_p(stream), B, T, C.c_void_p(x.data_ptr()), n_new,
                                               C.c_void_p(ids.data_ptr()),
                                               C.c_void_p(logits.data_ptr()) if logits is not None else None))
        return (ids, logits) if return_logits else ids

    def check(self):
        capi.check(self.lib.blurr_llm_check(self.handle, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def set_option(self, name: str, value: int):
        capi.check(self.lib.blurr_llm_set_option(self.handle, name.encode(), int(value)))

    def trace_report(self) -> str:
        buf = C.create_string_buffer(1 << 20)
        capi.check(self.lib.blurr_llm_trace_report(self.handle, buf, len(buf)))
        return buf.value.decode()

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.blurr_llm_last_launch_count(self.handle))

    @property
    def weight_bytes_per_token(self) -> int:
        return int(self.lib.blurr_llm_weight_bytes_per_token(self.handle))

    def close(self):
        if self.handle:
            self.lib.blurr_llm_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ActionDetokenizer:
    """OpenVLA's `predict_action` tail (public modeling_prismatic.py): the last `action_dim` generated ids ->
    `vocab_size - id` -> bin index -> bin centre on [-1, 1] -> de-normalised with the dataset's q01 / q99 where `mask`."""

    def __init__(self, vocab_size: int = 32000, n_bins: int = 256):
        self.vocab_size = vocab_size
        bins = np.linspace(-1.0, 1.0, n_bins)
        self.bin_centers = (bins[:-1] + bins[1:]) / 2.0

    def normalized(self, token_ids: np.ndarray) -> np.ndarray:
        discretized = self.vocab_size - np.asarray(token_ids, dtype=np.int64)
        discretized = np.clip(discretized - 1, a_min=0, a_max=self.bin_centers.shape[0] - 1)
        return self.bin_centers[discretized]

    def actions(self, token_ids: np.ndarray, q01: np.ndarray, q99: np.ndarray, mask: Optional[np.ndarray] = None) -> np.ndarray:
        norm = self.normalized(token_ids)
        q01, q99 = np.asarray(q01, dtype=np.float64), np.asarray(q99, dtype=np.float64)
        mask = np.ones_like(q01, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        return np.where(mask, 0.5 * (norm + 1) * (q99 - q01) + q01, norm)


def build_prompt_embeds(decoder: LlamaDecoder, input_ids: torch.Tensor, patch_embeds: torch.Tensor) -> torch.Tensor:
    """Prismatic's multimodal prompt: `[BOS] + projected patch embeddings + the remaining prompt tokens`.
    input_ids [B, S] (first column BOS), patch_embeds bf16 [B, P, hidden] -> bf16 [B, 1 + P + S - 1, hidden]."""
    tok = decoder.embed(input_ids)
    return torch.cat([tok[:, :1], patch_embeds.to(tok.dtype), tok[:, 1:]], dim=1).contiguous()


def synthetic_llama_state_dict(cfg: LlamaShapedConfig, device, seed: int = 0, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Random-init weights with transformers' key names and init scale (normal std 0.02; norms at 1), generated on the
    device layer by layer (13.5 GB for the 7B shape)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    QW, KVW = cfg.num_heads * cfg.head_dim, cfg.num_kv_heads * cfg.head_dim

    def w(rows, cols):
        return (torch.randn((rows, cols), device=device, dtype=torch.float32, generator=g) * 0.02).to(dtype)

    sd = {"model.embed_tokens.weight": w(cfg.vocab, cfg.hidden)}
    for l in range(cfg.num_layers):
        p = f"model.layers.{l}."
        sd[p + "self_attn.q_proj.weight"] = w(QW, cfg.hidden)
        sd[p + "self_attn.k_proj.weight"] = w(KVW, cfg.hidden)
        sd[p + "self_attn.v_proj.weight"] = w(KVW, cfg.hidden)
        sd[p + "self_attn.o_proj.weight"] = w(cfg.hidden, QW)
        sd[p + "mlp.gate_proj.weight"] = w(cfg.intermediate, cfg.hidden)
        sd[p + "mlp.up_proj.weight"] = w(cfg.intermediate, cfg.hidden)
        sd[p + "mlp.down_proj.weight"] = w(cfg.hidden, cfg.intermediate)
        sd[p + "input_layernorm.weight"] = torch.ones(cfg.hidden, device=device, dtype=dtype)
        sd[p + "post_attention_layernorm.weight"] = torch.ones(cfg.hidden, device=device, dtype=dtype)
    sd["model.norm.weight"] = torch.ones(cfg.hidden, device=device, dtype=dtype)
    sd["lm_head.weight"] = w(cfg.vocab, cfg.hidden)
    return sd
