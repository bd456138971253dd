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


# ---------------------------------------------------------------------------------------------------------------------
# Vision side: generic ViT towers + projector (include/blurr_vit.h)
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class VitShapedConfig:
    num_layers: int            # blocks to RUN (OpenVLA takes the second-to-last block's output: depth - 1)
    hidden: int
    num_heads: int
    mlp_dim: int
    num_prefix_tokens: int = 0
    use_layerscale: bool = False
    gelu_erf: bool = False
    image_size: int = 224
    patch_size: int = 14
    ln_eps: float = 1e-6


def dinov2_large_reg4_config(depth: int = 24) -> VitShapedConfig:
    """`vit_large_patch14_reg4_dinov2` at 224 px: 24 blocks, 1024 wide, cls + 4 registers, LayerScale, exact GELU."""
    return VitShapedConfig(num_layers=depth - 1, hidden=1024, num_heads=16, mlp_dim=4096, num_prefix_tokens=5,
                           use_layerscale=True, gelu_erf=True)


def siglip_so400m_config(depth: int = 27) -> VitShapedConfig:
    """`vit_so400m_patch14_siglip_224`: 27 blocks, 1152 wide, no prefix tokens, tanh GELU."""
    return VitShapedConfig(num_layers=depth - 1, hidden=1152, num_heads=16, mlp_dim=4304)


class VitEncoder:
    """One ViT tower on one B200 through `include/blurr_vit.h`; `forward` returns the patch-token features."""

    def __init__(self, cfg: VitShapedConfig, device, max_batch: int = 1):
        self.lib = capi.load_library()
        self.cfg = cfg
        self.device = torch.device(device)
        self.max_batch = max_batch
        self.handle = C.c_void_p()
        c = capi.VitConfigC(capi.VIT_ABI_VERSION, cfg.num_layers, cfg.hidden, cfg.num_heads, cfg.mlp_dim, cfg.image_size,
                            cfg.patch_size, cfg.num_prefix_tokens, int(cfg.use_layerscale), int(cfg.gelu_erf), cfg.ln_eps)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        capi.check(self.lib.blurr_vit_create(C.byref(c), index, max_batch, C.byref(self.handle)))
        self.n_patches = (cfg.image_size // cfg.patch_size) ** 2

    def set_weight(self, key: str, tensor: torch.Tensor):
        t = tensor.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        torch.cuda.current_stream(self.device).synchronize()
        shape = (C.c_int64 * t.dim())(*t.shape)
        capi.check(self.lib.blurr_vit_set_weight(self.handle, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))

    def finalize(self):
        capi.check(self.lib.blurr_vit_finalize(self.handle))

    def set_option(self, name: str, value: int):
        capi.check(self.lib.blurr_vit_set_option(self.handle, name.encode(), int(value)))

    @classmethod
    def from_hf_dinov2(cls, model, device, max_batch: int = 1, num_layers: Optional[int] = None) -> "VitEncoder":
        """From a transformers `Dinov2WithRegistersModel` (image_size 224 so that its position embeddings are native)."""
        hc = model.config
        depth = hc.num_hidden_layers
        cfg = VitShapedConfig(num_layers=num_layers or depth - 1, hidden=hc.hidden_size, num_heads=hc.num_attention_heads,
                              mlp_dim=int(hc.hidden_size * hc.mlp_ratio), num_prefix_tokens=1 + hc.num_register_tokens,
                              use_layerscale=True, gelu_erf=True, image_size=hc.image_size, patch_size=hc.patch_size,
                              ln_eps=hc.layer_norm_eps)
        self = cls(cfg, device, max_batch)
        sd = {k: v.to(torch.bfloat16) for k, v in model.state_dict().items()}
        e = "embeddings."
        self.set_weight("patch.weight", sd[e + "patch_embeddings.projection.weight"].flatten(1))
        self.set_weight("patch.bias", sd[e + "patch_embeddings.projection.bias"])
        pos = sd[e + "position_embeddings"][0]
        self.set_weight("pos", pos[1:])
        prefix = torch.cat([sd[e + "cls_token"][0] + pos[:1], sd[e + "register_tokens"][0]], dim=0)     # bf16 add, like HF
        self.set_weight("prefix", prefix)
        for l in range(cfg.num_layers):
            p, q = f"encoder.layer.{l}.", f"layers.{l}."
            for a, b in (("norm1", "ln1"), ("norm2", "ln2"), ("mlp.fc1", "fc1"), ("mlp.fc2", "fc2"),
                         ("attention.attention.query", "q"), ("attention.attention.key", "k"),
                         ("attention.attention.value", "v"), ("attention.output.dense", "o")):
                self.set_weight(q + b + ".weight", sd[p + a + ".weight"])
                self.set_weight(q + b + ".bias", sd[p + a + ".bias"])
            self.set_weight(q + "ls1", sd[p + "layer_scale1.lambda1"])
            self.set_weight(q + "ls2", sd[p + "layer_scale2.lambda1"])
        self.finalize()
        return self

    @classmethod
    def from_hf_siglip(cls, model, device, max_batch: int = 1, num_layers: Optional[int] = None) -> "VitEncoder":
        """From a transformers `SiglipVisionModel`."""
        hc = model.config
        cfg = VitShapedConfig(num_layers=num_layers or hc.num_hidden_layers - 1, hidden=hc.hidden_size,
                              num_heads=hc.num_attention_heads, mlp_dim=hc.intermediate_size, image_size=hc.image_size,
                              patch_size=hc.patch_size, ln_eps=hc.layer_norm_eps)
        self = cls(cfg, device, max_batch)
        sd = {k: v.to(torch.bfloat16) for k, v in model.state_dict().items()}
        e = "vision_model.embeddings."
        self.set_weight("patch.weight", sd[e + "patch_embedding.weight"].flatten(1))
        self.set_weight("patch.bias", sd[e + "patch_embedding.bias"])
        self.set_weight("pos", sd[e + "position_embedding.weight"])
        for l in range(cfg.num_layers):
            p, q = f"vision_model.encoder.layers.{l}.", f"layers.{l}."
            for a, b in (("layer_norm1", "ln1"), ("layer_norm2", "ln2"), ("mlp.fc1", "fc1"), ("mlp.fc2", "fc2"),
                         ("self_attn.q_proj", "q"), ("self_attn.k_proj", "k"), ("self_attn.v_proj", "v"),
                         ("self_attn.out_proj", "o")):
                self.set_weight(q + b + ".weight", sd[p + a + ".weight"])
                self.set_weight(q + b + ".bias", sd[p + a + ".bias"])
        self.finalize()
        return self

    @classmethod
    def synthetic(cls, cfg: VitShapedConfig, device, max_batch: int = 1, seed: int = 0) -> "VitEncoder":
        """Random-init tower of the given shape (benchmarks)."""
        self = cls(cfg, device, max_batch)
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        H = cfg.hidden

        def w(*shape, std=0.02):
            return (torch.randn(shape, device=device, generator=g) * std).to(torch.bfloat16)

        self.set_weight("patch.weight", w(H, 3 * cfg.patch_size ** 2))
        self.set_weight("patch.bias", w(H))
        self.set_weight("pos", w(self.n_patches, H))
        if cfg.num_prefix_tokens:
            self.set_weight("prefix", w(cfg.num_prefix_tokens, H))
        ones = torch.ones(H, device=device, dtype=torch.bfloat16)
        for l in range(cfg.num_layers):
            q = f"layers.{l}."
            for name, (n, k) in (("q", (H, H)), ("k", (H, H)), ("v", (H, H)), ("o", (H, H)), ("fc1", (cfg.mlp_dim, H)),
                                 ("fc2", (H, cfg.mlp_dim))):
                self.set_weight(q + name + ".weight", w(n, k))
                self.set_weight(q + name + ".bias", w(n))
            for name in ("ln1", "ln2"):
                self.set_weight(q + name + ".weight", ones)
                self.set_weight(q + name + ".bias", torch.zeros_like(ones))
            if cfg.use_layerscale:
                self.set_weight(q + "ls1", ones)
                self.set_weight(q + "ls2", ones)
        self.finalize()
        return self

    def forward(self, pixel_values: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pixel_values bf16 [B, 3, 224, 224] -> bf16 [B, n_patches, hidden] (or the first `hidden` columns of `out`)."""
        if pixel_values.dtype != torch.bfloat16 or pixel_values.device.type != "cuda" or pixel_values.dim() != 4:
            raise ValueError("pixel_values must be a bf16 CUDA tensor [B, 3, H, W]")
        B = pixel_values.shape[0]
        if out is None:
            out = torch.empty((B, self.n_patches, self.cfg.hidden), device=self.device, dtype=torch.bfloat16)
        if out.stride(-1) != 1 or out.stride(0) != self.n_patches * out.stride(1):
            raise ValueError("out must be [B, n_patches, >= hidden] with contiguous rows")
        strides = (C.c_int64 * 4)(*pixel_values.stride())
        stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self.lib.blurr_vit_forward(self.handle, C.c_void_p(stream), B, C.c_void_p(pixel_values.data_ptr()), strides,
                                              C.c_void_p(out.data_ptr()), out.stride(1)))
        return out

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.blurr_vit_last_launch_count(self.handle))

    def close(self):
        if self.handle:
            self.lib.blurr_vit_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MlpProjector:
    """OpenVLA's fused-backbone projector: Linear -> GELU -> Linear -> GELU -> Linear (exact GELU), `include/blurr_vit.h`."""

    def __init__(self, dims, device, max_rows: int):
        self.lib = capi.load_library()
        self.dims = list(dims)
        self.device = torch.device(device)
        self.handle = C.c_void_p()
        arr = (C.c_int32 * len(self.dims))(*self.dims)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        capi.check(self.lib.blurr_mlp_create(arr, len(self.dims) - 1, index, max_rows, C.byref(self.handle)))

    def set_layer(self, i: int, weight: torch.Tensor, bias: torch.Tensor):
        w = weight.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        b = bias.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        assert tuple(w.shape) == (self.dims[i + 1], self.dims[i]) and tuple(b.shape) == (self.dims[i + 1],)
        torch.cuda.current_stream(self.device).synchronize()
        capi.check(self.lib.blurr_mlp_set_weight(self.handle, i, C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x bf16 [..., dims[0]] contiguous -> bf16 [..., dims[-1]]."""
        x2 = x.reshape(-1, self.dims[0])
        if not x2.is_contiguous() or x2.dtype != torch.bfloat16:
            raise ValueError("x must be contiguous bf16")
        y = torch.empty((x2.shape[0], self.dims[-1]), device=self.device, dtype=torch.bfloat16)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self.lib.blurr_mlp_forward(self.handle, C.c_void_p(stream), x2.shape[0], C.c_void_p(x2.data_ptr()),
                                              self.dims[0], C.c_void_p(y.data_ptr()), self.dims[-1]))
        return y.reshape(*x.shape[:-1], self.dims[-1])

    def close(self):
        if self.handle:
            self.lib.blurr_mlp_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FusedVisionBackbone:
    """DINOv2 + SigLIP towers on the same image, patch features concatenated (DINOv2 first, as in Prismatic's
    `dinosiglip` backbone) and projected to the language model's width."""

    def __init__(self, dino: VitEncoder, siglip: VitEncoder, projector: MlpProjector, concurrent: bool = True):
        self.dino, self.siglip, self.projector = dino, siglip, projector
        self.width = dino.cfg.hidden + siglip.cfg.hidden
        # the two towers are independent chains of small, latency-bound kernels: run them side by side on two streams
        self._side = torch.cuda.Stream(device=dino.device) if concurrent else None

    def forward(self, pixel_values_dino: torch.Tensor, pixel_values_siglip: torch.Tensor) -> torch.Tensor:
        B = pixel_values_dino.shape[0]
        feats = torch.empty((B, self.dino.n_patches, self.width), device=pixel_values_dino.device, dtype=torch.bfloat16)
        sig_out = feats[:, :, self.dino.cfg.hidden:]
        if self._side is None:
            self.dino.forward(pixel_values_dino, out=feats)
            self.siglip.forward(pixel_values_siglip, out=sig_out)
        else:
            cur = torch.cuda.current_stream(self.dino.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.siglip.forward(pixel_values_siglip, out=sig_out)
            self.dino.forward(pixel_values_dino, out=feats)
            cur.wait_stream(self._side)
            pixel_values_siglip.record_stream(self._side)
            feats.record_stream(self._side)
        return self.projector.forward(feats)
