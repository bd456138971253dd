"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Drives the *unmodified* reference (`/root/reference/third_party/open_pi_zero`) so that the
restatement in `oracle/pi0_oracle.py` can be pinned against it and golden vectors can be
generated (`tests/golden/make_golden.py`).  The reference imports `hydra` and `omegaconf`
(`src/model/vla/pizero.py:13`, `src/model/vla/joint_model.py:18`), which are not installed
here; the two call sites (`hydra.utils.instantiate(cfg.x)`, `OmegaConf.merge(a, b)`) are
served by the small stubs below.  Nothing from the reference is copied: it is imported from
where it lies.  `/root/reference` exists only in the build container; for the GPU box
`__graft_entry__.build()` places an unmodified copy of the reference's `open_pi_zero/src` tree in the
git-ignored `baseline/_ref/` (it travels with the snapshot like the built `.so`), which only
`bench.py --impl reference` uses there: the reference itself timed on the box's host cores.
"""

from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Optional

import torch

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SHIPPED = os.path.join(_REPO_ROOT, "baseline", "_ref")      # git-ignored copy made by __graft_entry__.build(); travels to the GPU box


def _pick_ref_root() -> str:
    env = os.environ.get("BLURR_REF_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference"):
        return "/root/reference"
    return _SHIPPED


REF_ROOT = _pick_ref_root()
OPZ_ROOT = os.path.join(REF_ROOT, "third_party", "open_pi_zero")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(OPZ_ROOT, "src", "model", "vla", "pizero.py"))


def _install_stubs() -> None:
    from blurr_b200.config import AttrDict, merge

    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")

        class OmegaConf:  # only what joint_model.py:329 uses
            @staticmethod
            def merge(a, b):
                return merge(a, b)

        om.OmegaConf = OmegaConf
        om.DictConfig = AttrDict
        sys.modules["omegaconf"] = om

    if "hydra" not in sys.modules:
        hy = types.ModuleType("hydra")
        hu = types.ModuleType("hydra.utils")

        def instantiate(node):  # pizero.py:73-77: `_target_` import + cls(**kwargs)
            kwargs = {k: v for k, v in node.items() if k != "_target_"}
            mod_name, cls_name = node["_target_"].rsplit(".", 1)
            return getattr(importlib.import_module(mod_name), cls_name)(**kwargs)

        hu.instantiate = instantiate
        hy.utils = hu
        sys.modules["hydra"] = hy
        sys.modules["hydra.utils"] = hu


def import_reference():
    """Returns the reference `src.model.vla.pizero` module."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    if OPZ_ROOT not in sys.path:
        sys.path.insert(0, OPZ_ROOT)
    return importlib.import_module("src.model.vla.pizero")


def build_reference_model(cfg, seed: Optional[int] = 0, dtype: torch.dtype = torch.float32):
    """`PiZeroInference(cfg)` with PyTorch default inits under `torch.manual_seed(seed)`
    (SURVEY.md Appendix F recipe), frozen, eval, cast like `scripts/benchmark_pi0.py:140-146`."""
    pz = import_reference()
    if seed is not None:
        torch.manual_seed(seed)
    model = pz.PiZeroInference(cfg, use_ddp=False)
    model.freeze_all_weights()
    model.to(dtype)
    model.eval()
    return model


def load_reference_model(cfg, state_dict, dtype: torch.dtype = torch.float32):
    """The reference `PiZeroInference` carrying the given weights: built on the meta device (skips ~45 s of default
    initialisation), weights adopted with `load_state_dict(assign=True)`, the two non-persistent buffers rebuilt the way
    the reference's constructors compute them, then frozen / cast / eval like `scripts/benchmark_pi0.py:140-146`."""
    pz = import_reference()
    with torch.device("meta"):
        model = pz.PiZeroInference(cfg, use_ddp=False)
    model.load_state_dict(state_dict, strict=True, assign=True)
    for m in model.modules():   # non-persistent buffers are not in the state_dict
        if type(m).__name__ == "GemmaRotaryEmbedding":
            m.inv_freq = 1.0 / (m.base ** (torch.arange(0, m.dim, 2, dtype=torch.int64).float() / m.dim))
        if type(m).__name__ == "SiglipVisionEmbeddings":
            m.position_ids = torch.arange(m.num_positions).expand((1, -1))
    model.freeze_all_weights()
    model.to(dtype)
    model.eval()
    return model


class patched_randn:
    """Inject fixed flow noise: the reference draws it inside `infer_action`
    (`pizero.py:511-513`) via `torch.randn`, so `torch.randn` is swapped for the duration of
    the call.  Returns a clone because the tensor is updated in place (:537)."""

    def __init__(self, noise: torch.Tensor):
        self.noise = noise
        self._orig = None

    def __enter__(self):
        self._orig = torch.randn
        noise = self.noise

        def fake_randn(*size, device=None, dtype=None, **kw):
            return noise.to(device=device, dtype=dtype).clone()

        torch.randn = fake_randn
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig
        return False


def install_taps(model, tap):
    """Forward hooks mirroring the tap names of `pi0_oracle` (per-layer activations)."""
    handles = []
    vt = model.vision_tower.vision_model
    handles.append(vt.embeddings.register_forward_hook(
        lambda m, i, o: tap("siglip.embeddings", o)))
    for l, layer in enumerate(vt.encoder.layers):
        handles.append(layer.register_forward_hook(
            lambda m, i, o, l=l: tap(f"siglip.layer{l}", o)))
    handles.append(vt.post_layernorm.register_forward_hook(
        lambda m, i, o: tap("siglip.post_layernorm", o)))
    handles.append(model.multi_modal_projector.register_forward_hook(
        lambda m, i, o: tap("projector", o)))
    return handles
