"""Per-episode state + device-side observation preprocessing.

The reference's control loop (`third_party/open_pi_zero/src/agent/eval.py:162-239`) rebuilds, on the
host and at every step, everything that only depends on the instruction (token ids, attention mask,
the block-causal mask, the position ids), resizes the camera frame with cv2 on the CPU
(`src/agent/env_adapter/simpler.py:52-98`), normalises it in fp32 (`src/model/vla/processing.py:96-136`)
and copies all eight tensors to the GPU.  `Episode` keeps the instruction-dependent tensors resident on
the device for the whole episode and runs the per-step part — Lanczos resize, normalisation, bf16 cast,
proprio normalisation — as CUDA kernels with bit-identical results, so a control step is: one H2D copy
of the raw uint8 frame and the raw proprio vector, two small kernels, the model graph.

`FramePreprocessor` / `normalize_proprio` wrap the C ABI (`blurr_preproc_*`,
`blurr_op_normalize_proprio` in include/blurr_pi0.h).  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import capi


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class FramePreprocessor:
    """cv2.resize(INTER_LANCZOS4) + VLAProcessor normalisation + bf16 cast for one frame geometry."""

    def __init__(self, src_h: int, src_w: int, size: Sequence[int] = (224, 224), device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FramePreprocessor runs on a CUDA device only (no CPU fallback)")
        self.src_h, self.src_w = int(src_h), int(src_w)
        self.dst_w, self.dst_h = int(size[0]), int(size[1])          # cv2 convention: (width, height)
        self.lib = capi.load_library()
        handle = C.c_void_p()
        capi.check(self.lib.blurr_preproc_create(self.device.index if self.device.index is not None else torch.cuda.current_device(), self.src_h, self.src_w, self.dst_h, self.dst_w,
                                                 C.byref(handle)))
        self.handle = handle

    def close(self):
        if getattr(self, "handle", None):
            self.lib.blurr_preproc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tables(self):
        """(x_ofs, x_alpha[dst_w,8], y_ofs, y_alpha[dst_h,8]) as numpy arrays (OpenCV's fixed-point tables)."""
        xo = np.zeros(self.dst_w, np.int32); xa = np.zeros((self.dst_w, 8), np.int16)
        yo = np.zeros(self.dst_h, np.int32); ya = np.zeros((self.dst_h, 8), np.int16)
        capi.check(self.lib.blurr_preproc_tables(self.handle, xo.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 xa.ctypes.data_as(C.POINTER(C.c_int16)),
                                                 yo.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 ya.ctypes.data_as(C.POINTER(C.c_int16))))
        return xo, xa, yo, ya

    def __call__(self, frames_u8: torch.Tensor, out: Optional[torch.Tensor] = None, return_resized: bool = False):
        """frames_u8: CUDA uint8 [H,W,3] or [B,H,W,3] (rows contiguous).  Returns bf16 pixel_values
        [B,3,h,w]; with `return_resized` also the uint8 [B,h,w,3] cv2-equivalent resize."""
        if frames_u8.dtype != torch.uint8 or frames_u8.device != self.device:
            raise ValueError("frames must be uint8 tensors on the preprocessor's device")
        if frames_u8.dim() == 3:
            frames_u8 = frames_u8[None]
        B, H, W, Cn = frames_u8.shape
        if (H, W, Cn) != (self.src_h, self.src_w, 3):
            raise ValueError(f"frame geometry {(H, W, Cn)} != {(self.src_h, self.src_w, 3)}")
        if frames_u8.stride(3) != 1 or frames_u8.stride(2) != 3:
            frames_u8 = frames_u8.contiguous()
        if out is None:
            out = torch.empty((B, 3, self.dst_h, self.dst_w), device=self.device, dtype=torch.bfloat16)
        resized = torch.empty((B, self.dst_h, self.dst_w, 3), device=self.device, dtype=torch.uint8) if return_resized else None
        capi.check(self.lib.blurr_preproc_frame(
            self.handle, _stream_ptr(self.device), C.c_void_p(frames_u8.data_ptr()), frames_u8.stride(1),
            B, frames_u8.stride(0), C.c_void_p(out.data_ptr()),
            C.c_void_p(resized.data_ptr()) if resized is not None else None))
        return (out, resized) if return_resized else out


def normalize_proprio(raw: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, kind: str = "bound",
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """raw: CUDA float64 [n, dim]; (lo, hi) = (p01, p99) for "bound", (mean, std) for "gaussian".
    Returns bf16 [n, dim] = bf16(float32(normalise_float64(raw))) exactly as the reference's adapter."""
    if raw.dtype != torch.float64 or not raw.is_cuda:
        raise ValueError("raw proprio must be a CUDA float64 tensor")
    n, dim = raw.shape
    lib = capi.load_library()
    if out is None:
        out = torch.empty((n, dim), device=raw.device, dtype=torch.bfloat16)
    capi.check(lib.blurr_op_normalize_proprio(_stream_ptr(raw.device), C.c_void_p(raw.data_ptr()), C.c_void_p(lo.data_ptr()),
                                              C.c_void_p(hi.data_ptr()), {"bound": 0, "gaussian": 1}[kind], n, dim,
                                              C.c_void_p(out.data_ptr())))
    return out


class Episode:
    """Everything of a rollout that does not change between control steps, resident on the device.

    model            a `blurr_b200.pizero.PiZeroInference`
    input_ids        [B, max_image_text_tokens] int64, attention_mask [B, max_image_text_tokens]
                     (the VLAProcessor/tokenizer output for the instruction; tokenisation stays on the host
                     and happens once per episode)
    frame_hw         geometry of the raw camera frames
    proprio_stats    {"p01": [...], "p99": [...]} ("bound") or {"mean": [...], "std": [...]} ("gaussian")
    `step(frame_u8, raw_proprio)` = env_adapter.preprocess + mask/position building + `.to(device)` +
    `model(**inputs)` of eval.py:170-218, returning the [B, horizon, action_dim] actions on the device.
    """

    def __init__(self, model, input_ids: torch.Tensor, attention_mask: torch.Tensor, frame_hw: Sequence[int],
                 proprio_stats: dict, proprio_normalization_type: str = "bound", image_size: Sequence[int] = (224, 224)):
        self.model = model
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("Episode needs the model on a CUDA device")
        self.device = dev
        dtype = torch.bfloat16
        # eval.py:176-183 — once per episode instead of once per step
        causal_mask, vlm_pos, proprio_pos, action_pos = model.build_causal_mask_and_position_ids(attention_mask, dtype=dtype)
        itp_mask, action_mask = model.split_full_mask_into_submasks(causal_mask)
        self.input_ids = input_ids.to(dev)
        self.image_text_proprio_mask = itp_mask.to(dev)
        self.action_mask = action_mask.to(dev)
        self.vlm_position_ids = vlm_pos.to(dev)
        self.proprio_position_ids = proprio_pos.to(dev)
        self.action_position_ids = action_pos.to(dev)
        self.batch = int(input_ids.shape[0])
        self.kind = proprio_normalization_type
        keys = ("p01", "p99") if self.kind == "bound" else ("mean", "std")
        self.lo = torch.as_tensor(np.asarray(proprio_stats[keys[0]], dtype=np.float64), device=dev)
        self.hi = torch.as_tensor(np.asarray(proprio_stats[keys[1]], dtype=np.float64), device=dev)
        self.pre = FramePreprocessor(frame_hw[0], frame_hw[1], image_size, dev)
        dim = self.lo.numel()
        # pinned staging + device buffers reused every step (fixed addresses)
        self._frame_host = torch.empty((self.batch, frame_hw[0], frame_hw[1], 3), dtype=torch.uint8).pin_memory()
        self._frame_dev = torch.empty_like(self._frame_host, device=dev)
        self._prop_host = torch.empty((self.batch, dim), dtype=torch.float64).pin_memory()
        self._prop_dev = torch.empty_like(self._prop_host, device=dev)
        self.pixel_values = torch.empty((self.batch, 3, self.pre.dst_h, self.pre.dst_w), device=dev, dtype=dtype)
        self.proprios = torch.empty((self.batch, 1, dim), device=dev, dtype=dtype)
        # the pinned staging buffers are rewritten by the host every step: an event recorded after each step's H2D copies
        # is waited for before the next rewrite, so steps issued without a synchronise cannot corrupt a copy in flight
        self._staged = torch.cuda.Event()
        self._staged_pending = False

    def preprocess(self, frame_u8, raw_proprio):
        """Raw observation -> the model's `pixel_values` / `proprios` on the device (asynchronous)."""
        if self._staged_pending:
            self._staged.synchronize()
            self._staged_pending = False
        f = torch.as_tensor(frame_u8)
        if f.dim() == 3:
            f = f[None]
        if f.is_cuda:
            frames = f
        else:
            if not f.is_pinned():                   # pageable memory: stage through the pinned buffer
                self._frame_host.copy_(f)
                f = self._frame_host
            self._frame_dev.copy_(f, non_blocking=True)
            frames = self._frame_dev
        p = torch.as_tensor(raw_proprio, dtype=torch.float64)
        if p.dim() == 1:
            p = p[None]
        if p.is_cuda:
            prop = p
        else:
            if not p.is_pinned():
                self._prop_host.copy_(p)
                p = self._prop_host
            self._prop_dev.copy_(p, non_blocking=True)
            prop = self._prop_dev
        if frames is self._frame_dev or prop is self._prop_dev:
            self._staged.record(torch.cuda.current_stream(self.device))
            self._staged_pending = True
        self.pre(frames, out=self.pixel_values)
        normalize_proprio(prop, self.lo, self.hi, self.kind, out=self.proprios.view(self.batch, -1))
        return self.pixel_values, self.proprios

    def step(self, frame_u8, raw_proprio, noise: Optional[torch.Tensor] = None, check: bool = False) -> torch.Tensor:
        """One control step from a raw observation.  `check=True` synchronises and raises on a device-side error
        (`PiZero.check`); without it a failed step is still visible: its actions are NaN."""
        pixel_values, proprios = self.preprocess(frame_u8, raw_proprio)
        kwargs = {} if noise is None else {"noise": noise}
        with torch.inference_mode():
            out = self.model(self.input_ids, pixel_values, self.image_text_proprio_mask, self.action_mask,
                             self.vlm_position_ids, self.proprio_position_ids, self.action_position_ids, proprios,
                             **kwargs)
        if check:
            self.model.check()
        return out

    def close(self):
        self.pre.close()
