"""CPU restatement of the per-step observation preprocessing that feeds `PiZero.infer_action`
(SURVEY.md §8(f) row 1).  TEST INFRASTRUCTURE ONLY: imported by tests/, `__graft_entry__.smoke()` and
bench.py's CPU legs, never by the product path.

What it restates, and where the reference does it:
  * frame resize   — `cv2.resize(image, self.image_size, interpolation=cv2.INTER_LANCZOS4)`,
                     third_party/open_pi_zero/src/agent/env_adapter/simpler.py:59-64.  The algorithm is
                     OpenCV's (dependency `opencv-python`, unpinned in the reference's requirements.txt:16;
                     4.13.0 in this image): modules/imgproc/src/resize.cpp — `interpolateLanczos4`, the
                     8-tap fixed-point tables (`INTER_RESIZE_COEF_BITS` = 11, `saturate_cast<short>`),
                     `HResizeLanczos4<uchar,int,short>` with replicated borders and
                     `VResizeLanczos4<..., FixedPtCast<int, uchar, 22>>`.
                     PARITY PINNED: bit-exact against cv2 itself on random frames of many geometries
                     (tests/test_preprocess.py, runs wherever cv2 imports) and against the committed
                     digests in tests/golden/preprocess_golden.json (made by tests/golden/make_preprocess_golden.py).
  * normalisation  — `VLAProcessor.__call__` -> `process_images` (src/model/vla/processing.py:27-58,
                     112-117): uint8 * (1/255.0) -> (x - 0.5) / 0.5 in fp32, then `.to(dtype)`
                     (src/agent/eval.py:187,199); pinned against the reference function itself when
                     /root/reference is present.
  * proprio        — `normalize_bound` / `normalize_gaussian` in float64 numpy
                     (src/agent/env_adapter/base.py:8-18,33-40), `torch.as_tensor(..., float32)[None, None]`
                     (simpler.py:93-95) and `.to(dtype)` (eval.py:194).
"""

from __future__ import annotations

import math

import numpy as np
import torch

_S45 = 0.70710678118654752440084436210485
_CS = ((1.0, 0.0), (-_S45, -_S45), (0.0, 1.0), (_S45, -_S45), (-1.0, 0.0), (_S45, _S45), (0.0, -1.0), (-_S45, _S45))
COEF_BITS = 11            # INTER_RESIZE_COEF_BITS
KSIZE = 8


def lanczos4_coeffs(x: np.float32) -> np.ndarray:
    """resize.cpp `interpolateLanczos4(float x, float* coeffs)`: 8 float32 weights for phase x."""
    x = np.float32(x)
    y0 = -float(np.float32(x + np.float32(3))) * math.pi * 0.25      # `(x+3)` is a float32 sum in the C source
    s0, c0 = math.sin(y0), math.cos(y0)
    co = np.zeros(KSIZE, np.float32)
    total = np.float32(0)
    for i in range(KSIZE):
        y0_ = np.float32(np.float32(x + np.float32(3)) - np.float32(i))
        if abs(y0_) >= np.float32(1e-6):
            y = -float(y0_) * math.pi * 0.25
            co[i] = np.float32((_CS[i][0] * s0 + _CS[i][1] * c0) / (y * y))
        else:
            co[i] = np.float32(1e30)          # phase 0 / 1: a delta on this tap
        total = np.float32(total + co[i])
    inv = np.float32(np.float32(1.0) / total)
    return (co * inv).astype(np.float32)


def lanczos4_tables(src: int, dst: int):
    """Per destination index: first source tap offset (`sx`, taps are sx-3..sx+4) and the 8 int16
    fixed-point weights (resize.cpp, `resize()` table loop with `fixpt`)."""
    scale = 1.0 / (dst / src)             # resize(): scale_x = 1. / inv_scale_x
    ofs = np.zeros(dst, np.int32)
    alpha = np.zeros((dst, KSIZE), np.int16)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        w = lanczos4_coeffs(f) * np.float32(1 << COEF_BITS)
        alpha[d] = np.clip(np.rint(w), -32768, 32767).astype(np.int16)      # saturate_cast<short>(float) = cvRound
        ofs[d] = s
    return ofs, alpha


def resize_lanczos4_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """uint8 [H][W][C] -> uint8 [dst_h][dst_w][C], bit-exact cv2.resize(..., INTER_LANCZOS4)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    xo, xa = lanczos4_tables(W, dst_w)
    yo, ya = lanczos4_tables(H, dst_h)
    taps = np.arange(KSIZE)[None] - (KSIZE // 2 - 1)
    ix = np.clip(xo[:, None] + taps, 0, W - 1)                       # replicated border
    iy = np.clip(yo[:, None] + taps, 0, H - 1)
    hor = np.einsum("hwkc,wk->hwc", img[:, ix, :].astype(np.int64), xa.astype(np.int64))
    ver = np.einsum("hkwc,hk->hwc", hor[iy], ya.astype(np.int64))
    out = (ver + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS)       # FixedPtCast<int, uchar, 22>
    return np.clip(out, 0, 255).astype(np.uint8)


def process_images(images_u8: torch.Tensor) -> torch.Tensor:
    """processing.py:27-58 with rescale 1/255.0 and mean = std = 0.5: uint8 [B,3,H,W] -> fp32."""
    x = images_u8 * (1 / 255.0)
    mean = torch.tensor([0.5, 0.5, 0.5])[None, :, None, None]
    std = torch.tensor([0.5, 0.5, 0.5])[None, :, None, None]
    return (x - mean) / std


def preprocess_frame(frame_hwc_u8: np.ndarray, size=(224, 224), dtype=torch.bfloat16) -> torch.Tensor:
    """simpler.py:59-68 + processing.py + eval.py:187: HWC uint8 frame -> pixel_values [1,3,h,w] in `dtype`."""
    small = resize_lanczos4_u8(frame_hwc_u8, size[1], size[0])
    images = torch.as_tensor(small, dtype=torch.uint8).permute(2, 0, 1)[None]
    return process_images(images).to(dtype)


def normalize_bound(data, data_min, data_max, clip_min=-1, clip_max=1, eps=1e-8):
    """base.py:8-18 (float64 numpy)."""
    ndata = 2 * (data - data_min) / (data_max - data_min + eps) - 1
    return np.clip(ndata, clip_min, clip_max)


def normalize_gaussian(data, mean, std, eps=1e-8):
    """base.py:33-40."""
    return (data - mean) / (std + eps)


def preprocess_proprio(raw: np.ndarray, stats: dict, kind: str = "bound", dtype=torch.bfloat16) -> torch.Tensor:
    """simpler.py:73-95 + eval.py:194: raw float64 proprio [dim] -> [1,1,dim] in `dtype`."""
    raw = np.asarray(raw, dtype=np.float64)
    if kind == "bound":
        p = normalize_bound(raw, np.array(stats["p01"]), np.array(stats["p99"]), clip_min=-1, clip_max=1)
    elif kind == "gaussian":
        p = normalize_gaussian(raw, np.array(stats["mean"]), np.array(stats["std"]))
    else:
        raise ValueError(kind)
    return torch.as_tensor(p, dtype=torch.float32)[None, None].to(dtype)
