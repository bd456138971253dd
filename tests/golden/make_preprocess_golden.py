#!/usr/bin/env python3
"""Generates tests/golden/preprocess_golden.json with the REAL dependencies of the reference:
cv2.resize(INTER_LANCZOS4) (simpler.py:59-64) and, when /root/reference is present, the reference's own
`process_images` (src/model/vla/processing.py:47-58).  Inputs are seeded numpy frames, so only digests
and a few sampled values are stored.  Run in the build container:  python tests/golden/make_preprocess_golden.py"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/third_party/open_pi_zero"


from make_preprocess_golden_frames import frame  # noqa: E402

cases = []
for seed, (h, w), kind in [(0, (480, 640), "noise"), (1, (480, 640), "smooth"), (2, (512, 640), "smooth"),
                           (3, (256, 320), "noise"), (4, (128, 128), "smooth"), (5, (224, 224), "noise"),
                           (6, (720, 1280), "smooth"), (7, (333, 517), "noise")]:
    img = frame(seed, h, w, kind)
    small = cv2.resize(img, (224, 224), interpolation=cv2.INTER_LANCZOS4)
    entry = {"seed": seed, "h": h, "w": w, "kind": kind, "cv2_version": cv2.__version__,
             "resized_sha256": hashlib.sha256(small.tobytes()).hexdigest(),
             "resized_samples": small[::37, ::41].reshape(-1).tolist()}
    if os.path.isdir(REF):
        sys.path.insert(0, REF)
        from src.model.vla.processing import IMAGENET_STANDARD_MEAN, IMAGENET_STANDARD_STD, process_images
        px = process_images(torch.as_tensor(small, dtype=torch.uint8).permute(2, 0, 1)[None], rescale_factor=1 / 255.0,
                            image_mean=IMAGENET_STANDARD_MEAN, image_std=IMAGENET_STANDARD_STD).to(torch.bfloat16)
        entry["pixel_values_bf16_sha256"] = hashlib.sha256(px.view(torch.int16).numpy().tobytes()).hexdigest()
    cases.append(entry)
json.dump({"made_by": "tests/golden/make_preprocess_golden.py", "cases": cases},
          open(os.path.join(ROOT, "tests", "golden", "preprocess_golden.json"), "w"), indent=1)
print("wrote", len(cases), "cases")
