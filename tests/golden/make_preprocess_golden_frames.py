"""Seeded synthetic camera frames shared by make_preprocess_golden.py and tests/test_preprocess.py."""
import numpy as np


def frame(seed, h, w, kind):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 90 * np.sin(xx / 17.0 + seed) * np.cos(yy / 23.0) + rng.normal(0, 6, (h, w))
    return np.clip(np.stack([base, base * 0.8 + 30, 255 - base], -1), 0, 255).astype(np.uint8)
