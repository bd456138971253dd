#!/usr/bin/env python3
"""Golden vectors for the action post-processing, generated from the UNMODIFIED reference:
`src/utils/geometry.py:euler2axangle` (the vendored transforms3d functions `SimplerAdapter.postprocess` calls,
simpler.py:131) and `BaseEnvAdapter.denormalize_bound`.  Run in the build container (needs /root/reference):
    python tests/golden/make_postprocess_golden.py
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("BLURR_REF_ROOT", "/root/reference/third_party/open_pi_zero")
sys.path.insert(0, REF)
from src.utils.geometry import euler2axangle  # noqa: E402

rng = np.random.default_rng(20261018)
rpy = np.concatenate([
    rng.uniform(-np.pi, np.pi, (96, 3)),
    rng.uniform(-0.2, 0.2, (24, 3)),                   # the range de-normalised Bridge/Fractal actions live in
    [[0, 0, 0], [1e-9, 0, 0], [0, np.pi / 2, 0], [np.pi, 0, 0], [0, 0, -np.pi], [1e-17, 1e-17, 0], [3.0, -3.0, 3.0]],
])
rows = []
for r in rpy:
    ax, ang = euler2axangle(float(r[0]), float(r[1]), float(r[2]))
    rows.append({"rpy": [float(v).hex() for v in r], "axis": [float(v).hex() for v in ax], "angle": float(ang).hex()})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "postprocess_golden.json")
with open(out, "w") as f:
    json.dump({"source": "third_party/open_pi_zero/src/utils/geometry.py:261-291 (euler2axangle, axes='sxyz')", "cases": rows}, f)
print(f"wrote {len(rows)} cases to {out}")
