"""Action post-processing of the control loop: what the reference's env adapters do
with the model's [horizon, 7] action chunk before `env.step` — host-side float64 numpy, like the reference
(28 numbers per step; nothing here is worth a kernel).

Mirrors `third_party/open_pi_zero/src/agent/env_adapter/simpler.py`:
  * `SimplerAdapter.postprocess` (:100-141): de-normalise all but the gripper dimension
    (`denormalize_bound` / `denormalize_gaussian`, env_adapter/base.py:20-31,42-49), euler (sxyz) ->
    axis-angle, gripper post-processing;
  * `BridgeSimplerAdapter.postprocess_gripper` (:181-186): binarise to -1 / +1;
  * `EDRSimplerAdapter.postprocess_gripper` (:221-252) with its sticky-gripper state (`reset`, :197-202).
`euler2axangle` restates the transforms3d functions the reference vendors in `src/utils/geometry.py:261-291,
294-363,366-434` (`euler2quat(axes="sxyz")` + `quat2axangle`).  Parity is pinned: bit-exact against that file and
against golden vectors generated from it (tests/golden/postprocess_golden.json), and the whole `postprocess` incl. the
sticky-gripper state machine is bit-exact against the reference's `BridgeSimplerAdapter` / `EDRSimplerAdapter` objects
(tests/test_postprocess.py).
"""

from __future__ import annotations

import math
from typing import Sequence

import numpy as np

_FLOAT_EPS = np.finfo(np.float64).eps


def denormalize_bound(data, data_min, data_max, clip_min=-1, clip_max=1, eps=1e-8):
    """base.py:20-31 (eps is unused there as well)."""
    clip_range = clip_max - clip_min
    return (data - clip_min) / clip_range * (data_max - data_min) + data_min


def denormalize_gaussian(data, mean, std, eps=1e-8):
    """base.py:42-49."""
    return data * (std + eps) + mean


def euler2quat_sxyz(ai: float, aj: float, ak: float) -> np.ndarray:
    """transforms3d.euler.euler2quat(ai, aj, ak, axes='sxyz') -> (w, x, y, z)."""
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc])


def quat2axangle(quat: Sequence[float]):
    """transforms3d.quaternions.quat2axangle(quat) -> (unit axis [3], angle)."""
    w, x, y, z = (float(v) for v in quat)
    nq = w * w + x * x + y * y + z * z
    if not math.isfinite(nq):
        return np.array([1.0, 0.0, 0.0]), float("nan")
    identity_thresh = _FLOAT_EPS * 3
    if nq < _FLOAT_EPS ** 2:
        return np.array([1.0, 0.0, 0.0]), 0.0
    if nq != 1:
        s = math.sqrt(nq)
        w, x, y, z = w / s, x / s, y / s, z / s
    len2 = x * x + y * y + z * z
    if len2 < identity_thresh ** 2:
        return np.array([1.0, 0.0, 0.0]), 0.0
    theta = 2 * math.acos(max(min(w, 1), -1))
    return np.array([x, y, z]) / math.sqrt(len2), theta


def euler2axangle(ai: float, aj: float, ak: float):
    return quat2axangle(euler2quat_sxyz(ai, aj, ak))


class ActionPostprocessor:
    """`postprocess(actions)` of the reference adapters.  kind: "bridge" (binarised gripper) or "fractal"
    (sticky gripper, `reset()` at episode start).  `stats` = dataset_statistics["action"]."""

    def __init__(self, stats: dict, kind: str = "bridge", action_normalization_type: str = "bound",
                 sticky_gripper_num_repeat: int = 15):
        if kind not in ("bridge", "fractal"):
            raise ValueError(kind)
        self.stats, self.kind, self.norm = stats, kind, action_normalization_type
        self.sticky_gripper_num_repeat = sticky_gripper_num_repeat
        self.reset()

    def reset(self):
        self.sticky_action_is_on = False
        self.gripper_action_repeat = 0
        self.sticky_gripper_action = 0.0

    def postprocess_gripper(self, action: float) -> float:
        if self.kind == "bridge":
            return 2.0 * (action > 0.5) - 1.0
        action = (action * 2) - 1
        relative_gripper_action = -action
        if np.abs(relative_gripper_action) > 0.5 and self.sticky_action_is_on is False:
            self.sticky_action_is_on = True
            self.sticky_gripper_action = relative_gripper_action
        if self.sticky_action_is_on:
            self.gripper_action_repeat += 1
            relative_gripper_action = self.sticky_gripper_action
        if self.gripper_action_repeat == self.sticky_gripper_num_repeat:
            self.sticky_action_is_on = False
            self.gripper_action_repeat = 0
            self.sticky_gripper_action = 0.0
        return relative_gripper_action

    def postprocess(self, actions: np.ndarray) -> np.ndarray:
        """actions: [horizon, 7] (model output, float) -> [horizon, 7] env actions (xyz, axis*angle, gripper)."""
        actions = np.asarray(actions)
        if self.norm == "bound":
            raw = denormalize_bound(actions[:, :-1], np.array(self.stats["p01"])[:-1], np.array(self.stats["p99"])[:-1],
                                    clip_min=-1, clip_max=1)
        elif self.norm == "gaussian":
            raw = denormalize_gaussian(actions[:, :-1], np.array(self.stats["mean"])[:-1], np.array(self.stats["std"])[:-1])
        else:
            raise ValueError(self.norm)
        raw_actions = np.concatenate([raw, actions[:, -1:]], axis=1)
        out = np.zeros((len(raw_actions), 7))
        for idx, raw_action in enumerate(raw_actions):
            roll, pitch, yaw = raw_action[3:6]
            ax, angle = euler2axangle(roll, pitch, yaw)
            out[idx] = np.concatenate([raw_action[:3], ax * angle, [self.postprocess_gripper(raw_action[-1])]])
        return out
