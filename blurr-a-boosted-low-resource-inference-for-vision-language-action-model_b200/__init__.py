"""blurr_b200 — B200-native Pi-0 control-step path.

Host side (Python/PyTorch, mirrors the reference's `PiZeroInference` surface) over a C-ABI
CUDA library (`include/blurr_pi0.h`, built from `csrc/`).
"""

__version__ = "0.1.0"
