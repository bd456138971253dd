"""Episode sharding across the GPUs of one box.

The control step has no cross-episode interaction (masks are per sample, reference
`src/model/vla/pizero.py:353-357`; the reference evaluates one environment per process,
`src/agent/eval.py:93`), so the path shards by *episodes*: contiguous blocks of the batch per rank,
weights replicated, no collective inside a control step.  The only exchange is the gather of the
`[B_local, horizon, action_dim]` action chunks (and scalar metrics) at the end — one NCCL
`all_gather` over NVLink (gloo in the CPU tests).
"""

from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import torch
import torch.distributed as dist

BATCH_KEYS = ("input_ids", "pixel_values", "image_text_proprio_mask", "action_mask", "vlm_position_ids",
              "proprio_position_ids", "action_position_ids", "proprios", "noise")


def shard_range(n_episodes: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of episodes owned by `rank`; the first `n % world` ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_episodes, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_inputs(inputs: Dict[str, torch.Tensor], world_size: int, rank: int) -> Dict[str, torch.Tensor]:
    """Slice every per-episode tensor of a call dict to this rank's block (views, no copies)."""
    n = inputs["input_ids"].shape[0]
    lo, hi = shard_range(n, world_size, rank)
    return {k: (v[lo:hi] if k in BATCH_KEYS else v) for k, v in inputs.items()}


def gather_actions(local_actions: torch.Tensor, n_episodes: int) -> torch.Tensor:
    """All ranks receive the `[n_episodes, horizon, action_dim]` actions in episode order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_actions
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_episodes, world, r) for r in range(world)]
    max_len = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_len,) + tuple(local_actions.shape[1:]), dtype=local_actions.dtype,
                      device=local_actions.device)
    pad[: local_actions.shape[0]] = local_actions
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def infer_sharded(step_fn: Callable[..., torch.Tensor], inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Run `step_fn(**local_inputs)` on this rank's episodes and gather every rank's actions."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = inputs["input_ids"].shape[0]
    local = shard_inputs(inputs, world, rank)
    if local["input_ids"].shape[0] == 0:
        out = torch.zeros((0,) + tuple(inputs["noise"].shape[1:]), dtype=inputs["noise"].dtype,
                          device=inputs["noise"].device)
    else:
        out = step_fn(**local)
    return gather_actions(out, n)
