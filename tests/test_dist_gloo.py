"""N > 1 host logic on CPU with the gloo backend (world_size 2 and 3): episode sharding, the final
action gather, and equality with the single-process result.  The compute step is a deterministic
per-episode stand-in (the CUDA step cannot run without a GPU); the same `infer_sharded` drives the
real model in tests/test_gpu_full.py and bench.py."""

import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blurr_b200 import dist as bdist
from blurr_b200 import synth
from blurr_b200.config import bridge_config


def _fake_step(**kw):
    # depends on every per-episode input, episode by episode (no cross-episode mixing)
    b = kw["input_ids"].shape[0]
    s = (kw["input_ids"].float().sum(1) * 1e-6 + kw["pixel_values"].float().flatten(1).sum(1)
         + kw["proprios"].float().flatten(1).sum(1)
         + (kw["image_text_proprio_mask"] == 0).float().flatten(1).sum(1) * 1e-3).view(b, 1, 1)
    return kw["noise"].float() + s


def _worker(rank, world, port, n_episodes, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = bridge_config(1)
        inp = synth.synthetic_inputs(cfg, n_episodes, dtype=torch.float32, vary_text=True)
        out = bdist.infer_sharded(_fake_step, inp)
        ref = _fake_step(**inp)
        assert out.shape == ref.shape and torch.equal(out, ref), f"rank {rank}"
        if rank == 0:
            torch.save(out, path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_episodes", [(2, 6), (3, 7), (2, 1)])
def test_sharded_equals_single_process(tmp_path, world, n_episodes):
    path = str(tmp_path / "out.pt")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, n_episodes, path), nprocs=world, join=True)
    assert torch.load(path).shape[0] == n_episodes


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    with pytest.raises(ValueError):
        bdist.shard_range(4, 2, 2)
