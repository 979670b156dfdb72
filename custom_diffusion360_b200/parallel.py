"""Image-parallel sharding of the sampling path across the GPUs of one box.

Each image's 50-step trajectory is independent (its CFG triple stays on one GPU, weights are
replicated), so there is no collective inside the loop; the only exchange is one all_gather of the
final latents `[N, 4, L, L]`.  Works with any torch.distributed backend (NCCL over NVLink on the
B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def image_shard(n_images: int, rank: int, world: int) -> List[int]:
    """Indices of the images rank `rank` samples: r, r+W, r+2W, ... (round-robin keeps the shards
    within one image of each other for any N)."""
    return list(range(rank, n_images, world))


def gather_images(x_local: torch.Tensor, n_images: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """x_local: this rank's final latents, rows ordered like `image_shard`.  Returns the full
    `[n_images, ...]` tensor in global image order on every rank (one all_gather; shards are padded
    to the largest shard so a single fixed-size collective suffices)."""
    if not dist.is_available() or not dist.is_initialized():
        return x_local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_images + world - 1) // world
    pad = torch.zeros((per,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.empty((n_images,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    for r in range(world):
        idx = image_shard(n_images, r, world)
        if idx:
            out[idx] = bufs[r][: len(idx)]
    assert len(image_shard(n_images, rank, world)) == x_local.shape[0]
    return out
