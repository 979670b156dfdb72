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


def gather_references(unet, captured: dict, null_row: Optional[dict] = None,
                      group: Optional[dist.ProcessGroup] = None) -> dict:
    """Validation-epoch reference capture across ranks (reference main.py:594-602): every rank ran
    `UNetModel.capture_references` on ITS share of the validation views (`captured`: {pose block
    name: [r, hw, c]}, the same r on every rank); per pose block the shards are all-gathered,
    interleaved rank-minor like the reference's `rearrange(stack(output_list).transpose(0, 1),
    "b n ... -> (b n) ...")` (view j of rank k lands at row j * world + k), optionally followed by an
    explicit 'null' row — sampling uses `references[-1]` as the null reference of the unconditional CFG
    row (sample.py:92,96), i.e. whatever was captured last — and registered as that block's
    `references` buffer.
    Returns {name: references tensor}."""
    out = {}
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    for name, _ in unet.pose_blocks():
        t = captured[name].contiguous()
        if world > 1:
            bufs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(bufs, t, group=group)
            t = torch.stack(bufs).transpose(0, 1).reshape(-1, *t.shape[1:])
        if null_row is not None:
            t = torch.cat([t, null_row[name].to(t).reshape(1, *t.shape[1:])], 0)
        out[name] = t
    unet.register_references(out)
    return out
