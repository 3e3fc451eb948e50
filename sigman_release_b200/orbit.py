"""Multi-GPU orbit rendering: independent views shard across ranks, results are all-gathered
(BASELINE.json config 4; mirrors ``self.accelerator.gather(out['images_pred'])`` of
``/root/reference/core/loss/eval.py:81-82``).

One process per GPU; ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is only plumbing —
the renders themselves need no collective.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_views", "render_orbit_sharded", "to_wire", "from_wire"]


def shard_views(num_views: int, rank: int, world: int) -> list[int]:
    """Round-robin: rank r renders views r, r + world, r + 2*world, ..."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    return list(range(rank, num_views, world))


def to_wire(x: torch.Tensor, wire_dtype) -> torch.Tensor:
    """Wire format of the gather: fp32 (exact), fp16, or uint8 for images in [0, 1] (round to nearest
    of x * 255) — 2x / 4x fewer NVLink bytes for evaluation renders."""
    if wire_dtype in (None, torch.float32):
        return x
    if wire_dtype == torch.uint8:
        return (x.clamp(0, 1) * 255.0 + 0.5).to(torch.uint8)
    return x.to(wire_dtype)


def from_wire(x: torch.Tensor, wire_dtype) -> torch.Tensor:
    if wire_dtype in (None, torch.float32):
        return x
    if wire_dtype == torch.uint8:
        return x.to(torch.float32) / 255.0
    return x.to(torch.float32)


def render_orbit_sharded(render_fn: Callable[[Sequence[int]], torch.Tensor], num_views: int,
                         group=None, wire_dtype=None) -> torch.Tensor:
    """``render_fn(view_indices) -> [len(view_indices), C, H, W]`` is called with this rank's shard; returns the
    full ``[num_views, C, H, W]`` stack on every rank in view order.  Shards are padded to equal length so a single
    ``all_gather_into_tensor`` moves everything.  ``wire_dtype`` (fp16 / uint8) shrinks the gathered bytes; the result
    is converted back to fp32 (uint8 only for planes in [0, 1], e.g. clamped RGB + alpha)."""
    if not dist.is_available() or not dist.is_initialized():
        return from_wire(to_wire(render_fn(list(range(num_views))), wire_dtype), wire_dtype)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_views(num_views, rank, world)
    per = (num_views + world - 1) // world
    padded = mine + [mine[-1] if mine else 0] * (per - len(mine))
    local = render_fn(padded)
    if local.shape[0] != per:
        raise ValueError("render_fn must return one image stack per requested view")
    local = to_wire(local, wire_dtype).contiguous()
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    out = out.view(world, per, *local.shape[1:])
    # view v lives at [v % world, v // world]
    idx_r = torch.arange(num_views, device=local.device) % world
    idx_k = torch.arange(num_views, device=local.device) // world
    return from_wire(out[idx_r, idx_k], wire_dtype)
