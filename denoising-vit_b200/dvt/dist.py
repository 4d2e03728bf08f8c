"""Multi-GPU plumbing of stage 1: images shard across ranks with no data-path collective (the reference runs 8
unrelated processes over index ranges, sample_scripts/stage1.sh:8-19); the only exchange is one all-gather that
collates the per-rank stacks of denoised maps so every stage-2 rank holds the full set (replaces the .npy hand-off
of main_img_denoising.py:131-146 -> dvt/dataset/paired_list_dataset.py:30-37)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(num_items: int, rank: int, world: int) -> List[int]:
    """Rank-strided image assignment: rank r takes items r, r + world, ..."""
    return list(range(rank, num_items, world))


def collate_maps(local_maps: torch.Tensor, num_items: int) -> torch.Tensor:
    """local_maps [n_r, h, w, C] for the items of shard_indices(num_items, rank, world) -> [num_items, h, w, C] in
    original item order on every rank.  Ranks may hold different counts (num_items not divisible by world)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert local_maps.shape[0] == num_items
        return local_maps
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (num_items + world - 1) // world
    padded = local_maps.new_zeros((per,) + tuple(local_maps.shape[1:]))
    padded[: local_maps.shape[0]] = local_maps
    flat = local_maps.new_empty((world * per,) + tuple(local_maps.shape[1:]))  # ranks concatenated along dim 0
    dist.all_gather_into_tensor(flat, padded.contiguous())
    gathered = flat.view((world, per) + tuple(local_maps.shape[1:]))
    out = local_maps.new_empty((num_items,) + tuple(local_maps.shape[1:]))
    for r in range(world):
        idx = shard_indices(num_items, r, world)
        out[idx] = gathered[r, : len(idx)]
    return out
