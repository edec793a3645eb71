"""Data-parallel plumbing (one process per GPU, torch.distributed): videos are independent units, so the only
collectives on the path are (1) the gradient allreduce(mean) over the flat fp32 gradient buffer — bucketed so NCCL can
pipeline over NVLink/NVSwitch — and (2) the end-of-evaluation gather of the per-frame *integer* match sets, re-ordered to
dataset order on rank 0 so the recall floats are computed exactly as a single process would (no float allreduce)."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_indices(n_items: int, rank_: int, world_: int) -> List[int]:
    """Round-robin assignment of videos to ranks (video i -> rank i % world)."""
    return list(range(rank_, n_items, world_))


def allreduce_mean_(flat: torch.Tensor, bucket_elems: int = 32 * 1024 * 1024) -> torch.Tensor:
    """In-place mean over ranks of a flat buffer, issued as independent async buckets (128 MB of fp32 each)."""
    w = world()
    if w == 1:
        return flat
    works = [dist.all_reduce(flat[i:i + bucket_elems], async_op=True) for i in range(0, flat.numel(), bucket_elems)]
    for wk in works:
        wk.wait()
    flat.mul_(1.0 / w)
    return flat


def allreduce_sum_(flat: torch.Tensor, flag: torch.Tensor = None, bucket_elems: int = 32 * 1024 * 1024) -> torch.Tensor:
    """In-place SUM over ranks of the flat gradient buffer in independent async buckets (128 MB of fp32 each); the 1/world
    of the mean is folded into the optimiser kernel.  `flag` (int32[1], the skip-this-step flag of lib/utils.py:3-11) is
    max-reduced alongside so that all ranks skip together."""
    if world() == 1:
        return flat
    works = [dist.all_reduce(flat[i:i + bucket_elems], async_op=True) for i in range(0, flat.numel(), bucket_elems)]
    if flag is not None:
        works.append(dist.all_reduce(flag, op=dist.ReduceOp.MAX, async_op=True))
    for wk in works:
        wk.wait()
    return flat


def allreduce_sum_async(flat: torch.Tensor, bucket_elems: int = 32 * 1024 * 1024) -> list:
    """Start the in-place SUM over ranks of `flat` (independent async buckets) and return the work handles; the collective is
    ordered after everything already enqueued on the current stream and runs on NCCL's own stream."""
    if world() == 1 or flat.numel() == 0:
        return []
    return [dist.all_reduce(flat[i:i + bucket_elems], async_op=True) for i in range(0, flat.numel(), bucket_elems)]


def gather_frame_results(local_ids: torch.Tensor, local_masks: torch.Tensor):
    """all_gather of (global frame id, u32[3,3,8] match sets) with ragged sizes; returns (ids, masks) sorted by id
    on every rank.  local_ids: int64[F_local]; local_masks: int32[F_local,3,3,8]."""
    w = world()
    if w == 1:
        order = torch.argsort(local_ids)
        return local_ids[order], local_masks[order]
    n = torch.tensor([local_ids.numel()], dtype=torch.int64, device=local_ids.device)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n)
    mx = int(max(int(s) for s in sizes))
    pad_ids = torch.full((mx,), -1, dtype=torch.int64, device=local_ids.device)
    pad_ids[:local_ids.numel()] = local_ids
    pad_masks = torch.zeros((mx, 3, 3, 8), dtype=local_masks.dtype, device=local_masks.device)
    pad_masks[:local_masks.shape[0]] = local_masks
    all_ids = [torch.empty_like(pad_ids) for _ in range(w)]
    all_masks = [torch.empty_like(pad_masks) for _ in range(w)]
    dist.all_gather(all_ids, pad_ids)
    dist.all_gather(all_masks, pad_masks)
    ids = torch.cat([a[:int(s)] for a, s in zip(all_ids, sizes)])
    masks = torch.cat([m[:int(s)] for m, s in zip(all_masks, sizes)])
    order = torch.argsort(ids)
    return ids[order], masks[order]
