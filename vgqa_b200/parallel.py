"""Multi-GPU plumbing of the hot path: clips are independent units, so ranks own disjoint clip ranges and there is no
data-path collective.  torch.distributed is used only to merge the per-rank prediction dicts
(replaces the reference's pickled all_gather, vgqa/utils/distributed.py:45-80 used at vidstg_evaluator.py:190-198)."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch.distributed as dist


def partition_clips(n_clips: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) range of clip indices owned by `rank` (earlier ranks get the remainder)."""
    assert world_size >= 1 and 0 <= rank < world_size and n_clips >= 0
    base, rem = divmod(n_clips, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_predictions(local: Dict[Any, Any], group=None) -> Dict[Any, Any]:
    """Union of the per-rank {video_id: prediction} dicts on every rank; duplicate ids must agree."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    parts: List[Dict[Any, Any]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    merged: Dict[Any, Any] = {}
    for part in parts:
        for k, v in part.items():
            if k in merged:
                assert merged[k] == v, f"conflicting predictions for {k}"
            merged[k] = v
    return merged
