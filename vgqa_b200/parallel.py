"""Multi-GPU plumbing of the hot path: clips are independent units, so ranks own disjoint clip ranges and there is no
data-path collective.  torch.distributed is used only to merge the per-rank prediction dicts
(replaces the reference's pickled all_gather, vgqa/utils/distributed.py:45-80 used at vidstg_evaluator.py:190-198)."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch
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


# ----------------------------------------------------------------------------------------------------------------
# Frame sharding of ONE long clip: rank r owns the contiguous frames [r*T/G, (r+1)*T/G).
# ----------------------------------------------------------------------------------------------------------------
def shard_frames(T: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Equal contiguous frame shards (the library requires T % world_size == 0)."""
    assert T % world_size == 0, "frame sharding needs T divisible by the number of ranks"
    n = T // world_size
    return rank * n, (rank + 1) * n


class _DevArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (int(n),), "typestr": typestr, "version": 3}


def nccl_exchange(group=None):
    """Builds the `vgqa_exchange_fn` callback: the library's all-gathers / all-reduces run as NCCL collectives of
    torch.distributed on the library's stream.  Returns (ctypes callback, error list)."""
    from .engine import EXCHANGE_FN
    world = dist.get_world_size(group)
    errors: List[BaseException] = []

    def cb(user, op, send, recv, count, dtype, stream):
        try:
            ts, nb = ("<f4", 1) if dtype == 1 else ("|u1", 2)      # bf16 payloads travel as opaque bytes
            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
                if op == 1:
                    t = torch.as_tensor(_DevArray(recv, count, ts), device="cuda")
                    dist.all_reduce(t, group=group)
                else:
                    src = torch.as_tensor(_DevArray(send, count * nb, ts), device="cuda")
                    dst = torch.as_tensor(_DevArray(recv, count * nb * world, ts), device="cuda")
                    dist.all_gather_into_tensor(dst, src, group=group)
        except BaseException as e:  # exceptions cannot cross the C frame
            errors.append(e)

    return EXCHANGE_FN(cb), errors


def forward_sharded_clip(engine, vis, vid, text, pos, *, ori_size_hw, group=None, iteration_rate=-1, check_errors=True):
    """One long clip, frames sharded over the ranks of `group`.  vis/vid: THIS rank's frames [1, T_local, 256, H, W];
    text [1, L, 256] and pos [1, 256, H, W] replicated.  Returns the gathered outputs of the whole clip on every rank
    (pred_boxes [T,4], pred_sted [T,2], pred_actioness [T], att_sequences [T], choose1/2 [T], boxes_px [T,4], sted_idx [2]).
    `check_errors` reads the peer-memory exchange's error word after the call (one 4-byte D2H copy, a host sync): switch it
    off inside a timed loop and call engine.p2p_error() once at the end."""
    import ctypes
    from . import _lib
    from .engine import _declare
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if getattr(engine, "_shard_cfg", None) != (rank, world):   # not set up yet (engine.enable_p2p_sharding sets it, too)
        cb, errors = nccl_exchange(group)
        engine.set_sharding(rank, world, cb)
        engine._shard_cfg, engine._shard_errors = (rank, world), errors
    want = ["pred_boxes", "pred_sted", "pred_actioness", "att_sequences", "logits_r_a", "logits_r_m", "choose1", "choose2"]
    T_loc, H, W = vis.shape[1], vis.shape[3], vis.shape[4]
    key = (T_loc, H, W, text.shape[1])
    if getattr(engine, "_shard_outs_key", None) != key:     # outputs are allocated once per shape and reused
        engine._shard_outs, engine._shard_outs_key = engine.alloc_outputs(1, T_loc, H, W, text.shape[1], want), key
    o = engine.forward(vis, vid, text, pos, iteration_rate=iteration_rate, outs=engine._shard_outs)
    if engine._shard_errors:
        raise engine._shard_errors.pop()
    if check_errors and engine.p2p_error():                  # a peer never arrived within the kernel's ≈2 s bound
        raise RuntimeError("frame-sharded forward: a peer-memory exchange timed out (ranks out of step or a peer died)")
    full = {}
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "att_sequences", "choose1", "choose2"):
        loc = o[k][0].contiguous()
        buf = torch.empty((world,) + tuple(loc.shape), device=loc.device, dtype=loc.dtype)
        dist.all_gather_into_tensor(buf, loc, group=group)
        full[k] = buf.reshape((world * T_loc,) + tuple(loc.shape[1:]))
    full["logits_r_a"], full["logits_r_m"] = o["logits_r_a"][0], o["logits_r_m"][0]
    L_ = _lib.lib()
    _declare(L_)
    T = world * T_loc
    sizes = torch.tensor([[float(ori_size_hw[0]), float(ori_size_hw[1])]], device=vis.device)
    boxes_px = torch.empty(T, 4, device=vis.device)
    idx = torch.empty(1, 2, device=vis.device, dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L_.vgqa_postprocess(_lib.ptr(full["pred_boxes"]), _lib.ptr(full["pred_sted"]), _lib.ptr(sizes),
                                   _lib.ptr(boxes_px), _lib.ptr(idx), 1, T, ctypes.c_void_p(st)))
    full["boxes_px"], full["sted_idx"] = boxes_px, idx[0]
    return full
