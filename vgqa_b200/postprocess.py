"""Decode side of the hot path: PostProcess, single_forward's dict building, tube interpolation and the predict() merge.

Mirrors (same names, argument meaning and error behaviour):
  PostProcess.forward            vgqa/core/postprocessor.py:14-50     (device kernel `postprocess_kernel` via vgqa_postprocess)
  single_forward                 vgqa/training/evaluator.py:56-92
  linear_interp / _conf          vgqa/training/evaluator.py:10-54
  merge of the even/odd passes   vgqa/inference/grounding.py:214-244  (`merge_predictions`)
"""
from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Sequence

import torch

from . import _lib
from .engine import _declare


class PostProcess(torch.nn.Module):
    """build_postprocessors() → PostProcess (vgqa/core/__init__.py:52-54)."""

    @torch.no_grad()
    def forward(self, outputs, target_sizes, frames_id, durations):
        out_sted, out_bbox, kf_pr = outputs["pred_sted"], outputs["pred_boxes"], outputs["pr"]
        out_att = outputs["att_sequences"]
        assert len(out_bbox) == len(target_sizes)
        b, t, _ = out_sted.shape
        assert all(int(d) == t for d in durations), "vgqa_b200 decodes clips of equal length per call (reference: batch 1)"
        L = _lib.lib()
        _declare(L)
        dev = out_bbox.device
        sizes = target_sizes.to(dev, torch.float32).reshape(b, t, 2)[:, 0].contiguous()
        boxes = out_bbox.to(torch.float32).contiguous()
        sted = out_sted.to(torch.float32).contiguous()
        boxes_px = torch.empty(b * t, 4, device=dev, dtype=torch.float32)
        idx = torch.empty(b, 2, device=dev, dtype=torch.int32)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.vgqa_postprocess(_lib.ptr(boxes), _lib.ptr(sted), _lib.ptr(sizes), _lib.ptr(boxes_px), _lib.ptr(idx),
                                      b, t, ctypes.c_void_p(st)))
        idx_h = idx.cpu().tolist()   # the one device→host read of the decode (reference: `.item()`, postprocessor.py:44)
        pred_steds = [[frames_id[i][s], frames_id[i][e] + 1] for i, (s, e) in enumerate(idx_h)]
        return boxes_px, out_att, pred_steds, kf_pr


def build_postprocessors() -> PostProcess:
    return PostProcess()


@torch.no_grad()
def linear_interp(bbox_dict: Dict[int, List[List[float]]]):
    """Fill the integer frame ids between sampled ids with linearly interpolated boxes."""
    fids = sorted(bbox_dict)
    if len(fids) < 2:
        return bbox_dict
    for left, right in zip(fids[:-1], fids[1:]):
        gap = right - left
        if gap > 1:
            lo, hi = bbox_dict[left][0], bbox_dict[right][0]
            slope = [(h - l) / gap for l, h in zip(lo, hi)]
            for step in range(1, gap):
                bbox_dict[left + step] = [[l + step * s for l, s in zip(lo, slope)]]
    fids = sorted(bbox_dict)
    assert max(fids) - min(fids) + 1 == len(fids)
    return {f: bbox_dict[f] for f in fids}


@torch.no_grad()
def linear_interp_conf(conf_dict: Dict[int, Any]):
    """Nearest-hold fill of confidences: the left value up to the middle of a gap, the right value after it."""
    fids = sorted(conf_dict)
    if len(fids) < 2:
        return conf_dict
    for left, right in zip(fids[:-1], fids[1:]):
        gap = right - left
        for step in range(1, gap):
            conf_dict[left + step] = conf_dict[left] if step <= gap // 2 else conf_dict[right]
    fids = sorted(conf_dict)
    assert max(fids) - min(fids) + 1 == len(fids)
    return {f: conf_dict[f] for f in fids}


@torch.no_grad()
def single_forward(cfg, model, videos, texts, targets, device, postprocessor):
    """One forward + PostProcess → (bbox_pred, att_pred, temp_pred, kf_pred) dicts keyed like the reference's."""
    durations = videos.durations
    targets[0]["durations"] = durations
    outputs = model(videos, texts, targets)
    b, t = len(durations), max(durations)
    sizes = torch.tensor([list(tg["ori_size"]) for tg in targets for _ in range(t)], device=device)
    assert sizes.shape[0] == outputs["pred_boxes"].shape[0]
    frame_ids = [tg["frame_ids"] for tg in targets]
    boxes, att, steds, kf = postprocessor(outputs, sizes, frame_ids, durations)
    boxes = boxes.view(b, t, 4).cpu().tolist()        # one D2H copy instead of the reference's 2T per-frame copies
    att = att.reshape(b, t).cpu().tolist()
    vids = [tg["item_id"] for tg in targets]
    bbox_pred, att_pred, temp_pred, kf_pred = {}, {}, {}, {}
    for i in range(b):
        fids = frame_ids[i]
        assert durations[i] == len(fids)
        bbox_pred[vids[i]] = {fids[j]: [boxes[i][j]] for j in range(durations[i])}
        att_pred[vids[i]] = {fids[j]: [att[i][j]] for j in range(durations[i])}
    qtypes = [tg["qtype"] for tg in targets]
    assert len(steds) == len(qtypes)
    for i in range(b):
        temp_pred[vids[i]] = {"sted": steds[i], "qtype": qtypes[i]}
    kf_pred[vids[0]] = kf
    return bbox_pred, att_pred, temp_pred, kf_pred


def merge_predictions(pass1, pass2, fps: float, vid_key=0) -> Dict[str, Any]:
    """predict()'s merge of the even / odd frame passes into {"temporal": {...}, "tube": [...]}."""
    (bbox1, att1, temp1, _), (bbox2, att2, temp2, _) = pass1, pass2
    bbox = dict(bbox1[vid_key]); bbox.update(bbox2[vid_key])
    bbox_full = linear_interp(bbox)
    att = dict(att1[vid_key]); att.update(att2[vid_key])
    att_full = linear_interp_conf(att)
    s1, s2 = temp1[vid_key]["sted"], temp2[vid_key]["sted"]
    sted = [min(s1[0], s2[0]), max(s1[1], s2[1])]
    denom = max(fps, 1e-6)
    tube = []
    for fid in sorted(bbox_full):
        box = bbox_full[fid][0]
        conf = att_full.get(fid, 1.0)
        tube.append({"frame": int(fid), "bbox": [float(v) for v in box[:4]],
                     "score": float(conf[0] if isinstance(conf, list) else conf)})
    return {"temporal": {"start": float(sted[0]) / denom, "end": float(sted[1]) / denom, "score": 1.0}, "tube": tube}
