"""Batch evaluation over a dataset shard per GPU — the caller that exercises clip partitioning.

Mirrors, on top of the B200 engine:
  * `do_eval`                      vgqa/training/evaluator.py:95-150   (even / odd passes, merge, evaluator updates, gather, summarize)
  * `VidSTGiouEvaluator.evaluate`  vgqa/data/metrics/vidstg_evaluator.py:17-136  (tIoU, vIoU, gt_vIoU, recalls, key-frame P/R)
  * `VidSTGEvaluator`              vgqa/data/metrics/vidstg_evaluator.py:139-260 (update* / synchronize_between_processes / summarize)

Differences on purpose: items of a rank are batched (the even and odd passes of `clips_per_call` items are ONE engine call — clips
never interact), results leave the device in three copies per call instead of 2T per clip, and the per-rank dicts are merged with
`all_gather_object` (vgqa_b200/parallel.py) instead of the reference's padded pickled byte tensors (utils/distributed.py:45-80).
Items are hot-path-boundary tensors (or raw extractor outputs with raw=True); video decoding and the backbones stay the caller's.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from .parallel import gather_predictions, partition_clips
from .postprocess import linear_interp, linear_interp_conf


def np_box_iou(boxes1: np.ndarray, boxes2: np.ndarray) -> np.ndarray:
    """Pairwise IoU of xyxy boxes (vgqa/utils/box_ops.py:14-38)."""
    a1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    a2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    lt = np.maximum(boxes1[:, None, :2], boxes2[:, :2])
    rb = np.minimum(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clip(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (a1[:, None] + a2 - inter)


class VidSTGiouEvaluator:
    """Per-video temporal / spatio-temporal IoU metrics.  `gt_data`: list of {"item_id", "gt_temp_bound": [s, e],
    "bboxs": {frame_id: [x1, y1, x2, y2]}, "description"} — the records of the reference's annotation cache
    (vidstg_evaluator.py:24-39), or a path to that cache (`torch.load`)."""

    def __init__(self, gt_data, iou_thresholds: Optional[List[float]] = None):
        if isinstance(gt_data, str):
            gt_data = torch.load(gt_data)
        self.vid2steds, self.vid2box, self.vid2names, self.vid2sents = {}, {}, {}, {}
        for d in gt_data:
            i = d["item_id"]
            self.vid2names[i] = i
            self.vid2sents[i] = d.get("description", "")
            self.vid2box[i] = {fid: [box] for fid, box in d["bboxs"].items()}
            self.vid2steds[i] = d["gt_temp_bound"]
        self.iou_thresholds = iou_thresholds or [0.3, 0.5]

    def evaluate(self, predictions, video_predictions, pred_conf, pred_kf):
        vid_metrics: Dict[Any, Dict[str, Any]] = {}
        for vid, vp in video_predictions.items():
            gt, pr = self.vid2steds[vid], vp["sted"]
            max_start, min_end = max(gt[0], pr[0]), min(gt[1], pr[1])
            min_start, max_end = min(gt[0], pr[0]), max(gt[1], pr[1])
            if min_end <= max_start:
                tiou = 0
            else:
                inter = min_end - max_start
                tiou = inter / ((gt[1] - gt[0]) + (pr[1] - pr[0]) - inter)
            m = {"gt_sted": gt, "pred_sted": pr, "tiou": tiou, "qtype": vp["qtype"], "img_metrics": {}}
            n_union = max_end - min_start if max_end > min_start else 0
            viou = gt_viou = 0
            prediction = predictions.get(vid, {})
            for fid, gt_boxes in self.vid2box[vid].items():
                if fid not in prediction:
                    continue
                iou = np_box_iou(np.array(prediction[fid]), np.array(gt_boxes))[0][0]
                if max_start <= fid < min_end:
                    viou += iou
                gt_viou += iou
            viou = viou / max(n_union, 1)
            gt_viou = gt_viou / max(len(self.vid2box[vid]), 1)
            m["viou"], m["gt_viou"] = viou, gt_viou
            for th in self.iou_thresholds:
                m[f"viou@{th}"] = 1 if viou > th else 0
                m[f"gt_viou@{th}"] = 1 if gt_viou > th else 0
            vid_metrics[vid] = m
        for vid, kf in pred_kf.items():
            vid_metrics[vid]["kf_pr"] = kf
        return vid_metrics, self.vid2names, self.vid2sents


class VidSTGEvaluator:
    """Accumulates the per-rank prediction dicts, merges them across ranks and averages the metrics per question type."""

    def __init__(self, gt_data, iou_thresholds: Sequence[float] = (0.3, 0.5), logger=None, group=None, distributed: bool = True):
        self.evaluator = VidSTGiouEvaluator(gt_data, list(iou_thresholds))
        self.iou_thresholds = list(iou_thresholds)
        self.predictions, self.att_predictions, self.confs, self.video_predictions, self.kf_pred = {}, {}, {}, {}, {}
        self.results = None
        self.logger = logger or logging.getLogger("vgqa_b200.evaluate")
        self.group = group
        self.distributed = distributed   # False: a process-local evaluator (synchronize_between_processes is a no-op)

    def update(self, predictions): self.predictions.update(predictions)
    def update_att(self, predictions): self.att_predictions.update(predictions)
    def update_conf(self, confs): self.confs.update(confs)
    def update_kf_pr(self, kf_pr): self.kf_pred.update(kf_pr)
    def video_update(self, video_predictions): self.video_predictions.update(video_predictions)

    def synchronize_between_processes(self):
        if not self.distributed:
            return
        for name in ("predictions", "att_predictions", "confs", "kf_pred", "video_predictions"):
            setattr(self, name, gather_predictions(getattr(self, name), self.group))

    def summarize(self, main_process: bool = True):
        if not main_process:
            return None
        self.results, _, _ = self.evaluator.evaluate(self.predictions, self.video_predictions, self.confs, self.kf_pred)
        names = ["gt_viou", "tiou", "viou", "kf_p", "kf_r"]
        for th in self.iou_thresholds:
            names += [f"viou@{th}", f"gt_viou@{th}"]
        metrics: Dict[str, Dict[str, float]] = {}
        counter: Dict[str, int] = {}
        for x in self.results.values():
            q = x["qtype"]
            m = metrics.setdefault(q, {n: 0 for n in names})
            for n in names:
                m[n] += x["kf_pr"][0] if n == "kf_p" else x["kf_pr"][1] if n == "kf_r" else x[n]
            counter[q] = counter.get(q, 0) + 1
        out = {}
        for q, m in metrics.items():
            for n in names:
                out[f"{q}_{n}"] = m[n] / max(counter[q], 1)
                self.logger.info("%s %s: %.4f", q, n, out[f"{q}_{n}"])
        return out


def merge_even_odd(p1, p2):
    """do_eval's merge of the even / odd passes of one batch (evaluator.py:128-139): dict union → linear_interp(_conf),
    averaged key-frame precision / recall, sted = [min start, max end]."""
    (bbox1, att1, temp1, kf1), (bbox2, att2, temp2, kf2) = p1, p2
    bbox_pred, att_pred, temp_pred, kf_pred = {}, {}, {}, {}
    for vid in bbox1:
        b = dict(bbox1[vid]); b.update(bbox2[vid])
        bbox_pred[vid] = linear_interp(b)
        a = dict(att1[vid]); a.update(att2[vid])
        att_pred[vid] = linear_interp_conf(a)
        kf_pred[vid] = [(kf1[vid][0] + kf2[vid][0]) / 2, (kf1[vid][1] + kf2[vid][1]) / 2]
        temp_pred[vid] = {"sted": [min(temp1[vid]["sted"][0], temp2[vid]["sted"][0]),
                                   max(temp1[vid]["sted"][1], temp2[vid]["sted"][1])]}
        if "qtype" in temp1[vid]:
            temp_pred[vid]["qtype"] = temp1[vid]["qtype"]
    return bbox_pred, att_pred, temp_pred, kf_pred


def precision_recall(predicted: Sequence[int], true: Sequence[int]) -> Tuple[float, float]:
    """Key-frame precision / recall of the chosen frames against the annotated actioness (grounding_net.py:198-202)."""
    ps, ts = set(predicted), set(true)
    inter = len(ps & ts)
    return (0 if not ps else inter / len(ps)), (0 if not ts else inter / len(ts))


@torch.no_grad()
def do_eval(engine, items: Sequence[Mapping[str, Any]], evaluator: VidSTGEvaluator, *, clips_per_call: int = 1, rank: int = 0,
            world: int = 1, raw: bool = False):
    """items[i]: {"item_id", "vis" [2T,C,H,W], "vid" [2T,C,H,W], "text" [L,C], optional "pos" [1,256,H,W] (generated in the library when absent), "frame_ids" (2T ascending ints),
    "ori_size" (h, w), "qtype" (default 'none', evaluator.py:107-109), "actioness" [2T] 0/1}.  All items share (T, H, W, L).
    Rank `rank` evaluates its contiguous share of the items (no data-path collective), `clips_per_call` items per engine call
    (2 * clips_per_call clips: even and odd passes); afterwards the dicts are merged on every rank and rank 0 summarizes."""
    lo, hi = partition_clips(len(items), world, rank)
    dev = engine.device
    f32 = lambda t: torch.as_tensor(t, dtype=torch.float32, device=dev)
    for b0 in range(lo, hi, clips_per_call):
        batch = items[b0:min(b0 + clips_per_call, hi)]
        vis, vid, text, sizes = [], [], [], []
        for it in batch:
            v, w = f32(it["vis"]), f32(it["vid"])
            assert v.shape[0] % 2 == 0 and len(it["frame_ids"]) == v.shape[0], "2T sampled frames, one frame id each"
            for par in (0, 1):          # videos.subsample(2, start_idx=par) (evaluator.py:111,115)
                vis.append(v[par::2]); vid.append(w[par::2]); text.append(f32(it["text"]))
                sizes.append([float(it["ori_size"][0]), float(it["ori_size"][1])])
        T = vis[0].shape[0]
        o = engine.forward(torch.stack(vis).contiguous(), torch.stack(vid).contiguous(), torch.stack(text).contiguous(),
                           f32(batch[0]["pos"])[:1].contiguous() if batch[0].get("pos") is not None else None, ori_sizes_hw=torch.tensor(sizes, device=dev),
                           want=["att_sequences", "boxes_px", "sted_idx", "choose2"], raw=raw)
        boxes, att = o["boxes_px"].cpu().tolist(), o["att_sequences"].cpu().tolist()
        idx, chosen = o["sted_idx"].cpu().tolist(), (o["choose2"] > 0.5).cpu().numpy()
        passes = [({}, {}, {}, {}), ({}, {}, {}, {})]
        for q, it in enumerate(batch):
            vkey, fids = it["item_id"], [int(f) for f in it["frame_ids"]]
            act = np.asarray(it.get("actioness", np.ones(len(fids))))
            for par in (0, 1):
                c, pf = 2 * q + par, fids[par::2]
                s, e = idx[c]
                bb, aa, tt, kk = passes[par]
                bb[vkey] = {pf[j]: [boxes[c][j]] for j in range(T)}
                aa[vkey] = {pf[j]: [att[c][j]] for j in range(T)}
                tt[vkey] = {"sted": [pf[s], pf[e] + 1], "qtype": it.get("qtype", "none")}       # postprocessor.py:46-48
                kk[vkey] = precision_recall(np.nonzero(chosen[c])[0].tolist(), np.nonzero(act[par::2])[0].tolist())
        bbox_pred, att_pred, temp_pred, kf_pred = merge_even_odd(passes[0], passes[1])
        evaluator.update(bbox_pred)
        evaluator.update_att(att_pred)
        evaluator.update_kf_pr(kf_pred)
        evaluator.video_update(temp_pred)
    evaluator.synchronize_between_processes()
    return evaluator.summarize(main_process=(rank == 0))
