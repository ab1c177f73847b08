"""Persistent predict() pipeline over the hot path.

The reference's `vgqa.inference.grounding.predict` (grounding.py:142-244) rebuilds the model and reloads the checkpoint on
every call, samples 2*TRAIN_SAMPLE_NUM frames, runs TWO forwards (even / odd frames, grounding.py:180-212) and merges them
(grounding.py:214-244).  `GroundingPredictor` keeps one packed `GroundingEngine` alive, runs the even and the odd pass as ONE
two-clip batch (clips never interact — SURVEY §0 fact 5), decodes both passes with the library's PostProcess kernel and
returns the reference's output schema.  Several queries can be served per call (`predict_many`): 2*Q clips in one batch.

Inputs are the tensors at the hot-path boundary (outputs of input_proj / input_proj2 / the text resizer), or — raw_inputs=True —
the extractor outputs, or the sampled, resized and normalised FRAMES themselves plus the tokenizer's ids: with the extractor
weights in the state_dict (`vis_encoder.0.body.*`, `vid.*`, `text_encoder.body.*`) ResNet101, Video-Swin-T and RoBERTa run inside
the library as well.  Video decoding / resizing (decord, grounding.py:27-77) and the tokenizer stay the caller's.
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Sequence

import torch

from .engine import GroundingEngine
from .postprocess import merge_predictions


class GroundingPredictor:
    """One engine, many predict() calls.  `max_queries` bounds the number of (clip, query) pairs per call."""

    def __init__(self, state_dict: Mapping[str, Any], *, sample_num=64, max_hw=49, max_text=64, max_queries=1,
                 max_video_len=200, use_cuda_graph=True, raw_inputs=False, **engine_kw):
        """raw_inputs=True: items carry the extractor outputs ("vis" [2T,Cv,H,W] ResNet map, "vid" [2T,Cd,H,W] Video-Swin map,
        "text" [L,Ct] RoBERTa states) and input_proj / input_proj2 / the text resizer run fused inside the library."""
        self.sample_num = int(sample_num)
        self.raw_inputs = bool(raw_inputs)
        self.engine = GroundingEngine(state_dict, max_clips=2 * int(max_queries), max_frames=self.sample_num, max_hw=max_hw,
                                      max_text=max_text, max_video_len=max_video_len, use_cuda_graph=use_cuda_graph,
                                      **engine_kw)
        self.max_queries = int(max_queries)

    @torch.no_grad()
    def predict_many(self, items: Sequence[Mapping[str, Any]]) -> List[Dict[str, Any]]:
        """items[q]: {"vis": [2T,256,H,W], "vid": [2T,256,H,W], "text": [L,256], "frame_ids": 2T ints (ascending, as sampled
        by predict()), "ori_size": (h, w), "fps": float}.  All items share T, H, W, L.  "pos" ([1,256,H,W]) is optional: the
        library generates PositionEmbeddingSine itself (predict() resizes to a square, so nothing is padded, grounding.py:177).
        With raw_inputs=True the channel counts are those of the extractors (see __init__); an item may then carry
        "text_ids" ([L] RoBERTa token ids) instead of "text" — the text tower runs inside the library too (all items or none) —
        and "frames" ([2T,3,R,R] fp32: the sampled frames after the reference's resize + normalisation, R a multiple of 32 in 224..512, T >= 8) instead of "vis" / "vid": both extractors then run inside the library (all items or none).
        Returns one {"temporal": {...}, "tube": [...]} dict per item (grounding.py:227-244)."""
        Q = len(items)
        if Q == 0:
            return []
        if Q > self.max_queries:
            raise ValueError(f"{Q} queries exceed max_queries={self.max_queries}")
        dev = self.engine.device
        f32 = lambda t: torch.as_tensor(t, dtype=torch.float32, device=dev)
        vis, vid, text, sizes = [], [], [], []
        use_ids = self.raw_inputs and all("text_ids" in it for it in items)
        use_frames = self.raw_inputs and all("frames" in it for it in items)
        for it in items:
            v = f32(it["frames"] if use_frames else it["vis"])
            w = None if use_frames else f32(it["vid"])
            n = v.shape[0]
            if n < 2 or n % 2 != 0 or n // 2 > self.sample_num:
                raise ValueError("predict() samples an even number of frames, at most 2*sample_num")   # grounding.py:137-138,157
            assert len(it["frame_ids"]) == n, "one frame id per sampled frame"
            for par in (0, 1):                       # even pass, odd pass (grounding.py:163-168)
                vis.append(v[par::2])
                if not use_frames:
                    vid.append(w[par::2])
                text.append(torch.as_tensor(it["text_ids"], dtype=torch.int32, device=dev) if use_ids else f32(it["text"]))
                sizes.append([float(it["ori_size"][0]), float(it["ori_size"][1])])
        T = vis[0].shape[0]
        assert all(x.shape[0] == T for x in vis), "all queries of a call must sample the same number of frames"
        text = torch.stack(text).contiguous()
        if use_frames:   # every pass is a clip of its own for the extractors too (the Video-Swin windows span the frames of ONE pass)
            vis, vid = self.engine.extract_features(torch.cat(vis).contiguous(), 2 * Q)
        else:
            vis, vid = torch.stack(vis).contiguous(), torch.stack(vid).contiguous()
        pos = f32(items[0]["pos"])[:1].contiguous() if items[0].get("pos") is not None else None
        o = self.engine.forward(vis, vid, None if use_ids else text, pos, ori_sizes_hw=torch.tensor(sizes, device=dev),
                                want=["att_sequences", "boxes_px", "sted_idx"], raw=self.raw_inputs,
                                text_ids=text if use_ids else None)
        boxes = o["boxes_px"].reshape(2 * Q, T, 4).cpu().tolist()     # one D2H copy per output
        att = o["att_sequences"].reshape(2 * Q, T).cpu().tolist()
        idx = o["sted_idx"].reshape(2 * Q, 2).cpu().tolist()
        results = []
        for q, it in enumerate(items):
            fids = [int(f) for f in it["frame_ids"]]
            passes = []
            for par in (0, 1):
                c = 2 * q + par
                pf = fids[par::2]
                s, e = idx[c]
                passes.append(({0: {pf[j]: [boxes[c][j]] for j in range(T)}}, {0: {pf[j]: [att[c][j]] for j in range(T)}},
                               {0: {"sted": [pf[s], pf[e] + 1], "qtype": it.get("qtype", "declar")}}, {}))   # postprocessor.py:46-48
            results.append(merge_predictions(passes[0], passes[1], float(it.get("fps", 25.0))))
        return results

    def predict(self, vis, vid, text, pos=None, frame_ids=None, ori_size=None, fps=25.0, qtype="declar") -> Dict[str, Any]:
        return self.predict_many([{"vis": vis, "vid": vid, "text": text, "pos": pos, "frame_ids": frame_ids,
                                   "ori_size": ori_size, "fps": fps, "qtype": qtype}])[0]

    def close(self):
        self.engine.close()
