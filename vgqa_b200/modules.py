"""Python mirror of the reference's model seam for the hot path.

    build_hot_path(cfg, state_dict)        fused replacement of ground_encoder + *_clas + ground_decoder + MLP heads
    build_encoder(cfg, state_dict)         CrossModalEncoder drop-in: (videos, vis_pos, texts, vid) -> dict   (modal_encoder.py:41-85)
    B200VSTGNet                            VSTGNet drop-in: forward(videos, texts, targets, iteration_rate=-1) -> dict
                                           (grounding_net.py:88-204); the extractors are the caller's PyTorch modules — or, when
                                           their weights are in the state_dict, the library's own (ResNet101, Video-Swin-T, RoBERTa)
    NestedTensor                           (tensors, mask, durations) container (utils/training_utils.py:44-72)

All math of the path runs in libvgqa_b200.so (vgqa_b200/engine.py); there is no PyTorch fallback.
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Optional

import torch

from .engine import GroundingEngine


class NestedTensor:
    def __init__(self, tensors, mask, durations):
        self.tensors, self.mask, self.durations = tensors, mask, durations

    def to(self, *a, **k):
        return NestedTensor(self.tensors.to(*a, **k), None if self.mask is None else self.mask.to(*a, **k), self.durations)

    def decompose(self):
        return self.tensors, self.mask, self.durations

    def subsample(self, stride, start_idx=0):
        ts = [v[start_idx::stride] for v in torch.split(self.tensors, self.durations, dim=0)]
        ms = [m[start_idx::stride] for m in torch.split(self.mask, self.durations, dim=0)]
        return NestedTensor(torch.cat(ts, 0), torch.cat(ms, 0), [t.shape[0] for t in ts])


def _cfg_get(cfg, path, default):
    cur = cfg
    for k in path.split("."):
        cur = getattr(cur, k, None) if not isinstance(cur, dict) else cur.get(k)
        if cur is None:
            return default
    return cur


def _engine_from_cfg(cfg, state_dict, **cap) -> GroundingEngine:
    return GroundingEngine(
        state_dict,
        enc_layers=_cfg_get(cfg, "MODEL.VSTG.ENC_LAYERS", 6), dec_layers=_cfg_get(cfg, "MODEL.VSTG.DEC_LAYERS", 6),
        ffn_dim=_cfg_get(cfg, "MODEL.VSTG.FFN_DIM", 2048), app_num=_cfg_get(cfg, "DATASET.APP_NUM", 20),
        mot_num=_cfg_get(cfg, "DATASET.MOT_NUM", 34), max_video_len=_cfg_get(cfg, "INPUT.MAX_VIDEO_LEN", 200), **cap)


def precision_recall(predicted: List[int], true: List[int]):
    ps, ts = set(predicted), set(true)
    inter = len(ps & ts)
    return (0 if not ps else inter / len(ps)), (0 if not ts else inter / len(ts))


class HotPath(torch.nn.Module):
    """Everything VSTGNet.forward does between the feature extractors and the output dict (grounding_net.py:114-202)."""

    def __init__(self, cfg, state_dict: Mapping[str, Any], max_clips=1, max_frames=None, max_hw=196, max_text=64,
                 use_cuda_graph=False):
        super().__init__()
        assert _cfg_get(cfg, "MODEL.VSTG.HIDDEN", 256) == 256 and _cfg_get(cfg, "MODEL.VSTG.HEADS", 8) == 8
        assert _cfg_get(cfg, "MODEL.VSTG.FROM_SCRATCH", True), "only the FROM_SCRATCH cross-attention branch is built"
        assert not _cfg_get(cfg, "MODEL.VSTG.USE_LEARN_TIME_EMBED", False)
        mv = _cfg_get(cfg, "INPUT.MAX_VIDEO_LEN", 200)
        self.use_aux_loss = _cfg_get(cfg, "SOLVER.USE_AUX_LOSS", True)
        self.engine = _engine_from_cfg(cfg, state_dict, max_clips=max_clips, max_frames=max_frames or mv + 1,
                                       max_hw=max_hw, max_text=max_text, use_cuda_graph=use_cuda_graph)

    @torch.no_grad()
    def forward(self, vis_features, vis_mask, vis_pos, text_mask, text_features, vid_features, iteration_rate=-1, raw=False,
                text_ids=None, nhwc=False):
        """vis/vid_features [T,256,H,W], vis_mask [T,H,W] bool, vis_pos [T,256,H,W], text_features [L,1,256],
        text_mask [1,L] bool → the reference's output dict entries that depend on the hot path.
        raw=True: the features are the extractor outputs (ResNet map [T,Cv,H,W], Video-Swin map [T,Cd,H,W], RoBERTa states
        [L,1,Ct]) and input_proj / input_proj2 / text_encoder.resizer run fused inside the library (grounding_net.py:101,105;
        bert.py:73) — their weights must be in the state_dict.  text_ids [1, L] int32 (raw=True): RoBERTa token ids instead of
        text_features; the text tower (`text_encoder.body.*`) then runs inside the library too.  nhwc=True (raw only): the maps are
        already channels-last bf16 [T,H,W,C] (the library's own extractors).  vis_pos=None: PositionEmbeddingSine is generated in the
        library from vis_mask (position_encoding.py:50-91)."""
        if nhwc:
            T, H, W, d = vis_features.shape
        else:
            T, d, H, W = vis_features.shape
        assert vis_pos is None or vis_pos.shape[0] == T, "{} != {}".format(vis_pos.shape[0], T)          # modal_encoder.py:44
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        masked = bool(vis_mask is not None and vis_mask.any()) or bool(text_mask is not None and text_mask.any())
        kw = {}
        if masked:
            kw["vis_mask"] = vis_mask.reshape(T, H * W).to(torch.uint8).contiguous()
            kw["text_mask"] = text_mask.reshape(1, -1).to(torch.uint8).contiguous()
            pos = None if vis_pos is None else f32(vis_pos)
        else:
            pos = None if vis_pos is None else f32(vis_pos[:1])   # PositionEmbeddingSine of an all-False mask is identical on every frame
        if text_ids is not None:
            kw["text_ids"] = text_ids.reshape(1, -1).to(torch.int32).contiguous()
        fmap = f32
        if raw and vis_features.dtype == torch.bfloat16 and vid_features.dtype == torch.bfloat16:
            # bf16 backbones: hand the maps over as channels-last bf16 (raw_layout = 1).  A channels_last tensor already IS
            # [T, H, W, C] in memory, so this is a zero-copy view; a plain NCHW bf16 tensor is permuted once.
            fmap = lambda t: t.detach().permute(0, 2, 3, 1).contiguous()
        if nhwc:
            assert raw and vis_features.dtype == torch.bfloat16 and vid_features.dtype == torch.bfloat16
            fmap = lambda t: t.detach().contiguous()
        o = self.engine.forward(fmap(vis_features)[None], fmap(vid_features)[None],
                                None if text_ids is not None else f32(text_features[:, 0])[None], pos,
                                iteration_rate=iteration_rate, raw=raw, **kw)
        out = {"pred_boxes": o["pred_boxes"][0], "logits_f_m": o["logits_f_m"][0], "logits_f_a": o["logits_f_a"][0],
               "logits_r_a": o["logits_r_a"], "logits_r_m": o["logits_r_m"], "pred_sted": o["pred_sted"],
               "pred_actioness": o["pred_actioness"][..., None], "att_sequences": o["att_sequences"]}
        if self.use_aux_loss:
            out["aux_outputs"] = [{"pred_sted": o["aux_sted"][i], "pred_boxes": o["aux_boxes"][i, 0],
                                   "pred_actioness": o["aux_actioness"][i][..., None]}
                                  for i in range(o["aux_boxes"].shape[0] - 1)]
        out["_choose_index"] = torch.nonzero(o["choose2"][0] > 0.5).flatten()
        return out


def build_hot_path(cfg, state_dict, **kw) -> HotPath:
    return HotPath(cfg, state_dict, **kw)


class CrossModalEncoder(torch.nn.Module):
    """build_encoder(cfg) drop-in: `(videos=NestedTensor, vis_pos, texts=(mask, text, _), vid) -> dict` with the keys of
    modal_encoder.py:76-83.  `encoded_feature` is returned in the reference layout (S, T, 256)."""

    def __init__(self, cfg, state_dict, **cap):
        super().__init__()
        self.engine = _engine_from_cfg(cfg, state_dict, **cap)

    @torch.no_grad()
    def forward(self, videos=None, vis_pos=None, texts=None, vid=None):
        vis_features, vis_mask, vis_durations = videos.decompose()
        assert vis_pos.shape[0] == sum(vis_durations), "{} != {}".format(vis_pos.shape[0], sum(vis_durations))
        assert len(vis_durations) == 1, "the reference encoder handles one clip per call (modal_encoder.py:59)"
        T, _, H, W = vis_features.shape
        text_mask, text_features, _ = texts
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        vm = vis_mask.clone()
        vm[:, 0, 0] = False
        masked = bool(vm.any()) or bool(text_mask.any())
        kw = {}
        if masked:
            kw = {"vis_mask": vm.reshape(T, H * W).to(torch.uint8).contiguous(),
                  "text_mask": text_mask.reshape(1, -1).to(torch.uint8).contiguous()}
        o = self.engine.encode(f32(vis_features)[None], f32(vid)[None], f32(text_features[:, 0])[None],
                               f32(vis_pos) if masked else f32(vis_pos[:1]), **kw)
        L = text_features.shape[0]
        mask = torch.cat([vm.flatten(1), text_mask.expand(T, L), vm.flatten(1)], dim=1)
        return {"encoded_feature": o["encoded_feature"].permute(1, 0, 2), "encoded_mask": mask,
                "frames_cls": o["frames_cls"], "videos_cls": o["frames_cls"].mean(0), "durations": vis_durations,
                "fea_map_size": (H, W)}


def build_encoder(cfg, state_dict, **cap) -> CrossModalEncoder:
    return CrossModalEncoder(cfg, state_dict, **cap)


class B200VSTGNet(torch.nn.Module):
    """VSTGNet with the hot path swapped for the B200 library.  The feature extractors are the caller's modules with
    the reference's call signatures: vis_encoder(videos) -> (NestedTensor, pos); vid(tensors, T) -> {'3': feats};
    text_encoder(texts, device) -> ((mask, text, raw), cls); input_proj / input_proj2 are 1x1 convs."""

    def __init__(self, cfg, vis_encoder, vid, text_encoder, input_proj, input_proj2, state_dict, verb_label=None,
                 verb_label2=None, fused_front_end=None, **cap):
        super().__init__()
        self.vis_encoder, self.vid, self.text_encoder = vis_encoder, vid, text_encoder
        self.input_proj, self.input_proj2 = input_proj, input_proj2
        # fused front end (csrc/input_proj.cu): the 1x1 convs and the text resizer run inside the library when their weights
        # are in the state_dict; `input_proj` / `input_proj2` modules are then not called
        has = all(k in state_dict for k in ("input_proj.weight", "input_proj2.weight", "text_encoder.resizer.fc.weight"))
        self.fused_front_end = has if fused_front_end is None else bool(fused_front_end)
        assert has or not self.fused_front_end, "fused_front_end needs input_proj / input_proj2 / text_encoder.resizer weights"
        # fused text tower: only the TOKENIZER of the text encoder is called (bert.py:65), RoBERTa itself runs inside the library
        self.fused_text_tower = self.fused_front_end and hasattr(text_encoder, "tokenizer") and \
            "text_encoder.body.embeddings.word_embeddings.weight" in state_dict
        # fused extractors (csrc/resnet.cu, csrc/swin.cu): with the `vis_encoder.0.body.*` and `vid.*` weights in the state_dict the
        # ResNet101 / Video-Swin-T modules are not called either — `videos.tensors` go straight into the library
        self.fused_backbones = self.fused_front_end and "vis_encoder.0.body.conv1.weight" in state_dict and \
            "vid.patch_embed.proj.weight" in state_dict
        self.hot = HotPath(cfg, state_dict, **cap)
        self.verb_label = verb_label or {}
        self.verb_label2 = verb_label2 or {}

    @classmethod
    def from_reference(cls, model, cfg, **cap):
        """Wrap an already-built reference `VSTGNet` (shares its backbones, takes its state_dict)."""
        return cls(cfg, model.vis_encoder, model.vid, model.text_encoder, model.input_proj, model.input_proj2,
                   model.state_dict(), getattr(model, "verb_label", None), getattr(model, "verb_label2", None), **cap)

    @torch.no_grad()
    def forward(self, videos, texts, targets, iteration_rate: int = -1):
        frames = videos.tensors
        R = frames.shape[-1]
        nhwc = self.fused_backbones and frames.shape[0] >= 8 and frames.shape[-2] == R and R % 32 == 0 and 224 <= R <= 512
        if nhwc:
            # the library's extractors: layer4 map of ResNet101 and the last Video-Swin-T stage, both channels-last bf16; the mask is
            # interpolated as BackboneBase.forward does (backbone.py:92-96) and the positional encoding is generated from it
            eng = self.hot.engine
            frames = frames.detach().to(torch.float32).contiguous()
            vis_features, vid_features = (m[0] for m in eng.extract_features(frames, 1))
            vis_mask = torch.nn.functional.interpolate(videos.mask[None].float(), size=vis_features.shape[1:3]).to(torch.bool)[0]
            vis_pos = None
        else:
            vis_outputs, vis_pos = self.vis_encoder(videos)
            vis_res, vis_mask, vis_durations = vis_outputs.decompose()
            vid_res = self.vid(videos.tensors, len(videos.tensors))["3"]
            fused = self.fused_front_end
            vis_features = vis_res if fused else self.input_proj(vis_res)
            vid_features = vid_res if fused else self.input_proj2(vid_res)
        fused = self.fused_front_end
        info_key = str(targets[0]["item_id"])
        labels = self.verb_label if self.training else self.verb_label2
        texts = [labels[info_key]["sub"] + " " + texts[0]]
        vm = vis_mask.clone()
        vm[:, 0, 0] = False
        if self.fused_text_tower:
            tok = self.text_encoder.tokenizer(texts, padding="longest", return_tensors="pt")      # bert.py:65
            ids = tok["input_ids"].to(vis_features.device)
            text_mask = tok["attention_mask"].to(vis_features.device).ne(1)                       # bert.py:70
            out = self.hot(vis_features, vm, vis_pos, text_mask, None, vid_features, iteration_rate, raw=True, text_ids=ids, nhwc=nhwc)
        else:
            (text_mask, text_features, text_memory), _ = self.text_encoder(texts, vis_features.device)
            out = self.hot(vis_features, vm, vis_pos, text_mask, text_memory if fused else text_features, vid_features,
                           iteration_rate, raw=fused, nhwc=nhwc)
        choose_index = out.pop("_choose_index").tolist()
        out["verb_labels"] = labels.get(info_key, {}).get("verb_index_list", [])
        out["attr_labels"] = labels.get(info_key, {}).get("adj_index_list", [])
        gt_index = torch.nonzero(targets[0]["actioness"]).flatten().tolist()
        out["pr"] = precision_recall(choose_index, gt_index)
        return out
