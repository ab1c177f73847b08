"""GPU: the reference's WHOLE `VSTGNet.forward` (vgqa/core/grounding_net.py:88-203) reproduced from pixels and token ids by this
library alone — ResNet101 (csrc/resnet.cu), Video-Swin-T (csrc/swin.cu), the RoBERTa tower + resizer, input_proj / input_proj2,
PositionEmbeddingSine, the encoder, the classifiers, both decoder passes and the heads — against the golden of the unmodified
reference model run in fp32 on the same seeded frames, token ids and weights (tests/golden/make_golden_full.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_path
from make_golden_full import FRONT_END_CH, full_frames, full_mask, full_state_dict  # noqa: F401

# the second: right / bottom padding, masked (7x7 map: 2 columns, 1 row); the third: 256 px (8x8 maps) and 12 frames — the Video-Swin
# stages pad 12 → 16 frames and 64 / 32 / 16 / 8 → 70 / 35 / 21 / 14 positions
FIXTURES = ["full_vstgnet_T16_224_s0", "full_vstgnet_T8_224_masked_s1", "full_vstgnet_T12_256_s2"]

pytestmark = pytest.mark.gpu

# Tolerances (max-abs, outputs are O(1) logits / sigmoids): the hot path alone holds 2e-2 on exact inputs (test_parity_gpu.py); here
# its inputs come from 101 + 24 + 12 layers of bf16 extractors (ResNet map: mean deviation 0.8 % of mean |y|).  Measured on a B200:
# pred_boxes 4e-4, pred_sted 5.5e-3, pred_actioness 3.2e-3, logits_f_m 8.2e-3, logits_f_a 2.2e-2, logits_r_a / r_m 1.4e-2,
# att_sequences 3.2e-3, actioness_pass1 7e-4.
TOL = {"pred_boxes": 5e-3, "pred_sted": 4e-2, "pred_actioness": 4e-2, "logits_f_m": 4e-2, "logits_f_a": 4e-2, "logits_r_a": 4e-2,
       "logits_r_m": 4e-2, "att_sequences": 1e-2, "actioness_pass1": 1e-2}

@pytest.mark.parametrize("name", FIXTURES)
def test_whole_forward_from_pixels_matches_the_reference_model(name):
    from vgqa_b200.engine import GroundingEngine
    g = np.load(golden_path(name))
    T, R, seed = (int(g[k]) for k in ("T", "R", "seed"))
    pad = tuple(int(v) for v in g["pad"])
    ids = torch.from_numpy(g["text_ids"]).cuda()
    L = ids.shape[1]
    h = R // 32
    eng = GroundingEngine(full_state_dict(seed), max_clips=1, max_frames=T, max_hw=h * h, max_text=L)
    frames = torch.from_numpy(full_frames(seed, T, R, pad)).cuda()
    vis_map = eng.resnet_backbone(frames).view(1, T, h, h, 2048)        # `vis_res_features` (channels-last bf16)
    vid_map = eng.swin_backbone(frames, 1)                               # `vid_features_all['3']`
    kw = {}
    if any(pad):   # BackboneBase.forward (backbone.py:92-96): nearest interpolation of the pixel mask to the map; [:, 0, 0] as grounding_net does
        m = torch.nn.functional.interpolate(torch.from_numpy(full_mask(T, R, pad))[None].float(), size=(h, h)).bool()[0]
        m[:, 0, 0] = False
        kw = {"vis_mask": m.reshape(T, h * h).to(torch.uint8).cuda().contiguous(), "text_mask": torch.zeros(1, L, dtype=torch.uint8, device="cuda")}
    want = ["pred_boxes", "pred_sted", "pred_actioness", "logits_f_m", "logits_f_a", "logits_r_a", "logits_r_m", "att_sequences",
            "choose1", "choose2", "actioness_pass1", "aux_boxes", "aux_sted"]
    o = eng.forward(vis_map, vid_map, None, None, raw=True, text_ids=ids, want=want, **kw)      # free-running: no decision is forced
    torch.cuda.synchronize()
    o = {k: v.float().cpu().numpy() for k, v in o.items()}
    ref = {"pred_boxes": g["pred_boxes"], "pred_sted": g["pred_sted"][0], "pred_actioness": g["pred_actioness"][0, :, 0],
           "logits_f_m": g["logits_f_m"], "logits_f_a": g["logits_f_a"], "logits_r_a": g["logits_r_a"][0], "logits_r_m": g["logits_r_m"][0],
           "att_sequences": g["att_sequences"][0], "actioness_pass1": g["actioness_pass1"]}
    # the two frame selections are the reference's (its margins: |att - theta| >= 0.036, |actioness - 0.5| >= 0.04)
    sel = lambda w: np.flatnonzero(w.reshape(-1)[:T] > 0.5).tolist()
    assert sel(o["choose1"]) == g["choose1"].tolist() and sel(o["choose2"]) == g["choose2"].tolist()
    worst = {k: float(np.abs(o[k][0].reshape(r.shape) - r).max()) for k, r in ref.items()}
    print("whole-model max-abs errors:", {k: round(v, 4) for k, v in worst.items()})
    bad = {k: v for k, v in worst.items() if not v <= TOL[k]}
    assert not bad, f"max-abs errors over tolerance: {bad} (all: {worst})"
    # the decoded answer: same (start, end) frames as the reference's scores give, boxes within 2 % of the frame
    sted = g["pred_sted"][0]
    s_ref, e_ref = int(sted[:, 0].argmax()), int(sted[:, 1].argmax())
    s, e = int(o["pred_sted"][0][:, 0].argmax()), int(o["pred_sted"][0][:, 1].argmax())
    gap = lambda v: float(np.sort(v)[-1] - np.sort(v)[-2])
    if gap(sted[:, 0]) > 4 * worst["pred_sted"]:
        assert s == s_ref
    if gap(sted[:, 1]) > 4 * worst["pred_sted"]:
        assert e == e_ref
    # the auxiliary outputs (decoder layers 1..5 of 6; aux_boxes / aux_sted hold all six layers)
    for i in range(5):
        np.testing.assert_allclose(o["aux_boxes"][i][0].reshape(T, 4), g[f"aux{i}_pred_boxes"].reshape(T, 4), atol=5e-3, err_msg=f"aux {i}")
        np.testing.assert_allclose(o["aux_sted"][i][0].reshape(T, 2), g[f"aux{i}_pred_sted"].reshape(T, 2), atol=4e-2, err_msg=f"aux {i}")
    eng.close()


@pytest.mark.parametrize("name", FIXTURES)
def test_vstgnet_dropin_with_every_extractor_in_the_library(name):
    """`B200VSTGNet` given the whole state_dict: no PyTorch module of the model is called (the extractor modules passed in raise) —
    `videos.tensors` and the tokenizer's ids go into the library; output dict as grounding_net.py:164-203."""
    from make_golden import make_cfg
    from vgqa_b200 import modules as M
    g = np.load(golden_path(name))
    T, R, seed = (int(g[k]) for k in ("T", "R", "seed"))
    pad, sentence = tuple(int(v) for v in g["pad"]), str(g["sentence"])
    ids = torch.from_numpy(g["text_ids"]).long()

    class TextEncoder:            # what the drop-in calls of the reference's RoBERTa wrapper: the tokenizer (bert.py:50,65)
        @staticmethod
        def tokenizer(texts, padding, return_tensors):
            assert texts == [" " + sentence]       # grounding_net.py:110: subject + " " + sentence
            return {"input_ids": ids, "attention_mask": torch.ones_like(ids)}

    def boom(*a, **k):
        raise AssertionError("a PyTorch extractor module was called")

    model = M.B200VSTGNet(make_cfg(), boom, boom, TextEncoder(), boom, boom, full_state_dict(seed), verb_label2={"0": {"sub": ""}},
                          max_frames=T, max_hw=(R // 32) ** 2, max_text=ids.shape[1]).eval()
    assert model.fused_backbones and model.fused_front_end and model.fused_text_tower
    videos = M.NestedTensor(torch.from_numpy(full_frames(seed, T, R, pad)).cuda(), torch.from_numpy(full_mask(T, R, pad)).cuda(), [T])
    out = model(videos, [sentence], [{"item_id": 0, "actioness": torch.ones(T, device="cuda")}])
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "logits_f_m", "logits_f_a", "logits_r_a", "logits_r_m", "att_sequences"):
        got = out[k].float().cpu().numpy()
        assert got.shape == g[k].shape, (k, got.shape, g[k].shape)
        assert float(np.abs(got - g[k]).max()) <= TOL[k], (k, float(np.abs(got - g[k]).max()))
    assert len(out["aux_outputs"]) == 5 and out["pr"] == (1.0, 1.0)
    for i, a in enumerate(out["aux_outputs"]):
        assert float(np.abs(a["pred_boxes"].cpu().numpy() - g[f"aux{i}_pred_boxes"]).max()) <= TOL["pred_boxes"], i


def test_predictor_from_frames_and_token_ids():
    """`GroundingPredictor(raw_inputs=True)` fed with the sampled FRAMES and the token ids (grounding.py:142-244 after decoding /
    resizing / tokenising): identical to feeding it the maps of the library's extractors computed pass by pass."""
    from vgqa_b200.predict import GroundingPredictor
    g = np.load(golden_path("full_vstgnet_T16_224_s0"))
    T, R, seed = 8, 224, int(g["seed"])
    ids = g["text_ids"][0]
    pred = GroundingPredictor(full_state_dict(seed), sample_num=T, max_hw=49, max_text=len(ids), raw_inputs=True)
    frames = torch.from_numpy(full_frames(seed, 2 * T, R)).cuda()
    fids = list(range(5, 5 + 2 * T * 4, 4))
    got = pred.predict_many([{"frames": frames, "text_ids": ids, "frame_ids": fids, "ori_size": (360, 640), "fps": 25.0}])[0]
    eng = pred.engine
    maps = [eng.extract_features(frames[par::2].contiguous(), 1) for par in (0, 1)]
    nchw = lambda m: m[0].permute(0, 3, 1, 2).float()                      # [T, C, 7, 7] as the "vis" / "vid" items are laid out
    vis = torch.stack([nchw(maps[0][0]), nchw(maps[1][0])], 1).reshape(2 * T, 2048, 7, 7)     # interleave the passes again
    vid = torch.stack([nchw(maps[0][1]), nchw(maps[1][1])], 1).reshape(2 * T, 768, 7, 7)
    ref = pred.predict_many([{"vis": vis, "vid": vid, "text_ids": ids, "frame_ids": fids, "ori_size": (360, 640), "fps": 25.0}])[0]
    assert got["temporal"] == ref["temporal"] and got["temporal"]["score"] == 1.0
    assert [t["frame"] for t in got["tube"]] == [t["frame"] for t in ref["tube"]] == list(range(fids[0], fids[-1] + 1))
    np.testing.assert_allclose(np.asarray([t["bbox"] for t in got["tube"]]), np.asarray([t["bbox"] for t in ref["tube"]]), atol=1.0)   # pixels
    pred.close()
