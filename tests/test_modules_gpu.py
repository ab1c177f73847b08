"""GPU: the Python mirror of the reference seam (vgqa_b200/modules.py, postprocess.py) — output schema of the VSTGNet
drop-in, the CrossModalEncoder seam, PostProcess and the predict()-schema merge, checked against the goldens."""
import numpy as np
import pytest
import torch

from conftest import golden_path
from make_golden import make_cfg
from oracle import vgqa_oracle as O
from vgqa_b200 import modules as M
from vgqa_b200 import postprocess as PP

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _case(name="cfg1_T32_7x7_L20_s1"):
    g = np.load(golden_path(name))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    sd = O.synth_state_dict(seed)
    vis, vid, pos, text = O.synth_inputs(seed, T, H, W, L)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return g, sd, t(vis), t(vid), t(pos), t(text), T, H, W, L


def test_hot_path_output_dict_matches_reference_schema_and_values():
    g, sd, vis, vid, pos, text, T, H, W, L = _case()
    hot = M.build_hot_path(make_cfg(), sd, max_frames=64, max_hw=49, max_text=20)
    out = hot(vis, torch.zeros(T, H, W, dtype=torch.bool, device="cuda"), pos, torch.zeros(1, L, dtype=torch.bool, device="cuda"),
              text, vid)
    shapes = {"pred_boxes": (T, 4), "pred_sted": (1, T, 2), "pred_actioness": (1, T, 1), "logits_f_m": (T,), "logits_f_a": (T,),
              "logits_r_a": (1, 20), "logits_r_m": (1, 34), "att_sequences": (1, T)}
    for k, shp in shapes.items():
        assert tuple(out[k].shape) == shp, k
        assert float(np.abs(out[k].cpu().numpy() - g[k]).max()) <= TOL, k
    assert len(out["aux_outputs"]) == 5 and set(out["aux_outputs"][0]) == {"pred_sted", "pred_boxes", "pred_actioness"}
    np.testing.assert_allclose(out["aux_outputs"][2]["pred_boxes"].cpu().numpy(), g["aux_boxes"][2], atol=TOL)
    assert out["_choose_index"].tolist() == g["choose_pass2"].tolist()


def test_cross_modal_encoder_seam():
    g, sd, vis, vid, pos, text, T, H, W, L = _case("tiny_T3_3x4_L3")
    enc = M.build_encoder(make_cfg(), sd, max_frames=8, max_hw=16, max_text=8)
    videos = M.NestedTensor(vis, torch.zeros(T, H, W, dtype=torch.bool, device="cuda"), [T])
    info = enc(videos=videos, vis_pos=pos, texts=(torch.zeros(1, L, dtype=torch.bool, device="cuda"), text, None), vid=vid)
    S = 2 * H * W + L
    assert tuple(info["encoded_feature"].shape) == (S, T, 256) and tuple(info["encoded_mask"].shape) == (T, S)
    assert info["fea_map_size"] == (H, W) and info["durations"] == [T]
    np.testing.assert_allclose(info["encoded_feature"][:, 0].cpu().numpy(), g["enc_frame0"].astype(np.float32), atol=6e-2)
    np.testing.assert_allclose(info["frames_cls"].cpu().numpy(), g["frames_cls"], atol=TOL)
    with pytest.raises(AssertionError):
        enc(videos=videos, vis_pos=pos[:1], texts=(torch.zeros(1, L, dtype=torch.bool, device="cuda"), text, None), vid=vid)


class _FakeBackbone(torch.nn.Module):
    """Stands in for ResNet101+PositionEmbeddingSine / Video-Swin / RoBERTa: returns the synthetic features."""

    def __init__(self, vis, vid, pos, text):
        super().__init__()
        self.vis, self.vidf, self.pos, self.text = vis, vid, pos, text


def test_vstgnet_dropin_single_forward_and_predict_schema():
    g, sd, vis, vid, pos, text, T, H, W, L = _case()
    fb = _FakeBackbone(vis, vid, pos, text)
    vis_encoder = lambda videos: (M.NestedTensor(fb.vis, videos.mask[:, :H, :W], videos.durations), fb.pos)
    vid_model = lambda tensors, n: {"3": fb.vidf}
    text_encoder = lambda texts, device: ((torch.zeros(1, L, dtype=torch.bool, device=device), fb.text, None), None)
    ident = torch.nn.Identity()
    model = M.B200VSTGNet(make_cfg(), vis_encoder, vid_model, text_encoder, ident, ident, sd,
                          verb_label2={"0": {"sub": ""}}, max_frames=64, max_hw=49, max_text=20).eval()
    videos = M.NestedTensor(torch.zeros(T, 3, 224, 224, device="cuda"), torch.zeros(T, 224, 224, dtype=torch.bool, device="cuda"), [T])
    fids = g["frame_ids"].tolist()
    targets = [{"item_id": 0, "vid": "x", "ori_size": tuple(int(v) for v in g["ori_size"]), "qtype": "declar",
                "frame_ids": fids, "actioness": torch.ones(T, device="cuda")}]
    post = PP.build_postprocessors()
    bbox, att, temp, kf = PP.single_forward(None, model, videos, ["a person jumping"], targets, "cuda", post)
    assert sorted(bbox[0]) == fids and len(bbox[0][fids[0]][0]) == 4
    got = np.asarray([bbox[0][f][0] for f in fids])
    np.testing.assert_allclose(got, g["post_boxes"], atol=TOL * 640)
    assert temp[0]["qtype"] == "declar" and len(temp[0]["sted"]) == 2
    assert kf[0] == (1.0, 1.0) or 0 <= kf[0][0] <= 1          # precision/recall of the chosen frames vs all-ones actioness
    # even/odd merge → predict() schema
    odd = [{**targets[0], "frame_ids": [f + 1 for f in fids]}]
    p2 = PP.single_forward(None, model, videos, ["a person jumping"], odd, "cuda", post)
    res = PP.merge_predictions((bbox, att, temp, kf), p2, fps=25.0)
    assert [t["frame"] for t in res["tube"]] == list(range(fids[0], fids[-1] + 2))
    assert res["temporal"]["score"] == 1.0 and res["temporal"]["start"] <= res["temporal"]["end"]


def test_postprocess_module_matches_golden():
    g, sd, vis, vid, pos, text, T, H, W, L = _case("cfg2_T64_7x7_L20_s2")
    outputs = {"pred_sted": torch.from_numpy(g["pred_sted"]).cuda(), "pred_boxes": torch.from_numpy(g["pred_boxes"]).cuda(),
               "att_sequences": torch.from_numpy(g["att_sequences"]).cuda(), "pr": (0, 0)}
    sizes = torch.tensor([[int(g["ori_size"][0]), int(g["ori_size"][1])]] * T, device="cuda")
    boxes, att, steds, _ = PP.PostProcess()(outputs, sizes, [g["frame_ids"].tolist()], [T])
    assert steds == g["post_sted"].tolist()
    np.testing.assert_allclose(boxes.cpu().numpy(), g["post_boxes"], atol=1e-3)


@pytest.mark.parametrize("name", ["ev_cfg1_T32_7x7_L20_s0", "ev_masked_T32_7x7_L20_s20", "ev_ragged_T6_4x5_L7_masked_s0"])
def test_position_embedding_generated_in_the_library(name):
    """vgqa_inputs.pos == NULL: the library generates PositionEmbeddingSine(128, normalize=True) itself (vision/
    position_encoding.py:50-91) — from the padding mask when there is one.  The forward must give what it gives with the
    reference's own `pos` tensor (stored in the golden) passed in, and both must match the golden."""
    import numpy as np
    from conftest import golden_path
    from vgqa_b200.engine import GroundingEngine
    g = np.load(golden_path(name))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    masked = bool(g["masked"])
    sd = O.apply_calibration(O.synth_state_dict(seed, max_video_len=int(g["max_video_len"])), g)
    amp = float(g["event_amp"]) if "event_amp" in g.files else 0.0
    vis, vid, _, text = O.synth_event_inputs(seed, T, H, W, L, amp=amp) if amp > 0 else O.synth_inputs(seed, T, H, W, L)
    vm, tm = O.synth_masks(masked, T, H, W, L)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    kw = dict(vis_mask=t(vm.reshape(T, -1).astype(np.uint8)), text_mask=t(tm.astype(np.uint8))) if masked else {}
    eng = GroundingEngine(sd, max_clips=1, max_frames=T, max_hw=H * W, max_text=L)
    want = ["pred_boxes", "pred_sted", "logits_f_m", "frames_cls"]
    given = eng.forward(t(vis[None]), t(vid[None]), t(text[None, :, 0]), t(g["pos"]), want=want, **kw)
    given = {k: v.cpu().numpy() for k, v in given.items()}
    own = eng.forward(t(vis[None]), t(vid[None]), t(text[None, :, 0]), None, want=want, **kw)
    own = {k: v.cpu().numpy() for k, v in own.items()}
    for k in want:
        np.testing.assert_allclose(own[k], given[k], atol=8e-3, err_msg=k)   # pos differs by fp32 sin/cos rounding only (a few bf16 flips)
    assert float(np.abs(own["pred_boxes"][0] - g["pred_boxes"]).max()) <= 2e-2
    assert float(np.abs(own["frames_cls"] - g["frames_cls"]).max()) <= 2e-2
    eng.close()
