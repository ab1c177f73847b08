"""CPU: the oracle's front end (input_proj / input_proj2 / FeatureResizer — SURVEY.md §8f rank 2) against golden vectors
made by torch `nn.Conv2d` + the reference's own `FeatureResizer` class (tests/golden/make_golden_frontend.py), and the
front end → hot path chain against the reference modules' outputs.  fp32 vs fp32."""
import numpy as np
import pytest

from oracle import vgqa_oracle as O
from conftest import golden_path

FE_CASES = ["fe_tiny_T3_3x4_L3", "fe_cfg1_T32_7x7_L20_s0", "fe_yaml_T4_14x14_L20_s1"]


def load_case(name):
    g = np.load(golden_path(name))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    ch = tuple(int(x) for x in g["front_end_ch"])
    sd = O.synth_state_dict(seed, front_end_ch=ch)
    raw = O.synth_raw_inputs(seed, T, H, W, L, ch)
    return g, sd, raw, (T, H, W, L)


@pytest.mark.parametrize("name", FE_CASES)
def test_front_end_matches_reference_modules(name):
    g, sd, (vis_raw, vid_raw, text_raw), _ = load_case(name)
    vis, vid, text = O.front_end(sd, vis_raw, vid_raw, text_raw)
    np.testing.assert_allclose(text, g["fe_text"], atol=2e-5)
    np.testing.assert_allclose(vis[0], g["fe_vis_frame0"], atol=2e-5)
    np.testing.assert_allclose(vis[-1], g["fe_vis_frameN"], atol=2e-5)
    np.testing.assert_allclose(vid[0], g["fe_vid_frame0"], atol=2e-5)
    np.testing.assert_allclose(vid[-1], g["fe_vid_frameN"], atol=2e-5)
    assert abs(float(np.abs(vis).mean()) - float(g["fe_vis_abs_mean"])) < 1e-5


@pytest.mark.parametrize("name", ["fe_tiny_T3_3x4_L3", "fe_cfg1_T32_7x7_L20_s0"])
def test_front_end_then_hot_path_matches_reference(name):
    g, sd, raw, (T, H, W, L) = load_case(name)
    vis, vid, text = O.front_end(sd, *raw)
    pos = O.position_embedding_sine(np.zeros((T, H, W), bool))
    out = O.hot_path_forward(sd, vis, vid, pos, text, return_debug=True)
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "logits_f_m", "logits_f_a", "logits_r_a", "logits_r_m",
              "att_sequences"):
        np.testing.assert_allclose(out[k], g[k], atol=2e-4, err_msg=k)
    assert out["debug"]["choose_pass1"] == g["choose_pass1"].tolist()
    assert out["debug"]["choose_pass2"] == g["choose_pass2"].tolist()


TEXT_CASES = ["text_tiny_L5_2layers", "text_base_L20_12layers", "text_base_L64_12layers"]


@pytest.mark.parametrize("name", TEXT_CASES)
def test_roberta_oracle_matches_transformers(name):
    """oracle.roberta_encoder + feature_resizer vs transformers' RobertaModel + the reference FeatureResizer (make_golden_text.py)."""
    g = np.load(golden_path(name))
    seed, B, L, layers, vocab, pad_tail = (int(g[k]) for k in ("seed", "B", "L", "layers", "vocab", "pad_tail"))
    sd = O.synth_state_dict(seed, front_end_ch=tuple(int(x) for x in g["front_end_ch"]), text_tower=(layers, vocab))
    ids, pad = O.synth_text_ids(seed, B, L, vocab, pad_tail)
    hid = O.roberta_encoder(sd, ids, pad)
    keep = ~pad
    assert float(np.abs(hid - g["last_hidden_state"])[keep].max()) <= 5e-5
    text = O.feature_resizer(hid, sd).transpose(1, 0, 2)          # (L, B, 256) as bert.py:69,73
    assert float(np.abs(text - g["text_resized"])[keep.T].max()) <= 5e-5
