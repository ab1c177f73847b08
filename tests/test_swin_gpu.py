"""GPU: the last Video-Swin-T stage on this library's kernels (csrc/swin.cu: tcgen05 GEMMs + the multi-tile tcgen05 attention with the
relative position bias and the shift mask) against the golden of the reference's own `BasicLayer`
(vgqa/core/vision/video_swin_transformer.py:337-398; tests/golden/make_golden_swin.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_path
from make_golden_swin import swin_input
from oracle import vgqa_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["swin_stage4_T16_7x7_s0", "swin_stage4_B2_T24_7x7_s1"])
def test_swin_stage_matches_reference_golden(name):
    from vgqa_b200.engine import GroundingEngine
    g = np.load(golden_path(name))
    B, D, H, W, seed = (int(g[k]) for k in ("B", "D", "H", "W", "seed"))
    sd = O.synth_state_dict(0)
    sd.update(O.synth_swin_stage(seed))
    eng = GroundingEngine(sd, max_clips=B, max_frames=max(D, 8), max_hw=49, max_text=8)
    x = torch.from_numpy(swin_input(seed, B, D, H, W)).cuda()
    y16, y32 = eng.swin_stage(x, want_f32=True)
    torch.cuda.synchronize()
    got = y32.cpu().numpy()
    ref_rows = g["y_rows"]
    err_rows = np.abs(got.reshape(-1, 768)[::97] - ref_rows)
    err_all = np.abs(got - g["y"].astype(np.float32))
    # bf16 operands, fp32 residual stream; |y| is ~2 on average and up to ~12.  Yardstick: the deviation of torch's OWN bf16
    # autocast run of the same reference module from its fp32 run (stored in the fixture: mean 0.016, max 0.10) — the CUDA path
    # must not be worse than 1.25x that (it measures ≈0.75x: mean 0.012)
    assert float(err_rows.mean()) <= 1.25 * float(g["autocast_err_mean"]), (float(err_rows.mean()), float(g["autocast_err_mean"]))
    assert float(err_rows.max()) <= 1.25 * float(g["autocast_err_max"]), (float(err_rows.max()), float(g["autocast_err_max"]))
    assert float(err_all.max()) <= 1.25 * float(g["autocast_err_max"]) + 8e-3, float(err_all.max())   # + fp16 storage of the full map
    np.testing.assert_allclose(y16.float().cpu().numpy(), got, atol=4e-2)      # the bf16 copy handed to input_proj2
    # the rolled (odd) block really matters: without it the map differs by far more than the tolerance
    assert eng.last_launch_count >= 2 * 7 + 1
    eng.close()


def test_swin_stage_feeds_the_raw_forward():
    """The stage's channels-last bf16 output IS the `vid_raw` (raw_layout = 1) input of the forward."""
    from vgqa_b200.engine import GroundingEngine
    seed, T, H, W, L = 2, 8, 7, 7, 6
    ch = (128, 768, 64)
    sd = O.synth_state_dict(seed, front_end_ch=ch)
    sd.update(O.synth_swin_stage(seed))
    eng = GroundingEngine(sd, max_clips=1, max_frames=T, max_hw=H * W, max_text=L)
    x = torch.from_numpy(swin_input(seed, 1, T, H, W)).cuda()
    vid_map = eng.swin_stage(x)                                                   # [1, T, 7, 7, 768] bf16
    g = torch.Generator(device="cuda").manual_seed(3)
    vis_map = torch.randn(1, T, H, W, ch[0], device="cuda", generator=g).to(torch.bfloat16)
    text = torch.randn(1, L, ch[2], device="cuda", generator=g)
    o = eng.forward(vis_map, vid_map, text, None, raw=True, want=["pred_boxes", "pred_sted"])
    torch.cuda.synchronize()
    assert torch.isfinite(o["pred_boxes"]).all() and torch.isfinite(o["pred_sted"]).all()
    eng.close()


@pytest.mark.parametrize("name", ["swin_full_T16_224_s0", "swin_full_T12_256_s1"])
def test_video_swin_backbone_matches_reference_golden(name):
    """The second fixture exercises the padding path: 12 frames (→ 16) at 256 px (maps 64 / 32 / 16 / 8 → 70 / 35 / 21 / 14).
    The WHOLE Video-Swin-T extractor (patch embedding, four stages with shifted 3-D windows and their masks, PatchMerging) on the
    library's kernels against the golden of the reference's own `VideoSwinTransformerBackbone` (tests/golden/make_golden_swin_full.py):
    every 97th token row of all four stage outputs, and the last stage's map in full.  Yardstick per stage: the deviation of torch's
    own bf16-autocast run of the reference module from its fp32 run (stored in the fixture)."""
    from make_golden_swin_full import swin_frames
    from vgqa_b200.engine import GroundingEngine
    g = np.load(golden_path(name))
    clips, T, R, seed = (int(g[k]) for k in ("clips", "T", "R", "seed"))
    sd = O.synth_state_dict(0)
    sd.update(O.synth_swin_backbone(seed))
    eng = GroundingEngine(sd, max_clips=clips, max_frames=T, max_hw=64, max_text=8)
    frames = torch.from_numpy(swin_frames(seed, clips, T, R)).cuda()
    out, stages = eng.swin_backbone(frames, clips, want_stages=True)
    torch.cuda.synchronize()
    for s in range(4):
        got = stages[s].cpu().numpy().reshape(-1, 96 << s)[::97]
        err = np.abs(got - g[f"rows{s}"])
        assert float(err.mean()) <= 1.25 * float(g[f"autocast_err_mean{s}"]), (s, float(err.mean()), float(g[f"autocast_err_mean{s}"]))
        assert float(err.max()) <= 1.25 * float(g[f"autocast_err_max{s}"]), (s, float(err.max()), float(g[f"autocast_err_max{s}"]))
    y3 = g["y3"].astype(np.float32)
    e3 = np.abs(stages[3].cpu().numpy() - y3)
    assert float(e3.max()) <= 1.25 * float(g["autocast_err_max3"]) + 8e-3
    np.testing.assert_allclose(out.float().cpu().numpy(), stages[3].cpu().numpy(), atol=6e-2)      # the bf16 map handed to input_proj2
    eng.close()


def test_video_swin_backbone_at_384_px_matches_the_oracle():
    """BASELINE configs[4]'s resolution: 384 px → maps 96 / 48 / 24 / 12, none a multiple of 7 (padded to 98 / 49 / 28 / 14), against the
    numpy oracle on the same seeded frames: relative deviation of every stage output of the order of bf16 (measured 0.7 – 0.9 %)."""
    from make_golden_swin_full import swin_frames
    from vgqa_b200.engine import GroundingEngine
    seed, clips, T, R = 3, 1, 8, 384
    sw = O.synth_swin_backbone(seed)
    x = swin_frames(seed, clips, T, R)
    ref = O.video_swin_backbone(sw, x, clips)
    sd = O.synth_state_dict(0)
    sd.update(sw)
    eng = GroundingEngine(sd, max_clips=1, max_frames=T, max_hw=144, max_text=8)
    out, stages = eng.swin_backbone(torch.from_numpy(x).cuda(), clips, want_stages=True)
    torch.cuda.synchronize()
    for s in range(4):
        got = stages[s].cpu().numpy()
        assert got.shape == ref[s].shape
        rel = float(np.abs(got - ref[s]).mean() / np.abs(ref[s]).mean())
        assert rel <= 1.5e-2, (s, rel)
    assert out.shape == (1, T, 12, 12, 768)
    eng.close()
