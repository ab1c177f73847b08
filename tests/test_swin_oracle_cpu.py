"""CPU: the numpy restatement of the last Video-Swin-T stage (oracle.swin_stage) against the golden produced by the reference's own
`BasicLayer` (tests/golden/make_golden_swin.py) — fp32 vs fp32."""
import numpy as np
import pytest

from conftest import golden_path
from make_golden_swin import swin_input
from oracle import vgqa_oracle as O


@pytest.mark.parametrize("name", ["swin_stage4_T16_7x7_s0"])
def test_swin_stage_oracle_matches_reference_golden(name):
    g = np.load(golden_path(name))
    B, D, H, W, seed = (int(g[k]) for k in ("B", "D", "H", "W", "seed"))
    y = O.swin_stage(O.synth_swin_stage(seed), swin_input(seed, B, D, H, W))
    np.testing.assert_allclose(y.reshape(-1, 768)[::97], g["y_rows"], atol=2e-4)
    np.testing.assert_allclose(y, g["y"].astype(np.float32), atol=1e-2)       # the full map is stored in fp16


def test_relative_position_index_shape_and_range():
    idx = O.swin_relative_position_index((8, 7, 7))
    assert idx.shape == (392, 392) and idx.min() == 0 and idx.max() == 15 * 13 * 13 - 1
    assert (np.diag(idx) == idx[0, 0]).all()


@pytest.mark.parametrize("name", ["swin_full_T16_224_s0", "swin_full_T12_256_s1"])
def test_video_swin_backbone_oracle_matches_reference_golden(name):
    """The whole extractor (patch embedding, 4 stages with shifted 3-D windows + masks, PatchMerging) against the golden of the
    reference's own VideoSwinTransformerBackbone (every 97th token row of all four stage outputs)."""
    from make_golden_swin_full import swin_frames
    g = np.load(golden_path(name))
    clips, T, R, seed = (int(g[k]) for k in ("clips", "T", "R", "seed"))
    outs = O.video_swin_backbone(O.synth_swin_backbone(seed), swin_frames(seed, clips, T, R), clips)
    for s in range(4):
        np.testing.assert_allclose(outs[s].reshape(-1, 96 << s)[::97], g[f"rows{s}"], atol=3e-4, err_msg=f"stage {s}")
