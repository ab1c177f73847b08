"""CPU, build container only: the numpy oracle against the reference's own PyTorch modules run LIVE from /root/reference on shapes
and seeds that have no golden file (skipped where the reference tree is absent, e.g. on the GPU box — the committed goldens of
tests/golden/ cover that case)."""
import numpy as np
import pytest
import torch

from oracle import vgqa_oracle as O
from ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not mounted here")


@pytest.mark.parametrize("seed,T,H,W,L", [(11, 5, 3, 3, 4), (12, 9, 2, 5, 6)])
def test_hot_path_live(seed, T, H, W, L):
    from make_golden import RefHotPath, load_synth, make_cfg
    from ref_loader import load_reference
    R = load_reference()
    torch.manual_seed(0)
    model = RefHotPath(R, make_cfg()).eval()
    sd = O.synth_state_dict(seed)
    load_synth(model, sd)
    vis, vid, pos, text = O.synth_inputs(seed, T, H, W, L)
    vm, tm = O.synth_masks(False, T, H, W, L)
    out, dbg = model(torch.from_numpy(vis), torch.from_numpy(vid), torch.from_numpy(pos), torch.from_numpy(text),
                     torch.from_numpy(vm), torch.from_numpy(tm))
    mine = O.hot_path_forward(sd, vis, vid, pos, text, return_debug=True)
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "logits_f_m", "logits_f_a", "logits_r_a", "logits_r_m", "att_sequences"):
        np.testing.assert_allclose(mine[k], out[k].numpy(), atol=2e-4, err_msg=k)
    assert mine["debug"]["choose_pass1"] == dbg["choose_pass1"] and mine["debug"]["choose_pass2"] == dbg["choose_pass2"]


def test_front_end_live():
    from make_golden_frontend import run_front_end
    from ref_loader import load_feature_resizer
    seed, T, H, W, L, ch = 13, 3, 2, 3, 5, (128, 64, 64)
    sd = O.synth_state_dict(seed, front_end_ch=ch)
    raw = O.synth_raw_inputs(seed, T, H, W, L, ch)
    vis, vid, text = run_front_end(load_feature_resizer(), sd, *raw, ch)
    mv, md, mt = O.front_end(sd, *raw)
    np.testing.assert_allclose(mv, vis, atol=2e-5)
    np.testing.assert_allclose(md, vid, atol=2e-5)
    np.testing.assert_allclose(mt, text, atol=2e-5)
