"""Golden fixtures of the FRONT END + hot path chain, from the reference's own modules.

    python tests/golden/make_golden_frontend.py        # writes tests/golden/fe_*.npz   (build container only)

`input_proj` / `input_proj2` are `nn.Conv2d(C, 256, kernel_size=1)` exactly as VSTGNet.__init__ builds them
(vgqa/core/grounding_net.py:62,71); the text resizer is the reference's `FeatureResizer` class
(vgqa/core/language/bert.py:77-96) in eval mode.  They get the deterministic synthetic weights
`O.synth_state_dict(seed, front_end_ch=...)`, run on `O.synth_raw_inputs(...)`, and their outputs are fed to the
reference hot-path modules through `make_golden.run_case` (same record layout as the other fixtures, plus the
front end's own outputs).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import vgqa_oracle as O  # noqa: E402
from ref_loader import load_feature_resizer, load_reference  # noqa: E402
import make_golden as MG  # noqa: E402

# name, T, H, W, L, seed, (ResNet ch, Swin ch, RoBERTa hidden)
CASES = [
    ("fe_tiny_T3_3x4_L3", 3, 3, 4, 3, 0, (128, 64, 64)),
    ("fe_cfg1_T32_7x7_L20_s0", 32, 7, 7, 20, 0, (2048, 768, 768)),
    ("fe_yaml_T4_14x14_L20_s1", 4, 14, 14, 20, 1, (2048, 768, 768)),
]


@torch.no_grad()
def run_front_end(FeatureResizer, sd, vis_raw, vid_raw, text_raw, ch):
    ip = torch.nn.Conv2d(ch[0], 256, kernel_size=1).eval()
    ip2 = torch.nn.Conv2d(ch[1], 256, kernel_size=1).eval()
    rs = FeatureResizer(input_feat_size=ch[2], output_feat_size=256, dropout=0.1).eval()
    t = lambda k: torch.from_numpy(sd[k])
    ip.load_state_dict({"weight": t("input_proj.weight"), "bias": t("input_proj.bias")})
    ip2.load_state_dict({"weight": t("input_proj2.weight"), "bias": t("input_proj2.bias")})
    rs.load_state_dict({k[len("text_encoder.resizer."):]: torch.from_numpy(v) for k, v in sd.items()
                        if k.startswith("text_encoder.resizer.")})
    vis = ip(torch.from_numpy(vis_raw)).numpy()
    vid = ip2(torch.from_numpy(vid_raw)).numpy()
    # bert.py:70,73: last_hidden_state (1, L, C) .transpose(0, 1) → resizer → (L, 1, 256)
    text = rs(torch.from_numpy(text_raw)[None].transpose(0, 1)).numpy()
    return vis, vid, text


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    R = load_reference()
    FR = load_feature_resizer()
    for name, T, H, W, L, seed, ch in CASES:
        sd = O.synth_state_dict(seed, front_end_ch=ch)
        vis_raw, vid_raw, text_raw = O.synth_raw_inputs(seed, T, H, W, L, ch)
        vis, vid, text = run_front_end(FR, sd, vis_raw, vid_raw, text_raw, ch)
        extra = dict(front_end_ch=np.asarray(ch, np.int64), fe_text=text,
                     fe_vis_frame0=vis[0], fe_vis_frameN=vis[-1], fe_vid_frame0=vid[0], fe_vid_frameN=vid[-1],
                     fe_vis_abs_mean=np.float32(np.abs(vis).mean()), fe_vid_abs_mean=np.float32(np.abs(vid).mean()))
        MG.run_case(R, name, T, H, W, L, seed, 200, False, HERE, inputs=(vis, vid, text), extra=extra)
