"""Golden fixture of the last Video-Swin-T stage from the REAL reference class (build container only):

    python tests/golden/make_golden_swin.py        # writes tests/golden/swin_stage4_*.npz

`BasicLayer(dim=768, depth=2, num_heads=24, window_size=(8,7,7), mlp_ratio=4, qkv_bias=True)` of
vgqa/core/vision/video_swin_transformer.py (the `vid.layers[3]` of VSTGNet, grounding_net.py:67-71) is constructed, loaded with
`synth_swin_stage(seed)` and run on a seeded channels-last map; input (regenerated from the seed) and output are compared by
tests/test_swin_*.py.  The module imports `timm.models.layers` (absent here): a two-function stand-in is installed first."""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import vgqa_oracle as O  # noqa: E402
from ref_loader import load_reference  # noqa: E402

CASES = [("swin_stage4_T16_7x7_s0", 1, 16, 7, 7, 0), ("swin_stage4_B2_T24_7x7_s1", 2, 24, 7, 7, 1)]


def swin_input(seed, B, D, H, W, C=768):
    rng = np.random.Generator(np.random.PCG64(12000 + seed))
    return rng.standard_normal((B, D, H, W, C), dtype=np.float32)


def load_swin_module():
    load_reference()
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm"); models = types.ModuleType("timm.models"); layers = types.ModuleType("timm.models.layers")

        class DropPath(torch.nn.Module):
            def __init__(self, p=0.0):
                super().__init__()

            def forward(self, x):
                return x

        layers.DropPath = DropPath
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models; models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    return importlib.import_module("vgqa.core.vision.video_swin_transformer")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    M = load_swin_module()
    for name, B, D, H, W, seed in CASES:
        layer = M.BasicLayer(dim=768, depth=2, num_heads=24, window_size=(8, 7, 7), mlp_ratio=4.0, qkv_bias=True).eval()
        sd = O.synth_swin_stage(seed, prefix="")
        missing, unexpected = layer.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not unexpected and all("relative_position_index" in m for m in missing), (missing, unexpected)
        x = swin_input(seed, B, D, H, W)
        xt = torch.from_numpy(x).permute(0, 4, 1, 2, 3).contiguous()                    # (B, C, D, H, W) as the backbone feeds it
        with torch.no_grad():
            y = layer(xt)
            with torch.autocast("cpu", dtype=torch.bfloat16):                         # torch's own bf16 deviation: the yardstick
                yb = layer(xt)                                                        # of the CUDA path's tolerance
        ac = (yb.float() - y).abs()
        y = y.permute(0, 2, 3, 4, 1).contiguous().numpy()                             # back to channels-last
        mine = O.swin_stage({("vid.layers.3." + k): v for k, v in sd.items()}, x)
        print(name, "reference |y| mean", float(np.abs(y).mean()), "oracle max-abs diff", float(np.abs(mine - y).max()))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), B=B, D=D, H=H, W=W, seed=seed, torch_version=torch.__version__,
                            y=y.astype(np.float16), y_abs_mean=np.float32(np.abs(y).mean()),
                            autocast_err_mean=np.float32(ac.mean()), autocast_err_max=np.float32(ac.max()),
                            y_rows=y.reshape(-1, 768)[::97].copy())     # every 97th token row in fp32
