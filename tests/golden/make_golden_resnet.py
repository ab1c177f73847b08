"""Golden fixtures of the ResNet101 extractor from the reference's own construction (build container only):

    python tests/golden/make_golden_resnet.py        # writes tests/golden/resnet101_*.npz

`torchvision.models.resnet101(norm_layer=FrozenBatchNorm2d, replace_stride_with_dilation=[False, False, False])` wrapped in
`IntermediateLayerGetter(..., {"layer4": ...})` is what `Backbone.__init__` builds (vgqa/core/vision/backbone.py:104-113; its
`pretrained=True` download is replaced by `synth_resnet101(seed)`); FrozenBatchNorm2d is the reference's class (:13-57), imported
from /root/reference.  Stored: the layer4 map in full (fp16), every 61st position of the four layer outputs in fp32, and torch's own
bf16-autocast deviation (the yardstick of the CUDA path's tolerance)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import vgqa_oracle as O  # noqa: E402

CASES = [("resnet101_n2_224_s0", 2, 224, 0), ("resnet101_n3_64_s1", 3, 64, 1)]
ROW_STEP = 61


def resnet_frames(seed, n, R):
    rng = np.random.Generator(np.random.PCG64(16000 + seed))
    return rng.standard_normal((n, 3, R, R), dtype=np.float32)


def build_reference_body():
    import torchvision
    from torchvision.models._utils import IntermediateLayerGetter
    import importlib
    from ref_loader import load_reference
    load_reference()
    FrozenBN = importlib.import_module("vgqa.core.vision.backbone").FrozenBatchNorm2d
    net = torchvision.models.resnet101(weights=None, norm_layer=FrozenBN, replace_stride_with_dilation=[False, False, False])
    return IntermediateLayerGetter(net, return_layers={"layer1": "0", "layer2": "1", "layer3": "2", "layer4": "3"}).eval()


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for name, n, R, seed in CASES:
        body = build_reference_body()
        sd = O.synth_resnet101(seed, prefix="")
        missing, unexpected = body.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        x = resnet_frames(seed, n, R)
        with torch.no_grad():
            out = body(torch.from_numpy(x))
            with torch.autocast("cpu", dtype=torch.bfloat16):
                outb = body(torch.from_numpy(x))
        mine = O.resnet101_backbone({("vis_encoder.0.body." + k): v for k, v in sd.items()}, x)
        rec = dict(n=n, R=R, seed=seed, torch_version=torch.__version__)
        for l in range(4):
            y = out[str(l)]                                                       # (n, C, H, W)
            ycl = y.permute(0, 2, 3, 1).contiguous().numpy()
            ac = (outb[str(l)].float() - y).abs()
            rec[f"rows{l}"] = ycl.reshape(-1, ycl.shape[-1])[::ROW_STEP].copy()
            rec[f"autocast_err_mean{l}"] = np.float32(ac.mean())
            rec[f"autocast_err_max{l}"] = np.float32(ac.max())
            rec[f"abs_mean{l}"] = np.float32(np.abs(ycl).mean())
            print(name, "layer", l + 1, ycl.shape, "|y| mean", float(np.abs(ycl).mean()), "max", float(np.abs(ycl).max()),
                  "oracle max-abs diff", float(np.abs(mine[l] - ycl).max()), "autocast err mean/max", float(ac.mean()), float(ac.max()))
            if l == 3:
                rec["y3"] = ycl.astype(np.float16)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
